"""disyolo_b200 -- the DIS-YOLO hot path (YOLOv3/Darknet-53 + position-sensitive mask subnet:
forward, decode, NMS, mask assembly) on NVIDIA B200, as hand-written sm_100a CUDA behind a C ABI.

The directory is named ``dis-yolo_b200`` (the project's name); Python cannot import a hyphenated
name, so the importable alias package ``disyolo_b200`` (repo root) points its ``__path__`` here.

Host-side mirror of the reference's interface:
    disyolo_b200.yolo.config          <- yolo/config.py           (same constant names)
    disyolo_b200.yolo.yolo3_net_pos   <- yolo/yolo3_net_pos.py    (class YOLONet, same attributes)
    disyolo_b200.Session              <- the tf.Session.run(fetches, feed_dict) protocol the
                                         reference's drivers use (train_yolo3_mask.py:158-216,
                                         calculate_test_map.py:214-218)
PyTorch is used only for device/pinned buffers and streams.  There is no CPU fallback: importing
works anywhere, but constructing a net without the built library or without a GPU raises.
"""
from . import _lib                                    # noqa: F401
from .engine import Engine, TrainData, layer_table    # noqa: F401
from .weights import init_weights, save_npz, load_npz, variable_names  # noqa: F401
from .yolo.yolo3_net_pos import YOLONet, Session, AdamOptimizer      # noqa: F401
from .parallel import DataParallelTrainer, plan_buckets, BucketedAllReduce   # noqa: F401
from .pipeline import ImagePipeline                   # noqa: F401
from . import tf_checkpoint                           # noqa: F401

__all__ = ['Engine', 'TrainData', 'ImagePipeline', 'YOLONet', 'Session', 'AdamOptimizer', 'DataParallelTrainer', 'plan_buckets', 'init_weights', 'save_npz', 'load_npz',
           'variable_names', 'layer_table']
