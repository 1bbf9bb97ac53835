"""``YOLONet`` + ``Session``: the reference's builder surface on top of libdisyolo_b200.

The reference's ``YOLONet`` (yolo/yolo3_net_pos.py:12-65) builds a TensorFlow graph whose
placeholders and fetchable tensors are attributes of the object; its drivers then call
``sess.run(fetches, feed_dict)`` (train_yolo3_mask.py:146-176, calculate_test_map.py:214-218).
This module keeps exactly that calling convention:

    import disyolo_b200.yolo.config as cfg
    from disyolo_b200.yolo.yolo3_net_pos import YOLONet, Session
    cfg.BATCH_SIZE = 1
    net = YOLONet(False)
    sess = Session(net); sess.restore(weights)          # dict keyed by TF variable names
    det_box, det_mask = sess.run(net.evaluation, feed_dict={net.is_training: False,
                                 net.det_thresh: [cfg.OBJ_THRESHOLD], net.clip_window: windows,
                                 net.images: images})

Attributes are opaque handles (``Handle``) instead of tf.Tensors; ``Session.run`` resolves them.
The per-layer ``lock`` flags, which are literals inside the reference's ``build_network``
(:155-156), are exposed as ``YOLONet.lock`` (default: 1-52 locked, 53-82 trainable).
"""
import numpy as np

from . import config as cfg
from ..engine import Engine


class Handle(object):
    """Stand-in for a tf.placeholder / fetchable tf.Tensor."""

    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    def __repr__(self):
        return '<disyolo %s %s>' % (self.kind, self.name)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class AdamOptimizer(object):
    """tf.train.AdamOptimizer(learning_rate).minimize(net.total_loss) (train_yolo3_mask.py:55)."""

    def __init__(self, learning_rate=1e-4):
        self.learning_rate = float(learning_rate)

    def minimize(self, loss, global_step=None):
        if not isinstance(loss, Handle) or loss.name != 'total_loss':
            raise ValueError('minimize() expects net.total_loss')
        h = Handle('adam_minimize', 'op')
        h.optimizer = self
        return h


class YOLONet(object):
    def __init__(self, training=False, precision=None, device=None, lock=None):
        # 1. parameters (yolo3_net_pos.py:15-37)
        self.batchsize = cfg.BATCH_SIZE
        self.classes = cfg.CLASSES
        self.num_class = len(self.classes)
        self.anchors = cfg.ANCHORS
        self.num_anchor = 3
        self.output_depth = (self.num_class + 5) * self.num_anchor
        self.k = cfg.K_MAP
        self.k_mapout = self.k * self.k
        self.object_scale = cfg.OBJECT_SCALE
        self.noobject_scale = cfg.NOOBJECT_SCALE
        self.class_scale = cfg.CLASS_SCALE
        self.coord_scale = cfg.COORD_SCALE
        self.mask_scale = cfg.MASK_SCALE
        self.training = bool(training)
        self.image_size = int(cfg.IMAGE_SIZE)
        self.lock = list(lock) if lock is not None else [True] * 52 + [False] * 30

        # placeholders (:41-44, :52-57)
        self.is_training = Handle('training', 'placeholder')
        self.det_thresh = Handle('object_threshold', 'placeholder')
        self.clip_window = Handle('clip_window', 'placeholder')
        self.images = Handle('images', 'placeholder')
        if training:
            self.yolo1 = Handle('yolo1', 'placeholder')
            self.yolo2 = Handle('yolo2', 'placeholder')
            self.yolo3 = Handle('yolo3', 'placeholder')
            self.labels_value = [self.yolo3, self.yolo2, self.yolo1]
            self.true_boxes = Handle('true_boxes', 'placeholder')
            self.true_masks = Handle('true_masks', 'placeholder')
            self.total_loss = Handle('total_loss', 'fetch')

        # fetchables: logits = [predictions, detections, mask_pos] (:47, :463), evaluation (:65)
        self.predictions = Handle('predictions', 'fetch')
        self.detections = Handle('detections', 'fetch')
        self.mask_pos = Handle('mask_pos', 'fetch')
        self.logits = [self.predictions, self.detections, self.mask_pos]
        self.evaluation = Handle('evaluation', 'fetch')

        if precision is None:
            # inference and training both run on the tensor-core (bf16) engine, whatever the lock pattern: the
            # reference's stage 1 (backbone locked) and stage 2 ("lock=False for all layers", :155-156);
            # precision='fp32' selects the CUDA-core verification engine
            precision = 'bf16'
        dev = int(cfg.GPU) if device is None else int(device)
        if int(cfg.MAX_BOX_PER_IMAGE) != 20:
            # true_boxes [B,1,1,1,20,5] / true_masks [B,20,H,W] (:56-57): the loss kernels are compiled for 20
            raise ValueError('cfg.MAX_BOX_PER_IMAGE must be 20 (the shape of true_boxes / true_masks)')
        self.engine = Engine(image_size=self.image_size, max_batch=int(self.batchsize), precision=precision,
                             device=dev, anchors=self.anchors, num_classes=self.num_class, k_map=self.k,
                             alpha=cfg.ALPHA, iou_threshold=cfg.IOU_THRESHOLD,
                             max_detection=cfg.MAX_DETECTION, lock=self.lock)
        # loss scales / ignore threshold as read from cfg at construction (:30-35, loss_yolo :631-747)
        self.engine.set_loss_params(self.object_scale, self.noobject_scale, self.class_scale, self.coord_scale,
                                    self.mask_scale, cfg.IGNORE_THRESH)


class Session(object):
    """The subset of tf.Session the reference's drivers use."""

    def __init__(self, net, seed=0):
        self.net = net
        self.restored = False
        self.rng = np.random.default_rng(seed)      # stands in for the unseeded tf.random_shuffle (:781-782)
        self.init_seed = seed                       # ... and for the unseeded variable initialisers
        self.initialized_from_init = []
        self.train_ready = False
        self.last_losses = None

    def restore(self, weights):
        """weights: dict {tf variable name: ndarray} (what Saver.restore would read,
        train_yolo3_mask.py:104-111), the prefix of a TensorFlow checkpoint-V2 bundle (`model.ckpt-500`:
        read without TensorFlow by tf_checkpoint.read_checkpoint; variables the net does not have are
        ignored like slim's ignore_missing_vars=True, :106), or a path to an .npz written by
        weights.save_npz."""
        if isinstance(weights, str):
            from .. import tf_checkpoint
            if tf_checkpoint.is_checkpoint_prefix(weights):
                weights = tf_checkpoint.network_variables(tf_checkpoint.read_checkpoint(weights))
            else:
                from ..weights import load_npz
                weights = load_npz(weights)
        if not self.restored:
            # The reference runs global_variables_initializer() and THEN assign_from_checkpoint_fn(...,
            # ignore_missing_vars=True) (train_yolo3_mask.py:63,104-107): variables the checkpoint lacks
            # (yolov3_3class_coco.ckpt has no convolutional76..82) keep their initial value -- xavier /
            # zero bias for unlocked layers, truncated normal for locked ones, BatchNorm at identity.
            from ..weights import init_weights
            init = init_weights('reference', seed=self.init_seed, lock=self.net.lock)
            missing = {k: v for k, v in init.items() if k not in weights}
            if missing:
                weights = dict(weights)
                weights.update(missing)
            self.initialized_from_init = sorted(missing)
        self.net.engine.load_weights(weights)
        self.restored = True

    def save(self, prefix):
        """Saver.save counterpart (train_yolo3_mask.py:221-226): the current value of every
        yolo/convolutional{N}/... variable as a TensorFlow checkpoint-V2 bundle at `prefix`."""
        from .. import tf_checkpoint
        from ..weights import variable_names
        from ..engine import layer_table
        eng, out = self.net.engine, {}
        for L in layer_table():
            k, cin, cout = L['k'], L['cin'], L['cout']
            for name in variable_names(L['id'], L['bn']):
                shape = (k, k, cin, cout) if name.endswith('/weights') else (cout,)
                out[name] = eng.get_weights(name, shape)
        tf_checkpoint.write_checkpoint(prefix, out)
        return prefix

    def run(self, fetches, feed_dict=None):
        if not self.restored:
            raise RuntimeError('Session.restore(weights) must be called before run')
        net, feed = self.net, (feed_dict or {})
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        if any(getattr(f, 'name', '') in ('total_loss', 'adam_minimize') for f in flist):
            return self._run_train(flist, feed, single)
        images = np.ascontiguousarray(feed[net.images], np.float32)
        if images.shape[1] != net.image_size or images.shape[2] != net.image_size:
            raise ValueError('images must be [B,%d,%d,3]' % (net.image_size, net.image_size))
        windows = np.ascontiguousarray(feed[net.clip_window], np.float32)
        thresh = float(np.asarray(feed[net.det_thresh]).reshape(-1)[0])
        B = images.shape[0]
        eng = net.engine
        cache = {}

        def evaluation():
            if 'eval' not in cache:
                raw, box, cnt, msk = eng.forward_host(images, windows, thresh, want_masks=True)
                cnt_np = cnt.numpy()
                det_box = [box[b, :cnt_np[b]].numpy().copy() for b in range(B)]
                # no proposal -> scalar 0.0 (tf.constant(0.0), yolo3_net_pos.py:933)
                det_mask = [msk[b, :cnt_np[b]].numpy().copy() if cnt_np[b] > 0 else np.float32(0.0)
                            for b in range(B)]
                cache['eval'] = [det_box, det_mask]
                cache['raw'] = raw.numpy().copy()
            return cache['eval']

        out = []
        for f in flist:
            if f is net.evaluation:
                out.append(evaluation())
            elif f is net.detections:
                evaluation()
                out.append(cache['raw'])
            elif f is net.mask_pos:
                evaluation()
                out.append(eng.mask_pos(B).cpu().numpy())
            elif f is net.predictions:
                evaluation()
                out.append([eng.yolo(s, B).cpu().numpy() for s in range(3)])
            elif isinstance(f, list) and f is net.logits:
                evaluation()
                out.append([[eng.yolo(s, B).cpu().numpy() for s in range(3)], cache['raw'],
                            eng.mask_pos(B).cpu().numpy()])
            else:
                raise KeyError('unknown fetch %r' % (f,))
        return out[0] if single else out

    def _run_train(self, flist, feed, single):
        """sess.run([net.total_loss, optimizer], feed_dict) (train_yolo3_mask.py:146-149,216)."""
        net, eng = self.net, self.net.engine
        if not net.training:
            raise RuntimeError('YOLONet(training=True) is required for loss / optimizer fetches')
        if not self.train_ready:
            eng.train_init()
            self.train_ready = True
        images = np.ascontiguousarray(feed[net.images], np.float32)
        B = images.shape[0]
        tb = np.ascontiguousarray(feed[net.true_boxes], np.float32).reshape(B, -1, 5)
        tm = np.ascontiguousarray(feed[net.true_masks]).astype(np.uint8)
        labels = [feed[net.yolo3], feed[net.yolo2], feed[net.yolo1]]
        thresh = float(np.asarray(feed[net.det_thresh]).reshape(-1)[0])
        pp = np.stack([self.rng.permutation(eng.max_detection) for _ in range(B)]).astype(np.int32)
        pg = np.stack([self.rng.permutation(20) for _ in range(B)]).astype(np.int32)
        losses = eng.train_forward(images, labels, tb, tm, pp, pg, thresh)
        self.last_losses = dict(zip(('total', 'object', 'noobject', 'class', 'xy', 'wh', 'mask', 'l2'),
                                    [float(v) for v in losses]))
        out = []
        for f in flist:
            if f.name == 'total_loss':
                out.append(float(losses[0]))
            elif f.name == 'adam_minimize':
                eng.train_backward(82, 1)
                eng.train_apply(f.optimizer.learning_rate, 1.0)
                out.append(None)
            else:
                raise KeyError('fetch %r cannot be combined with a training fetch' % (f,))
        return out[0] if single else out
