"""Engine (bf16, training-mode forward) vs the bf16-emulating oracle, layer by layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import disyolo_b200 as dy
from oracle import dis_oracle as O, dis_oracle_train as T
from tests.test_gpu_train import _setup
from tests.util import rel_err

W, img, labels, tb, tm, pp, pg, thresh = _setup()
B, size = img.shape[0], img.shape[1]
lock = O.default_lock_flags()
perms = [(pp[b].tolist(), pg[b].tolist()) for b in range(B)]
T.NP_DT = np.float64
plain = T.train_step(img, W, lock, labels, tb, tm, perms, det_thresh=thresh, apply=False)[4]
T.EMULATE_BF16 = True
emul = T.train_step(img, W, lock, labels, tb, tm, perms, det_thresh=thresh, apply=False)[4]
eng = dy.Engine(image_size=size, max_batch=B, precision='bf16')
eng.load_weights(W)
eng.train_init()
eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, thresh)
for n in range(1, 83):
    y = eng.activation(n, B).cpu().numpy()
    print('%2d  eng-emul %.2e  eng-plain %.2e  emul-plain %.2e' % (n, rel_err(y, emul['acts'][n]), rel_err(y, plain['acts'][n]),
                                                       rel_err(emul['acts'][n], plain['acts'][n])))
