"""Tiny forward + training step + I/O kernels for compute-sanitizer (memcheck): never a bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import disyolo_b200 as dy
from oracle import dis_oracle as O, dis_oracle_train as T
size, B = 96, 2
W = O.make_weights('lively', 0)
rng = np.random.default_rng(0)
img = rng.random((B, size, size, 3), dtype=np.float32)
win = np.tile(np.array([[0, 0, 1, 1]], np.float32), (B, 1))
eng = dy.Engine(image_size=size, max_batch=B, precision='bf16')
eng.load_weights(W)
out = eng.forward(torch.from_numpy(img).cuda(), torch.from_numpy(win).cuda(), 0.1)
torch.cuda.synchronize()
print('forward ok, dets', out['det_count'].tolist())
labels, tb, tm = T.make_labels(rng, B, size)
pp = np.stack([rng.permutation(30) for _ in range(B)]).astype(np.int32)
pg = np.stack([rng.permutation(20) for _ in range(B)]).astype(np.int32)
eng.train_init()
l = eng.train_forward(img, labels, tb[:, 0, 0, 0], tm, pp, pg, 0.1)
eng.train_backward()
eng.train_apply(1e-4)
torch.cuda.synchronize()
print('train step ok, loss', float(l[0]))
rgb = rng.integers(0, 256, (70, 130, 3), dtype=np.uint8)
lb, w = eng.letterbox(rgb)
r = eng.postprocess(out['det_box'][0], out['det_count'][0:1], out['masks'][0], 70, 130)
torch.cuda.synchronize()
print('io ok', int(r['valid'].sum()))
