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

# CTA-pair plans (tcgen05.mma.cta_group::2 on 2-CTA clusters), split-N epilogue and halo'd boxes forced at this small size
from disyolo_b200.engine import set_option
for k, v in (('tc_cta2', 1), ('tc_split_n', 1), ('tc_halo', 1)):
    set_option(k, v)
e2 = dy.Engine(image_size=size, max_batch=B, precision='bf16')
e2.load_weights(W)
o2 = e2.forward(torch.from_numpy(img).cuda(), torch.from_numpy(win).cuda(), 0.1)
torch.cuda.synchronize()
for k in ('tc_cta2', 'tc_split_n', 'tc_halo'):
    set_option(k, -1)
print('forced CTA-pair forward ok, dets', o2['det_count'].tolist(), 'same detections as default plans:',
      bool((o2['det_count'] == out['det_count']).all()))
e2.close()
# batched NMS: several thousand candidates per class (radix-select / multi-batch path), 300 selections per class
e3 = dy.Engine(image_size=416, max_batch=1, precision='fp32', max_detection=300)
g = (52, 26, 13)
yol = [rng.standard_normal((1, s_, s_, 3, 8)).astype(np.float32) for s_ in g]
for y in yol:
    y[..., 2:4] -= 2.0
raw, box, cnt = e3.detect(yol, np.array([[0, 0, 1, 1]], np.float32), 0.01)
torch.cuda.synchronize()
print('big NMS ok, detections', cnt.tolist())
e3.close()
