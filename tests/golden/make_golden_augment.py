"""Generate tests/golden/augment_kat.json by RUNNING the reference's own training-data methods
(utils/train_data.py, class defect_train) on seeded inputs.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden_augment.py
The GPU box never runs this; it only reads the committed JSON.  utils/train_data.py imports pyblur and
skimage.draw (absent here): both are stubbed; no golden case calls into them (load_mask and
linearmotion_blur3C are therefore NOT pinned, see oracle/dis_oracle_augment.py).

Every case stores the generator seed and parameters plus the sha256 of the reference's output (and a few
sampled values); tests regenerate the inputs from the seed, so the fixture stays a few kilobytes.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference'
sys.path.insert(0, REF)
for name in ('skimage', 'skimage.draw', 'pyblur'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['skimage'].draw = sys.modules['skimage.draw']
sys.modules['skimage.draw'].polygon = None
sys.modules['pyblur'].__all__ = []

from utils import train_data  # noqa: E402

dt = train_data.defect_train.__new__(train_data.defect_train)
dt.image_size = 576
dt.max_box_per_image = 20


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def make_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def make_masks(seed, n, h, w):
    """n blob masks (unions of ellipses) as float32 0/1, [20, h, w] with the first n non-empty."""
    rng = np.random.default_rng(seed)
    m = np.zeros((20, h, w), np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):
        for _ in range(int(rng.integers(1, 4))):
            cy, cx = rng.uniform(0.15, 0.85) * h, rng.uniform(0.15, 0.85) * w
            ry, rx = rng.uniform(0.03, 0.2) * h, rng.uniform(0.03, 0.2) * w
            m[i][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = 1.0
    return m


out = {}
# (seed, h, w, new_w, new_h, dx, dy): letterbox placement, up-scale with crop, down-scale with offsets, identity
PLACE = [(1, 348, 620, 576, 323, 0, 126), (2, 620, 348, 323, 576, 126, 0), (3, 300, 400, 864, 648, -150, -40),
         (4, 480, 640, 432, 324, 100, 200), (5, 240, 320, 691, 518, -60, 30), (6, 576, 576, 576, 576, 0, 0),
         (7, 97, 131, 700, 518, -124, 58)]
cases = []
for seed, h, w, nw, nh, dx, dy in PLACE:
    img = make_image(seed, h, w)
    placed = dt.apply_random_scale_and_crop(img, nw, nh, dx, dy, 'image')
    rec = dict(seed=seed, h=h, w=w, new_w=nw, new_h=nh, dx=dx, dy=dy, placed_sha=sha(placed), reads={})
    for flip in (1, 2, 3):
        res = dt.image_read(img.copy(), [0, nw, nh, dx, dy], flip, 1, 'image')
        assert res.dtype == np.float32 and res.shape == (576, 576, 3)
        rec['reads'][str(flip)] = dict(sha=sha(res), sample=[float(v) for v in res[::97, ::89, 1].ravel()[:12]])
    masks = make_masks(100 + seed, 3, h, w)
    rec['masks'] = {}
    for flip in (1, 2, 3):
        rm = dt.resize_mask(masks, [0, nw, nh, dx, dy], flip, np.array([0, 1, 2]), 'mask')
        assert rm.dtype == bool and rm.shape == (20, 576, 576)
        rec['masks'][str(flip)] = dict(sha=sha(rm.astype(np.uint8)), count=int(rm.sum()))
    cases.append(rec)
out['place'] = cases

# extract_bboxes / load_box
boxes = []
for seed in (11, 12, 13):
    m = make_masks(seed, 4, 200, 260)
    boxes.append(dict(seed=seed, boxes=[[int(v) for v in dt.extract_bboxes(m[i].astype(np.uint8))] for i in range(4)]))
out['bboxes'] = boxes

# noise and light on a placed 576 x 576 image, with the reference's own np.random draws
noise = []
for seed in (21, 22):
    img = make_image(seed, 576, 576)
    np.random.seed(seed)
    sp = dt.add_salt_pepper_noise(img.copy())
    np.random.seed(seed)
    lt = dt.change_light(img.copy())
    noise.append(dict(seed=seed, salt_pepper_sha=sha(sp), light_sha=sha(lt), ones=int((sp == 1).all(-1).sum()),
                      zeros=int((sp == 0).all(-1).sum())))
out['noise'] = noise

# image_read with the noise / light switches (bnl 2 = salt & pepper, 3 = light in image_read's own mapping)
reads = []
for seed, bnl in ((31, 2), (32, 3)):
    img = make_image(seed, 300, 400)
    np.random.seed(seed)
    res = dt.image_read(img.copy(), [0, 576, 432, 0, 72], 2, bnl, 'image')
    reads.append(dict(seed=seed, bnl=bnl, sha=sha(res)))
out['image_read_bnl'] = reads

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'augment_kat.json')
json.dump(out, open(dst, 'w'), indent=1, sort_keys=True)
print('wrote', dst)
