"""Generate tests/golden/ref_kat.json by RUNNING the parts of the reference that import here.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
The GPU box never runs this; it only reads the committed JSON.

Sources (paths relative to /root/reference):
  utils/val_data.py:36-63          defect_val.image_read      (letterbox + clip window)
  utils/validation_map.py:200-226  MAP.correct_yolo_boxes, MAP.sigmoid
  utils/voc_eval_mask.py:11-56     voc_ap, compute_overlaps_masks
  yolo/config.py                   constants
"""
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference'
sys.path.insert(0, REF)

# utils/validation_map.py imports skimage (absent); the functions we call do not use it.
for name in ('skimage', 'skimage.draw'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['skimage.draw'].polygon = lambda *a, **k: None
sys.modules['skimage'].draw = sys.modules['skimage.draw']

import yolo.config as cfg                                  # noqa: E402
from utils import val_data, validation_map, voc_eval_mask  # noqa: E402

out = {}
out['config'] = dict(
    CLASSES=list(cfg.CLASSES), ANCHORS=np.asarray(cfg.ANCHORS).tolist(), ALPHA=cfg.ALPHA,
    BATCH_SIZE=cfg.BATCH_SIZE, IMAGE_SIZE=cfg.IMAGE_SIZE, K_MAP=cfg.K_MAP, BASE_GRID=cfg.BASE_GRID,
    OBJECT_SCALE=cfg.OBJECT_SCALE, NOOBJECT_SCALE=cfg.NOOBJECT_SCALE, CLASS_SCALE=cfg.CLASS_SCALE,
    COORD_SCALE=cfg.COORD_SCALE, MASK_SCALE=cfg.MASK_SCALE, SCORE_SCALE=cfg.SCORE_SCALE,
    IGNORE_THRESH=cfg.IGNORE_THRESH, OBJ_THRESHOLD=cfg.OBJ_THRESHOLD, IOU_THRESHOLD=cfg.IOU_THRESHOLD,
    TEST_SIZE=cfg.TEST_SIZE, MAX_BOX_PER_IMAGE=cfg.MAX_BOX_PER_IMAGE, MAX_DETECTION=cfg.MAX_DETECTION,
    MAX_ITER=cfg.MAX_ITER, SUMMARY_ITER=cfg.SUMMARY_ITER, SAVE_ITER=cfg.SAVE_ITER,
    FLIPPED=cfg.FLIPPED, BLUR_NOISE_LIGHT=cfg.BLUR_NOISE_LIGHT, GPU=cfg.GPU)

# --- letterbox (image_read does not touch self beyond image_size) ---
dv = val_data.defect_val.__new__(val_data.defect_val)
dv.image_size = 576
lb = []
imdir = os.path.join(REF, 'data/train_sample/images')
import cv2  # noqa: E402
for nm in sorted(os.listdir(imdir)):
    img, win = dv.image_read(os.path.join(imdir, nm))
    h, w = cv2.imread(os.path.join(imdir, nm)).shape[:2]
    lb.append(dict(name=nm, h=int(h), w=int(w), shape=list(img.shape), window=[float(v) for v in win],
                   pad_value=float(img[0, 0, 0]) if win[0] > 0 else float(img[0, 0, 0]),
                   mean=float(img.mean()), corner=float(img[0, 0, 0])))
out['letterbox'] = lb

# --- correct_yolo_boxes / sigmoid ---
m = validation_map.MAP.__new__(validation_map.MAP)
cases = [(.2, .3, .9, .7, 348, 620, 576, 576), (0., 0., 1., 1., 348, 620, 576, 576),
         (.1, .25, .5, .75, 620, 348, 576, 576), (.3, .3, .31, .9, 480, 640, 576, 576),
         (.05, .6, .95, .99, 1000, 1000, 576, 576)]
out['correct_yolo_boxes'] = [dict(args=list(c), out=[int(v) for v in m.correct_yolo_boxes(*c)]) for c in cases]
xs = [0., 100., -100., 1.5, -3.25, 50., -50.]
out['sigmoid'] = dict(x=xs, y=[float(v) for v in m.sigmoid(np.array(xs))])

# --- mask IoU / AP ---
a = np.zeros((8, 8, 1), bool); a[2:6, 2:6, 0] = True
b = np.zeros((8, 8, 2), bool); b[2:6, 2:4, 0] = True; b[:, :, 1] = True
out['compute_overlaps_masks'] = voc_eval_mask.compute_overlaps_masks(a, b).tolist()
out['voc_ap'] = dict(rec=[.5, .5, 1.], prec=[1., .5, 2. / 3],
                     ap=float(voc_eval_mask.voc_ap(np.array([.5, .5, 1.]), np.array([1., .5, 2. / 3]), False)))

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_kat.json')
json.dump(out, open(dst, 'w'), indent=1, sort_keys=True)
print('wrote', dst)

# --- letterbox on seeded synthetic images through the reference's own image_read (bit-level pin for
# oracle/dis_oracle_io.letterbox; the GPU kernel dy_letterbox is then compared with that oracle) ---
import hashlib  # noqa: E402
import tempfile  # noqa: E402
syn = []
for seed, (h, w) in enumerate([(348, 620), (700, 500), (576, 576), (1000, 1001), (97, 1300)]):
    rgb = np.random.default_rng(100 + seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    path = os.path.join(tempfile.gettempdir(), 'dy_golden_%d.png' % seed)
    cv2.imwrite(path, rgb[:, :, ::-1])                      # image_read reads BGR and converts to RGB
    img, win = dv.image_read(path)
    f32 = np.ascontiguousarray(img, np.float32)
    ys, xs = [0, 100, 288, 400, 575], [0, 57, 288, 431, 575]
    syn.append(dict(seed=100 + seed, h=h, w=w, window=[float(v) for v in win],
                    sha256_f32=hashlib.sha256(f32.tobytes()).hexdigest(),
                    samples=[[y, x, [float(v) for v in f32[y, x]]] for y, x in zip(ys, xs)]))
out['letterbox_synthetic'] = syn
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_kat.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
print('letterbox_synthetic:', len(syn), 'cases')

# --- mask-level mAP: the reference's own voc_eval (utils/voc_eval_mask.py:58-134) on seeded synthetic data
# sets (oracle/dis_oracle_eval.synthetic_dataset); pins oracle/dis_oracle_eval.voc_eval ---
if not hasattr(np, 'bool'):
    np.bool = bool            # the reference predates NumPy 1.24 (np.bool alias, voc_eval_mask.py:80)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dis_oracle_eval as OE  # noqa: E402
ve = []
for seed in (1, 2, 3):
    names, recs, dets = OE.synthetic_dataset(seed)
    setfile = os.path.join(tempfile.gettempdir(), 'dy_golden_set_%d.txt' % seed)
    with open(setfile, 'w') as f:
        f.write('\n'.join(names) + '\n')
    for c in range(3):
        for use07 in (False, True):
            if not dets[c]:
                continue
            import copy  # noqa: E402
            r, p, ap = voc_eval_mask.voc_eval(copy.deepcopy(dets[c]), copy.deepcopy(recs), setfile, c, ovthresh=0.5,
                                              use_07_metric=use07)
            ve.append(dict(seed=seed, classid=c, use_07_metric=use07, recall=float(r), precision=float(p), ap=float(ap)))
out['voc_eval_synthetic'] = ve
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_kat.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
print('voc_eval_synthetic:', len(ve), 'cases', ve[:2])
