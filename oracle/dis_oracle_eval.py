"""CPU oracle for the mask-level mAP  --  TEST INFRASTRUCTURE ONLY (see dis_oracle.py).

NumPy restatement of utils/voc_eval_mask.py: compute_overlaps_masks (:38-56), voc_ap (:9-36), voc_eval
(:58-134).  PARITY STATUS: PINNED -- the reference's own voc_eval / compute_overlaps_masks (pure NumPy,
importable) were run in the build container on seeded synthetic data sets and their outputs are committed
in tests/golden/ref_kat.json ('voc_eval_synthetic', 'compute_overlaps_masks').
"""
import numpy as np


def compute_overlaps_masks(masks1, masks2):
    """masks [H, W, instances] -> IoU [n1, n2] (float32 arithmetic, like the reference)."""
    if masks1.shape[-1] == 0 or masks2.shape[-1] == 0:
        return np.zeros((masks1.shape[-1], masks2.shape[-1]))
    masks1 = np.reshape(masks1 > .5, (-1, masks1.shape[-1])).astype(np.float32)
    masks2 = np.reshape(masks2 > .5, (-1, masks2.shape[-1])).astype(np.float32)
    area1, area2 = np.sum(masks1, axis=0), np.sum(masks2, axis=0)
    intersections = np.dot(masks1.T, masks2)
    union = area1[:, None] + area2[None, :] - intersections
    with np.errstate(invalid='ignore', divide='ignore'):
        return intersections / union


def voc_ap(rec, prec, use_07_metric=False):
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def voc_eval(detfile, recs, imagenames, classid, ovthresh=0.5, use_07_metric=False):
    class_recs, npos = {}, 0
    for imagename in imagenames:
        R = [obj for obj in recs[imagename] if obj['classid'] == classid]
        bbox = np.concatenate([np.expand_dims(x['mask'], -1) for x in R], -1) if len(R) > 0 else np.array([])
        difficult = np.array([x['difficult'] for x in R]).astype(bool)
        npos = npos + sum(~difficult)
        class_recs[imagename] = {'mask': bbox, 'difficult': difficult, 'det': [False] * len(R)}
    image_ids = [x['imageid'] for x in detfile]
    confidence = np.array([float(x['score']) for x in detfile])
    BB = [np.expand_dims(x['mask'], -1) for x in detfile]
    sorted_ind = np.argsort(-confidence)
    if len(sorted_ind) == 0:
        return 0., 0., 0.
    BB = [BB[x] for x in sorted_ind]
    image_ids = [image_ids[x] for x in sorted_ind]
    nd = len(image_ids)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        ovmax, jmax = -np.inf, -1
        BBGT = R['mask'].astype(float)
        if BBGT.size > 0:
            overlaps = compute_overlaps_masks(BB[d].astype(float), BBGT)
            ovmax, jmax = np.max(overlaps[0]), np.argmax(overlaps[0])
        if ovmax > ovthresh:
            if not R['difficult'][jmax]:
                if not R['det'][jmax]:
                    tp[d] = 1.
                    R['det'][jmax] = 1
                else:
                    fp[d] = 1.
        else:
            fp[d] = 1.
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return tp[-1] / float(npos), tp[-1] / np.maximum(tp[-1] + fp[-1], np.finfo(np.float64).eps), ap


def synthetic_dataset(seed, n_images=6, h=96, w=128, num_class=3):
    """Seeded ground truth + detections: blobs, jittered / dropped / duplicated detections, a few difficult
    objects.  Returns (imagenames, recs, detfiles {classid: detfile})."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    names = ['img%03d' % i for i in range(n_images)]
    recs, dets = {}, {c: [] for c in range(num_class)}
    for name in names:
        objs = []
        for _ in range(int(rng.integers(1, 5))):
            cy, cx, ry, rx = rng.uniform(10, h - 10), rng.uniform(10, w - 10), rng.uniform(5, 25), rng.uniform(5, 30)
            m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            c = int(rng.integers(0, num_class))
            objs.append({'classid': c, 'difficult': int(rng.random() < 0.15), 'mask': m, 'imageid': name})
            for _ in range(int(rng.integers(0, 3))):          # 0-2 detections per object (misses, duplicates)
                dy_, dx_ = rng.integers(-6, 7, 2)
                dm = np.roll(np.roll(m, int(dy_), 0), int(dx_), 1)
                dets[c].append({'imageid': name, 'score': float(rng.uniform(0.2, 1.0)), 'mask': dm})
        if rng.random() < 0.5:                                 # a false positive somewhere
            c = int(rng.integers(0, num_class))
            fpm = np.zeros((h, w), bool)
            y0, x0 = int(rng.integers(0, h - 12)), int(rng.integers(0, w - 12))
            fpm[y0:y0 + 10, x0:x0 + 10] = True
            dets[c].append({'imageid': name, 'score': float(rng.uniform(0.2, 1.0)), 'mask': fpm})
        recs[name] = objs
    return names, recs, dets
