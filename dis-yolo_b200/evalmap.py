"""Mask-level VOC AP with the mask IoUs computed on the GPU  (evaluation part).

Counterpart of utils/voc_eval_mask.py: `voc_ap` (:9-36) and `voc_eval` (:58-134) keep the reference's
signatures and results -- a detection is a true positive iff its best-IoU ground-truth mask (> ovthresh,
strict) is not "difficult" and no higher-scored detection of the same image claimed it; VOC-2010
area-under-envelope AP (or the 11-point VOC-2007 one).  The host side is written image by image on whole
arrays (no per-detection Python loop); the one hot operation, `compute_overlaps_masks` (:38-56: a
[pixels x n_det]^T [pixels x n_gt] product per detection), is ONE dy_mask_overlaps launch per image
covering all of that image's detections of the class.
There is no CPU fallback: an Engine (GPU) is required.
"""
import numpy as np


def voc_ap(rec, prec, use_07_metric=False):
    """Area under the precision envelope (VOC 2010+) or the 11-point VOC-2007 average; same results as
    utils/voc_eval_mask.py:9-36, computed without Python loops."""
    rec = np.asarray(rec, np.float64)
    prec = np.asarray(prec, np.float64)
    if use_07_metric:
        # max precision at recall >= t for t = 0, 0.1, ..., 1.0: suffix maximum looked up by binary search
        order = np.argsort(rec, kind='stable')
        r_sorted = rec[order]
        suffix_max = np.maximum.accumulate(prec[order][::-1])[::-1] if rec.size else prec
        pos = np.searchsorted(r_sorted, np.arange(0., 1.1, 0.1), side='left')
        p = np.where(pos < rec.size, suffix_max[np.minimum(pos, max(rec.size - 1, 0))] if rec.size else 0., 0.)
        ap = 0.
        for v in p:                      # the reference's accumulation order (11 additions of p / 11.)
            ap = ap + v / 11.
        return ap
    r = np.concatenate(([0.], rec, [1.]))
    envelope = np.maximum.accumulate(np.concatenate(([0.], prec, [0.]))[::-1])[::-1]
    step = np.flatnonzero(r[1:] != r[:-1])
    return np.sum((r[step + 1] - r[step]) * envelope[step + 1])


def _mark_image(order, best_iou, best_gt, difficult, ovthresh):
    """True / false positives of ONE image's detections.  `order`: their ranks in the global descending-score
    order; a ground-truth mask is claimed by the first (highest ranked) detection whose best IoU exceeds the
    threshold, later claimants are false positives, matches to "difficult" objects count as neither."""
    n = len(order)
    tp, fp = np.zeros(n, bool), np.zeros(n, bool)
    if best_iou is None:                     # the image has no ground truth of this class
        fp[:] = True
        return tp, fp
    hit = best_iou > ovthresh
    fp[~hit] = True
    cand = hit & ~difficult[best_gt]
    by_rank = np.argsort(order, kind='stable')
    claimed = best_gt[by_rank][cand[by_rank]]
    first = np.zeros(claimed.size, bool)
    first[np.unique(claimed, return_index=True)[1]] = True
    idx = by_rank[cand[by_rank]]
    tp[idx[first]] = True
    fp[idx[~first]] = True
    return tp, fp


def voc_eval(detfile, recs, imagesetfile, classid, ovthresh=0.5, use_07_metric=False, engine=None):
    """Signature and results of utils/voc_eval_mask.py:58-134.  detfile: list of {'imageid', 'score', 'mask'
    [h,w] bool (numpy or cuda tensor)}; recs: {imagename: [{'classid', 'difficult', 'mask' [h,w]}, ...]};
    imagesetfile: path of the image list (one name per line) or the list itself.  Returns (recall, precision,
    ap).  Detections only interact with detections of the same image, so the marking runs image by image on
    the IoU matrix one dy_mask_overlaps launch produced for that image."""
    if engine is None:
        raise RuntimeError('voc_eval needs an Engine: the mask IoUs are computed on the GPU (no CPU fallback)')
    if isinstance(imagesetfile, str):
        with open(imagesetfile, 'r') as f:
            imagenames = [line.strip() for line in f]
    else:
        imagenames = list(imagesetfile)
    t = engine.torch
    gt = {}
    for name in imagenames:
        objs = [o for o in recs[name] if o['classid'] == classid]
        gt[name] = ([o['mask'] for o in objs], np.array([o['difficult'] for o in objs], bool))
    npos = int(sum(int((~d).sum()) for _, d in gt.values()))
    nd = len(detfile)
    if nd == 0:
        return 0., 0., 0.
    ranking = np.argsort(-np.array([float(x['score']) for x in detfile]))     # the reference's sort (and tie order)
    rank_of = np.empty(nd, np.int64)
    rank_of[ranking] = np.arange(nd)
    members = {}
    for d, x in enumerate(detfile):
        members.setdefault(x['imageid'], []).append(d)

    def dev(m):
        m = m if isinstance(m, t.Tensor) else t.from_numpy(np.ascontiguousarray(m))
        return m.to(engine.device).to(t.uint8)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for name, dets in members.items():
        masks, difficult = gt[name]
        order = rank_of[dets]
        if masks:
            iou = engine.mask_overlaps(t.stack([dev(detfile[d]['mask']) for d in dets]),
                                       t.stack([dev(m) for m in masks])).cpu().numpy()
            t_img, f_img = _mark_image(order, iou.max(axis=1), iou.argmax(axis=1), difficult, ovthresh)
        else:
            t_img, f_img = _mark_image(order, None, None, difficult, ovthresh)
        tp[order] = t_img
        fp[order] = f_img
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    eps = np.finfo(np.float64).eps
    ap = voc_ap(tp / float(npos), tp / np.maximum(tp + fp, eps), use_07_metric)
    return tp[-1] / float(npos), tp[-1] / np.maximum(tp[-1] + fp[-1], eps), ap
