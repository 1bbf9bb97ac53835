"""Mask-level VOC AP with the mask IoUs computed on the GPU  (evaluation part).

Mirror of utils/voc_eval_mask.py: `voc_ap` (:9-36) and `voc_eval` (:58-134) keep the reference's
signatures and semantics -- detections sorted by descending confidence, greedy true/false-positive marking
against the not-yet-claimed ground-truth mask of highest IoU (> ovthresh, strict), "difficult" objects
ignored, the VOC-2010 area-under-envelope AP (or the 11-point VOC-2007 one).  The one hot operation,
`compute_overlaps_masks` (:38-56: a [pixels x n_det]^T [pixels x n_gt] product per detection), is replaced
by ONE dy_mask_overlaps launch per image covering all of that image's detections of the class.
There is no CPU fallback: an Engine (GPU) is required.
"""
import numpy as np


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


def voc_eval(detfile, recs, imagesetfile, classid, ovthresh=0.5, use_07_metric=False, engine=None):
    """detfile: list of {'imageid', 'score', 'mask' [h,w] bool (numpy or cuda tensor)};
    recs: {imagename: [{'classid', 'difficult', 'mask' [h,w]}, ...]}; imagesetfile: path of the image list
    (one name per line) or the list itself.  Returns (recall, precision, ap) like the reference."""
    if engine is None:
        raise RuntimeError('voc_eval needs an Engine: the mask IoUs are computed on the GPU (no CPU fallback)')
    if isinstance(imagesetfile, str):
        with open(imagesetfile, 'r') as f:
            imagenames = [x.strip() for x in f.readlines()]
    else:
        imagenames = list(imagesetfile)
    t = engine.torch
    class_recs, npos = {}, 0
    for imagename in imagenames:
        R = [obj for obj in recs[imagename] if obj['classid'] == classid]
        difficult = np.array([x['difficult'] for x in R]).astype(bool)
        npos = npos + int(sum(~difficult))
        class_recs[imagename] = {'masks': [x['mask'] for x in R], 'difficult': difficult, 'det': [False] * len(R)}
    image_ids = [x['imageid'] for x in detfile]
    confidence = np.array([float(x['score']) for x in detfile])
    sorted_ind = np.argsort(-confidence)
    if len(sorted_ind) == 0:
        return 0., 0., 0.
    # one IoU matrix per image: rows = that image's detections (any order), columns = its ground truth
    by_image = {}
    for d in sorted_ind:
        by_image.setdefault(image_ids[d], []).append(int(d))

    def dev(m):
        m = m if isinstance(m, t.Tensor) else t.from_numpy(np.ascontiguousarray(m))
        return m.to(engine.device).to(t.uint8)
    overlaps = {}
    for imageid, dets in by_image.items():
        R = class_recs[imageid]
        if not R['masks']:
            continue
        ov = engine.mask_overlaps(t.stack([dev(detfile[d]['mask']) for d in dets]),
                                  t.stack([dev(m) for m in R['masks']])).cpu().numpy()
        for row, d in enumerate(dets):
            overlaps[d] = ov[row]
    nd = len(sorted_ind)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for k, d in enumerate(sorted_ind):
        R = class_recs[image_ids[d]]
        ovmax, jmax = -np.inf, -1
        if int(d) in overlaps:
            ovmax = np.max(overlaps[int(d)])
            jmax = int(np.argmax(overlaps[int(d)]))
        if ovmax > ovthresh:
            if not R['difficult'][jmax]:
                if not R['det'][jmax]:
                    tp[k] = 1.
                    R['det'][jmax] = 1
                else:
                    fp[k] = 1.
        else:
            fp[k] = 1.
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    recall = tp[-1] / float(npos)
    precision = tp[-1] / np.maximum(tp[-1] + fp[-1], np.finfo(np.float64).eps)
    return recall, precision, ap
