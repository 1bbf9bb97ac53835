"""CPU: the mAP oracle (oracle/dis_oracle_eval.py) against the reference's own voc_eval / voc_ap /
compute_overlaps_masks run in the build container (tests/golden/ref_kat.json)."""
import copy

import numpy as np

from oracle import dis_oracle_eval as OE


def test_voc_eval_oracle_matches_reference(golden):
    cache = {}
    for case in golden['voc_eval_synthetic']:
        if case['seed'] not in cache:
            cache[case['seed']] = OE.synthetic_dataset(case['seed'])
        names, recs, dets = cache[case['seed']]
        r, p, ap = OE.voc_eval(copy.deepcopy(dets[case['classid']]), copy.deepcopy(recs), names, case['classid'], 0.5,
                               case['use_07_metric'])
        assert (float(r), float(p), float(ap)) == (case['recall'], case['precision'], case['ap']), case
    assert len({c['ap'] for c in golden['voc_eval_synthetic']}) > 8        # the fixture is not degenerate


def test_overlaps_and_ap_known_answers(golden):
    a = np.zeros((8, 8, 1), bool); a[:4, :4, 0] = True
    b = np.zeros((8, 8, 2), bool); b[:4, :2, 0] = True; b[:, :, 1] = True
    assert OE.compute_overlaps_masks(a, b).tolist() == [[0.5, 0.25]]           # SURVEY 8c known answer (4)
    assert abs(OE.voc_ap(np.array([.5, .5, 1.]), np.array([1., .5, 2. / 3])) - golden['voc_ap']['ap']) < 1e-15


def test_product_voc_ap_matches_reference(golden):
    """evalmap.voc_ap (host part of the product's voc_eval mirror) against the reference's voc_ap output."""
    from disyolo_b200 import evalmap
    rec, prec = np.array([.5, .5, 1.]), np.array([1., .5, 2. / 3])
    assert abs(evalmap.voc_ap(rec, prec, False) - golden['voc_ap']['ap']) < 1e-15
    assert evalmap.voc_ap(rec, prec, True) == OE.voc_ap(rec, prec, True)


class _HostIoU(object):
    """Test double for Engine.mask_overlaps (the GPU kernel is checked bit-for-bit in test_gpu_eval.py): lets the
    CPU suite run the product's host-side marking / AP logic against the reference's own outputs."""

    def __init__(self):
        import torch
        self.torch, self.device = torch, torch.device('cpu')

    def mask_overlaps(self, a, b):
        t = self.torch
        a, b = a.reshape(a.shape[0], -1).float(), b.reshape(b.shape[0], -1).float()
        inter = a @ b.t()
        return inter / (a.sum(1)[:, None] + b.sum(1)[None, :] - inter)


def test_product_voc_eval_host_logic_matches_reference(golden):
    """evalmap.voc_eval (image-wise vectorised marking) == the reference's sequential loop on every golden case."""
    from disyolo_b200 import evalmap
    eng, cache = _HostIoU(), {}
    for case in golden['voc_eval_synthetic']:
        if case['seed'] not in cache:
            cache[case['seed']] = OE.synthetic_dataset(case['seed'])
        names, recs, dets = cache[case['seed']]
        r, p, ap = evalmap.voc_eval(copy.deepcopy(dets[case['classid']]), copy.deepcopy(recs), names, case['classid'],
                                    0.5, case['use_07_metric'], engine=eng)
        assert (float(r), float(p), float(ap)) == (case['recall'], case['precision'], case['ap']), case


def test_product_voc_ap_random_against_oracle():
    from disyolo_b200 import evalmap
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 50, 400):
        tp = np.cumsum(rng.random(n) < 0.6)
        fp = np.cumsum(np.ones(n)) - tp
        rec, prec = tp / max(tp[-1], 1) * rng.uniform(0.3, 1.0), tp / np.maximum(tp + fp, 1e-9)
        for m07 in (False, True):
            assert abs(evalmap.voc_ap(rec, prec, m07) - OE.voc_ap(rec, prec, m07)) < 1e-14, (n, m07)
