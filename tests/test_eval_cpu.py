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
