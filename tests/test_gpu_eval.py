"""GPU: dy_mask_overlaps and the GPU-backed voc_eval mirror (evaluation part)
against the mAP oracle, which is pinned to the reference's own voc_eval (tests/test_eval_cpu.py).
Integer counts -> the IoUs are bit-identical to NumPy's float32 arithmetic; recall / precision / AP equal
the committed outputs of the reference exactly."""
import copy

import numpy as np
import pytest

from oracle import dis_oracle_eval as OE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    import disyolo_b200 as dy
    e = dy.Engine(image_size=64, max_batch=1, precision='fp32')
    yield e
    e.close()


@pytest.mark.parametrize('shape', [(96, 128, 5, 7), (348, 620, 30, 20), (33, 41, 1, 1), (1200, 1600, 12, 9)],
                         ids=lambda s: '%dx%d_%dx%d' % s)
def test_mask_overlaps_bit_exact(eng, shape):
    h, w, n1, n2 = shape
    rng = np.random.default_rng(h + n1)
    yy, xx = np.mgrid[0:h, 0:w]

    def blobs(n):
        out = np.zeros((n, h, w), bool)
        for i in range(n):
            cy, cx, ry, rx = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(2, h / 2), rng.uniform(2, w / 2)
            out[i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        return out
    a, b = blobs(n1), blobs(n2)
    if n1 > 2:
        a[1] = False                                   # an empty mask: 0 / (0 + area - 0) = 0
    want = OE.compute_overlaps_masks(np.moveaxis(a, 0, -1), np.moveaxis(b, 0, -1)).astype(np.float32)
    got = eng.mask_overlaps(a, b).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    # two empty masks: 0/0 = NaN on both sides
    z = np.zeros((1, h, w), bool)
    assert np.isnan(eng.mask_overlaps(z, z).cpu().numpy()[0, 0])


def test_voc_eval_matches_reference_outputs(eng, golden):
    from disyolo_b200 import evalmap
    cache = {}
    for case in golden['voc_eval_synthetic']:
        if case['seed'] not in cache:
            cache[case['seed']] = OE.synthetic_dataset(case['seed'])
        names, recs, dets = cache[case['seed']]
        r, p, ap = evalmap.voc_eval(copy.deepcopy(dets[case['classid']]), copy.deepcopy(recs), names, case['classid'],
                                    0.5, case['use_07_metric'], engine=eng)
        assert (float(r), float(p), float(ap)) == (case['recall'], case['precision'], case['ap']), case


def test_voc_eval_requires_an_engine():
    from disyolo_b200 import evalmap
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        evalmap.voc_eval([], {}, [], 0)
