"""CPU: the training-data oracle (oracle/dis_oracle_augment.py) against outputs of the reference's own
utils/train_data.py methods run in the build container (tests/golden/augment_kat.json), plus known answers for
the two restated third-party pieces (skimage.draw.polygon, pyblur's 3x3 line kernels)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import dis_oracle_augment as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def make_masks(seed, n, h, w):
    rng = np.random.default_rng(seed)
    m = np.zeros((20, h, w), np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):
        for _ in range(int(rng.integers(1, 4))):
            cy, cx = rng.uniform(0.15, 0.85) * h, rng.uniform(0.15, 0.85) * w
            ry, rx = rng.uniform(0.03, 0.2) * h, rng.uniform(0.03, 0.2) * w
            m[i][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = 1.0
    return m


@pytest.fixture(scope='module')
def kat():
    with open(os.path.join(ROOT, 'tests', 'golden', 'augment_kat.json')) as f:
        return json.load(f)


def test_scale_crop_flip_matches_reference(kat):
    pytest.importorskip('cv2')
    for c in kat['place']:
        img = make_image(c['seed'], c['h'], c['w'])
        args = (c['new_w'], c['new_h'], c['dx'], c['dy'])
        assert sha(A.apply_random_scale_and_crop(img, *args, 'image', 576)) == c['placed_sha'], c['seed']
        masks = make_masks(100 + c['seed'], 3, c['h'], c['w'])
        for flip in (1, 2, 3):
            res = A.image_read(img, *args, flip, 1, 576)
            assert res.dtype == np.float32 and sha(res) == c['reads'][str(flip)]['sha'], (c['seed'], flip)
            rm = A.resize_mask(masks, *args, flip, [0, 1, 2], 576)
            assert sha(rm.astype(np.uint8)) == c['masks'][str(flip)]['sha'], (c['seed'], flip)
            assert int(rm.sum()) == c['masks'][str(flip)]['count']


def test_bboxes_noise_light_match_reference(kat):
    pytest.importorskip('cv2')
    for c in kat['bboxes']:
        m = make_masks(c['seed'], 4, 200, 260)
        assert [list(A.extract_bboxes(m[i].astype(np.uint8))) for i in range(4)] == c['boxes']
    for c in kat['noise']:
        img = make_image(c['seed'], 576, 576)
        s, p = A.replay_salt_pepper(c['seed'], img.shape)
        assert sha(A.add_salt_pepper_noise(img, s, p)) == c['salt_pepper_sha']
        assert sha(A.change_light(img, A.replay_light_coeff(c['seed']))) == c['light_sha']
    for c in kat['image_read_bnl']:
        img = make_image(c['seed'], 300, 400)
        s, p = A.replay_salt_pepper(c['seed'], (576, 576, 3))
        res = A.image_read(img, 576, 432, 0, 72, 2, c['bnl'], 576, salt_rc=s, pepper_rc=p,
                           coeff=A.replay_light_coeff(c['seed']))
        assert sha(res) == c['sha'], c


def test_polygon_known_answers():
    """skimage.draw.polygon semantics (interior + boundary): the documented example of skimage
    (`polygon([1, 2, 8], [1, 7, 4])` on a 10 x 10 image) and axis-aligned shapes."""
    img = np.zeros((10, 10), np.uint8)
    rr, cc = A.polygon([1, 2, 8], [1, 7, 4])
    img[rr, cc] = 1
    want = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                     [0, 1, 0, 0, 0, 0, 0, 0, 0, 0],
                     [0, 0, 1, 1, 1, 1, 1, 1, 0, 0],
                     [0, 0, 1, 1, 1, 1, 1, 0, 0, 0],
                     [0, 0, 0, 1, 1, 1, 1, 0, 0, 0],
                     [0, 0, 0, 1, 1, 1, 0, 0, 0, 0],
                     [0, 0, 0, 0, 1, 1, 0, 0, 0, 0],
                     [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
                     [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
                     [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.uint8)
    assert np.array_equal(img, want)
    rr, cc = A.polygon([2, 2, 6, 6], [3, 8, 8, 3])          # rectangle: boundary included
    sq = np.zeros((10, 10), np.uint8)
    sq[rr, cc] = 1
    assert sq.sum() == 5 * 6 and sq[2:7, 3:9].all()
    m = A.load_mask(20, 12, 12, [[dict(type='out', all_points_x=[1, 10, 10, 1], all_points_y=[1, 1, 10, 10]),
                                  dict(type='in', all_points_x=[4, 7, 7, 4], all_points_y=[4, 4, 7, 7])]])
    assert m[0, 5, 5] == 0 and m[0, 2, 2] == 1 and m[0, 4, 4] == 1 and m[0, 7, 7] == 1     # hole, ring, hole vertices
    assert m[1:].sum() == 0


def test_line_kernels():
    k = A.line_kernel3(0, 'full')
    assert np.allclose(k, np.array([[0, 0, 0], [1, 1, 1], [0, 0, 0]], np.float32) / 3)
    k = A.line_kernel3(45, 'full')
    assert np.allclose(k, np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0]], np.float32) / 3)
    k = A.line_kernel3(90, 'right')
    assert np.count_nonzero(k) == 2 and k[1, 1] == 0.5 and k[2, 1] == 0.5
    k = A.line_kernel3(135, 'left')
    assert np.count_nonzero(k) == 2 and k[0, 0] == 0.5 and k[1, 1] == 0.5
