"""GPU: the training data pipeline kernels (dy_polygon_masks, dy_mask_boxes, dy_augment_image, dy_augment_masks,
dy_salt_pepper, dy_change_light, dy_motion_blur3, dy_u8_to_unit_float) against the oracle that is pinned to the
reference's own utils/train_data.py outputs (tests/test_augment_cpu.py).  Integer / byte work: bit-exact.
The one documented exception: cv2's 8-bit RGB->HLS differs from the restatement by one hue unit on 3 of the
2^24 RGB triples, so change_light allows a 1e-5 fraction of differing pixels."""
import numpy as np
import pytest

from oracle import dis_oracle_augment as A
from tests.test_augment_cpu import make_image, make_masks

pytestmark = pytest.mark.gpu

PLACE = [(1, 348, 620, 576, 323, 0, 126), (2, 620, 348, 323, 576, 126, 0), (3, 300, 400, 864, 648, -150, -40),
         (4, 480, 640, 432, 324, 100, 200), (5, 240, 320, 691, 518, -60, 30), (6, 576, 576, 576, 576, 0, 0),
         (7, 97, 131, 700, 518, -124, 58), (8, 1080, 1920, 576, 324, 0, 126)]


@pytest.fixture(scope='module')
def td():
    import disyolo_b200 as dy
    return dy.TrainData(image_size=576)


@pytest.mark.parametrize('case', PLACE, ids=lambda c: 'seed%d_%dx%d_to_%dx%d' % c[:5])
def test_place_image_and_masks_bit_exact(td, case):
    pytest.importorskip('cv2')
    seed, h, w, nw, nh, dx, dy_ = case
    img = make_image(seed, h, w)
    masks = make_masks(100 + seed, 3, h, w)
    for flip in (1, 2, 3):
        want = A.flip_image(A.apply_random_scale_and_crop(img, nw, nh, dx, dy_, 'image', 576), flip)
        got = td.place_image(img, nw, nh, dx, dy_, flip).cpu().numpy()
        assert np.array_equal(got, want), 'image: %d differing bytes' % int((got != want).sum())
        wm = A.resize_mask(masks, nw, nh, dx, dy_, flip, [0, 1, 2], 576)[:3]
        gm = td.place_masks(masks[:3], nw, nh, dx, dy_, flip).cpu().numpy().astype(bool)
        assert np.array_equal(gm, wm), 'masks: %d differing pixels' % int((gm != wm).sum())


def test_image_read_all_effects(td):
    pytest.importorskip('cv2')
    img = make_image(41, 300, 400)
    args = (576, 432, 0, 72)
    s, p = A.replay_salt_pepper(41, (576, 576, 3))
    coeff = A.replay_light_coeff(41)
    for bnl, kw in ((1, {}), (2, dict(salt_rc=s, pepper_rc=p)), (3, dict(coeff=coeff)),
                    (4, dict(kernel=A.line_kernel3(45, 'full'))), (4, dict(kernel=A.line_kernel3(90, 'right')))):
        want = A.image_read(img, *args, 2, bnl, 576, **kw)
        got = td.image_read(img, *args, 2, bnl, **kw).cpu().numpy()
        assert got.dtype == np.float32 and got.shape == (576, 576, 3)
        bad = float(np.mean((got != want).any(-1)))
        assert bad <= (1e-5 if bnl == 3 else 0.0), 'bnl %d: %.3g of the pixels differ' % (bnl, bad)


def test_change_light_exhaustive_sample(td):
    """A dense sample of the RGB cube through RGB->HLS->scale->RGB against cv2 (the oracle calls cv2)."""
    import torch
    pytest.importorskip('cv2')
    v = np.arange(0, 256, 3, dtype=np.uint8)
    R, G, B = np.meshgrid(v, v, v, indexing='ij')
    rgb = np.stack([R, G, B], -1).reshape(-1, 86, 3)          # 86^3 colours as an image
    for coeff in (0.5, 0.8371, 1.0, 1.3, 1.499):
        want = A.change_light(rgb, coeff)
        got = td.change_light(torch.from_numpy(rgb.copy()).cuda(), coeff).cpu().numpy()
        bad = int((got != want).any(-1).sum())
        assert bad <= 3, 'coeff %.3f: %d of %d colours differ' % (coeff, bad, rgb.shape[0] * rgb.shape[1])


def test_polygon_masks_and_boxes(td):
    rng = np.random.default_rng(5)
    h, w = 240, 320
    polys = []
    for i in range(6):
        n = int(rng.integers(3, 12))
        cx, cy, r = rng.uniform(60, w - 60), rng.uniform(60, h - 60), rng.uniform(15, 55)
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        xs = np.clip(np.rint(cx + r * np.cos(ang) * rng.uniform(0.5, 1.0, n)), 0, w - 1).astype(int).tolist()
        ys = np.clip(np.rint(cy + r * np.sin(ang) * rng.uniform(0.5, 1.0, n)), 0, h - 1).astype(int).tolist()
        inst = [dict(type='out', all_points_x=xs, all_points_y=ys)]
        if i % 2 == 0:                                         # an inner background region
            hx = np.clip(np.rint(cx + 0.3 * r * np.array([-1, 1, 1, -1])), 0, w - 1).astype(int).tolist()
            hy = np.clip(np.rint(cy + 0.3 * r * np.array([-1, -1, 1, 1])), 0, h - 1).astype(int).tolist()
            inst.append(dict(type='in', all_points_x=hx, all_points_y=hy))
        polys.append(inst)
    polys.append([dict(type='out', all_points_x=[10, 10, 10], all_points_y=[5, 9, 7])])      # degenerate: a segment
    want = A.load_mask(20, h, w, polys)[:len(polys)].astype(np.uint8)
    got = td.polygon_masks(polys, h, w)
    assert np.array_equal(got.cpu().numpy(), want), int((got.cpu().numpy() != want).sum())
    boxes = td.mask_boxes(got).cpu().numpy()
    for i in range(len(polys)):
        assert list(boxes[i]) == list(A.extract_bboxes(want[i])), i
    import torch
    empty = torch.zeros((2, 50, 60), dtype=torch.uint8).cuda()
    empty[1, 7, 9] = 1
    b = td.mask_boxes(empty).cpu().numpy()
    assert list(b[0]) == [0, 0, 0, 0] and list(b[1]) == [9, 7, 10, 8]
