"""GPU: decode / threshold / NMS / top-k / mask assembly kernels against the oracle, fed identical
inputs (north_star: boxes within 0.5 px, NMS keep-sets bit-exact on identical decoded boxes,
masks IoU >= 0.99)."""
import numpy as np
import pytest

from oracle import dis_oracle as O
from tests.util import mask_iou, synthetic_heads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    import disyolo_b200 as dy
    e = dy.Engine(image_size=576, max_batch=4, precision='fp32', device=0)
    yield e
    e.close()


def _windows(B):
    w = np.tile(np.array([[0, 0, 1, 1]], np.float32), (B, 1))
    if B > 1:
        w[1] = [0.21875, 0.0, 0.7795139, 1.0]          # letterbox window of 00044.jpg (golden)
    return w


def test_decode_matches_oracle(eng):
    rng = np.random.default_rng(21)
    B = 3
    yolos = synthetic_heads(rng, B, 576)
    win = _windows(B)
    box, cls, score = [t.cpu().numpy() for t in eng.decode(yolos, win)]
    pred = O.interpret_output(yolos)
    for b in range(B):
        ob, oc, osc = O.decode_candidates(pred, b, win[b])
        assert box[b].shape == ob.shape == (20412, 4)
        assert np.max(np.abs(box[b] - ob)) * 576 < 0.5          # pixels
        assert np.max(np.abs(box[b] - ob)) < 2e-5               # and in fact ~1 ulp
        assert np.max(np.abs(score[b] - osc)) < 1e-6
        amb = np.sort(np.sort(pred['cls'][0][b].reshape(-1, 3), 1)[:, -2:], 1)
        assert (cls[b] != oc).sum() == 0 or np.min(amb[:, 1] - amb[:, 0]) < 1e-6


def test_nms_keep_set_bit_exact(eng):
    """Identical decoded boxes in -> identical keep set (indices AND order) out."""
    rng = np.random.default_rng(22)
    B = 4
    yolos = synthetic_heads(rng, B, 576, obj_bias=-1.5)
    win = _windows(B)
    box, cls, score = eng.decode(yolos, win)
    idx, cnt, raw = [t.cpu().numpy() for t in eng.nms(box, cls, score, 0.25)]
    box, cls, score = box.cpu().numpy(), cls.cpu().numpy(), score.cpu().numpy()
    for b in range(B):
        rows, kept = O.select_detections(box[b], cls[b], score[b], 0.25)
        assert cnt[b] == len(kept) > 0
        assert idx[b, :cnt[b]].tolist() == kept.tolist()
        assert np.all(idx[b, cnt[b]:] == -1)
        assert np.array_equal(raw[b, :cnt[b]], rows) and np.all(raw[b, cnt[b]:] == 0)


def test_nms_with_ties_and_degenerate_boxes(eng):
    rng = np.random.default_rng(23)
    N = 3000
    c = rng.random((N, 2)).astype(np.float32)
    wh = (rng.random((N, 2)) * 0.2).astype(np.float32)
    box = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    box[::17, 2] = box[::17, 0]                                   # zero-area boxes
    box[5::29] = box[5::29][:, [2, 3, 0, 1]]                      # flipped corners
    score = rng.choice(np.linspace(0.26, 0.95, 12), N).astype(np.float32)   # heavy ties
    cls = rng.integers(0, 3, N).astype(np.int32)
    idx, cnt, raw = [t.cpu().numpy() for t in eng.nms(box[None], cls[None], score[None], 0.25)]
    rows, kept = O.select_detections(box, cls, score, 0.25)
    assert idx[0, :cnt[0]].tolist() == kept.tolist()
    assert np.array_equal(raw[0, :cnt[0]], rows)


def test_nms_empty_and_single(eng):
    box = np.zeros((1, 64, 4), np.float32); box[0, :, 2:] = 0.5
    cls = np.zeros((1, 64), np.int32)
    score = np.full((1, 64), 0.1, np.float32)
    idx, cnt, raw = [t.cpu().numpy() for t in eng.nms(box, cls, score, 0.25)]
    assert cnt[0] == 0 and np.all(idx == -1) and np.all(raw == 0)
    score[0, 7] = 0.9
    idx, cnt, raw = [t.cpu().numpy() for t in eng.nms(box, cls, score, 0.25)]
    assert cnt[0] == 1 and idx[0, 0] == 7


def test_detect_from_heads(eng):
    rng = np.random.default_rng(24)
    B = 2
    yolos = synthetic_heads(rng, B, 576, obj_bias=-2.0)
    win = _windows(B)
    raw, box, cnt = [t.cpu().numpy() for t in eng.detect(yolos, win, 0.25)]
    pred = O.interpret_output(yolos)
    want = O.filter_detections(pred, win, 0.25)
    assert np.max(np.abs(raw[..., :4] - want[..., :4])) * 576 < 0.5
    assert np.array_equal(raw[..., 4], want[..., 4])
    assert np.max(np.abs(raw[..., 5] - want[..., 5])) < 1e-6
    db, _ = O.val_test(want, np.zeros((B, 288, 288, 9), np.float32))
    for b in range(B):
        assert cnt[b] == len(db[b])
        assert np.max(np.abs(box[b, :cnt[b]] - db[b])) < 2e-5 and np.all(box[b, cnt[b]:] == 0)


@pytest.mark.parametrize('layout', ['nhwc', 'planar'])
def test_mask_assembly(eng, layout):
    rng = np.random.default_rng(25)
    B, S, md = 2, 288, 30
    sm = (rng.standard_normal((B, S, S, 9)) * 3).astype(np.float32)
    det = np.zeros((B, md, 6), np.float32)
    cnt = np.array([7, 0], np.int32)
    kat = [[.1, .2, .5, .9], [.21875, 0, .7795139, 1], [0, 0, .00868, .0295], [0, 0, 1, 1],
           [.5, .5, .51, .9], [.3, .31, .9, .33], [.02, .6, .98, .97]]
    det[0, :7, :4] = np.array(kat, np.float32)
    det[0, :7, 5] = 0.9
    arg = sm if layout == 'nhwc' else np.ascontiguousarray(sm.transpose(0, 3, 1, 2))
    got = eng.assemble_masks(arg, det, cnt, layout).cpu().numpy()
    props, want = O.assemble_masks(det[0], sm[0], 3)
    assert len(props) == 7
    assert np.max(np.abs(got[0, :7] - want)) < 1e-6
    for d in range(7):
        assert mask_iou(got[0, d], want[d]) >= 0.99
    assert got[0, 3, 0, 0] != 0.5 and got[0, 0, 0, 0] == 0.5     # inside / outside


def test_detect_stress_many_boxes():
    """Stress config: 1152 px, low threshold, max_detection 1000 (BASELINE configs[4])."""
    import disyolo_b200 as dy
    e = dy.Engine(image_size=1152, max_batch=1, precision='fp32', device=0, max_detection=1000)
    rng = np.random.default_rng(26)
    yolos = synthetic_heads(rng, 1, 1152, obj_bias=0.0)
    for y in yolos:
        y[..., 2:4] -= 1.5                                        # small boxes so that many survive NMS
    win = _windows(1)
    raw, box, cnt = [t.cpu().numpy() for t in e.detect(yolos, win, 0.05)]
    pred = O.interpret_output(yolos)
    want = O.filter_detections(pred, win, 0.05, max_detection=1000)
    n = int((want[0, :, 5] > 0).sum())
    assert n >= 1000, n                                        # the stress bar: >= 1,000 boxes through NMS + masks
    assert np.array_equal(raw[0, :, 4], want[0, :, 4])
    assert np.max(np.abs(raw[0] - want[0])) < 2e-5
    sm = (rng.standard_normal((1, 576, 576, 9))).astype(np.float32)
    got = e.assemble_masks(sm, box, cnt, 'nhwc').cpu().numpy()
    props, wm = O.assemble_masks(want[0], sm[0], 3)
    assert cnt[0] == len(props)
    sel = rng.choice(cnt[0], 12, replace=False)
    for d in sel:
        assert np.max(np.abs(got[0, d] - wm[d])) < 1e-6
    e.close()


def test_nms_multi_batch_with_massive_ties():
    """The batched NMS beyond one shared-memory batch (2048 candidates per class): thousands of candidates that
    share a handful of score values (the radix select has to descend into the candidate-index bits of the key),
    more selections than fit one batch's survivors, every class; keep sets and order bit-exact against the oracle."""
    import disyolo_b200 as dy
    e = dy.Engine(image_size=576, max_batch=2, precision='fp32', device=0, max_detection=400)
    rng = np.random.default_rng(31)
    N = 18000
    box = np.zeros((2, N, 4), np.float32)
    cy, cx = rng.random((2, N)).astype(np.float32), rng.random((2, N)).astype(np.float32)
    h, w = (0.01 + 0.03 * rng.random((2, N))).astype(np.float32), (0.01 + 0.03 * rng.random((2, N))).astype(np.float32)
    h[0] *= 4.0                                      # image 0: large boxes -> most candidates are suppressed, every
    w[0] *= 4.0                                      # class walks through several 2048-candidate batches
    box[..., 0], box[..., 1], box[..., 2], box[..., 3] = cy - h, cx - w, cy + h, cx + w
    box = np.clip(box, 0.0, 1.0)
    cls = rng.integers(0, 3, (2, N)).astype(np.int32)
    score = rng.choice(np.array([0.9, 0.6, 0.6000001, 0.3], np.float32), (2, N)).astype(np.float32)
    score[1] = 0.5                                   # image 1: ONE score value for all 18,000 candidates
    idx, cnt, raw = [t.cpu().numpy() for t in e.nms(box, cls, score, 0.25)]
    for b in range(2):
        rows, kept = O.select_detections(box[b], cls[b], score[b], 0.25, max_detection=400)
        n = len(kept)
        assert cnt[b] == n and n >= 150
        assert idx[b, :n].tolist() == kept.tolist(), 'image %d: keep set / order differs' % b
        assert np.array_equal(raw[b, :n], rows)
    e.close()
