"""CPU: pin the oracle against the reference's runnable pieces (tests/golden/ref_kat.json, made by
tests/golden/make_golden.py from /root/reference) and the hand-derived known answers of
SURVEY.md section 8(c)."""
import numpy as np
import pytest

from oracle import dis_oracle as O


def test_config_matches_reference(golden):
    g = golden['config']
    assert np.allclose(O.ANCHORS, np.array(g['ANCHORS'], np.float32))
    assert O.NUM_CLASS == len(g['CLASSES']) and O.ALPHA == g['ALPHA'] and O.K_MAP == g['K_MAP']
    assert O.OBJ_THRESHOLD == g['OBJ_THRESHOLD'] and O.IOU_THRESHOLD == g['IOU_THRESHOLD']
    assert O.MAX_DETECTION == g['MAX_DETECTION']


def test_correct_yolo_boxes_kat(golden):
    for case in golden['correct_yolo_boxes']:
        assert list(O.correct_yolo_boxes(*case['args'])) == case['out']


def test_letterbox_window_kat(golden):
    for lb in golden['letterbox']:
        w = O.letterbox_window(lb['h'], lb['w'], 576)
        assert np.allclose(w, np.array(lb['window'], np.float32), atol=1e-7)
        assert abs(lb['pad_value'] - 127.0 / 255.0) < 1e-12 or lb['window'][0] == 0.0


def test_sigmoid_kat(golden):
    x = np.array(golden['sigmoid']['x'], np.float32)
    y = np.array(golden['sigmoid']['y'])
    got = O.sigmoid(x)
    # the reference's host sigmoid clips at +-50; inside that range both agree to fp32 precision
    inside = np.abs(x) < 50
    assert np.allclose(got[inside], y[inside], rtol=2e-7, atol=1e-9)


def test_bin_edges_known_answers():
    # SURVEY.md 8(c), hand-derived from yolo3_net_pos.py:876-897 with k=3, S=288
    S = 288
    cases = [([.1, .2, .5, .9], [29, 58, 144, 259], [58, 125, 192, 259], [29, 67, 106, 144]),
             ([.21875, 0, .7795139, 1], [63, 0, 224, 288], [0, 96, 192, 288], [63, 117, 170, 224]),
             ([0, 0, .00868, .0295], [0, 0, 2, 8], [0, 3, 5, 8], [0, 1, 1, 2])]
    for box, pb_want, gx_want, gy_want in cases:
        pb = np.rint(np.array(box, np.float32) * np.float32(S))
        assert pb.astype(int).tolist() == pb_want
        gx, gy = O.bin_edges(pb, 3)
        assert gx == gx_want and gy == gy_want


def test_candidate_layout_and_flops():
    assert 3 * (72 * 72 + 36 * 36 + 18 * 18) == 20412
    assert O.flops_per_image(576) == 132683857920          # 132.684 GFLOP
    assert O.flops_per_image(1152) == 4 * O.flops_per_image(576)
    t = O.layer_table()
    assert len(t) == 82 and sum(1 for v in t.values() if v['bn']) == 78
    nweights = sum(v['k'] ** 2 * v['cin'] * v['cout'] for v in t.values())
    assert nweights == 61602208                            # SURVEY 8a totals


def test_same_padding_is_asymmetric_for_stride_2():
    assert O.same_pad(576, 3, 2) == (0, 1)
    assert O.same_pad(576, 3, 1) == (1, 1)
    assert O.same_pad(18, 1, 1) == (0, 0)


@pytest.mark.parametrize('k,s,cin,cout', [(3, 1, 5, 7), (3, 2, 4, 6), (1, 1, 8, 3)])
def test_conv_restatements_agree(k, s, cin, cout):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 12, 10, cin)).astype(np.float32)
    w = rng.standard_normal((k, k, cin, cout)).astype(np.float32)
    a = O.conv2d_same(x, w, s)
    b = O.conv2d_same_numpy(x, w, s)
    assert a.shape == b.shape == (2, 12 // s, 10 // s, cout)
    assert np.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_stride2_taps_start_at_2i():
    # one-hot input: out[i,j] must see in[2i+kh, 2j+kw] (pad 0 before, 1 after)
    x = np.zeros((1, 8, 8, 1), np.float32); x[0, 3, 5, 0] = 1
    w = np.arange(9, dtype=np.float32).reshape(3, 3, 1, 1)
    y = O.conv2d_same(x, w, 2)[0, :, :, 0]
    want = np.zeros((4, 4), np.float32)
    for oy in range(4):
        for ox in range(4):
            kh, kw = 3 - 2 * oy, 5 - 2 * ox
            if 0 <= kh < 3 and 0 <= kw < 3:
                want[oy, ox] = w[kh, kw, 0, 0]
    assert np.array_equal(y, want)


def test_nms_variants_agree_and_respect_ties():
    rng = np.random.default_rng(2)
    for trial in range(20):
        n = int(rng.integers(1, 200))
        c = rng.random((n, 2)).astype(np.float32)
        wh = (rng.random((n, 2)) * 0.3).astype(np.float32)
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        scores = rng.choice(np.linspace(0.3, 0.9, 7), n).astype(np.float32)     # many ties
        idx = rng.permutation(1000)[:n]
        a = O.nms_tf(boxes, scores, idx, 30, 0.3)
        b = O.nms_tf_vectorised(boxes, scores, idx, 30, 0.3)
        assert a == b
        assert len(a) <= 30
        for p, q in zip(a, a[1:]):       # selection order: score desc, then index asc
            assert (scores[p], -idx[p]) >= (scores[q], -idx[q])


def test_iou_tf_degenerate_and_flipped():
    assert O.iou_tf([0, 0, 0, 1], [0, 0, 1, 1]) == 0          # zero area
    assert O.iou_tf([1, 1, 0, 0], [0, 0, 1, 1]) == 1          # flipped corners are normalised
    assert abs(O.iou_tf([0, 0, 1, 1], [0, 0.5, 1, 1.5]) - 1 / 3) < 1e-6


def test_mask_assembly_semantics():
    rng = np.random.default_rng(3)
    S = 32
    sm = rng.standard_normal((S, S, 9)).astype(np.float32)
    boxes = np.zeros((5, 6), np.float32)
    boxes[0, :4] = [.1, .2, .5, .9]
    boxes[1, :4] = [.5, .5, .5, .9]            # zero height -> dropped (:877-880)
    boxes[2, :4] = [0, 0, 1, 1]
    props, m = O.assemble_masks(boxes, sm, 3)
    assert props.shape == (2, 6) and m.shape == (2, S, S)
    pb = np.rint(boxes[0, :4] * S)
    gx, gy = O.bin_edges(pb, 3)
    assert m[0, 0, 0] == np.float32(0.5)                                   # outside box
    y, x = gy[1], gx[2]                                                    # bin (1,2) -> ch 5
    assert np.isclose(m[0, y, x], 1 / (1 + np.exp(-sm[y, x, 5])), rtol=1e-6)
    # whole-image box: bin (by,bx) covers thirds
    assert np.isclose(m[1, S - 1, 0], 1 / (1 + np.exp(-sm[S - 1, 0, 6])), rtol=1e-6)
    none_props, none_m = O.assemble_masks(np.zeros((3, 6), np.float32), sm, 3)
    assert none_props.shape == (0, 6) and none_m == 0.0


def test_select_detections_topk_and_padding():
    rng = np.random.default_rng(4)
    yolos = None
    from tests.util import synthetic_heads
    yolos = synthetic_heads(rng, 1, 96, obj_bias=-1.0)
    pred = O.interpret_output(yolos)
    det = O.filter_detections(pred, np.array([[0, 0, 1, 1]], np.float32), 0.25)
    assert det.shape == (1, 30, 6)
    sc = det[0, :, 5]
    n = int((sc > 0).sum())
    assert n > 0 and np.all(sc[:n - 1] >= sc[1:n]) and np.all(det[0, n:] == 0)
    assert np.all(det[0, :n, :4] >= 0) and np.all(det[0, :n, :4] <= 1)


def test_forward_shapes_small():
    W = O.make_weights('lively', 0)
    img = np.random.default_rng(0).random((1, 64, 64, 3), dtype=np.float32)
    acts = {}
    yolos, mp = O.forward_network(img, W, acts=acts)
    assert [y.shape for y in yolos] == [(1, 8, 8, 3, 8), (1, 4, 4, 3, 8), (1, 2, 2, 3, 8)]
    assert mp.shape == (1, 32, 32, 9) and len(acts) == 82
    assert np.isfinite(mp).all() and float(np.abs(acts[52]).mean()) < 100.0


# --------------------------------------------------------------------------------------------------
# Independent cross-checks of the TF-side arithmetic the reference keeps no vectors for:
# the oracle's restatements against torch / torchvision implementations of the same operators.  They do not
# pin TensorFlow's results, but they break the oracle's self-consistency: a wrong restatement would
# have to be wrong in the same way in an unrelated code base.
# --------------------------------------------------------------------------------------------------
def test_oracle_nms_against_torchvision():
    import torch
    torchvision = pytest.importorskip('torchvision')
    from torchvision.ops import nms
    rng = np.random.default_rng(12)
    for trial in range(6):
        n = int(rng.integers(40, 400))
        yx = rng.random((n, 2)).astype(np.float32) * 0.8
        hw = (rng.random((n, 2)).astype(np.float32) * 0.3 + 0.02)
        boxes = np.concatenate([yx, yx + hw], axis=1)                     # (y1,x1,y2,x2), the reference's order
        scores = rng.permutation(n).astype(np.float32) / n                # distinct scores: no tie-order question
        keep = O.nms_tf(boxes, scores, np.arange(n), n, 0.3)
        tv = nms(torch.from_numpy(boxes[:, [1, 0, 3, 2]].copy()), torch.from_numpy(scores), 0.3).numpy()
        assert list(keep) == list(tv), trial
        assert list(O.nms_tf_vectorised(boxes, scores, np.arange(n), n, 0.3)) == list(tv)


def test_oracle_batchnorm_and_conv_against_torch_functional():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(13)
    x = rng.standard_normal((2, 9, 11, 16)).astype(np.float32)
    gamma, beta = rng.uniform(0.5, 1.5, 16).astype(np.float32), rng.standard_normal(16).astype(np.float32)
    mean, var = rng.standard_normal(16).astype(np.float32), rng.uniform(0.5, 2, 16).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    want = F.batch_norm(xt, torch.from_numpy(mean), torch.from_numpy(var), torch.from_numpy(gamma),
                        torch.from_numpy(beta), training=False, eps=O.BN_EPS).permute(0, 2, 3, 1).numpy()
    assert np.allclose(O.batch_norm_infer(x, gamma, beta, mean, var), want, rtol=1e-5, atol=1e-6)
    got, m, v = O.batch_norm_train(x, gamma, beta)
    want = F.batch_norm(xt, None, None, torch.from_numpy(gamma), torch.from_numpy(beta), training=True,
                        eps=O.BN_EPS).permute(0, 2, 3, 1).numpy()
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5)
    assert np.allclose(m, x.mean((0, 1, 2)), atol=1e-6) and np.allclose(v, x.var((0, 1, 2)), rtol=1e-5)
    # TF 'SAME' stride-2 on an even extent = pad (0 before, 1 after), i.e. F.pad(0,1,0,1) + a VALID conv
    w = rng.standard_normal((3, 3, 16, 8)).astype(np.float32)
    wt = torch.from_numpy(w).permute(3, 2, 0, 1)
    x2 = rng.standard_normal((1, 10, 12, 16)).astype(np.float32)
    want = F.conv2d(F.pad(torch.from_numpy(x2).permute(0, 3, 1, 2), (0, 1, 0, 1)), wt, stride=2).permute(0, 2, 3, 1)
    assert np.allclose(O.conv2d_same_numpy(x2, w, 2), want.numpy(), rtol=1e-4, atol=1e-5)
    assert np.allclose(O.conv2d_same(x2, w, 2), want.numpy(), rtol=1e-4, atol=1e-5)


def test_product_and_oracle_lively_weights_are_bit_identical():
    """bench.py / smoke() use disyolo_b200.init_weights('lively', s), the parity tests O.make_weights('lively', s):
    two independent generators of the same stream -- they must never diverge."""
    import disyolo_b200 as dy
    for seed in (0, 3):
        a, b = dy.init_weights('lively', seed), O.make_weights('lively', seed)
        assert sorted(a) == sorted(b)
        for k in a:
            assert a[k].dtype == b[k].dtype == np.float32 and np.array_equal(a[k], b[k]), k


def test_overlaps_graph_degenerate_boxes():
    """overlaps_graph (yolo3_net_pos.py:954-975) has no epsilon: two zero-area boxes give 0/0 = NaN, which can never
    satisfy `IoU >= 0.5`, so such a RoI is not a positive (the device kernel maps NaN to -1 in the same place,
    csrc/train.cu mask_roi_kernel); a zero-area box against a proper one is a plain 0."""
    from oracle import dis_oracle_train as T
    deg = np.array([[0.3, 0.3, 0.3, 0.3]], np.float32)
    box = np.array([[0.1, 0.1, 0.5, 0.5], [0.3, 0.3, 0.3, 0.3]], np.float32)
    ov = T.overlaps(deg, box)
    assert ov.shape == (1, 2) and ov[0, 0] == 0.0 and np.isnan(ov[0, 1])
    # through mask_rois: the only proposal and the only ground-truth box are the same zero-area box -> no positive RoI
    det = np.zeros((30, 6), np.float32)
    det[0, :4] = deg[0]
    tb = np.zeros((20, 5), np.float32)
    tb[0] = [0.3, 0.3, 0.0, 0.0, 1.0]            # xc, yc, w, h, class: kept (non-zero row), zero area
    rois, assign, keep = T.mask_rois(det, tb, list(range(30)), list(range(20)))
    assert len(rois) == 0 and keep.tolist() == [0]
    # ... while a proper ground-truth box is its own positive (IoU 1 with itself)
    tb[0] = [0.3, 0.3, 0.2, 0.2, 1.0]
    rois, assign, keep = T.mask_rois(det, tb, list(range(30)), list(range(20)))
    assert len(rois) == 1 and assign.tolist() == [0]
