"""GPU: dy_assign_labels (label part) against the NumPy restatement of
utils/train_data.py:134-178,189-228,258-262.  The arithmetic is reproduced op for op (float64 box transform,
float32 anchor IoU, float32 normalisation), so the label tensors must be bit-identical; cases cover every
flip mode, clamping at the borders, boxes competing for one cell, zero-size boxes and empty images."""
import numpy as np
import pytest

from oracle import dis_oracle_labels as OL

pytestmark = pytest.mark.gpu


def _case(rng, B, net, crowded=False):
    boxes = np.zeros((B, 20, 5), np.float32)
    nbox = rng.integers(0, 21, B).astype(np.int32)
    place = np.zeros((B, 4), np.float32)
    for b in range(B):
        ih, iw = rng.integers(200, 1200, 2)
        new_w = int(rng.uniform(0.6, 1.0) * net)
        new_h = int(rng.uniform(0.6, 1.0) * net)
        place[b] = [new_w / iw, new_h / ih, int(rng.uniform(0, net - new_w)), int(rng.uniform(0, net - new_h))]
        for i in range(nbox[b]):
            if crowded:                                   # many boxes around one point: cell conflicts
                cx, cy = iw * 0.5 + rng.uniform(-6, 6), ih * 0.5 + rng.uniform(-6, 6)
            else:
                cx, cy = rng.uniform(-20, iw + 20), rng.uniform(-20, ih + 20)      # partly outside: clamping
            w, h = rng.uniform(0, iw * 0.8), rng.uniform(0, ih * 0.8)
            if i % 7 == 6:
                w = 0.0                                   # degenerate box: IoU 0 with every anchor
            boxes[b, i] = [cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, rng.integers(0, 3)]
    flip = rng.integers(1, 4, B).astype(np.int32)
    return boxes, nbox, place, flip


@pytest.mark.parametrize('net,crowded', [(576, False), (576, True), (160, False), (1152, False)])
def test_assign_labels_bit_exact(net, crowded):
    import disyolo_b200 as dy
    rng = np.random.default_rng(net + crowded)
    B = 6
    boxes, nbox, place, flip = _case(rng, B, net, crowded)
    nbox[0] = 0                                            # an image without objects
    want = OL.assign_labels(boxes, nbox, place, flip, net)
    eng = dy.Engine(image_size=net, max_batch=1, precision='fp32')
    got = [t.cpu().numpy() for t in eng.assign_labels(boxes, nbox, place, flip)]
    eng.close()
    assert sum(int((w[..., 4] == 1).sum()) for w in want[:3]) > 0
    for name, g, w in zip(('yolo3', 'yolo2', 'yolo1', 'true_boxes'), got, want):
        assert g.shape == w.shape and np.array_equal(g, w), name
    if crowded:                                            # conflicts did occur: fewer cells than valid boxes
        assigned = sum(int((w[..., 4] == 1).sum()) for w in want[:3])
        assert assigned < int(nbox.sum())


def test_labels_feed_the_training_step():
    """Labels assigned on the GPU are consumed by dy_train_forward without leaving the device."""
    import torch
    import disyolo_b200 as dy
    from oracle import dis_oracle as O
    net, B = 160, 2
    rng = np.random.default_rng(4)
    boxes, nbox, place, flip = _case(rng, B, net)
    nbox[:] = np.maximum(nbox, 2)
    eng = dy.Engine(image_size=net, max_batch=B, precision='bf16')
    eng.load_weights(O.make_weights('lively', 0))
    eng.train_init()
    y3, y2, y1, tb = eng.assign_labels(boxes, nbox, place, flip)
    img = torch.from_numpy(rng.random((B, net, net, 3), dtype=np.float32)).cuda()
    tm = torch.zeros((B, 20, net, net), dtype=torch.uint8, device='cuda')
    pp = torch.from_numpy(np.stack([rng.permutation(30) for _ in range(B)]).astype(np.int32)).cuda()
    pg = torch.from_numpy(np.stack([rng.permutation(20) for _ in range(B)]).astype(np.int32)).cuda()
    losses = eng.train_forward(img, [y3, y2, y1], tb, tm, pp, pg, 0.2)
    assert np.isfinite(losses).all() and losses[0] > 0
    eng.close()
