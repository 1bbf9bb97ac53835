"""GPU: dy_letterbox / dy_postprocess against the I/O oracle, which calls
cv2.resize exactly as the reference does and is pinned to the reference's image_read bit for bit
(tests/test_io_oracle_cpu.py).

Tolerances: letterbox 4e-7 absolute on values in [0,1] (= 3 ulp of the 0..255 pixel value / 255: cv2's SIMD
kernels order / contract the two-tap sums differently; a NumPy restatement of the kernel's arithmetic
shows the same 3-ulp spread against cv2); corrected boxes exact; boolean masks: IoU >= 0.99
per detection and < 0.05 % differing pixels overall (pixels whose interpolated value is within an
ulp of 0.5 may flip); merged semantic mask >= 99.95 % equal."""
import numpy as np
import pytest

from oracle import dis_oracle_io as IO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    import disyolo_b200 as dy
    e = dy.Engine(image_size=576, max_batch=1, precision='fp32')
    yield e
    e.close()


@pytest.mark.parametrize('hw', [(348, 620), (700, 500), (576, 576), (1000, 1001), (97, 1300), (33, 40), (2000, 1504)],
                         ids=lambda s: '%dx%d' % s)
def test_letterbox_matches_image_read(eng, hw):
    pytest.importorskip('cv2')
    h, w = hw
    rgb = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want, win = IO.letterbox(rgb, 576)
    got, gwin = eng.letterbox(rgb)
    got = got.cpu().numpy()
    assert np.array_equal(gwin, win)
    d = np.abs(got.astype(np.float64) - want)
    print('max abs diff %.3g' % d.max())
    assert d.max() < 4e-7
    # the padding is exactly 127/255 and the embedded region is where the reference puts it
    top, left = int(round(win[0] * 576)), int(round(win[1] * 576))
    if top > 0:
        assert np.all(got[:top] == np.float32(127.0 / 255.0))
    if left > 0:
        assert np.all(got[:, :left] == np.float32(127.0 / 255.0))


def _detections(rng, n, S, window):
    """n boxes inside the letterbox window + smooth random masks (so that the 0.5 level set is a curve)."""
    det = np.zeros((30, 6), np.float32)
    wy1, wx1, wy2, wx2 = window
    for k in range(n):
        cy, cx = rng.uniform(wy1, wy2), rng.uniform(wx1, wx2)
        hh, hw = rng.uniform(0.01, 0.3), rng.uniform(0.01, 0.3)
        det[k, :4] = [max(cy - hh, wy1), max(cx - hw, wx1), min(cy + hh, wy2), min(cx + hw, wx2)]
        det[k, 4] = rng.integers(0, 3)
        det[k, 5] = rng.uniform(0.3, 1.0)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32) / S
    masks = np.zeros((30, S, S), np.float32)
    for k in range(n):
        a, b, c = rng.uniform(2, 9, 3)
        masks[k] = 0.5 + 0.45 * np.sin(a * yy * 6.28 + c) * np.cos(b * xx * 6.28)
    return det, masks


@pytest.mark.parametrize('hw', [(348, 620), (700, 500), (1200, 1600)], ids=lambda s: '%dx%d' % s)
def test_postprocess_matches_reference_loop(eng, hw):
    pytest.importorskip('cv2')
    import torch
    h, w = hw
    rng = np.random.default_rng(h + w)
    _, win = IO.letterbox(np.zeros((h, w, 3), np.uint8), 576)
    n = 17
    det, masks = _detections(rng, n, 288, win)
    det[5, :4] = [0.5, 0.5, 0.5, 0.7]            # zero-height box: skipped by the reference (`continue`)
    boxes, valid, full, merged = IO.postprocess(det[:n], masks[:n], h, w, 576)
    out = eng.postprocess(torch.from_numpy(det).cuda(), torch.tensor([n], dtype=torch.int32).cuda(),
                          torch.from_numpy(masks).cuda(), h, w)
    gb, gv = out['boxes'].cpu().numpy(), out['valid'].cpu().numpy()
    gf, gm = out['full_masks'].cpu().numpy().astype(bool), out['merged'].cpu().numpy()
    assert np.array_equal(gb[:n], boxes)
    assert np.array_equal(gv[:n].astype(bool), valid) and not gv[n:].any()
    # (full_masks planes at or beyond det_count are unspecified, like dy_forward's mask rows)
    diff = 0
    for k in range(n):
        inter, union = np.logical_and(gf[k], full[k]).sum(), np.logical_or(gf[k], full[k]).sum()
        assert union == 0 or inter / union >= 0.99, (k, inter, union)
        diff += int(np.logical_xor(gf[k], full[k]).sum())
    print('differing mask pixels: %d of %d' % (diff, n * h * w))
    assert diff <= 5e-4 * n * h * w
    assert np.mean(gm == merged) >= 0.9995
    assert valid.sum() == n - 1


def test_letterbox_feeds_the_network(eng):
    """uint8 image -> dy_letterbox -> dy_forward on the same stream: identical to feeding the oracle's
    letterboxed image (the reference's calculate_test_map.py:208-218 sequence)."""
    pytest.importorskip('cv2')
    import torch
    import disyolo_b200 as dy
    from oracle import dis_oracle as O
    e = dy.Engine(image_size=160, max_batch=1, precision='bf16')
    e.load_weights(O.make_weights('lively', 0))
    rgb = np.random.default_rng(3).integers(0, 256, (120, 200, 3), dtype=np.uint8)
    img, win = e.letterbox(rgb)
    want_img, want_win = IO.letterbox(rgb, 160)
    a = e.forward(img[None], torch.from_numpy(win[None]).cuda(), 0.2)
    b = e.forward(torch.from_numpy(want_img.astype(np.float32)[None]).cuda(), torch.from_numpy(want_win[None]).cuda(), 0.2)
    torch.cuda.synchronize()
    assert int(a['det_count'][0]) == int(b['det_count'][0])
    assert torch.allclose(a['det_raw'], b['det_raw'], atol=1e-5)
    e.close()


@pytest.mark.parametrize('sizes', [[(120, 200), (240, 180), (160, 160)], [(200, 180)] * 3],
                         ids=['mixed_shapes', 'one_shape_batched_launches'])
def test_image_pipeline_matches_stepwise_calls(sizes):
    """ImagePipeline (uint8 images in, boxes + merged masks out, two batches in flight) returns exactly what
    letterbox -> forward -> postprocess return when called one after the other, and its instance masks
    agree with the oracle's loop on the same detections.  A batch of same-shape frames goes through
    dy_letterbox_batch / dy_postprocess_batch (one launch each), mixed shapes through the per-image calls."""
    pytest.importorskip('cv2')
    import torch
    import disyolo_b200 as dy
    from oracle import dis_oracle as O
    B, S = 3, 160
    e = dy.Engine(image_size=S, max_batch=B, precision='bf16')
    e.load_weights(O.make_weights('lively', 0))
    rng = np.random.default_rng(9)
    batches = [[rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes] for _ in range(3)]
    pipe = dy.ImagePipeline(e, 240, 200, depth=2, want_instance_masks=True)
    tickets = [pipe.submit(batches[0], 0.2), pipe.submit(batches[1], 0.2)]
    with pytest.raises(Exception):
        pipe.submit(batches[2], 0.2)
    results = [pipe.result(tickets[0])]
    results[0] = [{k: (v.copy() if v is not None else None) for k, v in r.items()} for r in results[0]]
    tickets.append(pipe.submit(batches[2], 0.2))
    for tk in tickets[1:]:
        results.append([{k: (v.copy() if v is not None else None) for k, v in r.items()} for r in pipe.result(tk)])
    total = 0
    for imgs, res in zip(batches, results):
        lb = [e.letterbox(im) for im in imgs]
        batch = torch.stack([x[0] for x in lb])
        win = torch.from_numpy(np.stack([x[1] for x in lb])).cuda()
        out = e.forward(batch, win, 0.2)
        torch.cuda.synchronize()
        for b, (im, r) in enumerate(zip(imgs, res)):
            h, w = im.shape[:2]
            n = int(out['det_count'][b])
            assert len(r['boxes']) == n
            total += n
            pp = e.postprocess(out['det_box'][b], out['det_count'][b:b + 1], out['masks'][b], h, w)
            assert np.array_equal(r['boxes'], pp['boxes'].cpu().numpy()[:n])
            assert np.array_equal(r['merged'], pp['merged'].cpu().numpy())
            assert np.array_equal(r['masks'], pp['full_masks'].cpu().numpy()[:n].astype(bool))
            if n:
                ob, ov, of, om = IO.postprocess(out['det_box'][b, :n].cpu().numpy(), out['masks'][b, :n].cpu().numpy(), h, w, S)
                assert np.array_equal(r['boxes'], ob) and np.array_equal(r['valid'], ov)
                assert np.mean(r['merged'] == om) >= 0.999
    assert total > 0
    e.close()
