"""GPU: every conv flavour of the reference's builders (conv / conv_bn / res_conv_bn,
yolo3_net_pos.py:109-151) through dy_conv_layer, against the oracle on the same seeded inputs.

bf16 engine (tcgen05): inputs and weights are rounded to bf16 on both sides, so the remaining
difference is fp32 accumulation order + the bf16 rounding of the stored output: tolerance 1e-2
relative (north_star: per-layer activations within 1e-2 in bf16).
fp32 verification kernel: tolerance 1e-4 relative (north_star), measured ~1e-6."""
import numpy as np
import pytest

from oracle import dis_oracle as O
from tests.util import bf16_round, rel_err

pytestmark = pytest.mark.gpu

# (B, H, W, cin, cout, k, s, act, residual)
CASES = [
    (2, 18, 18, 1024, 512, 1, 1, True, False),     # bottleneck 1x1 (conv45..)
    (2, 36, 36, 64, 32, 1, 1, True, False),        # cout=32 tile (conv3)
    (3, 20, 12, 96, 32, 1, 1, True, False),        # 96 = 64+32 channels -> 32-wide K chunks (conv80)
    (2, 18, 18, 512, 1024, 3, 1, True, True),      # res_conv_bn (conv46..)
    (2, 24, 24, 32, 64, 3, 1, True, True),         # Cin=32 3x3 (conv4, conv81)
    (1, 16, 20, 64, 128, 3, 1, True, False),       # non-square
    (2, 32, 32, 32, 64, 3, 2, True, False),        # stride 2, Cin=32 (conv2)
    (2, 36, 36, 64, 128, 3, 2, True, False),       # stride 2 (conv5)
    (1, 36, 36, 256, 512, 3, 2, True, False),      # stride 2 (conv27)
    (2, 18, 18, 1024, 32, 1, 1, False, False),     # linear biased head shape (conv59, padded cout)
    (5, 9, 9, 128, 256, 3, 1, True, False),        # many tiny images: tiles straddle image borders
]


def _make(case, seed):
    B, H, W, cin, cout, k, s, act, res = case
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, W, cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = (rng.standard_normal(cout) * 0.3).astype(np.float32)
    r = rng.standard_normal((B, H // s, W // s, cout)).astype(np.float32) if res else None
    return x, w, scale, shift, r


def _oracle(x, w, s, scale, shift, act, r):
    y = O.conv2d_same(x, w, s) * scale + shift
    if act:
        y = O.leaky_relu(y)
    if r is not None:
        y = y + r
    return y


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'B%d_%dx%d_%d-%d_k%ds%d%s%s' % (
    c[0], c[1], c[2], c[3], c[4], c[5], c[6], '_act' if c[7] else '', '_res' if c[8] else ''))
def test_conv_bf16_tcgen05(case):
    import torch
    from disyolo_b200.engine import conv_layer
    x, w, scale, shift, r = _make(case, 11)
    xb, wb = bf16_round(x), bf16_round(w)
    rb = bf16_round(r) if r is not None else None
    want = _oracle(xb, wb, case[6], scale, shift, case[7], rb)
    got = conv_layer(torch.from_numpy(x).cuda(), w, case[6], scale, shift, case[7], 0.1,
                     torch.from_numpy(r).cuda() if r is not None else None, 'bf16').cpu().numpy()
    assert got.shape == want.shape
    e = rel_err(got, want)
    worst = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 0.25)))
    print('bf16 conv %s: rel %.3g worst %.3g' % (str(case), e, worst))
    assert e < 5e-3 and worst < 1e-2


@pytest.mark.parametrize('case', CASES[:9], ids=lambda c: 'B%d_%dx%d_%d-%d_k%ds%d' % c[:7])
def test_conv_fp32_verify(case):
    import torch
    from disyolo_b200.engine import conv_layer
    x, w, scale, shift, r = _make(case, 12)
    want = _oracle(x, w, case[6], scale, shift, case[7], r)
    got = conv_layer(torch.from_numpy(x).cuda(), w, case[6], scale, shift, case[7], 0.1,
                     torch.from_numpy(r).cuda() if r is not None else None, 'fp32').cpu().numpy()
    e = rel_err(got, want)
    print('fp32 conv %s: rel %.3g' % (str(case), e))
    assert e < 1e-5
    assert float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0))) < 1e-4


@pytest.mark.parametrize('shape', [(2, 64, 48), (3, 40, 96), (1, 576, 576), (2, 33, 32)],
                         ids=lambda s: '%dx%dx%d' % s)
def test_conv1_stem(shape):
    """convolutional1 (3->32): W % 32 == 0 runs the tcgen05 im2col stem kernel (bf16 operands),
    other widths the CUDA-core fallback (fp32 operands)."""
    import torch
    from disyolo_b200.engine import conv_layer
    rng = np.random.default_rng(5)
    x = rng.random((shape[0], shape[1], shape[2], 3), dtype=np.float32)
    w = (rng.standard_normal((3, 3, 3, 32)) * 0.3).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 32).astype(np.float32)
    shift = (rng.standard_normal(32) * 0.3).astype(np.float32)
    want = _oracle(x, w, 1, scale, shift, True, None)
    got = conv_layer(torch.from_numpy(x).cuda(), w, 1, scale, shift, True, 0.1, None, 'bf16').cpu().numpy()
    assert rel_err(got, want) < 6e-3          # bf16 operands (tensor-core stem) + bf16 store
    if shape[2] % 32 == 0:
        want_b = _oracle(bf16_round(x), bf16_round(w), 1, scale, shift, True, None)
        assert rel_err(got, want_b) < 4e-3    # same operand rounding on both sides
    got32 = conv_layer(torch.from_numpy(x).cuda(), w, 1, scale, shift, True, 0.1, None, 'fp32').cpu().numpy()
    assert rel_err(got32, want) < 1e-5


def test_conv_linearity_property():
    """conv(a*x) == a*conv(x) for the linear (no activation, zero shift) conv: a size-independent
    property checked at a full-size layer shape (conv54 at batch 4: 18x18x512 -> 1024)."""
    import torch
    from disyolo_b200.engine import conv_layer
    rng = np.random.default_rng(6)
    x = bf16_round(rng.standard_normal((4, 18, 18, 512)).astype(np.float32))
    w = (rng.standard_normal((3, 3, 512, 1024)) / 68.0).astype(np.float32)
    one, zero = np.ones(1024, np.float32), np.zeros(1024, np.float32)
    xc = torch.from_numpy(x).cuda()
    y1 = conv_layer(xc, w, 1, one, zero, False, 0.1, None, 'bf16').cpu().numpy()
    y2 = conv_layer(xc * 2.0, w, 1, one, zero, False, 0.1, None, 'bf16').cpu().numpy()
    assert np.array_equal(y2, 2.0 * y1)       # power-of-two scaling is exact in bf16/fp32


@pytest.mark.parametrize('resident,halo,staged,tma,dual', [(1, 1, 1, 1, 1), (0, 1, 0, 0, 0), (1, 0, 0, 0, 1),
                                                           (0, 0, 1, 0, 0), (0, 1, 1, 1, 1), (0, 0, 1, 1, 0),
                                                           (0, 0, 0, 0, 1)])
@pytest.mark.parametrize('case', [CASES[0], CASES[1], CASES[2], CASES[3], CASES[4], CASES[5], CASES[6], CASES[7],
                                  CASES[8], CASES[10]],
                         ids=lambda c: 'B%d_%dx%d_%d-%d_k%ds%d' % c[:7])
def test_conv_bf16_forced_modes(case, resident, halo, staged, tma, dual):
    """Every planning mode of the tensor-core engine (weights resident in smem or streamed; one
    halo'd activation box shared by the three horizontal taps or one box per tap) must give the
    same result as the default plan -- bit for bit, since the MMA order along K is unchanged."""
    import torch
    from disyolo_b200.engine import conv_layer, set_option
    x, w, scale, shift, r = _make(case, 13)
    xc = torch.from_numpy(x).cuda()
    rc = torch.from_numpy(r).cuda() if r is not None else None
    want = _oracle(bf16_round(x), bf16_round(w), case[6], scale, shift, case[7],
                   bf16_round(r) if r is not None else None)
    try:
        set_option('tc_resident', resident)
        set_option('tc_halo', halo)
        set_option('tc_staged', staged)
        set_option('tc_tma_epi', tma)
        set_option('tc_dual_issue', dual)
        got = conv_layer(xc, w, case[6], scale, shift, case[7], 0.1, rc, 'bf16').cpu().numpy()
    finally:
        set_option('tc_resident', -1)
        set_option('tc_halo', -1)
        set_option('tc_staged', -1)
        set_option('tc_tma_epi', -1)
        set_option('tc_dual_issue', -1)
    e = rel_err(got, want)
    assert e < 5e-3, 'resident=%d halo=%d staged=%d tma=%d rel err %.3g' % (resident, halo, staged, tma, e)
    assert float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 0.25))) < 1e-2


@pytest.mark.parametrize('halo,split,slab', [(0, 0, -1), (1, 1, -1), (0, 1, 32), (1, 0, 32)])
@pytest.mark.parametrize('case', [CASES[0], CASES[3], CASES[5], CASES[7], CASES[8], CASES[10],
                                  (3, 19, 17, 128, 256, 3, 1, True, True),      # odd tile count, residual
                                  (1, 8, 8, 256, 128, 1, 1, True, False)],      # a single (half-empty) pair tile
                         ids=lambda c: 'B%d_%dx%d_%d-%d_k%ds%d' % c[:7])
def test_conv_bf16_cta_pairs(case, halo, split, slab):
    """CTA-pair plans (2-CTA clusters, tcgen05.mma.cta_group::2 with M = 256, each CTA staging half of the weight
    tile) forced on small shapes -- odd numbers of 128-row tiles (the second CTA of the last pair works on rows
    beyond the tensor), residuals, stride 2, with and without the shared halo box / the column-split epilogue /
    32-column staging slabs -- against the oracle, and bit for bit against the single-CTA plan with the same K order."""
    import torch
    from disyolo_b200.engine import conv_layer, set_option
    x, w, scale, shift, r = _make(case, 17)
    xc = torch.from_numpy(x).cuda()
    rc = torch.from_numpy(r).cuda() if r is not None else None
    want = _oracle(bf16_round(x), bf16_round(w), case[6], scale, shift, case[7],
                   bf16_round(r) if r is not None else None)
    opts = dict(tc_cta2=1, tc_halo=halo, tc_split_n=split, tc_slab=slab)
    try:
        for k, v in opts.items():
            set_option(k, v)
        got = conv_layer(xc, w, case[6], scale, shift, case[7], 0.1, rc, 'bf16').cpu().numpy()
        set_option('tc_cta2', 0)
        single = conv_layer(xc, w, case[6], scale, shift, case[7], 0.1, rc, 'bf16').cpu().numpy()
    finally:
        for k in opts:
            set_option(k, -1)
    e = rel_err(got, want)
    assert e < 5e-3, 'cta2 halo=%d split=%d slab=%d rel err %.3g' % (halo, split, slab, e)
    assert float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 0.25))) < 1e-2
    assert np.array_equal(got, single), 'CTA-pair plan differs from the single-CTA plan (same K order)'
