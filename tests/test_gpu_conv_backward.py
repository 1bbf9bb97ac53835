"""GPU: dgrad / wgrad of the tensor-core training engine (dy_conv_backward) against the float64
autograd oracle of the same 'SAME' convolution (oracle.dis_oracle_train.conv_backward).

Operands are rounded to bf16 on both sides, so what remains is fp32 accumulation order (dw, fp32
output: tolerance 2e-3 norm-wise, measured ~1e-5..1e-4) and the bf16 rounding of the stored dx
(tolerance 1e-2, the per-layer bf16 bound of north_star)."""
import numpy as np
import pytest

from oracle import dis_oracle_train as T
from tests.util import bf16_round, rel_err

pytestmark = pytest.mark.gpu

# (B, H, W, cin, cout, k) -- the stage-1 shapes of layers 53..82 plus edge geometries
CASES = [
    (2, 18, 18, 128, 256, 3),      # conv70/72/74
    (2, 18, 18, 256, 128, 1),      # conv71/73
    (2, 10, 10, 512, 1024, 3),     # conv54/56/58: 4 ci blocks x 4 N tiles x 9 taps
    (2, 10, 10, 1024, 512, 1),     # conv53/55/57
    (2, 36, 36, 32, 64, 3),        # conv81: 32-channel source (SWIZZLE_64B boxes, zero-filled ci rows)
    (2, 36, 36, 64, 9, 1),         # conv82: cout 9 padded to 32 gradient channels
    (2, 18, 18, 256, 24, 1),       # conv75 head: cout 24 padded to 32
    (1, 20, 12, 96, 32, 1),        # conv80: 96 = 3 boxes of 32 channels, non-square
    (3, 9, 9, 192, 64, 1),         # conv77: 192 channels = 1.5 blocks of 128
    (5, 7, 7, 64, 128, 3),         # conv78: many tiny images, K tail (rows not a multiple of 64)
]


# the five down-sampling convs (3x3 stride 2, TF 'SAME' = pad 0 before / 1 after on even extents): wgrad over the
# space-to-depth copy of the input, dgrad as four parity-block GEMMs scattered back (OUT_UNS2D)
CASES_S2 = [
    (2, 32, 32, 32, 64),           # conv2 (32-channel source, SWIZZLE_64B boxes)
    (2, 24, 24, 64, 128),          # conv5
    (1, 20, 12, 128, 256),         # conv10, non-square
    (3, 12, 12, 256, 512),         # conv27
    (2, 8, 8, 512, 1024),          # conv44
]


@pytest.mark.parametrize('case', CASES_S2, ids=lambda c: 'B%d_%dx%d_%d-%d_s2' % c)
def test_conv_backward_stride2_tc(case):
    import torch
    import disyolo_b200.engine as E
    B, H, W, cin, cout = case
    rng = np.random.default_rng(sum(case))
    x = bf16_round(rng.standard_normal((B, H, W, cin)).astype(np.float32))
    dz = bf16_round((rng.standard_normal((B, H // 2, W // 2, cout)) * 0.1).astype(np.float32))
    w = bf16_round((rng.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32))
    dx_ref, dw_ref = T.conv_backward(x, dz, w, stride=2)
    dx, dw = E.conv_backward(torch.from_numpy(x).cuda(), torch.from_numpy(dz).cuda(), w, stride=2)
    torch.cuda.synchronize()
    e_dw = rel_err(dw.cpu().numpy(), dw_ref)
    e_dx = rel_err(dx.cpu().numpy(), dx_ref)
    print('stride 2: dw rel err %.3g  dx rel err %.3g' % (e_dw, e_dx))
    assert e_dw < 2e-3, 'wgrad rel err %.3g' % e_dw
    assert e_dx < 1e-2, 'dgrad rel err %.3g' % e_dx


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'B%d_%dx%d_%d-%d_k%d' % c)
def test_conv_backward_tc(case):
    import torch
    import disyolo_b200.engine as E
    B, H, W, cin, cout, k = case
    rng = np.random.default_rng(sum(case))
    x = bf16_round(rng.standard_normal((B, H, W, cin)).astype(np.float32))
    dz = bf16_round((rng.standard_normal((B, H, W, cout)) * 0.1).astype(np.float32))
    w = bf16_round((rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32))
    dx_ref, dw_ref = T.conv_backward(x, dz, w)
    dx, dw = E.conv_backward(torch.from_numpy(x).cuda(), torch.from_numpy(dz).cuda(), w)
    torch.cuda.synchronize()
    e_dw = rel_err(dw.cpu().numpy(), dw_ref)
    e_dx = rel_err(dx.cpu().numpy(), dx_ref)
    print('dw rel err %.3g  dx rel err %.3g' % (e_dw, e_dx))
    assert e_dw < 2e-3, 'wgrad rel err %.3g' % e_dw
    assert e_dx < 1e-2, 'dgrad rel err %.3g' % e_dx


def test_wgrad_is_linear_in_dz():
    """Size-independent property at a full-size shape (conv81 at 288^2, batch 2): wgrad(dz1 + dz2)
    = wgrad(dz1) + wgrad(dz2) up to fp32 summation order."""
    import torch
    import disyolo_b200.engine as E
    rng = np.random.default_rng(11)
    B, H, cin, cout = 2, 288, 32, 64
    x = torch.from_numpy(bf16_round(rng.standard_normal((B, H, H, cin)).astype(np.float32))).cuda()
    # multiples of 1/8 in [-1, 1]: dz1 + dz2 is exact in bf16
    d1 = torch.from_numpy((rng.integers(-8, 9, (B, H, H, cout)) / 8.0).astype(np.float32)).cuda()
    d2 = torch.from_numpy((rng.integers(-8, 9, (B, H, H, cout)) / 8.0).astype(np.float32)).cuda()
    w = np.zeros((3, 3, cin, cout), np.float32)
    _, a = E.conv_backward(x, d1, w, want_dx=False)
    _, b = E.conv_backward(x, d2, w, want_dx=False)
    _, c = E.conv_backward(x, d1 + d2, w, want_dx=False)
    torch.cuda.synchronize()
    e = rel_err(c.cpu().numpy(), (a + b).cpu().numpy())
    print('linearity rel err %.3g' % e)
    assert e < 1e-4
