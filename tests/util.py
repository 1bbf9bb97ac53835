"""Shared helpers for the parity tests (oracle on one side, the CUDA library on the other)."""
import numpy as np


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b||."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def bf16_round(x):
    """Round fp32 -> bf16 -> fp32 (round-to-nearest-even), NumPy only."""
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def mask_iou(a, b, thr=0.5):
    a = np.asarray(a) > thr; b = np.asarray(b) > thr
    u = np.logical_or(a, b).sum()
    return 1.0 if u == 0 else float(np.logical_and(a, b).sum()) / float(u)


def synthetic_heads(rng, B, size, num_class=3, obj_bias=-3.0):
    """Random head maps [B,g,g,3,5+C] for strides 8/16/32 with a sparse objectness."""
    out = []
    for s in (8, 16, 32):
        g = size // s
        y = rng.standard_normal((B, g, g, 3, 5 + num_class)).astype(np.float32)
        y[..., 2:4] *= 0.5
        y[..., 4] = y[..., 4] * 1.5 + obj_bias
        y[..., 5:] *= 2.0
        out.append(y)
    return out
