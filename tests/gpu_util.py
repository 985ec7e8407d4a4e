"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle / golden vectors)."""
import numpy as np
import torch

import smoe_stereo_b200 as sb


def dev():
    return torch.device("cuda", 0)


def fused_from_levels(levels, dtype=torch.float32):
    """Upload reference pyramid levels [(B,H,W1,W2_i)] into a FusedPyramid (poisoned padding)."""
    B, H, W1, W2 = levels[0].shape
    pyr = sb.FusedPyramid(B, H, W1, W2, len(levels), dtype, dev())
    pyr.buffer.fill_(float("nan"))
    for i, lv in enumerate(levels):
        assert lv.shape[-1] == pyr.widths[i]
        pyr.level(i).copy_(torch.from_numpy(np.ascontiguousarray(lv)).to(dev()).to(dtype))
    return pyr


def make_coords(B, H, W1, W2, T, seed, oob_frac=0.02):
    """(T,B,2,H,W1) fp32: grid minus a smooth-ish disparity plus noise; >=1% out of range on each
    side, some exact integers."""
    rs = np.random.RandomState(seed)
    xs = np.broadcast_to(np.arange(W1, dtype=np.float32), (B, H, W1))
    ys = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None], (B, H, W1))
    out = np.empty((T, B, 2, H, W1), dtype=np.float32)
    for t in range(T):
        disp = rs.uniform(0, min(64.0, W2 / 3), size=(B, H, 1)).astype(np.float32) \
            + rs.standard_normal((B, H, W1)).astype(np.float32) * (0.5 * t + 0.5)
        x = (xs - disp).astype(np.float32)
        m = rs.uniform(size=x.shape)
        x = np.where(m < oob_frac, -rs.uniform(1, 40, size=x.shape), x)
        x = np.where(m > 1 - oob_frac, W2 + rs.uniform(0, 40, size=x.shape), x)
        x = np.where((m > 0.5) & (m < 0.52), np.round(x), x)
        out[t, :, 0] = x.astype(np.float32)
        out[t, :, 1] = ys
    return out
