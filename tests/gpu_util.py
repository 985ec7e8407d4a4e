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


XCHECK_LIB = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)),
                                        "_build", "libsmoecorr_xcheck.so")


class xcheck_lib:
    """Context manager: run the package on the CROSS-CHECK library (tests/_build/libsmoecorr_xcheck.so:
    the product kernels plus the superseded ones and the debug selectors, csrc/xcheck_api.h) instead
    of libsmoecorr.so.  `with xcheck_lib(bwd_staged=1) as L:` also sets Tuning fields for the block."""

    def __init__(self, **tuning):
        self.tuning = tuning
        self.saved = {}

    def __enter__(self):
        import ctypes
        import os
        import pytest
        from smoe_stereo_b200 import _lib, corr
        if not os.path.exists(XCHECK_LIB):
            pytest.skip("tests/_build/libsmoecorr_xcheck.so not built (python smoe-stereo_b200/build.py --xcheck)")
        L = ctypes.CDLL(XCHECK_LIB)
        missing = [s for s in _lib.SYMBOLS + _lib.XCHECK_SYMBOLS if not hasattr(L, s)]
        assert not missing, missing
        _lib._declare(L)
        self.L, self.old = L, _lib._lib
        _lib._lib = L
        _lib._fast = None
        corr._lookup_fn = None
        for k, v in self.tuning.items():
            assert L.smoe_corr_debug_set_tuning(k.encode(), int(v)) == 0, k
        return L

    def __exit__(self, *exc):
        from smoe_stereo_b200 import _lib, corr
        defaults = dict(pdl=2, fwd_kernel=0, px_tpb=32, px_split=1, px_ld256=-1, px_hint=-1, lookup_tile=32,
                        bwd_staged=0, bwd_kernel=1, bwd_rowmod=2, build_pair=0, build_stages=0, build_debug=0,
                        conv_lanes=0)
        for k in self.tuning:
            self.L.smoe_corr_debug_set_tuning(k.encode(), defaults[k])
        self.L.smoe_corr_debug_set_lookup_kernel(-1)
        _lib._lib = self.old
        _lib._fast = None
        corr._lookup_fn = None
