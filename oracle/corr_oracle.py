"""CPU oracle for the SMoE-Stereo 1-D correlation path -- TEST INFRASTRUCTURE ONLY.

Two independent restatements of the reference algorithm live here:

* ``numpy_*``  : vectorised numpy (fp64 accumulate for the volume), used by the tests.
* ``c_*``      : ctypes bindings of ``corr_oracle.c`` (fp32, OpenMP), used by the tests
                 and as bench.py's timed CPU baseline ("port").

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``smoe-stereo_b200/``) never does.

Parity pinning: the reference has no tests or golden vectors for this path
(SURVEY.md section 4); both restatements are pinned against outputs of the reference's
own ``core/corr.py`` executed in the build container, stored in ``tests/golden/``
(generator: ``tests/golden/make_golden.py``).

Reference citations (relative to the reference repo root):
  volume   core/corr.py:148-156     pyramid  core/corr.py:122-125
  lookup   core/corr.py:127-146 + core/utils/utils.py:84-99 (== sampler/sampler_kernel.cu:39-59)
  backward sampler/sampler_kernel.cu:83-103
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcorr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile corr_oracle.c with the committed Makefile (gcc, no external deps)."""
    src = os.path.join(_HERE, "corr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        i = ctypes.c_int
        L.oracle_num_threads.restype = i
        L.oracle_set_num_threads.argtypes = [i]
        L.oracle_level_width.argtypes = [i, i]
        L.oracle_level_width.restype = i
        L.oracle_corr_volume.argtypes = [fp, fp, fp, i, i, i, i, i]
        L.oracle_pool_level.argtypes = [fp, fp, ctypes.c_long, i]
        L.oracle_lookup_level.argtypes = [fp, fp, i, fp, i, i, i, i, i, ctypes.c_float]
        L.oracle_lookup_level_bwd.argtypes = [fp, i, fp, fp, i, i, i, i, i, ctypes.c_float, i]
        L.oracle_fold_pyramid_grad.argtypes = [fp, fp, ctypes.c_long, i]
        L.oracle_corr_volume_bwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, i]
        L.oracle_corr_path.argtypes = [fp, fp, fp, i, fp, fp, fp, i, i, i, i, i, i, i, i]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def level_widths(w2: int, num_levels: int) -> list[int]:
    """Width of each pyramid level: floor-halving (core/corr.py:124)."""
    out = []
    for _ in range(num_levels):
        out.append(w2)
        w2 //= 2
    return out


# --------------------------------------------------------------------------- numpy
def numpy_corr_volume(f1: np.ndarray, f2: np.ndarray) -> np.ndarray:
    """(B,C,H,W1),(B,C,H,W2) -> (B,H,W1,W2); core/corr.py:148-156."""
    C = f1.shape[1]
    v = np.einsum("bchk,bchw->bhkw", f1.astype(np.float64), f2.astype(np.float64))
    return (v / np.sqrt(np.float64(C))).astype(np.float32)


def numpy_pyramid(vol: np.ndarray, num_levels: int) -> list[np.ndarray]:
    """[(B,H,W1,W2_i)] with W2_{i+1} = floor(W2_i/2); core/corr.py:122-125."""
    levels = [vol]
    for _ in range(num_levels - 1):
        cur = levels[-1]
        wo = cur.shape[-1] // 2
        nxt = ((cur[..., 0:2 * wo:2] + cur[..., 1:2 * wo:2]) / np.float32(2.0)).astype(np.float32)
        levels.append(nxt)
    return levels


def numpy_lookup_level(vol: np.ndarray, coords_x: np.ndarray, radius: int) -> np.ndarray:
    """vol (B,H,W1,W2), coords_x (B,H,W1) already scaled to this level -> (B,2r+1,H,W1).

    sampler/sampler_kernel.cu:39-59: out[k] = (1-dx)*V[fl-r+k] + dx*V[fl-r+k+1], zero outside.
    """
    B, H, W1, W2 = vol.shape
    x0 = coords_x.astype(np.float32)
    fl = np.floor(x0)
    dx = (x0 - fl).astype(np.float32)
    ok = np.isfinite(fl) & (np.abs(fl) < 1.0e8)
    base = np.where(ok, fl, -1.0e7).astype(np.int64) - radius
    out = np.zeros((B, 2 * radius + 1, H, W1), dtype=np.float32)
    for k in range(2 * radius + 1):
        acc = np.zeros((B, H, W1), dtype=np.float32)
        for t, w in ((k, np.float32(1.0) - dx), (k + 1, dx)):
            idx = base + t
            inb = (idx >= 0) & (idx < W2) & ok
            g = np.take_along_axis(vol, np.clip(idx, 0, W2 - 1)[..., None], axis=-1)[..., 0]
            acc = acc + np.where(inb, g * w, np.float32(0)).astype(np.float32)
        out[:, k] = acc
    return out


def numpy_lookup(levels: list[np.ndarray], coords: np.ndarray, radius: int) -> np.ndarray:
    """levels[i] (B,H,W1,W2_i), coords (B,>=1,H,W1) -> (B,L*(2r+1),H,W1); channel = level*(2r+1)+tap
    (core/corr.py:143-146); level i samples at coords/2**i (core/corr.py:137)."""
    outs = []
    for i, lv in enumerate(levels):
        outs.append(numpy_lookup_level(lv, coords[:, 0] / np.float32(2 ** i), radius))
    return np.concatenate(outs, axis=1)


def numpy_conv1x1(corr: np.ndarray, weight: np.ndarray, bias, relu: bool = True) -> np.ndarray:
    """The layer that consumes the lookup: `F.relu(self.convc1(corr))`, a 1x1 Conv2d
    (core/update.py:70,77).  corr (B,Cin,H,W), weight (Cout,Cin), bias (Cout,) or None; fp32
    products accumulated in fp64 and rounded once (the reference accumulates in fp32)."""
    y = np.einsum("oc,bchw->bohw", weight.astype(np.float64), corr.astype(np.float64))
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)[None, :, None, None]
    if relu:
        y = np.maximum(y, 0.0)
    return y.astype(np.float32)


def numpy_lookup_level_bwd(coords_x: np.ndarray, grad: np.ndarray, W2: int, radius: int) -> np.ndarray:
    """Transpose of numpy_lookup_level: (B,2r+1,H,W1) -> (B,H,W1,W2); sampler_kernel.cu:83-103."""
    B, _, H, W1 = grad.shape
    x0 = coords_x.astype(np.float32)
    fl = np.floor(x0)
    dx = (x0 - fl).astype(np.float32)
    ok = np.isfinite(fl) & (np.abs(fl) < 1.0e8)
    base = np.where(ok, fl, -1.0e7).astype(np.int64) - radius
    out = np.zeros((B, H, W1, W2), dtype=np.float32)
    bi, hi, wi = np.meshgrid(np.arange(B), np.arange(H), np.arange(W1), indexing="ij")
    for k in range(2 * radius + 1):
        g = grad[:, k]
        for t, w in ((k, np.float32(1.0) - dx), (k + 1, dx)):
            idx = base + t
            inb = (idx >= 0) & (idx < W2) & ok
            np.add.at(out, (bi[inb], hi[inb], wi[inb], idx[inb]), (g * w)[inb])
    return out


def numpy_fold_pyramid_grad(glevels: list[np.ndarray]) -> np.ndarray:
    """Fold per-level gradients onto level 0 (avg_pool2d backward chain, core/corr.py:122-125)."""
    g = [x.astype(np.float32).copy() for x in glevels]
    for i in range(len(g) - 1, 0, -1):
        wo = g[i].shape[-1]
        g[i - 1][..., 0:2 * wo:2] += np.float32(0.5) * g[i]
        g[i - 1][..., 1:2 * wo:2] += np.float32(0.5) * g[i]
    return g[0]


def numpy_corr_volume_bwd(f1, f2, G):
    """G (B,H,W1,W2) -> (df1, df2); autograd of core/corr.py:154-156."""
    C = f1.shape[1]
    s = 1.0 / np.sqrt(np.float64(C))
    df1 = np.einsum("bhkw,bchw->bchk", G.astype(np.float64), f2.astype(np.float64)) * s
    df2 = np.einsum("bhkw,bchk->bchw", G.astype(np.float64), f1.astype(np.float64)) * s
    return df1.astype(np.float32), df2.astype(np.float32)


# --------------------------------------------------------------------------- C (ctypes)
def c_corr_volume(f1: np.ndarray, f2: np.ndarray) -> np.ndarray:
    B, C, H, W1 = f1.shape
    W2 = f2.shape[3]
    out = np.empty((B, H, W1, W2), dtype=np.float32)
    lib().oracle_corr_volume(_p(f1), _p(f2), _p(out), B, C, H, W1, W2)
    return out


def c_pyramid(vol: np.ndarray, num_levels: int) -> list[np.ndarray]:
    B, H, W1, W2 = vol.shape
    levels = [vol]
    for _ in range(num_levels - 1):
        cur = levels[-1]
        nxt = np.empty((B, H, W1, cur.shape[-1] // 2), dtype=np.float32)
        lib().oracle_pool_level(_p(cur), _p(nxt), B * H * W1, cur.shape[-1])
        levels.append(nxt)
    return levels


def c_lookup(levels: list[np.ndarray], coords: np.ndarray, radius: int) -> np.ndarray:
    B, H, W1, _ = levels[0].shape
    rd = 2 * radius + 1
    outs = []
    for i, lv in enumerate(levels):
        o = np.empty((B, rd, H, W1), dtype=np.float32)
        lib().oracle_lookup_level(_p(lv), _p(coords), coords.shape[1], _p(o), B, H, W1,
                                  lv.shape[-1], radius, ctypes.c_float(0.5 ** i))
        outs.append(o)
    return np.concatenate(outs, axis=1)


def c_lookup_level_bwd(coords: np.ndarray, grad: np.ndarray, W2: int, radius: int,
                       coord_scale: float = 1.0, into: np.ndarray | None = None) -> np.ndarray:
    B, _, H, W1 = grad.shape
    acc = into is not None
    out = into if acc else np.empty((B, H, W1, W2), dtype=np.float32)
    lib().oracle_lookup_level_bwd(_p(coords), coords.shape[1], _p(grad), _p(out), B, H, W1, W2,
                                  radius, ctypes.c_float(coord_scale), 1 if acc else 0)
    return out


def c_fold_pyramid_grad(glevels: list[np.ndarray]) -> np.ndarray:
    g = [np.ascontiguousarray(x, dtype=np.float32).copy() for x in glevels]
    for i in range(len(g) - 1, 0, -1):
        rows = int(np.prod(g[i - 1].shape[:-1]))
        lib().oracle_fold_pyramid_grad(_p(g[i - 1]), _p(g[i]), rows, g[i - 1].shape[-1])
    return g[0]


def c_corr_volume_bwd(f1, f2, G):
    B, C, H, W1 = f1.shape
    W2 = f2.shape[3]
    df1 = np.empty_like(f1)
    df2 = np.empty_like(f2)
    lib().oracle_corr_volume_bwd(_p(f1), _p(f2), _p(G), _p(df1), _p(df2), B, C, H, W1, W2)
    return df1, df2


class CorrPathBaseline:
    """The reference's CPU hot path (build + `iters` lookups) through the C port; reusable
    buffers so that the timed region contains only the arithmetic (bench.py cpu_baseline)."""

    def __init__(self, B, C, H, W1, W2, num_levels, radius, iters, coords_ch=2):
        self.shape = (B, C, H, W1, W2, num_levels, radius, iters, coords_ch)
        rows = B * H * W1
        self.pyramid = np.empty(rows * sum(level_widths(W2, num_levels)), dtype=np.float32)
        rd = 2 * radius + 1
        self.level_out = np.empty(B * rd * H * W1, dtype=np.float32)
        self.out = np.empty((iters, B, num_levels * rd, H, W1), dtype=np.float32)

    def run(self, f1, f2, coords):
        B, C, H, W1, W2, L, r, iters, cch = self.shape
        assert coords.shape == (iters, B, cch, H, W1)
        lib().oracle_corr_path(_p(f1), _p(f2), _p(coords), cch, _p(self.pyramid),
                               _p(self.level_out), _p(self.out), B, C, H, W1, W2, L, r, iters)
        return self.out


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def use_all_host_threads() -> int:
    """Give the OpenMP port every host core this process may run on (launchers such as torchrun
    export OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().oracle_set_num_threads(n)
    return num_threads()
