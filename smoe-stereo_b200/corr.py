"""Drop-in replacement for the reference's ``core/corr.py`` on B200.

Same names, constructor signatures and call semantics as the reference
(``/root/reference/core/corr.py``):

* ``CorrBlock1D(fmap1, fmap2, num_levels=4, radius=4)`` + ``__call__(coords)``     core/corr.py:110-156
* ``CorrBlockFast1D(...)``                                                         core/corr.py:31-61
* ``PytorchAlternateCorrBlock1D(...)``                                             core/corr.py:64-107
* ``CorrSampler`` autograd.Function over ``corr_sampler.forward/backward``         core/corr.py:17-29

but the arithmetic runs in hand-written sm_100a kernels behind the C ABI of libsmoecorr.so
(include/smoe_corr.h): the volume + 4-level pyramid is ONE tcgen05/TMA kernel, every lookup is
ONE fused kernel over all levels, and the lookup backward accumulates into one persistent
gradient pyramid per training step instead of a dense volume per call.

Host code is plumbing only: allocation (torch caching allocator), stream/device selection,
argument checks that raise RuntimeError like TORCH_CHECK (sampler/sampler.cpp:20-22).  There is
no CPU or PyTorch fallback -- tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.nn.functional as F

from . import _lib
from . import corr_sampler

__all__ = ["CorrBlock1D", "CorrBlockFast1D", "PytorchAlternateCorrBlock1D", "CorrSampler",
           "AlternateCorrBlock", "set_default_output_ring", "set_default_volume_dtype"]

_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16}

# Opt-in storage format of the volume/pyramid for fp32 feature maps in CorrBlock1D / the "alt"
# alias: None = fp32 (the reference's "reg" semantics, default); torch.float16 = round the fp32
# tensor-core result once to fp16 (~5e-4 relative, inside north_star's 1e-3), halving the pyramid
# so that e.g. the B=8 training volume (166 MB) becomes L2-resident.  Lookup outputs stay fp32.
# Also settable with SMOE_CORR_VOLUME_DTYPE=fp16 in the environment.
_default_volume_dtype = torch.float16 if os.environ.get("SMOE_CORR_VOLUME_DTYPE", "").lower() in (
    "fp16", "float16", "half") else None


# SMOE_CORR_ALL_LEVELS=1: always write all four pyramid levels; =0: never write levels 1 and 3 for
# radius-4 blocks (the default, see _CorrBlockBase.__init__)
_skip_odd_levels = {"": None, "0": True, "1": False}.get(os.environ.get("SMOE_CORR_ALL_LEVELS", ""), None)
_SKIP_ODD_MIN_BYTES = 0             # None above: skip levels 1/3 for volumes of at least this size


# Opt-in (set_default_output_ring / install(output_ring=n)): every block serves its no-grad lookups from
# a ring of n preallocated output buffers (see _CorrBlockBase.reuse_outputs).  0 = a fresh tensor per call.
_default_output_ring = 0


def set_default_output_ring(depth: int) -> None:
    """depth > 0: blocks constructed from now on call reuse_outputs(depth) on themselves -- for update
    loops that consume `corr` inside the iteration that produced it (core/raft_stereo.py:205-214,
    core/SMoEStereo.py:233-248).  Only affects calls made while autograd is not recording."""
    global _default_output_ring
    _default_output_ring = max(int(depth), 0)


def set_default_volume_dtype(dtype) -> None:
    """None / torch.float32 -> fp32 pyramids (default); torch.float16 -> fp16 storage (see above)."""
    global _default_volume_dtype
    if dtype not in (None, torch.float32, torch.float16):
        raise ValueError("volume dtype must be None, torch.float32 or torch.float16")
    _default_volume_dtype = None if dtype in (None, torch.float32) else dtype


# ------------------------------------------------------------------------------------ helpers
def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


# torch.cuda.current_stream() builds a Python Stream object (~6 us per call, half of an eager lookup's
# host cost); the raw getters are what torch itself uses underneath.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    if _raw_stream is not None:
        idx = device.index if device.index is not None else _raw_device()
        return ctypes.c_void_p(_raw_stream(idx))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _on_device:
    """Make `device` current for the duration of a C-ABI call (no-op when it already is)."""

    __slots__ = ("idx", "prev")

    def __init__(self, device: torch.device):
        self.idx = device.index if device.index is not None else _raw_device()
        self.prev = None

    def __enter__(self):
        cur = _raw_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


class FusedPyramid:
    """Device buffer holding every level of every pixel row contiguously (include/smoe_corr.h).

    buffer: (B*H*W1, pitch) ; level i of a pixel occupies [offset_i, offset_i + W2_i).
    """

    def __init__(self, B, H, W1, W2, num_levels, dtype, device, zero=False):
        if not 1 <= num_levels <= _lib.MAX_LEVELS:
            raise RuntimeError(f"num_levels={num_levels} outside [1,{_lib.MAX_LEVELS}]")
        self.B, self.H, self.W1, self.W2 = B, H, W1, W2
        self.num_levels, self.dtype, self.device = num_levels, dtype, device
        self.widths, self.offsets, self.pitch = _lib.pyramid_layout(W2, num_levels, _DTYPE_CODE[dtype])
        rows = B * H * W1
        alloc = torch.zeros if zero else torch.empty
        self.buffer = alloc((rows, self.pitch), dtype=dtype, device=device)
        self.c = _lib.Pyramid()
        es = self.buffer.element_size()
        base = self.buffer.data_ptr()
        for i in range(num_levels):
            self.c.level_ptr[i] = base + self.offsets[i] * es
            self.c.width[i] = self.widths[i]
            self.c.pitch[i] = self.pitch
        self.c.num_levels = num_levels
        self.c.dtype = _DTYPE_CODE[dtype]
        self.c.B, self.c.H, self.c.W1 = B, H, W1
        self.c.levels_missing = 0
        self.ref = ctypes.byref(self.c)
        self.addr = ctypes.addressof(self.c)
        self.out_ring = None                            # [buffers, next]: see _CorrBlockBase.reuse_outputs
        self.ring_radius = 0
        # cached for the eager per-call path (_lookup): device ordinal, coords plane geometry
        self.dev_idx = device.index if device.index is not None else _raw_device()
        self._coords_fast = (H * W1, W1, 1)          # strides (ch, H, W) of a contiguous coords tensor
        # (shape, stride) of the contiguous 2- and 1-channel coords tensors the models pass -> batch stride
        self._coords_ok = {(torch.Size((B, c, H, W1)), (c * H * W1, H * W1, W1, 1)): c * H * W1 for c in (1, 2)}
        self._out_shape = {}

    def level(self, i: int) -> torch.Tensor:
        """Strided (B,H,W1,W2_i) view of level i (no copy)."""
        if (self.c.levels_missing >> i) & 1:
            self.materialize()
        return self._level_view(i)

    def _level_view(self, i: int) -> torch.Tensor:
        v = self.buffer[:, self.offsets[i]:self.offsets[i] + self.widths[i]]
        return v.view(self.B, self.H, self.W1, self.widths[i])

    def mark_missing(self, mask: int) -> None:
        """Levels the build did not write (SMOE_BUILD_SKIP_ODD_LEVELS)."""
        self.c.levels_missing = mask & ((1 << self.num_levels) - 1)

    def materialize(self) -> None:
        """Fill in the levels the build skipped (smoe_corr_fill_levels): level i = pair average of
        level i-1 with the very arithmetic of the build epilogue / F.avg_pool2d (fp32 sum, * 0.5,
        one rounding), so the result is bit-identical to a build without
        SMOE_BUILD_SKIP_ODD_LEVELS.  Only needed by consumers other than the radius-4 lookup
        kernels (`corr_pyramid`)."""
        if self.c.levels_missing:
            with _on_device(self.device):
                rc = _lib.lib().smoe_corr_fill_levels(self.ref, _stream_ptr(self.device))
            _lib.check(rc, "smoe_corr_fill_levels")
            self.c.levels_missing = 0


def _is_channels_last(t: torch.Tensor) -> bool:
    """(B,C,H,W)-shaped tensor stored (B,H,W,C) -- and not also NCHW-contiguous (C == 1 etc.)."""
    return (t.dim() == 4 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last))


def _layout_for_build(fmap1: torch.Tensor, fmap2: torch.Tensor):
    """Pick the memory layout the build kernel reads in place.

    Feature maps that arrive channels-last -- what cuDNN emits for a model run in
    ``memory_format=torch.channels_last``, e.g. the decoder's last 1x1 convolution
    (core/extractor.py:565-570) -- are consumed as they are (K-major operands, flag
    SMOE_BUILD_CHANNELS_LAST) when C*itemsize is a multiple of 16; everything else is made NCHW
    contiguous, as the reference's einsum would do internally."""
    if (_is_channels_last(fmap1) or _is_channels_last(fmap2)) and \
            (fmap1.shape[1] * fmap1.element_size()) % 16 == 0:
        f1 = fmap1.contiguous(memory_format=torch.channels_last)
        f2 = fmap2.contiguous(memory_format=torch.channels_last)
        if f1.data_ptr() % 16 == 0 and f2.data_ptr() % 16 == 0:
            return f1, f2
    return _aligned16(fmap1.contiguous()), _aligned16(fmap2.contiguous())


def _aligned16(t: torch.Tensor) -> torch.Tensor:
    """TMA wants 16-byte-aligned bases; a contiguous view that starts mid-allocation is cloned."""
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def _build_into(pyr: FusedPyramid, fmap1: torch.Tensor, fmap2: torch.Tensor, flags: int = 0) -> None:
    B, C, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    L = _lib.lib()
    code = _DTYPE_CODE[fmap1.dtype]
    if _is_channels_last(fmap1):
        if not _is_channels_last(fmap2):
            raise RuntimeError("fmap1 and fmap2 must share one memory layout")
        flags |= _lib.BUILD_CHANNELS_LAST
    elif not (fmap1.is_contiguous() and fmap2.is_contiguous()):
        raise RuntimeError("feature maps must be contiguous (NCHW or channels-last)")
    ws_bytes = 0 if flags & _lib.BUILD_CHANNELS_LAST else \
        L.smoe_corr_build_workspace_bytes(B, C, H, W1, W2, code)
    pyr.mark_missing(0b1010 if flags & _lib.BUILD_SKIP_ODD_LEVELS else 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=fmap1.device) if ws_bytes else None
    with _on_device(fmap1.device):
        rc = L.smoe_corr_build(fmap1.data_ptr(), fmap2.data_ptr(), code, B, C, H, W1, W2, pyr.ref,
                               math.sqrt(C), ws.data_ptr() if ws is not None else None,
                               ws_bytes, flags, _stream_ptr(fmap1.device))
    _lib.check(rc, "smoe_corr_build")
    if ws is not None:
        ws.record_stream(torch.cuda.current_stream(fmap1.device))


def _pitched_f32(t: torch.Tensor):
    """(tensor whose image rows are `pitch` elements apart with pitch % 4 == 0, pitch, is_copy):
    the operand layout smoe_corr_build_bwd reads/writes by TMA (16-byte strides)."""
    W = t.shape[3]
    t = t.float()
    if W % 4 == 0:
        return _aligned16(t.contiguous()), W, False
    Wp = (W + 3) // 4 * 4
    buf = t.new_zeros(t.shape[:3] + (Wp,))
    buf[..., :W].copy_(t)
    return buf, Wp, True


def _build_backward(gp: FusedPyramid, fmap1: torch.Tensor, fmap2: torch.Tensor):
    """Autograd of core/corr.py:154-156 given the folded level-0 gradient in `gp`:
    d fmap1 = G . fmap2^T / sqrt(C), d fmap2 = G^T . fmap1^T / sqrt(C)  (smoe_corr_build_bwd,
    tcgen05 kind::tf32).  Widths that are not multiples of 4 (never produced by the model: images
    are padded to /32, dataloader/utils.py:260-267) violate TMA's 16-byte stride rule for NCHW
    operands; they run the same kernel on pitch-padded copies."""
    B, C, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    f1, p1, _ = _pitched_f32(fmap1)
    f2, p2, _ = _pitched_f32(fmap2)
    df1 = torch.empty((B, C, H, p1), dtype=torch.float32, device=f1.device)
    df2 = torch.empty((B, C, H, p2), dtype=torch.float32, device=f2.device)
    with _on_device(gp.device):
        rc = _lib.lib().smoe_corr_build_bwd(f1.data_ptr(), f2.data_ptr(), _lib.DTYPE_F32, B, C, H, W1, W2,
                                            p1, p2, gp.ref, math.sqrt(C), df1.data_ptr(), df2.data_ptr(),
                                            _stream_ptr(gp.device))
    _lib.check(rc, "smoe_corr_build_bwd")
    if p1 != W1:
        df1 = df1[..., :W1].contiguous()
    if p2 != W2:
        df2 = df2[..., :W2].contiguous()
    return df1.to(fmap1.dtype), df2.to(fmap2.dtype)


def _batched_smem_bytes(W2: int, num_levels: int) -> int:
    """Upper bound of the shared memory smoe_corr_lookup_bwd_batched needs per CTA: the gradient
    rows of a 32-pixel tile (all levels, each padded to 4 words, row padded to a bank offset)."""
    off, w = 0, W2
    for _ in range(num_levels):
        off += (w + 3) // 4 * 4
        w //= 2
    return 32 * (off + 32) * 4


def _batchable(pyr: FusedPyramid, coords: torch.Tensor, grad_out: torch.Tensor, radius: int):
    """(coords, batch stride, grad_out) if this lookup's backward can join the batched kernel
    (smoe_corr_lookup_bwd_batched), else None -> per-call accumulate path."""
    if radius != 4 or grad_out.dtype not in _DTYPE_CODE:
        return None
    if _batched_smem_bytes(pyr.W2, pyr.num_levels) > 227 * 1024:
        return None
    grad_out = grad_out.contiguous()
    coords, bstride = _coords_plane(coords, pyr.B, pyr.H, pyr.W1)
    return coords, bstride, grad_out


def _batched_lookup_backward(pyr: FusedPyramid, pending: list, radius: int, gp):
    """Run the deferred lookup backwards of a step; returns a 1-level FusedPyramid holding the folded
    level-0 gradient (adds to `gp`'s level 0 when some lookups took the per-call path)."""
    accumulate = 1
    if gp is None:
        gp = FusedPyramid(pyr.B, pyr.H, pyr.W1, pyr.W2, 1, torch.float32, pyr.device)
        accumulate = 0
    L = _lib.lib()
    by_dtype: dict = {}
    for item in pending:
        by_dtype.setdefault(item[2].dtype, []).append(item)
    for dtype, items in by_dtype.items():
        for i in range(0, len(items), _lib.MAX_BATCHED_ITERS):
            chunk = items[i:i + _lib.MAX_BATCHED_ITERS]
            n = len(chunk)
            cptr = (ctypes.c_void_p * n)(*[c.data_ptr() for c, _, _ in chunk])
            cstr = (ctypes.c_int64 * n)(*[s for _, s, _ in chunk])
            gptr = (ctypes.c_void_p * n)(*[g.data_ptr() for _, _, g in chunk])
            with _on_device(pyr.device):
                rc = L.smoe_corr_lookup_bwd_batched(gp.ref, pyr.num_levels, n, cptr, cstr, gptr,
                                                    _DTYPE_CODE[dtype], radius, accumulate,
                                                    _stream_ptr(pyr.device))
            _lib.check(rc, "smoe_corr_lookup_bwd_batched")
            accumulate = 1
    return gp


def _coords_plane(coords: torch.Tensor, B: int, H: int, W1: int):
    """Validate coords (B,>=1,H,W1) fp32 and return (tensor kept alive, batch stride in elements).
    Only channel 0 (x) is read (core/corr.py:129, :47)."""
    _require_cuda(coords, "coords")
    shp = coords.shape
    if len(shp) != 4 or shp[0] != B or shp[2] != H or shp[3] != W1:
        raise RuntimeError(f"coords must have shape ({B},>=1,{H},{W1}), got {tuple(shp)}")
    if coords.dtype != torch.float32:
        coords = coords.float()
    st = coords.stride()
    if st[3] != 1 or st[2] != W1 or (B > 1 and st[0] < H * W1):      # strided, or a broadcast batch
        coords = coords[:, :1].contiguous()
        st = coords.stride()
    return coords, (st[0] if B > 1 else max(st[0], H * W1))


_lookup_fn = None
_warned_coords_grad = False


def _warn_coords_grad() -> None:
    """The reference's CorrBlock1D / PytorchAlternateCorrBlock1D sample with grid_sample and are
    therefore differentiable w.r.t. coords; the models never use that (coords1 is detached right
    before every lookup, core/SMoEStereo.py:234, core/raft_stereo.py:206) and neither does
    corr_sampler (core/corr.py:29 returns None).  Here coords never receive a gradient."""
    global _warned_coords_grad
    if not _warned_coords_grad:
        _warned_coords_grad = True
        import warnings
        warnings.warn("smoe_stereo_b200: coords.requires_grad is set, but the correlation lookup does not "
                      "propagate gradients to coords (as corr_sampler, core/corr.py:29); detach them like "
                      "core/SMoEStereo.py:234 does", RuntimeWarning, stacklevel=3)


def _check_out(pyr: FusedPyramid, out: torch.Tensor, radius: int, out_dtype) -> None:
    """`out=` of CorrBlock*.__call__: the caller's own (B, L*(2r+1), H, W1) contiguous buffer."""
    want = (pyr.B, pyr.num_levels * (2 * radius + 1), pyr.H, pyr.W1)
    if not isinstance(out, torch.Tensor) or not out.is_cuda or out.device != pyr.device:
        raise RuntimeError("out must be a CUDA tensor on the device of the feature maps")
    if tuple(out.shape) != want or out.dtype is not out_dtype or not out.is_contiguous():
        raise RuntimeError(f"out must be a contiguous {out_dtype} tensor of shape {want}, got "
                           f"{out.dtype} {tuple(out.shape)}")


def _lookup(pyr: FusedPyramid, coords: torch.Tensor, radius: int, out_dtype, use_ring: bool = False,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """One eager lookup.  This is the call an unmodified model makes once per GRU iteration, and its
    kernel runs ~3 us, so the host side is kept short (tools/eager_overhead.py): a contiguous fp32
    CUDA coords tensor of the right shape skips the general validation, the device switch and the
    stream wrapper objects, and the C entry point is reached through the CPython binding
    (_lib.fast_lookup) instead of ctypes."""
    global _lookup_fn
    if _lookup_fn is None:
        _lookup_fn = _lib.fast_lookup()
    # fast path: a contiguous fp32 CUDA (B, 1|2, H, W1) tensor -- one tuple comparison per property
    bstride = pyr._coords_ok.get((coords.shape, coords.stride())) if coords.dtype is torch.float32 and coords.is_cuda else None
    if bstride is None:
        shp = coords.shape
        if (coords.dtype is torch.float32 and coords.is_cuda and len(shp) == 4 and shp[0] == pyr.B
                and shp[2] == pyr.H and shp[3] == pyr.W1 and coords.stride()[1:] == pyr._coords_fast
                and (pyr.B == 1 or coords.stride(0) >= pyr._coords_fast[0])):
            bstride = coords.stride(0) if pyr.B > 1 else shp[1] * pyr._coords_fast[0]
        else:
            coords, bstride = _coords_plane(coords, pyr.B, pyr.H, pyr.W1)
    ring = pyr.out_ring if use_ring else None
    if out is not None:
        _check_out(pyr, out, radius, out_dtype)
    elif ring is not None and ring[0][0].dtype is out_dtype and radius == pyr.ring_radius:
        bufs, pos = ring
        out = bufs[pos]
        ring[1] = pos + 1 if pos + 1 < len(bufs) else 0
    else:
        shape = pyr._out_shape.get(radius)
        if shape is None:
            shape = pyr._out_shape[radius] = (pyr.B, pyr.num_levels * (2 * radius + 1), pyr.H, pyr.W1)
        out = torch.empty(shape, dtype=out_dtype, device=pyr.device)
    idx = pyr.dev_idx
    if _raw_stream is not None and _raw_device() == idx:
        rc = _lookup_fn(pyr.addr, coords.data_ptr(), bstride, radius, out.data_ptr(),
                        _DTYPE_CODE[out_dtype], _raw_stream(idx))
    else:
        with _on_device(pyr.device):
            rc = _lookup_fn(pyr.addr, coords.data_ptr(), bstride, radius, out.data_ptr(),
                            _DTYPE_CODE[out_dtype], _stream_ptr(pyr.device).value or 0)
    if rc:
        _lib.check(rc, "smoe_corr_lookup")
    return out


# ---------------------------------------------------------------------------- autograd plumbing
# Deferred lookup backwards are flushed into the folded level-0 gradient every _FLUSH_EVERY items, so
# at most that many (coords, grad_out) pairs -- B*36*H*W1 values each, 17 MB at cfg2 -- are kept
# alive beyond the point where the reference would have freed them (it consumes each grad_out
# inside its own sampler backward).  32 covers the reference's 16 / 22 / 32-iteration steps in ONE
# launch (a flush re-reads and re-writes the level-0 gradient); longer loops stay bounded.
_FLUSH_EVERY = 32


class _State:
    """Per-block state shared by the build and lookup autograd nodes."""

    def __init__(self):
        self.pyr: FusedPyramid | None = None
        self.radius = 4
        self.build_flags = 0
        self.reset_backward()

    def reset_backward(self):
        """Forget everything a (possibly aborted) backward pass has accumulated."""
        self.grad_pyr: FusedPyramid | None = None      # per-call (accumulating) backward path, all levels
        self.g0: FusedPyramid | None = None            # folded level-0 gradient of flushed batches
        self.pending: list = []                        # deferred lookup backwards (batched path)
        self.seen: set = set()                         # lookup nodes that already ran in this pass

    def flush_pending(self):
        if self.pending:
            self.g0 = _batched_lookup_backward(self.pyr, self.pending, self.radius, self.g0)
            self.pending = []


class _BuildFn(torch.autograd.Function):
    """Volume + pyramid build.  Returns a 1-element token; every lookup takes the token as an
    input, so autograd runs this node's backward ONCE, after all lookup backwards have
    accumulated into the state."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, state):
        pyr = FusedPyramid(fmap1.shape[0], fmap1.shape[2], fmap1.shape[3], fmap2.shape[3],
                           state.num_levels, state.pyr_dtype, fmap1.device)
        _build_into(pyr, fmap1, fmap2, state.build_flags)
        state.pyr = pyr
        ctx.state = state
        ctx.save_for_backward(fmap1, fmap2)
        return torch.zeros(1, dtype=torch.float32, device=fmap1.device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_token):
        fmap1, fmap2 = ctx.saved_tensors
        state = ctx.state
        try:
            state.flush_pending()
            gp, g0 = state.grad_pyr, state.g0
            if gp is None and g0 is None:
                return torch.zeros_like(fmap1), torch.zeros_like(fmap2), None
            if gp is not None:          # lookups that went through the per-call accumulate path
                with _on_device(gp.device):
                    rc = _lib.lib().smoe_corr_fold_pyramid_grad(gp.ref, _stream_ptr(gp.device))
                _lib.check(rc, "smoe_corr_fold_pyramid_grad")
                if g0 is not None:      # a step that mixed both paths (different radii / dtypes)
                    gp.level(0).add_(g0.level(0))
            else:
                gp = g0
            df1, df2 = _build_backward(gp, fmap1, fmap2)
        finally:
            state.reset_backward()
        return df1, df2, None


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, token, coords, state, out_dtype):
        ctx.state = state
        ctx.save_for_backward(coords)
        return _lookup(state.pyr, coords, state.radius, out_dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        state = ctx.state
        pyr = state.pyr
        # A lookup node runs once per backward pass.  Seeing it again means the previous pass never
        # reached the build node (an exception or OOM further down the graph, with retain_graph):
        # what it left behind must not leak into this pass.
        if id(ctx) in state.seen:
            state.reset_backward()
        state.seen.add(id(ctx))
        item = _batchable(pyr, coords, grad_out, state.radius)
        if item is not None:            # defer: the build backward handles the whole step at once
            state.pending.append(item)
            if len(state.pending) >= _FLUSH_EVERY:
                state.flush_pending()
            return torch.zeros(1, dtype=torch.float32, device=pyr.device), None, None, None
        if state.grad_pyr is None:      # zeroed once per backward pass, not once per call
            state.grad_pyr = FusedPyramid(pyr.B, pyr.H, pyr.W1, pyr.W2, pyr.num_levels,
                                          torch.float32, pyr.device, zero=True)
        gp = state.grad_pyr
        grad_out = grad_out.contiguous()
        if grad_out.dtype not in _DTYPE_CODE:
            grad_out = grad_out.float()
        coords, bstride = _coords_plane(coords, pyr.B, pyr.H, pyr.W1)
        with _on_device(pyr.device):
            rc = _lib.lib().smoe_corr_lookup_bwd(gp.ref, coords.data_ptr(), bstride, 1.0,
                                                 state.radius, grad_out.data_ptr(),
                                                 _DTYPE_CODE[grad_out.dtype], _lib.BWD_ACCUMULATE,
                                                 _stream_ptr(pyr.device))
        _lib.check(rc, "smoe_corr_lookup_bwd")
        return torch.zeros(1, dtype=torch.float32, device=pyr.device), None, None, None


# --------------------------------------------------------------------------------- the blocks
class _CorrBlockBase:
    """Shared implementation; subclasses fix the output dtype rule of the reference class."""

    _float_output = True

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        _require_cuda(fmap1, "fmap1")
        _require_cuda(fmap2, "fmap2")
        if fmap1.dim() != 4 or fmap2.dim() != 4 or fmap1.shape[:3] != fmap2.shape[:3]:
            raise RuntimeError("fmap1/fmap2 must be (B,C,H,W1)/(B,C,H,W2) with equal B,C,H, got "
                               f"{tuple(fmap1.shape)} and {tuple(fmap2.shape)}")
        if fmap1.device != fmap2.device:
            raise RuntimeError("fmap1 and fmap2 must be on the same device")
        if fmap1.dtype != fmap2.dtype:
            fmap2 = fmap2.to(fmap1.dtype)
        if fmap1.dtype not in _DTYPE_CODE:      # bf16 / fp64 inputs: compute like the fp32 path
            fmap1, fmap2 = fmap1.float(), fmap2.float()
        self.num_levels = num_levels
        self.radius = radius
        fmap1, fmap2 = _layout_for_build(fmap1, fmap2)
        st = _State()
        pyr_dtype = fmap1.dtype
        if self._float_output and fmap1.dtype == torch.float32 and _default_volume_dtype is not None:
            pyr_dtype = _default_volume_dtype
        st.num_levels, st.radius, st.pyr_dtype = num_levels, radius, pyr_dtype
        # The radius-4 lookup kernels read levels 0 and 2 only and recompute 1 and 3 (bit-exactly),
        # so the build does not write them (a third of the pyramid bytes: cfg1 build 29.6 -> 25.6 us,
        # cfg3 208 -> 190 us).  `corr_pyramid` and the cross-check kernels materialise them on
        # demand (FusedPyramid.materialize).  (With the one-thread-per-pixel lookup mapping the
        # shorter build was offset by slower lookups at cfg1 -- 141.1 vs 138.5 us per step; with
        # the per-level-pair mapping that is now the default it is a net gain, 127.1 vs 129.0 us.)
        st.build_flags = 0
        if radius == 4 and _skip_odd_levels is not False:
            vol_bytes = (fmap1.shape[0] * fmap1.shape[2] * fmap1.shape[3] * fmap2.shape[3] * 15 // 8
                         * (2 if pyr_dtype == torch.float16 else 4))
            if _skip_odd_levels is True or vol_bytes > _SKIP_ODD_MIN_BYTES:
                st.build_flags = _lib.BUILD_SKIP_ODD_LEVELS
        self._state = st
        self._token = None
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            self._token = _BuildFn.apply(fmap1, fmap2, st)
        else:
            st.pyr = FusedPyramid(fmap1.shape[0], fmap1.shape[2], fmap1.shape[3], fmap2.shape[3],
                                  num_levels, pyr_dtype, fmap1.device)
            _build_into(st.pyr, fmap1, fmap2, st.build_flags)
            if _default_output_ring:
                self.reuse_outputs(_default_output_ring)

    @property
    def corr_pyramid(self):
        """Reference-shaped views (B,H,W1,1,W2_i) of the fused buffer (core/corr.py:41)."""
        p = self._state.pyr
        return [p.level(i).unsqueeze(3) for i in range(p.num_levels)]

    def __call__(self, coords, out=None):
        """corr_fn(coords) of the reference (core/corr.py:127-146).  `out=` (an extension, optional):
        write the result into the caller's own contiguous (B, L*(2r+1), H, W1) buffer instead of a
        fresh tensor -- what a serving loop that stages results for a device-to-host copy wants
        (no allocator traffic, no record_stream bookkeeping); not available while autograd records."""
        p = self._state.pyr
        out_dtype = torch.float32 if self._float_output else p.dtype
        if torch.is_grad_enabled():
            if coords.requires_grad:
                _warn_coords_grad()
            if self._token is not None:
                if out is not None:
                    raise RuntimeError("out= cannot be used while autograd is recording the lookup")
                return _LookupFn.apply(self._token, coords, self._state, out_dtype)
            return _lookup(p, coords, self.radius, out_dtype, False, out)
        return _lookup(p, coords, self.radius, out_dtype, True, out)

    def reuse_outputs(self, depth: int = 2):
        """Opt-in: serve `__call__` from a ring of `depth` preallocated output buffers instead of a
        fresh torch.empty per call (~3 us of host time per lookup, as long as the kernel itself at
        B=1).  The result of a call is then only valid until `depth` further calls have been made --
        fine for the GRU loop, which consumes corr inside the same iteration
        (core/raft_stereo.py:207-214), not for callers that keep every lookup.  depth=0 restores
        per-call allocation.  Ignored while autograd is recording."""
        p = self._state.pyr
        if depth <= 0:
            p.out_ring = None
            return self
        dt = torch.float32 if self._float_output else p.dtype
        p.out_ring = [[torch.empty((p.B, p.num_levels * (2 * self.radius + 1), p.H, p.W1), dtype=dt,
                                   device=p.device) for _ in range(depth)], 0]
        p.ring_radius = self.radius
        return self

    def lookup_update(self, coords1, delta_flow=None, coords0=None):
        """Opt-in fused GRU-loop step.  Equivalent to the reference loop body
        around the lookup (core/SMoEStereo.py:234-248):

            delta_flow[:, 1] = 0; coords1 = coords1 + delta_flow      # skipped if delta_flow is None
            corr = corr_fn(coords1); flow = coords1 - coords0          # flow None if coords0 is None

        returns (corr, coords1_new, flow).  Without autograd this is ONE kernel launch instead of
        four; when gradients are needed (training) it composes the differentiable pieces."""
        p = self._state.pyr
        out_dtype = torch.float32 if self._float_output else p.dtype
        needs_grad = torch.is_grad_enabled() and (
            self._token is not None or (delta_flow is not None and delta_flow.requires_grad)
            or coords1.requires_grad)
        if needs_grad:
            if delta_flow is not None:
                mask = delta_flow.new_tensor([1.0, 0.0]).view(1, 2, 1, 1)
                coords1 = coords1 + delta_flow * mask
            corr = self(coords1.detach())
            flow = coords1 - coords0 if coords0 is not None else None
            return corr, coords1, flow
        for name, t in (("coords1", coords1), ("delta_flow", delta_flow), ("coords0", coords0)):
            if t is None:
                continue
            _require_cuda(t, name)
            if tuple(t.shape) != (p.B, 2, p.H, p.W1) or t.dtype != torch.float32:
                raise RuntimeError(f"{name} must be a float32 tensor of shape ({p.B},2,{p.H},{p.W1})")
        coords1 = coords1.contiguous()
        delta_flow = delta_flow.contiguous() if delta_flow is not None else None
        coords0 = coords0.contiguous() if coords0 is not None else None
        c1_new = torch.empty_like(coords1)
        flow = torch.empty_like(coords1) if coords0 is not None else None
        out = torch.empty((p.B, p.num_levels * (2 * self.radius + 1), p.H, p.W1), dtype=out_dtype,
                          device=p.device)
        with _on_device(p.device):
            rc = _lib.lib().smoe_corr_lookup_update(
                p.ref, coords1.data_ptr(), delta_flow.data_ptr() if delta_flow is not None else None,
                coords0.data_ptr() if coords0 is not None else None, c1_new.data_ptr(),
                flow.data_ptr() if flow is not None else None, self.radius, out.data_ptr(),
                _DTYPE_CODE[out_dtype], _stream_ptr(p.device))
        _lib.check(rc, "smoe_corr_lookup_update")
        return out, c1_new, flow

    def lookup_conv(self, coords, weight, bias=None, relu=True, out_dtype=None):
        """Opt-in fused lookup + 1x1 convolution (+bias, +ReLU):

            cor = F.relu(F.conv2d(corr_fn(coords), weight, bias))     # core/update.py:70,77 (convc1)

        in ONE kernel; the L*(2r+1)-channel lookup result is never written.  `weight` is
        (Cout, L*(2r+1)) or Conv2d-shaped (Cout, L*(2r+1), 1, 1); products run in tf32 on the
        tensor cores with fp32 accumulation.  `out_dtype` defaults to the dtype an autocast-free
        F.conv2d on the lookup result would produce.  Shapes the fused kernel does not cover
        (radius != 4, Cout not a multiple of 16 or > 256) and calls that need gradients run the
        lookup kernel followed by F.conv2d."""
        p = self._state.pyr
        lookup_dtype = torch.float32 if self._float_output else p.dtype
        cin = p.num_levels * (2 * self.radius + 1)
        if weight.dim() == 4:
            if weight.shape[2:] != (1, 1):
                raise RuntimeError("lookup_conv fuses a 1x1 convolution; weight must be (Cout,Cin,1,1)")
            weight2d = weight.reshape(weight.shape[0], weight.shape[1])
        else:
            weight2d = weight
        if weight2d.dim() != 2 or weight2d.shape[1] != cin:
            raise RuntimeError(f"weight must have {cin} input channels, got {tuple(weight.shape)}")
        cout = weight2d.shape[0]
        if bias is not None and tuple(bias.shape) != (cout,):
            raise RuntimeError(f"bias must have shape ({cout},)")
        if out_dtype is None:
            out_dtype = lookup_dtype
        needs_grad = torch.is_grad_enabled() and (
            self._token is not None or weight.requires_grad or (bias is not None and bias.requires_grad))
        fusable = (self.radius == 4 and cout % 16 == 0 and 16 <= cout <= 256
                   and out_dtype in _DTYPE_CODE and not needs_grad)
        if not fusable:
            corr = self(coords)
            w4 = weight2d.reshape(cout, cin, 1, 1).to(corr.dtype)
            y = F.conv2d(corr, w4, bias.to(corr.dtype) if bias is not None else None)
            return (F.relu(y) if relu else y).to(out_dtype)
        _require_cuda(weight2d, "weight")
        if p.c.levels_missing and p.dtype == torch.float32:
            p.materialize()      # fp32 volumes beyond L2 run the lane-group version, which reads all levels
        w = weight2d.detach().to(torch.float32).contiguous()
        bvec = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        coords, bstride = _coords_plane(coords, p.B, p.H, p.W1)
        out = torch.empty((p.B, cout, p.H, p.W1), dtype=out_dtype, device=p.device)
        with _on_device(p.device):
            rc = _lib.lib().smoe_corr_lookup_conv1x1(
                p.ref, coords.data_ptr(), bstride, 1.0, self.radius, _DTYPE_CODE[lookup_dtype],
                w.data_ptr(), bvec.data_ptr() if bvec is not None else None, cout, 1 if relu else 0,
                out.data_ptr(), _DTYPE_CODE[out_dtype], _stream_ptr(p.device))
        _lib.check(rc, "smoe_corr_lookup_conv1x1")
        return out

    @staticmethod
    def corr(fmap1, fmap2):
        """(B,D,H,W1),(B,D,H,W2) -> (B,H,W1,1,W2), same dtype (core/corr.py:148-156, :53-61).
        Differentiable: backward = smoe_corr_build_bwd (see _build_backward)."""
        return _CorrVolumeFn.apply(fmap1, fmap2)


class _CorrVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2):
        _require_cuda(fmap1, "fmap1")
        _require_cuda(fmap2, "fmap2")
        f1, f2 = fmap1, fmap2
        if f1.dtype not in _DTYPE_CODE:
            f1, f2 = f1.float(), f2.float()
        f1, f2 = _layout_for_build(f1, f2)
        pyr = FusedPyramid(f1.shape[0], f1.shape[2], f1.shape[3], f2.shape[3], 1, f1.dtype, f1.device)
        _build_into(pyr, f1, f2)
        ctx.save_for_backward(f1, f2)
        return pyr.level(0).unsqueeze(3).contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        f1, f2 = ctx.saved_tensors
        gp = FusedPyramid(f1.shape[0], f1.shape[2], f1.shape[3], f2.shape[3], 1, torch.float32, f1.device)
        gp.level(0).copy_(g.squeeze(3))               # (B,H,W1,W2) -> pitched rows of the C ABI
        return _build_backward(gp, f1, f2)


class CorrBlock1D(_CorrBlockBase):
    """"reg" implementation (core/corr.py:110-156): output is always fp32 (:146)."""
    _float_output = True


class CorrBlockFast1D(_CorrBlockBase):
    """"reg_cuda" implementation (core/corr.py:31-61): output keeps the volume dtype (:49-51),
    i.e. fp16 feature maps give an fp16 pyramid and fp16 lookups with the reference's
    half-precision rounding order."""
    _float_output = False


class PytorchAlternateCorrBlock1D(_CorrBlockBase):
    """"alt" implementation (core/corr.py:64-107).  The reference trades memory for recompute
    (no volume; 36 grid_samples of fmap2 per call).  With 180 GB of HBM per B200 the volume of
    even a 2000x3000 pair is ~2 GB, so this class is an API-compatible alias of the volume
    path: same constructor, same (B,36,H,W) fp32 output (agrees with "reg" to 3.5e-5,
    SURVEY.md 2).  As in every real call site, coords[:,1] is the pixel's own row."""
    _float_output = True

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        super().__init__(fmap1, fmap2, num_levels=num_levels, radius=radius)
        self.fmap1 = fmap1
        self.fmap2 = fmap2


class CorrSampler(torch.autograd.Function):
    """Same contract as core/corr.py:17-29 on top of corr_sampler.forward/backward."""

    @staticmethod
    def forward(ctx, volume, coords, radius):
        ctx.save_for_backward(volume, coords)
        ctx.radius = radius
        corr, = corr_sampler.forward(volume, coords, radius)
        return corr

    @staticmethod
    def backward(ctx, grad_output):
        volume, coords = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_volume, = corr_sampler.backward(volume, coords, grad_output, ctx.radius)
        return grad_volume, None, None


class AlternateCorrBlock:
    """"alt_cuda" is dead code in the reference (core/corr.py:159-161 raises first thing)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        raise NotImplementedError
