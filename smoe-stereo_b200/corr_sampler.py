"""``corr_sampler`` -- same module surface as the reference's CUDA extension
(sampler/sampler.cpp:48-51, sampler/sampler_kernel.cu), served by libsmoecorr.so:

    forward(volume, coords, radius)              -> [corr]          (B, 2r+1, H, W1)
    backward(volume, coords, corr_grad, radius)  -> [volume_grad]   same shape/dtype as volume

volume: (B,H,W1,W2) contiguous CUDA tensor, fp32 or fp16; coords: (B,1,H,W1) fp32 (already
scaled to the level, core/corr.py:49).  Results come back as a one-element list because the
callers unpack them with ``corr, = ...`` (core/corr.py:22,28).  Argument errors raise
RuntimeError with the TORCH_CHECK messages of sampler/sampler.cpp:20-22.

Deliberate differences from the reference kernel:
  * launches on the current torch stream, not the legacy default stream (sampler_kernel.cu:127);
  * does not read the non-existent coords[:,1] channel (sampler_kernel.cu:40);
  * backward writes the dense gradient in one pass instead of zeros_like + scatter (:148-163).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16}


def _check_input(x, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _row_pitched(v: torch.Tensor) -> bool:
    """(B,H,W1,W2) whose pixel rows are evenly `pitch` >= W2 elements apart -- what the
    `corr_pyramid[i]` views of this package's own blocks are (one level of a fused pyramid row)."""
    if v.dim() != 4 or v.stride(3) != 1:
        return False
    B, H, W1, W2 = v.shape
    pitch = v.stride(2)
    return pitch >= W2 and (H == 1 or v.stride(1) == W1 * pitch) and (B == 1 or v.stride(0) == H * W1 * pitch)


def _check_volume(v):
    """CHECK_INPUT(volume) of sampler/sampler.cpp:20-22, relaxed to admit row-pitched views."""
    if not isinstance(v, torch.Tensor) or not v.is_cuda:
        raise RuntimeError("volume must be a CUDA tensor")
    if not (v.is_contiguous() or _row_pitched(v)):
        raise RuntimeError("volume must be contiguous")


def _planar(volume: torch.Tensor) -> _lib.Pyramid:
    B, H, W1, W2 = volume.shape
    p = _lib.Pyramid()
    p.level_ptr[0] = volume.data_ptr()
    p.width[0] = W2
    p.pitch[0] = W2 if volume.is_contiguous() else volume.stride(2)
    p.num_levels = 1
    p.dtype = _DTYPE_CODE[volume.dtype]
    p.B, p.H, p.W1 = B, H, W1
    return p


def _check_shapes(volume, coords):
    if volume.dim() != 4:
        raise RuntimeError(f"volume must be (B,H,W1,W2), got {tuple(volume.shape)}")
    if volume.dtype not in _DTYPE_CODE:
        raise RuntimeError(f"volume dtype {volume.dtype} not supported (float32 / float16)")
    B, H, W1, _ = volume.shape
    if coords.dim() != 4 or coords.shape[0] != B or coords.shape[2] != H or coords.shape[3] != W1:
        raise RuntimeError(f"coords must be ({B},C,{H},{W1}), got {tuple(coords.shape)}")
    if coords.dtype != torch.float32:
        raise RuntimeError("coords must be float32")
    if volume.device != coords.device:
        raise RuntimeError("volume and coords must be on the same device")


def _launch_ctx(device):
    from .corr import _on_device, _stream_ptr
    return _on_device(device), _stream_ptr(device)


def forward(volume, coords, radius):
    """sampler_forward (sampler/sampler.cpp:24-32)."""
    _check_volume(volume)
    _check_input(coords, "coords")
    _check_shapes(volume, coords)
    B, H, W1, _ = volume.shape
    radius = int(radius)
    out = torch.empty((B, 2 * radius + 1, H, W1), dtype=volume.dtype, device=volume.device)
    p = _planar(volume)
    guard, stream = _launch_ctx(volume.device)
    with guard:
        rc = _lib.lib().smoe_corr_lookup(ctypes.byref(p), coords.data_ptr(),
                                         max(coords.stride(0), H * W1), 1.0, radius,
                                         out.data_ptr(), _DTYPE_CODE[volume.dtype], stream)
    _lib.check(rc, "smoe_corr_lookup")
    return [out]


def backward(volume, coords, corr_grad, radius):
    """sampler_backward (sampler/sampler.cpp:34-45): fresh dense gradient w.r.t. volume."""
    _check_volume(volume)
    _check_input(coords, "coords")
    _check_input(corr_grad, "corr_grad")
    _check_shapes(volume, coords)
    B, H, W1, _ = volume.shape
    radius = int(radius)
    if tuple(corr_grad.shape) != (B, 2 * radius + 1, H, W1):
        raise RuntimeError(f"corr_grad must be ({B},{2 * radius + 1},{H},{W1}), got "
                           f"{tuple(corr_grad.shape)}")
    if corr_grad.dtype != volume.dtype:
        raise RuntimeError("corr_grad must have the volume's dtype")
    grad = torch.empty(volume.shape, dtype=volume.dtype, device=volume.device)
    p = _planar(grad)
    guard, stream = _launch_ctx(volume.device)
    with guard:
        rc = _lib.lib().smoe_corr_lookup_bwd(ctypes.byref(p), coords.data_ptr(),
                                             max(coords.stride(0), H * W1), 1.0, radius,
                                             corr_grad.data_ptr(), _DTYPE_CODE[corr_grad.dtype],
                                             _lib.BWD_OVERWRITE, stream)
    _lib.check(rc, "smoe_corr_lookup_bwd")
    return [grad]
