"""ctypes binding of libsmoecorr.so (the C ABI declared in include/smoe_corr.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, a
RuntimeError is raised (mirroring TORCH_CHECK in the reference's sampler/sampler.cpp:20-22).
"""
from __future__ import annotations

import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsmoecorr.so")

MAX_LEVELS = 4
ABI_VERSION = 2
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
DTYPE_F32, DTYPE_F16 = 0, 1
BUILD_TF32_TRUNCATE = 1
BUILD_CHANNELS_LAST = 2
BUILD_SKIP_ODD_LEVELS = 4
BWD_OVERWRITE, BWD_ACCUMULATE = 0, 1
OUT_MULTIMEM, OUT_SCALAR = 1, 2
MAX_BATCHED_ITERS = 48

# every symbol include/smoe_corr.h declares
SYMBOLS = (
    "smoe_corr_abi_version", "smoe_corr_last_error", "smoe_corr_device_supported",
    "smoe_corr_pyramid_layout", "smoe_corr_build_workspace_bytes", "smoe_corr_build",
    "smoe_corr_lookup", "smoe_corr_lookup_bwd", "smoe_corr_fold_pyramid_grad",
    "smoe_corr_fill_levels", "smoe_corr_build_bwd", "smoe_corr_lookup_update",
    "smoe_corr_lookup_bwd_batched", "smoe_corr_lookup_conv1x1", "smoe_corr_lookup_to",
)
# only in the cross-check library (tests/_build/libsmoecorr_xcheck.so, csrc/xcheck_api.h)
XCHECK_SYMBOLS = ("smoe_corr_debug_set_timeline", "smoe_corr_debug_set_lookup_kernel",
                  "smoe_corr_debug_set_tuning")


class Pyramid(ctypes.Structure):
    """smoe_pyramid_t"""
    _fields_ = [
        ("level_ptr", ctypes.c_void_p * MAX_LEVELS),
        ("width", ctypes.c_int32 * MAX_LEVELS),
        ("pitch", ctypes.c_int64 * MAX_LEVELS),
        ("num_levels", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("B", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("W1", ctypes.c_int32),
        ("levels_missing", ctypes.c_int32),
    ]


_lock = threading.Lock()
_lib = None


def _declare(L: ctypes.CDLL) -> None:
    i32, i64, f32, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
    pp = ctypes.POINTER(Pyramid)
    L.smoe_corr_abi_version.restype = i32
    L.smoe_corr_last_error.restype = ctypes.c_char_p
    L.smoe_corr_device_supported.restype = i32
    L.smoe_corr_pyramid_layout.argtypes = [i32, i32, i32, ctypes.POINTER(ctypes.c_int32),
                                           ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.smoe_corr_build_workspace_bytes.argtypes = [i32] * 6
    L.smoe_corr_build_workspace_bytes.restype = i64
    L.smoe_corr_build.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, pp, f32, vp, i64,
                                  ctypes.c_uint, vp]
    L.smoe_corr_lookup.argtypes = [pp, vp, i64, f32, i32, vp, i32, vp]
    L.smoe_corr_lookup_bwd.argtypes = [pp, vp, i64, f32, i32, vp, i32, i32, vp]
    L.smoe_corr_lookup_update.argtypes = [pp, vp, vp, vp, vp, vp, i32, vp, i32, vp]
    L.smoe_corr_lookup_update.restype = i32
    L.smoe_corr_lookup_to.argtypes = [pp, vp, i64, f32, i32, vp, i64, i64, i32, ctypes.c_uint, vp]
    L.smoe_corr_lookup_to.restype = i32
    L.smoe_corr_lookup_conv1x1.argtypes = [pp, vp, i64, f32, i32, i32, vp, vp, i32, i32, vp, i32, vp]
    L.smoe_corr_lookup_conv1x1.restype = i32
    L.smoe_corr_lookup_bwd_batched.argtypes = [pp, i32, i32, vp, vp, vp, i32, i32, i32, vp]
    L.smoe_corr_lookup_bwd_batched.restype = i32
    L.smoe_corr_fold_pyramid_grad.argtypes = [pp, vp]
    L.smoe_corr_fill_levels.argtypes = [pp, vp]
    L.smoe_corr_fill_levels.restype = i32
    L.smoe_corr_build_bwd.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i64, i64, pp, f32, vp, vp, vp]
    L.smoe_corr_build_bwd.restype = i32
    if hasattr(L, "smoe_corr_debug_set_timeline"):        # cross-check library only
        L.smoe_corr_debug_set_timeline.argtypes = [vp]
        L.smoe_corr_debug_set_timeline.restype = None
        L.smoe_corr_debug_set_lookup_kernel.argtypes = [i32]
        L.smoe_corr_debug_set_lookup_kernel.restype = None
        L.smoe_corr_debug_set_tuning.argtypes = [ctypes.c_char_p, i32]
        L.smoe_corr_debug_set_tuning.restype = i32
    for name in ("smoe_corr_pyramid_layout", "smoe_corr_build", "smoe_corr_lookup",
                 "smoe_corr_lookup_bwd", "smoe_corr_fold_pyramid_grad"):
        getattr(L, name).restype = i32


def lib() -> ctypes.CDLL:
    """Load libsmoecorr.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build the CUDA extension first "
                        "(python -c 'import __graft_entry__ as g; g.build()'). "
                        "smoe_stereo_b200 has no CPU fallback.")
                L = ctypes.CDLL(LIB_PATH)
                missing = [s for s in SYMBOLS if not hasattr(L, s)]
                if missing:
                    raise RuntimeError(f"{LIB_PATH} does not export {missing}")
                _declare(L)
                if L.smoe_corr_abi_version() != ABI_VERSION:
                    raise RuntimeError("libsmoecorr.so ABI version mismatch; rebuild it")
                _lib = L
    return _lib


_fast = None


def fast_lookup():
    """`lookup(pyr_addr, coords_ptr, bstride, radius, out_ptr, out_dtype, stream) -> rc`: the CPython
    binding of smoe_corr_lookup (csrc/fastcall.c, ~0.2 us per call instead of ~2 us through
    ctypes), bound to the loaded library; falls back to an equivalent ctypes closure when the
    extension module has not been built.  Same C entry point either way."""
    global _fast
    if _fast is None:
        L = lib()
        try:
            import importlib.util
            import sysconfig
            path = os.path.join(HERE, "_smoe_fastcall" + sysconfig.get_config_var("EXT_SUFFIX"))
            spec = importlib.util.spec_from_file_location("_smoe_fastcall", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.bind(ctypes.cast(L.smoe_corr_lookup, ctypes.c_void_p).value)
            _fast = mod.lookup
        except Exception:           # noqa: BLE001  (not built: ctypes does the same call)
            fn = L.smoe_corr_lookup

            def _fast(pyr, coords, bstride, radius, out, out_dtype, stream):
                return fn(ctypes.cast(pyr, ctypes.POINTER(Pyramid)), coords, bstride, 1.0, radius, out,
                          out_dtype, stream)
    return _fast


def last_error() -> str:
    return lib().smoe_corr_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != OK:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def pyramid_layout(W2: int, num_levels: int, dtype_code: int):
    """-> (widths, level_offsets, pitch) in elements; smoe_corr_pyramid_layout."""
    widths = (ctypes.c_int32 * MAX_LEVELS)()
    offs = (ctypes.c_int64 * MAX_LEVELS)()
    pitch = ctypes.c_int64(0)
    check(lib().smoe_corr_pyramid_layout(W2, num_levels, dtype_code, widths, offs,
                                         ctypes.byref(pitch)), "smoe_corr_pyramid_layout")
    return list(widths)[:num_levels], list(offs)[:num_levels], int(pitch.value)
