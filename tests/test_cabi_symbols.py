"""CPU: the C-ABI library loads and exports every symbol include/smoe_corr.h declares; the pure
host-side entry points behave (no GPU compute is called here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from smoe_stereo_b200 import _lib
    return _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "smoe_corr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smoe_corr_\w+)\s*\(", src)))


def test_header_symbols_exported(built):
    L = ctypes.CDLL(built.LIB_PATH)
    names = header_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/smoe_corr.h but not exported"
    assert sorted(built.SYMBOLS) == names, "python binding list out of sync with the header"
    assert L.smoe_corr_abi_version() == built.ABI_VERSION


def test_struct_layout_matches_header(built):
    # void*[4] + int32[4] + int64[4] + 6*int32
    assert ctypes.sizeof(built.Pyramid) == 32 + 16 + 32 + 24
    assert built.Pyramid.pitch.offset == 48 and built.Pyramid.num_levels.offset == 80


def test_pyramid_layout_host_arithmetic(built):
    w, off, pitch = built.pyramid_layout(240, 4, built.DTYPE_F32)
    assert w == [240, 120, 60, 30] and off == [0, 240, 360, 420] and pitch == 480
    w, off, pitch = built.pyramid_layout(37, 4, built.DTYPE_F32)     # floor-halving, 16-B aligned levels
    assert w == [37, 18, 9, 4] and off == [0, 40, 60, 72] and pitch % 32 == 0
    w, off, pitch = built.pyramid_layout(750, 4, built.DTYPE_F16)
    assert w == [750, 375, 187, 93] and all(o % 8 == 0 for o in off) and pitch % 64 == 0
    with pytest.raises(RuntimeError, match="num_levels"):
        built.pyramid_layout(64, 5, built.DTYPE_F32)
    with pytest.raises(RuntimeError, match="dtype"):
        built.pyramid_layout(64, 4, 7)


def test_argument_errors_are_reported_without_a_gpu(built):
    L = built.lib()
    p = built.Pyramid()
    p.num_levels = 0
    rc = L.smoe_corr_lookup(ctypes.byref(p), None, 0, 1.0, 4, None, built.DTYPE_F32, None)
    assert rc == built.ERR_INVALID_ARG and "num_levels" in built.last_error()
    p.num_levels = 1
    p.dtype = built.DTYPE_F32
    p.B, p.H, p.W1 = 1, 2, 3
    p.width[0], p.pitch[0] = 8, 4
    rc = L.smoe_corr_lookup(ctypes.byref(p), None, 0, 1.0, 4, None, built.DTYPE_F32, None)
    assert rc == built.ERR_INVALID_ARG and "pitch" in built.last_error()
    assert L.smoe_corr_build_workspace_bytes(1, 256, 135, 240, 240, built.DTYPE_F32) == 0
    assert L.smoe_corr_build_workspace_bytes(1, 256, 500, 750, 750, built.DTYPE_F32) == \
        2 * ((256 * 500 * 752 * 4 + 255) // 256 * 256)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "smoe-stereo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU", ""), f"{f} mentions the oracle"


def _fake_pyramid(built, W2=240, levels=4, missing=0):
    """A descriptor with plausible (never dereferenced) pointers: argument validation only."""
    w, off, pitch = built.pyramid_layout(W2, levels, built.DTYPE_F32)
    p = built.Pyramid()
    for i in range(levels):
        p.level_ptr[i] = 0x10000000 + off[i] * 4
        p.width[i] = w[i]
        p.pitch[i] = pitch
    p.num_levels, p.dtype = levels, built.DTYPE_F32
    p.B, p.H, p.W1 = 1, 2, 8
    p.levels_missing = missing
    return p


def test_unbuilt_levels_and_layout_flags_are_validated_without_a_gpu(built):
    """smoe_pyramid_t.levels_missing (SMOE_BUILD_SKIP_ODD_LEVELS): entry points that would read a
    level that was not built refuse before anything is launched; so does a channels-last build
    whose channel count cannot be a TMA stride."""
    L = built.lib()
    fake = ctypes.c_void_p(0x20000000)
    # radius 3 goes through the generic kernel, which reads every level
    p = _fake_pyramid(built, missing=0b1010)
    rc = L.smoe_corr_lookup(ctypes.byref(p), fake, 16, 1.0, 3, fake, built.DTYPE_F32, None)
    assert rc == built.ERR_INVALID_ARG and "were not built" in built.last_error()
    # an even level can never be marked missing
    p = _fake_pyramid(built, missing=0b0100)
    rc = L.smoe_corr_lookup(ctypes.byref(p), fake, 16, 1.0, 4, fake, built.DTYPE_F32, None)
    assert rc == built.ERR_INVALID_ARG and "even level" in built.last_error()
    # smoe_corr_fill_levels cannot recompute level 0
    p = _fake_pyramid(built, missing=0b0001)
    rc = L.smoe_corr_fill_levels(ctypes.byref(p), None)
    assert rc == built.ERR_INVALID_ARG and "level 0" in built.last_error()
    # channels-last needs C*sizeof(T) % 16 == 0
    p = _fake_pyramid(built, W2=40)
    p.H, p.W1 = 3, 40
    rc = L.smoe_corr_build(fake, fake, built.DTYPE_F32, 1, 6, 3, 40, 40, ctypes.byref(p), 2.449,
                           None, 0, built.BUILD_CHANNELS_LAST, None)
    assert rc == built.ERR_UNSUPPORTED and "channels-last" in built.last_error()
