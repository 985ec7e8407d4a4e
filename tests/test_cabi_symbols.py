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
