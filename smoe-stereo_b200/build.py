"""Compile libsmoecorr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python smoe-stereo_b200/build.py [--force] [--verbose] [--xcheck]

No torch headers, no JIT cache: the .so lands next to the sources so that it travels to the
GPU box with the repo snapshot.

--xcheck additionally builds tests/_build/libsmoecorr_xcheck.so from the same sources with
-DSMOE_CROSSCHECK: the product kernels PLUS the superseded ones (csrc/xcheck_*.cuh), the
smoe_corr_debug_* selectors and the SMOE_CORR_* experiment switches.  Only tests/ and tools/ load
it (tests/gpu_util.py::xcheck_lib); the package never does.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "lookup.cu", "lookup_conv.cu", "build.cu", "build_bwd.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "px_gather.cuh", "xcheck_lookup.cuh", "xcheck_lookup_conv.cuh",
           os.path.join("..", "..", "include", "smoe_corr.h")]
LIB = os.path.join(HERE, "libsmoecorr.so")
XCHECK_DIR = os.path.join(ROOT, "tests", "_build")
XCHECK_LIB = os.path.join(XCHECK_DIR, "libsmoecorr_xcheck.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--compress-mode=size",        # the line tables of ~150 kernel instantiations compress 15x
    "-Xcompiler", "-fPIC",
    *(["-Xptxas", "-v"] if os.environ.get("SMOE_PTXAS_V") else []),
    *os.environ.get("SMOE_NVCC_DEFS", "").split(),      # experiment switches (-DNAME=VALUE ...)
]


def _needs_build(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _needs_build(LIB)


def _compile(lib: str, objdir: str, extra: list[str], verbose: bool) -> str:
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stdout.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return lib


def build_fastcall(force: bool = False, verbose: bool = False) -> str:
    """csrc/fastcall.c -> _smoe_fastcall.<abi>.so: CPython binding of the per-iteration lookup call."""
    import sysconfig
    src = os.path.join(CSRC, "fastcall.c")
    out = os.path.join(HERE, "_smoe_fastcall" + sysconfig.get_config_var("EXT_SUFFIX"))
    if force or not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-Wall",
               "-I" + sysconfig.get_paths()["include"], src, "-o", out]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """The product library (+ the optional fast Python binding of smoe_corr_lookup)."""
    if force or _needs_build(LIB):
        _compile(LIB, CSRC, [], verbose)
    build_fastcall(force, verbose)
    return LIB


def build_xcheck(force: bool = False, verbose: bool = False) -> str:
    """The cross-check library (tests / tools only)."""
    if force or _needs_build(XCHECK_LIB):
        _compile(XCHECK_LIB, os.path.join(XCHECK_DIR, "obj"), ["-DSMOE_CROSSCHECK"], verbose)
    return XCHECK_LIB


if __name__ == "__main__":
    force, verbose = "--force" in sys.argv, "--verbose" in sys.argv
    print(build(force=force, verbose=verbose))
    if "--xcheck" in sys.argv:
        print(build_xcheck(force=force, verbose=verbose))
