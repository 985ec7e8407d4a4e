"""Compile libsmoecorr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python smoe-stereo_b200/build.py [--force] [--verbose]

No torch headers, no JIT cache: the .so lands next to the sources so that it travels to the
GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "lookup.cu", "lookup_conv.cu", "build.cu", "build_bwd.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "px_gather.cuh", os.path.join("..", "..", "include", "smoe_corr.h")]
LIB = os.path.join(HERE, "libsmoecorr.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    *(["-Xptxas", "-v"] if os.environ.get("SMOE_PTXAS_V") else []),
    *os.environ.get("SMOE_NVCC_DEFS", "").split(),      # experiment switches (-DNAME=VALUE ...)
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
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
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static",
           "-Xlinker", "--exclude-libs,ALL"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
