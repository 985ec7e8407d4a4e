#!/usr/bin/env python
"""Build the REFERENCE's own CUDA sampler (sampler/sampler.cpp + sampler/sampler_kernel.cu) into
oracle/_ref/corr_sampler_ref.so -- TEST INFRASTRUCTURE, never loaded by the product.

    python oracle/build_ref_sampler.py [--reference /root/reference] [--force]

The sources are compiled where they lie: they are copied to a scratch directory under /tmp
(never into this repository), the two `AT_DISPATCH_FLOATING_TYPES_AND_HALF(volume.type(), ...)`
call sites (sampler/sampler_kernel.cu:126,157) get the 2-token compatibility patch
`.type()` -> `.scalar_type()` that torch >= 2.x requires, and
torch.utils.cpp_extension builds them for sm_100a.  Only the resulting shared object lands in
oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun like our own .so).  The kernels are
otherwise the unmodified reference: tests/test_gpu_ref_sampler.py uses them to pin
corr_sampler.forward/backward and CorrBlockFast1D (fp32 and fp16) to the reference's arithmetic.

Needs /root/reference (build container only); on the GPU box the prebuilt .so is used.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
MODULE = "corr_sampler_ref"
OUT = os.path.join(OUT_DIR, MODULE + ".so")


def build(reference: str = "/root/reference", force: bool = False, verbose: bool = False) -> str | None:
    """Returns the path of the built module, or None when the reference checkout is absent."""
    src_dir = os.path.join(reference, "sampler")
    srcs = [os.path.join(src_dir, n) for n in ("sampler.cpp", "sampler_kernel.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    scratch = tempfile.mkdtemp(prefix="ref_sampler_")
    try:
        patched = []
        for s in srcs:
            text = open(s).read()
            if s.endswith(".cu"):
                n = text.count("AT_DISPATCH_FLOATING_TYPES_AND_HALF(volume.type(),")
                if n != 2:
                    raise RuntimeError(f"expected 2 dispatch sites in {s}, found {n}")
                text = text.replace("AT_DISPATCH_FLOATING_TYPES_AND_HALF(volume.type(),",
                                    "AT_DISPATCH_FLOATING_TYPES_AND_HALF(volume.scalar_type(),")
            dst = os.path.join(scratch, os.path.basename(s))
            open(dst, "w").write(text)
            patched.append(dst)
        bdir = os.path.join(scratch, "build")
        os.makedirs(bdir)
        load(name=MODULE, sources=patched, build_directory=bdir, verbose=verbose, is_python_module=False,
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-O3"])
        os.makedirs(OUT_DIR, exist_ok=True)
        shutil.copyfile(os.path.join(bdir, MODULE + ".so"), OUT)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return OUT


def load_module():
    """Import oracle/_ref/corr_sampler_ref.so (tests only).  None if it has not been built."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(MODULE, OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.reference, a.force, a.verbose))
