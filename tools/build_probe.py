"""Event-timed corr_build_kernel alone, for 1..4 pyramid levels, on rotating (L2-cold) inputs.

    python tools/build_probe.py [B C H W]

With a probe build of the library (SMOE_NVCC_DEFS=-DSMOE_BUILD_PROBES python
smoe-stereo_b200/build.py --force) SMOE_CORR_BUILD_DEBUG=1|2|4 times the kernel without its
operand loads / result stores / MMAs; that is how profiles/r01/README.md splits the build time.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _build_into

dev = torch.device("cuda", 0)
shapes = [tuple(int(x) for x in sys.argv[1:5])] if len(sys.argv) > 4 else [(1, 256, 135, 240), (8, 256, 96, 312)]
for (B, C, H, W) in shapes:
    nset = max(2, int(600e6 // (2 * B * C * H * W * 4 + B * H * W * W * 4 * 1.4)) + 1)
    nset = min(nset, 6)
    fs = [(torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)) for _ in range(nset)]
    for nl in (1, 2, 3, 4):
        pyrs = [sb.FusedPyramid(B, H, W, W, nl, torch.float32, dev) for _ in range(nset)]
        for i in range(nset):
            _build_into(pyrs[i], *fs[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 40
        e0.record()
        for i in range(n):
            _build_into(pyrs[i % nset], *fs[i % nset])
        e1.record()
        torch.cuda.synchronize()
        print("probe=%s B=%d W=%d levels=%d build %.1f us" % (
            os.environ.get("SMOE_CORR_BUILD_DEBUG", "0"), B, W, nl, e0.elapsed_time(e1) * 1e3 / n))
