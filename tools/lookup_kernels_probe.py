"""Forward lookup kernels side by side (SMOE debug selector): four-window lane groups (vec),
level-pair lane groups (pair), thread per pixel (px).  32 lookups per CUDA graph, rotating coords."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200 import _lib
from smoe_stereo_b200.corr import _lookup
dev = torch.device("cuda", 0)
SHAPES = {"cfg1": (1, 135, 240), "cfg2": (8, 80, 184), "cfg3": (8, 96, 312), "cfg4": (1, 500, 750),
          "cfg1x8": (8, 135, 240)}
L = _lib.lib()
side = torch.cuda.Stream(dev)
for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4"]):
    B, H, W = SHAPES[name]
    for dt in (torch.float32, torch.float16):
        pyr = sb.FusedPyramid(B, H, W, W, 4, dt, dev)
        pyr.buffer.normal_()
        g = torch.Generator(device="cpu").manual_seed(0)
        base = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
        coords = [(base - torch.rand(B, 1, H, 1, generator=g) * 60 - torch.randn(B, 1, H, W, generator=g) * (1 + 0.5 * t)).to(dev)
                  for t in range(8)]
        res = {}
        for which, kn in ((3, "vec"), (2, "pair"), (1, "px")):
            L.smoe_corr_debug_set_lookup_kernel(which)
            keep_all = os.environ.get("PROBE_KEEP") is not None   # 32 distinct output buffers (as bench.py):
            def lk():                                              # writes stream to DRAM instead of staying in L2
                outs = []
                o = None
                for t in range(32):
                    o = _lookup(pyr, coords[t % 8], 4, dt)
                    if keep_all:
                        outs.append(o)
                return outs if keep_all else o
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                lk(); torch.cuda.synchronize()
                with torch.cuda.graph(gr, stream=side):
                    keep = lk()
            for _ in range(3): gr.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): gr.replay()
            b.record(); torch.cuda.synchronize()
            res[kn] = a.elapsed_time(b) * 1e3 / 20 / 32
        L.smoe_corr_debug_set_lookup_kernel(-1)
        es = 2 if dt == torch.float16 else 4
        alg = B * H * W * (4 + 4 * 10 * es + 36 * es)
        print(f"{name} {str(dt)[6:]:8s} volume {pyr.buffer.numel()*es/1e6:7.1f} MB: " +
              ", ".join(f"{k} {v:.2f} us ({alg/v/1e3:.0f} GB/s)" for k, v in res.items()))
        del pyr
