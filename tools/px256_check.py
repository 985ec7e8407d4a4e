import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import smoe_stereo_b200 as sb
from smoe_stereo_b200 import _lib
from smoe_stereo_b200.corr import _lookup
from gpu_util import dev, make_coords
L = _lib.lib()
ok = True
for levels in (4, 3, 2, 1):
    for (B, C, H, W1, W2) in [(2, 32, 9, 77, 77), (1, 16, 5, 40, 312), (1, 16, 3, 33, 22), (2, 16, 4, 50, 9), (1, 16, 4, 96, 240)]:
        g = torch.Generator(device="cpu").manual_seed(W2 + levels)
        f1 = torch.randn(B, C, H, W1, generator=g).to(dev()); f2 = torch.randn(B, C, H, W2, generator=g).to(dev())
        blk = sb.CorrBlock1D(f1, f2, num_levels=levels)
        coords = torch.from_numpy(make_coords(B, H, W1, W2, 3, seed=W1)).to(dev())
        coords[0, 0, 0, 0, :6] = torch.tensor([float("nan"), float("inf"), -float("inf"), 3e9, -0.0, W2 - 1.0])
        coords[1, 0, 0, 1, :4] = torch.tensor([-4.0, -4.5, W2 + 3.999, W2 + 4.0])
        for t in range(3):
            L.smoe_corr_debug_set_lookup_kernel(1); a = blk(coords[t]).clone()
            L.smoe_corr_debug_set_lookup_kernel(4); b = blk(coords[t]).clone()
            L.smoe_corr_debug_set_lookup_kernel(-1)
            if not torch.equal(a, b):
                ok = False
                print("MISMATCH", levels, (B, C, H, W1, W2), t, (a - b).abs().max().item(), torch.isnan(b).any().item())
print("px256 bitwise:", "OK" if ok else "FAILED")
# timing, 32 distinct outputs per graph
side = torch.cuda.Stream(dev())
for name, (B, H, W) in (("cfg1", (1, 135, 240)), ("cfg2", (8, 80, 184)), ("cfg3", (8, 96, 312))):
    pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev()); pyr.buffer.normal_()
    g = torch.Generator(device="cpu").manual_seed(0)
    base = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    coords = [(base - torch.rand(B, 1, H, 1, generator=g) * 60 - torch.randn(B, 1, H, W, generator=g) * (1 + 0.5 * t)).to(dev()) for t in range(8)]
    res = {}
    for which, kn in ((1, "px"), (4, "px256")):
        L.smoe_corr_debug_set_lookup_kernel(which)
        def lk():
            return [_lookup(pyr, coords[t % 8], 4, torch.float32) for t in range(32)]
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
        del gr, keep
    L.smoe_corr_debug_set_lookup_kernel(-1)
    print(name, res)
    del pyr
