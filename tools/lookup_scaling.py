"""Lookup launch time vs problem size (fixed cost vs per-pixel cost), pyramid L2-resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _lookup
dev = torch.device("cuda", 0)
W = 240
side = torch.cuda.Stream(dev)
for H in (4, 20, 40, 80, 135, 200, 270):
    pyr = sb.FusedPyramid(1, H, W, W, 4, torch.float32, dev)
    pyr.buffer.normal_()
    coords = torch.zeros(32, 1, 2, H, W, device=dev)
    coords[:, :, 0] = torch.arange(W, device=dev).float() - torch.rand(32, 1, H, W, device=dev) * 50
    def fn():
        o = None
        for t in range(32):
            o = _lookup(pyr, coords[t], 4, torch.float32)
        return o
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            keep = fn()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 100 / 32
    px = H * W
    print(f"H={H:4d} px={px:6d} CTAs={(px + 31) // 32:5d} ({(px + 31) // 32 / 148:5.2f}/SM): {us:5.2f} us per lookup")
