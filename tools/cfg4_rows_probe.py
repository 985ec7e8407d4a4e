"""Streaming lookup at the Middlebury width (W=750) for different row counts: is the lower DRAM
efficiency of the full 2.2 GB volume (config 4 on one GPU) a size effect (TLB reach, 2 MB pages)?
32 lookups with 32 distinct outputs per CUDA graph, as bench.py's measure_config."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import smoe_stereo_b200 as sb  # noqa: E402
from smoe_stereo_b200.corr import _lookup  # noqa: E402

dev = torch.device("cuda", 0)
side = torch.cuda.Stream(dev)
PEAK = 6549.4
W = int(sys.argv[1]) if len(sys.argv) > 1 else 750
for H in ((500, 250, 125, 63, 32) if len(sys.argv) < 3 else (int(sys.argv[2]),)):
    pyr = sb.FusedPyramid(1, H, W, W, 4, torch.float32, dev)
    pyr.buffer.normal_()
    pyr.mark_missing(0b1010)
    g = torch.Generator(device="cpu").manual_seed(0)
    base = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    coords = [(base - torch.rand(1, 1, H, 1, generator=g) * 60 - torch.randn(1, 1, H, W, generator=g) * (1 + 0.5 * t)).to(dev)
              for t in range(8)]

    def lk():
        return [_lookup(pyr, coords[t % 8], 4, torch.float32) for t in range(32)]
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        lk()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            keep = lk()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20 / 32
    alg = H * W * 308
    print(f"W={W} H={H:3d}: pyramid {pyr.buffer.numel() * 4 / 1e6:7.0f} MB  {us:6.2f} us/lookup  {us * 1e3 / (H * W):6.3f} ns/px  "
          f"{alg / us / 1e3:6.0f} GB/s algorithmic = {alg / us / 1e3 / PEAK * 100:4.1f}% of peak", flush=True)
    del gr, keep, pyr, coords
    torch.cuda.empty_cache()
