"""Fused lookup+convc1 (smoe_corr_lookup_conv1x1) against lookup followed by torch's conv2d+relu,
BASELINE config 1 plane (B=1, 135x240), 32 iterations per CUDA graph replay.

    python tools/lookup_conv_probe.py [B H W]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import smoe_stereo_b200 as sb

B, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 135, 240)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
IT = 32


def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(IT):
                fn()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * IT)


for name, cls, fdtype in (("CorrBlock1D fp32", sb.CorrBlock1D, torch.float32),
                          ("CorrBlockFast1D fp16", sb.CorrBlockFast1D, torch.float16)):
    f1 = torch.randn(B, 256, H, W, device=dev).to(fdtype)
    f2 = torch.randn(B, 256, H, W, device=dev).to(fdtype)
    blk = cls(f1, f2, num_levels=4, radius=4)
    coords = torch.zeros(B, 2, H, W, device=dev)
    coords[:, 0] = torch.arange(W, device=dev, dtype=torch.float32)[None, None, :] - torch.rand(B, H, W, device=dev) * 60
    w = torch.randn(64, 36, 1, 1, device=dev) * 0.2
    b = torch.randn(64, device=dev)
    w16, b16 = w.half(), b.half()
    with torch.no_grad():
        t_lookup = graph_time(lambda: blk(coords))
        t_unfused32 = graph_time(lambda: F.relu(F.conv2d(blk(coords).float(), w, b)))
        t_unfused16 = graph_time(lambda: F.relu(F.conv2d(blk(coords).half(), w16, b16)))     # what autocast runs
        t_fused32 = graph_time(lambda: blk.lookup_conv(coords, w, b, out_dtype=torch.float32))
        t_fused16 = graph_time(lambda: blk.lookup_conv(coords, w, b, out_dtype=torch.float16))
    print(f"{name} B={B} {H}x{W}: lookup alone {t_lookup:.2f} us | lookup+conv+relu fp32 {t_unfused32:.2f} us, "
          f"autocast-fp16 {t_unfused16:.2f} us | fused fp32-out {t_fused32:.2f} us, fp16-out {t_fused16:.2f} us")
