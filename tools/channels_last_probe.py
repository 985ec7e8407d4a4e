"""Build kernel: NCHW (MN-major operands) vs channels-last (K-major operands),
fp32 and fp16 feature maps.  Event-timed over rotating input sets larger than L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _build_into
dev = torch.device("cuda", 0)
SHAPES = {"cfg1": (1, 256, 135, 240), "cfg2": (8, 256, 80, 184), "cfg3": (8, 256, 96, 312),
          "cfg4": (1, 256, 500, 750)}


def timeit(fns, reps):
    for f in fns: f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fns[i % len(fns)]()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


for name in (sys.argv[1:] or list(SHAPES)):
    B, C, H, W = SHAPES[name]
    px = B * H * W
    for dt in (torch.float32, torch.float16):
        es = 2 if dt == torch.float16 else 4
        nset = max(2, int(300e6 // (2 * px * C * es)) + 1)
        sets = [(torch.randn(B, C, H, W, device=dev).to(dt), torch.randn(B, C, H, W, device=dev).to(dt))
                for _ in range(min(nset, 6))]
        pyr = sb.FusedPyramid(B, H, W, W, 4, dt, dev)
        bytes_ = 2 * px * C * es + px * sum(pyr.widths) * es
        res = {}
        for layout in ("nchw", "channels_last"):
            ss = sets if layout == "nchw" else [tuple(t.contiguous(memory_format=torch.channels_last) for t in s) for s in sets]
            fns = [(lambda s=s: _build_into(pyr, s[0], s[1])) for s in ss]
            res[layout] = timeit(fns, 40 if px < 200000 else 10)
            del ss, fns
        print(f"{name} {str(dt)[6:]}: nchw {res['nchw']:.1f} us ({bytes_/res['nchw']/1e3:.0f} GB/s), "
              f"channels_last {res['channels_last']:.1f} us ({bytes_/res['channels_last']/1e3:.0f} GB/s algorithmic)")
        del sets, pyr
