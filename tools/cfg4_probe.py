import torch, sys
sys.path.insert(0, ".")
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _build_into
dev = torch.device("cuda", 0)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) * 1e3 / n
for (B, C, H, W) in [(1, 256, 500, 750), (1, 256, 500, 752), (1, 256, 63, 750), (1, 256, 63, 752)]:
    f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
    pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
    tb = t(lambda: _build_into(pyr, f1, f2))
    px = B * H * W
    bytes_ = 2 * B * C * H * W * 4 + px * sum(pyr.widths) * 4
    print(f"{(B,C,H,W)}: build {tb:.0f} us, {bytes_/tb/1e3:.0f} GB/s algorithmic ({bytes_/tb/1e3/65.49:.0f}% of HBM peak)")
    del f1, f2, pyr
