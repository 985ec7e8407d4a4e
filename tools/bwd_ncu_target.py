"""ncu target: the batched lookup backward at BASELINE config 2, one variant of the cross-check
library, launched eagerly a few times.   python tools/bwd_ncu_target.py <bwd_kernel id> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import smoe_stereo_b200 as sb  # noqa: E402
from smoe_stereo_b200.corr import _batchable, _batched_lookup_backward  # noqa: E402
from gpu_util import xcheck_lib  # noqa: E402

which = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
B, H, W, iters, r = 8, 80, 184, 22, 4
g = torch.Generator(device="cpu").manual_seed(0)
xs = torch.arange(W).float().view(1, 1, W).expand(B, H, W)
coords, gouts = [], []
for t in range(iters):
    c = torch.zeros(B, 2, H, W)
    c[:, 0] = xs - torch.rand(B, H, 1, generator=g) * 48 - torch.randn(B, H, W, generator=g) * (0.5 * t)
    coords.append(c.to(dev))
    gouts.append(torch.randn(B, 36, H, W, generator=g).to(dev))
pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
items = [_batchable(pyr, coords[t], gouts[t], r) for t in range(iters)]
with xcheck_lib() as L:
    L.smoe_corr_debug_set_tuning(b"bwd_kernel", which)
    for _ in range(reps):
        g0 = _batched_lookup_backward(pyr, items, r, None)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g0 = _batched_lookup_backward(pyr, items, r, None)
    b.record()
    torch.cuda.synchronize()
    print(f"bwd_kernel={which}: {a.elapsed_time(b) * 1e3:.1f} us (eager, one launch)")
    L.smoe_corr_debug_set_tuning(b"bwd_kernel", 1)
