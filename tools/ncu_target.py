"""ncu target for profiles/r02: ONE kernel class of ONE BASELINE config, launched eagerly through the
PRODUCT library with the same arguments bench.py's measure_config uses.

    python tools/ncu_target.py <cfg1|cfg2|cfg3|cfg4> <build|lookup|lookup_bwd|build_bwd> [launches]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import smoe_stereo_b200 as sb  # noqa: E402
from bench import WORKLOADS, synth_inputs_gpu  # noqa: E402
from smoe_stereo_b200.corr import (_batchable, _batched_lookup_backward, _build_backward, _build_into,  # noqa: E402
                                   _lookup)

cfg, what = sys.argv[1], sys.argv[2]
half = "--fp16" in sys.argv            # reg_cuda: fp16 feature maps, fp16 pyramid, fp16 lookups
argv = [a for a in sys.argv if a != "--fp16"]
n = int(argv[3]) if len(argv) > 3 else 4
wl = WORKLOADS[cfg]
dev = torch.device("cuda", 0)
B, C, H, W, L, r, iters = (wl[k] for k in ("B", "C", "H", "W", "L", "r", "iters"))
f1, f2, coords = synth_inputs_gpu(torch, wl, B, H, seed=7000, dev=dev)
vdt = torch.float16 if half else torch.float32
if half:
    f1, f2 = f1.half(), f2.half()
pyr = sb.FusedPyramid(B, H, W, W, L, vdt, dev)
flags = sb._lib.BUILD_SKIP_ODD_LEVELS
_build_into(pyr, f1, f2, flags)
torch.cuda.synchronize()
keep = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if what == "build":
    a.record()
    for _ in range(n):
        _build_into(pyr, f1, f2, flags)
elif what == "lookup":
    a.record()
    for t in range(n):
        keep.append(_lookup(pyr, coords[t % iters], r, vdt))
elif what in ("lookup_bwd", "build_bwd"):
    g = torch.Generator(device=dev).manual_seed(9000)
    gouts = [torch.randn((B, L * (2 * r + 1), H, W), device=dev, generator=g) for _ in range(iters)]
    items = [_batchable(pyr, coords[t], gouts[t], r) for t in range(iters)]
    g0 = _batched_lookup_backward(pyr, items, r, None)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        if what == "lookup_bwd":
            g0 = _batched_lookup_backward(pyr, items, r, None)
        else:
            keep.append(_build_backward(g0, f1, f2))
else:
    raise SystemExit(f"unknown kernel class {what}")
b.record()
torch.cuda.synchronize()
print(f"{cfg} {what}: {a.elapsed_time(b) * 1e3 / n:.1f} us per launch (eager, {n} launches)")
