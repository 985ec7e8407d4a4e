"""Per-CTA timeline of corr_build_kernel (debug aid): prints when each role reached each
milestone, in microseconds since the CTA started (SM clock assumed 1.965 GHz)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _build_into

B, C, H, W = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 256, 135, 240)))
dev = torch.device("cuda", 0)
f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    _build_into(pyr, f1, f2)
tl = torch.zeros(148 * 3 * 16, dtype=torch.int64, device=dev)
flush.zero_(); torch.cuda.synchronize()
sb._lib.lib().smoe_corr_debug_set_timeline(tl.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _build_into(pyr, f1, f2); e1.record()
sb._lib.lib().smoe_corr_debug_set_timeline(None)
torch.cuda.synchronize()
print("kernel (cold L2) %.1f us" % (e0.elapsed_time(e1) * 1e3))
t = tl.cpu().numpy().reshape(148, 3, 16) / 1965.0
np.set_printoptions(precision=1, suppress=True, linewidth=200)
for name, role, slots in (("producer: start, item0 issued, item1 issued", 0, [0, 1, 2, 3]),
                          ("mma: item0 first/last stage landed, committed; item1 ...", 1, [0, 1, 2, 3, 4, 5]),
                          ("epilogue: item0 acc ready, item0 done, item1 acc ready, item1 done, all stores done", 2, [0, 1, 2, 3, 15])):
    a = t[:, role, slots]
    print(name)
    print("  median", np.median(a, axis=0), "\n  max   ", a.max(axis=0), "\n  cta0  ", a[0], "\n  cta147", a[147])
if os.environ.get("TL_RAW"):
    for cta in (0, 73, 147):
        print("cta", cta)
        print("  producer", t[cta, 0])
        print("  mma     ", t[cta, 1])
        print("  epilogue", t[cta, 2])
if os.environ.get("TL_PROBE"):
    # probe builds (-DSMOE_BUILD_PROBES): warp 2's per-chunk cycle sums, slots 9..13 of role 2
    raw = tl.cpu().numpy().reshape(148, 3, 16)[:, 2, 9:14].astype(np.float64)
    raw = raw[raw[:, 4] > 0]
    per = raw[:, :4] / raw[:, 4:5]
    print("epilogue cycles per chunk (median over CTAs): store-drain wait %.0f | tmem ld %.0f | scale+pool+stage+fence %.0f | store issue %.0f | chunks/CTA %.0f"
          % (*np.median(per, axis=0), np.median(raw[:, 4])))
