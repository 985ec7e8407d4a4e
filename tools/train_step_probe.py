"""BASELINE config 2 (SceneFlow training shape): B=8, 320x736 crop -> 80x184, 22 GRU iterations.
Times the correlation part of one training step, kernel class by kernel class, against the
algorithmic-byte roofline of SURVEY.md 8d."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200 import _lib
from smoe_stereo_b200.corr import _build_into, _lookup, _build_backward, _stream_ptr

dev = torch.device("cuda", 0)
B, C, H, W, L, r, iters = 8, 256, 80, 184, 4, 4, 22
if len(sys.argv) > 1:
    B = int(sys.argv[1])
px = B * H * W
g = torch.Generator(device="cpu").manual_seed(0)
f1 = torch.randn(B, C, H, W, generator=g).to(dev)
f2 = torch.randn(B, C, H, W, generator=g).to(dev)
xs = torch.arange(W).float().view(1, 1, W).expand(B, H, W)
coords = []
for t in range(iters):
    c = torch.zeros(B, 2, H, W)
    c[:, 0] = xs - torch.rand(B, H, 1, generator=g) * 48 - torch.randn(B, H, W, generator=g) * (0.5 * t)
    coords.append(c.to(dev))
gouts = [torch.randn(B, 36, H, W, generator=g).to(dev) for _ in range(iters)]
VOL = torch.float16 if os.environ.get("SMOE_CORR_VOLUME_DTYPE") else torch.float32
pyr = sb.FusedPyramid(B, H, W, W, L, VOL, dev)
gp = sb.FusedPyramid(B, H, W, W, L, torch.float32, dev, zero=True)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6549.4
widths = pyr.widths

def bwd(t):
    rc = _lib.lib().smoe_corr_lookup_bwd(gp.ref, coords[t].data_ptr(), coords[t].stride(0), 1.0, r,
                                         gouts[t].data_ptr(), _lib.DTYPE_F32, _lib.BWD_ACCUMULATE, _stream_ptr(dev))
    _lib.check(rc, "bwd")

def fold():
    _lib.check(_lib.lib().smoe_corr_fold_pyramid_grad(gp.ref, _stream_ptr(dev)), "fold")

side = torch.cuda.Stream(dev)
def graph_of(fn):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            keep = fn()
    return gr, keep

def timeit(gr, reps=30):
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps

def lookups():
    o = None
    for t in range(iters):
        o = _lookup(pyr, coords[t], r, torch.float32)
    return o

rows = []
def report(name, us, bytes_):
    rows.append((name, us, bytes_))
    print(f"{name:34s} {us:9.1f} us   {bytes_/1e6:9.1f} MB algorithmic   {bytes_/us/1e3:7.0f} GB/s  ({bytes_/us/1e3/peak*100:5.1f}% of {peak:.0f})")

s_pyr = px * sum(widths) * 4
report("build (volume + pyramid)", timeit(graph_of(lambda: _build_into(pyr, f1, f2))[0]), 2 * B * C * H * W * 4 + s_pyr)
report(f"{iters} x lookup forward", timeit(graph_of(lookups)[0]), iters * px * 308)
report("zero gradient pyramid", timeit(graph_of(lambda: gp.buffer.zero_())[0]), gp.buffer.numel() * 4)
report(f"{iters} x lookup backward (accumulate)", timeit(graph_of(lambda: [bwd(t) for t in range(iters)])[0]), iters * px * 468)
report("fold pyramid gradient", timeit(graph_of(fold)[0]), px * (2 * widths[0] + sum(widths[1:])) * 4)
report("build backward (2 GEMMs)", timeit(graph_of(lambda: _build_backward(gp, f1, f2))[0]), px * widths[0] * 4 + 4 * B * C * H * W * 4)
from smoe_stereo_b200.corr import _batchable, _batched_lookup_backward
items = [_batchable(pyr, coords[t], gouts[t], r) for t in range(iters)]
g0holder = {}
def batched():
    g0holder["g"] = _batched_lookup_backward(pyr, items, r, None)
    return g0holder["g"]
t_batched = timeit(graph_of(batched)[0])
print(f"{'-> batched backward (all iters + fold)':34s} {t_batched:9.1f} us   replaces zero + {iters} x backward + fold = "
      f"{rows[2][1] + rows[3][1] + rows[4][1]:.1f} us")
tot = sum(r_[1] for r_ in rows)
print(f"{'sum with batched backward':34s} {rows[0][1] + rows[1][1] + t_batched + rows[5][1]:9.1f} us")
print(f"{'sum':34s} {tot:9.1f} us")
# reference-style dense backward for comparison: zeros_like + scatter per level per iteration
vol = pyr.level(0).contiguous()
c0 = coords[0][:, :1].contiguous(); g0 = gouts[0][:, :9].contiguous()
report("corr_sampler.backward, level 0 dense", timeit(graph_of(lambda: sb.corr_sampler.backward(vol, c0, g0, r))[0]), px * widths[0] * 4)
