"""Round-2 kernel variants, side by side, in ONE process on the cross-check library
(tests/_build/libsmoecorr_xcheck.so; variants are selected with smoe_corr_debug_set_tuning).

    python tools/variants_probe.py [bwd] [lookup] [conv]

bwd    : batched lookup backward at BASELINE config 2 (B=8, 80x184, 22 iterations):
         threads per (pixel, level) NS in {1,2,3} x shared-memory row length mod 32.
lookup : streaming forward lookup (fp32 volume beyond L2) at cfg2 / cfg3 / cfg4: L2 eviction hints x
         pixels per CTA, 32 lookups per CUDA graph with 32 DISTINCT outputs (as a consumer has them).
conv   : fused lookup + convc1 at cfg3: thread-per-pixel with 256-bit loads vs the lane-group version.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import smoe_stereo_b200 as sb  # noqa: E402
from smoe_stereo_b200.corr import _lookup, _batchable, _batched_lookup_backward  # noqa: E402
from gpu_util import xcheck_lib  # noqa: E402

dev = torch.device("cuda", 0)
side = torch.cuda.Stream(dev)
PEAK = 6549.4


def graph_of(fn):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            keep = fn()
    return gr, keep


def timeit(gr, reps=20):
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def probe_bwd(L):
    B, H, W, iters, r = 8, 80, 184, 22, 4
    px = B * H * W
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
    floor_mb = iters * px * 148 / 1e6 + px * W * 4 / 1e6
    ref = None
    for which, name in ((1, "lvl   [pixel][column], register ring"), (2, "cp    [column][pixel], register ring"),
                        (3, "async [column][pixel], cp.async ring  "),
                        (4, "tma   [column][pixel], TMA producer   "),
                        (21, "cp + L2 prefetch 2 iterations ahead  "), (22, "cp + L2 prefetch 4 iterations ahead  "),
                        (23, "cp + L2 prefetch 8 iterations ahead  "), (24, "PROBE cp+pf4: gradient loads only    "),
                        (11, "PROBE cp: gradient loads only        "), (12, "PROBE cp: scatter + fold only        ")):
        L.smoe_corr_debug_set_tuning(b"bwd_kernel", which)
        hold = {}

        def run():
            hold["g"] = _batched_lookup_backward(pyr, items, r, None)
            return hold["g"]
        us = timeit(graph_of(run)[0])
        out = hold["g"].level(0).clone()
        if ref is None:
            ref = out
        if which in (11, 12, 24):
            out = ref
        print(f"bwd cfg2 {name}: {us:7.1f} us  ({floor_mb / us:6.2f} TB/s on the {floor_mb:.0f} MB DRAM floor, "
              f"{floor_mb / us / PEAK * 1e3 * 100:4.1f}% of peak)  bit-identical to lvl: {torch.equal(out, ref)}", flush=True)
    L.smoe_corr_debug_set_tuning(b"bwd_staged", 1)
    hold = {}

    def run_staged():
        hold["g"] = _batched_lookup_backward(pyr, items, r, None)
        return hold["g"]
    us = timeit(graph_of(run_staged)[0])
    print(f"bwd cfg2 staged (first version): {us:7.1f} us  bit-identical: {torch.equal(hold['g'].level(0), ref)}")
    L.smoe_corr_debug_set_tuning(b"bwd_staged", 0)
    L.smoe_corr_debug_set_tuning(b"bwd_kernel", 1)


SHAPES = {"cfg2": (8, 80, 184), "cfg3": (8, 96, 312), "cfg4": (1, 500, 750)}


def probe_lookup(L):
    for name, (B, H, W) in SHAPES.items():
        pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
        pyr.buffer.normal_()
        pyr.mark_missing(0b1010)
        g = torch.Generator(device="cpu").manual_seed(0)
        base = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
        coords = [(base - torch.rand(B, 1, H, 1, generator=g) * 60 - torch.randn(B, 1, H, W, generator=g) * (1 + 0.5 * t)).to(dev)
                  for t in range(8)]
        alg = B * H * W * 308
        ref = None
        for ld256, hint, tpb in ((1, 0, 32), (1, 0, 64), (1, 1, 32), (1, 1, 64), (1, 4, 32), (1, 4, 64), (1, 5, 32),
                                 (1, 5, 64), (1, 5, 128), (1, 7, 32), (1, 7, 64), (1, 3, 64)):
            L.smoe_corr_debug_set_tuning(b"px_ld256", ld256)
            L.smoe_corr_debug_set_tuning(b"px_hint", hint)
            L.smoe_corr_debug_set_tuning(b"px_tpb", tpb)

            def lk():
                return [_lookup(pyr, coords[t % 8], 4, torch.float32) for t in range(32)]
            gr, keep = graph_of(lk)
            us = timeit(gr) / 32
            if ref is None:
                ref = keep[5].clone()
            print(f"lookup {name} ld256={ld256} hint={hint} tpb={tpb:3d}: {us:6.2f} us  {alg / us / 1e3:6.0f} GB/s algorithmic "
                  f"= {alg / us / 1e3 / PEAK * 100:4.1f}% of peak   bit-identical: {torch.equal(keep[5], ref)}", flush=True)
            del gr, keep
        for k, v in ((b"px_ld256", -1), (b"px_hint", -1), (b"px_tpb", 32)):
            L.smoe_corr_debug_set_tuning(k, v)
        del pyr
        torch.cuda.empty_cache()


def probe_conv(L):
    B, H, W = SHAPES["cfg3"]
    pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
    pyr.buffer.normal_()
    g = torch.Generator(device="cpu").manual_seed(0)
    base = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    coords = [(base - torch.rand(B, 1, H, 1, generator=g) * 60 - torch.randn(B, 1, H, W, generator=g)).to(dev) for t in range(8)]
    w = torch.randn(64, 36, device=dev) * 0.1
    bias = torch.randn(64, device=dev)
    blk = sb.CorrBlock1D.__new__(sb.CorrBlock1D)
    st = sb.corr._State()
    st.pyr, st.radius = pyr, 4
    blk._state, blk._token, blk.radius, blk.num_levels = st, None, 4, 4
    ref = None
    for lanes, ld256 in ((1, 0), (0, 0), (0, -1)):
        L.smoe_corr_debug_set_tuning(b"conv_lanes", lanes)
        L.smoe_corr_debug_set_tuning(b"px_ld256", ld256)

        def run():
            return [blk.lookup_conv(coords[t % 8], w, bias) for t in range(16)]
        gr, keep = graph_of(run)
        us = timeit(gr) / 16
        if ref is None:
            ref = keep[3].clone()
        err = (keep[3] - ref).abs().max().item()
        print(f"lookup+conv cfg3 lanes={lanes} ld256={ld256}: {us:6.2f} us per iteration  max|diff vs lanes| {err:.2e}", flush=True)
    L.smoe_corr_debug_set_tuning(b"conv_lanes", 0)
    L.smoe_corr_debug_set_tuning(b"px_ld256", -1)


if __name__ == "__main__":
    what = sys.argv[1:] or ["bwd", "lookup", "conv"]
    with xcheck_lib() as L:
        if "bwd" in what:
            probe_bwd(L)
        if "lookup" in what:
            probe_lookup(L)
        if "conv" in what:
            probe_conv(L)
