#!/usr/bin/env python
"""bench.py -- correlation hot path benchmark (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1]
                    [--configs cfg2,cfg3,cfg4,cfg5 | none]

metric  : "corr build + 32-iter lookup pairs/sec at 540x960" (BASELINE.json configs[0]:
          B=1, C=256, H=135, W=240, 4 levels, radius 4, 32 lookup iterations).
step    : one pass of the hot path over one stereo pair: volume+pyramid build (1 tcgen05 kernel)
          + 32 fused pyramid lookups (32 kernels), i.e. 33 launches of our own kernels.
value   : pairs/s with inputs resident in HBM, the step replayed as a CUDA graph, timed with CUDA
          events on the launching stream; steps rotate over several independent input sets whose
          combined footprint exceeds the 126 MB L2 (stated in config.l2 / run.l2).  value_stats repeats the
          timed region so that one line is self-validating.
e2e     : pairs/s through the public API (CorrBlock1D(...) + __call__) with HOST buffers: pinned
          host feature maps / coords copied H2D and every lookup result copied D2H inside the
          timed region; e2e.host_copy_ceiling = the same bytes moved with no compute at all.
roofline: the lookup kernel of the headline workload (dominant: 32 of 33 launches) -- algorithmic
          bytes per launch (px * 308 B) / average launch duration from CUDA events.
          roofline_hbm = the same kernel at BASELINE config 3, whose 600 MB pyramid streams from
          HBM (the headline pyramid is L2-resident).
configs : BASELINE configs 2 (training step: build + 22 lookups + batched lookup backward + build
          backward), 3, 4 and 5 (end-to-end SMoE-Stereo frames/s with the corr path swapped in),
          each with its own rooflines and clock window.  Under torchrun the B=8 configs are
          batch-sharded and config 4 is row-sharded over the ranks (strong scaling, no data-path
          collective; the optional output all-gather is timed separately).
cpu_baseline / --impl reference: the UNMODIFIED reference core/corr.py CorrBlock1D on PyTorch CPU
          (baseline/_ref, installed by tools/install_ref.sh) on the host cores of the same box;
          the oracle's C/OpenMP port of it is reported next to it (cpu_baseline_port).

Under torchrun (N>1) the headline is weak scaling: every rank runs the same per-GPU workload
(batch-partitioned pairs, no collective in the path); the time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: B, C, H, W, levels, radius, iters, description
    "cfg1": dict(B=1, C=256, H=135, W=240, L=4, r=4, iters=32,
                 desc="540x960 pair at 1/4 res: B=1,C=256,H=135,W=240, 4 levels, radius 4, 32 lookups"),
    "cfg2": dict(B=8, C=256, H=80, W=184, L=4, r=4, iters=22, train=True,
                 desc="SceneFlow training crop 320x736, B=8: C=256,H=80,W=184, build + 22 lookups + lookup backward + build backward"),
    "cfg3": dict(B=8, C=256, H=96, W=312, L=4, r=4, iters=32,
                 desc="KITTI 384x1248, B=8: C=256,H=96,W=312, 4 levels, radius 4, 32 lookups"),
    "cfg4": dict(B=1, C=256, H=500, W=750, L=4, r=4, iters=32,
                 desc="Middlebury full resolution ~2000x3000: B=1,C=256,H=500,W=750, 4 levels, radius 4, 32 lookups"),
    "cfg1x8": dict(B=8, C=256, H=135, W=240, L=4, r=4, iters=32,
                   desc="8 x 540x960 pairs per step: B=8,C=256,H=135,W=240, 4 levels, radius 4, 32 lookups"),
}
METRIC = "corr build+32-iter lookup pairs/sec at 540x960"
UNIT = "pairs/s"
L2_BYTES = 126e6


def level_widths(w, L):
    out = []
    for _ in range(L):
        out.append(w)
        w //= 2
    return out


def algorithmic_bytes(wl, vol_bytes=4, in_bytes=4):
    """build = read each fmap once + write each used level once;
    lookup per iteration = px * (4 + L*(2r+2)*s_vol + L*(2r+1)*4)."""
    px = wl["B"] * wl["H"] * wl["W"]
    build = 2 * wl["B"] * wl["C"] * wl["H"] * wl["W"] * in_bytes + px * sum(level_widths(wl["W"], wl["L"])) * vol_bytes
    lookup = px * (4 + wl["L"] * (2 * wl["r"] + 2) * vol_bytes + wl["L"] * (2 * wl["r"] + 1) * 4)
    flops = 2 * wl["C"] * px * wl["W"]
    return build, lookup, flops


def training_bytes(wl):
    """Algorithmic bytes of the backward half of a training step as THIS design moves them
    (DESIGN.md 3.3/3.4): the batched lookup backward reads every grad_out and coordinate once and
    writes the folded level-0 gradient once; the build backward reads that gradient and the two
    feature maps and writes the two feature-map gradients.  (SURVEY 8d's per-iteration figure,
    px*468 B, describes the reference-style read-modify-write of a gradient pyramid in HBM.)"""
    px = wl["B"] * wl["H"] * wl["W"]
    ncorr = wl["L"] * (2 * wl["r"] + 1)
    lookup_bwd = wl["iters"] * px * (ncorr * 4 + 4) + px * wl["W"] * 4
    build_bwd = px * wl["W"] * 4 + 4 * wl["B"] * wl["C"] * wl["H"] * wl["W"] * 4
    return lookup_bwd, build_bwd


def synth_inputs(wl, seed):
    """Synthetic feature maps and per-iteration coordinates: smooth disparity +
    noise growing with the iteration, >=1% of pixels out of range on each side."""
    rs = np.random.RandomState(seed)
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    xs = np.broadcast_to(np.arange(W, dtype=np.float32), (B, H, W))
    ys = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None], (B, H, W))
    base = rs.uniform(0, 64, size=(B, H, 1)).astype(np.float32)
    coords = np.empty((wl["iters"], B, 2, H, W), dtype=np.float32)
    for t in range(wl["iters"]):
        x = xs - base - rs.standard_normal((B, H, W)).astype(np.float32) * (0.5 * t)
        m = rs.uniform(size=x.shape)
        x = np.where(m < 0.015, -rs.uniform(1, 30, size=x.shape), x)
        x = np.where(m > 0.985, W + rs.uniform(0, 30, size=x.shape), x)
        coords[t, :, 0] = x.astype(np.float32)
        coords[t, :, 1] = ys
    return f1, f2, coords


def synth_inputs_gpu(torch, wl, B, H, seed, dev):
    """Same recipe as synth_inputs, generated on the device (the large configs would spend seconds
    in numpy): feature maps ~ N(0,1), coordinates = grid - smooth disparity - growing noise, 1.5 %
    of the pixels out of range on each side."""
    C, W, iters = wl["C"], wl["W"], wl["iters"]
    g = torch.Generator(device=dev).manual_seed(seed)
    f1 = torch.randn((B, C, H, W), device=dev, generator=g)
    f2 = torch.randn((B, C, H, W), device=dev, generator=g)
    xs = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W)
    ys = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1).expand(B, H, W)
    base = torch.rand((B, H, 1), device=dev, generator=g) * 64
    coords = torch.empty((iters, B, 2, H, W), device=dev)
    for t in range(iters):
        x = xs - base - torch.randn((B, H, W), device=dev, generator=g) * (0.5 * t)
        m = torch.rand((B, H, W), device=dev, generator=g)
        x = torch.where(m < 0.015, -(1 + 29 * torch.rand((B, H, W), device=dev, generator=g)), x)
        x = torch.where(m > 0.985, W + 30 * torch.rand((B, H, W), device=dev, generator=g), x)
        coords[t, :, 0] = x
        coords[t, :, 1] = ys
    return f1, f2, coords


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:           # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.15 or ts > t1 + 0.15:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ------------------------------------------------------------------------------- CPU baselines
def cpu_port_timer(wl, budget_s, min_runs=2):
    """The oracle's C/OpenMP port of the reference's CPU path (build + iters lookups)."""
    from oracle import corr_oracle as o
    o.build()
    o.use_all_host_threads()
    f1, f2, coords = synth_inputs(wl, seed=0)
    base = o.CorrPathBaseline(wl["B"], wl["C"], wl["H"], wl["W"], wl["W"], wl["L"], wl["r"], wl["iters"])
    base.run(f1, f2, coords)            # warm-up (page faults, thread pool)
    times = []
    t_start = time.perf_counter()
    while len(times) < min_runs or (time.perf_counter() - t_start < budget_s and len(times) < 200):
        t0 = time.perf_counter()
        base.run(f1, f2, coords)
        times.append(time.perf_counter() - t0)
    return times, o.num_threads()


class PytorchReferencePath:
    """The unmodified reference: core/corr.py:110-156 CorrBlock1D on PyTorch CPU (baseline/_ref)."""

    def __init__(self, wl):
        import torch
        from baseline import ref_models as R
        if not R.available():
            raise RuntimeError("baseline/_ref/core is not installed")
        self.torch = torch
        torch.set_num_threads(os.cpu_count() or 1)    # torchrun exports OMP_NUM_THREADS=1
        self.block = R.import_core(with_ref_sampler=False).CorrBlock1D
        assert self.block.__module__ == "core.corr"
        f1, f2, coords = synth_inputs(wl, seed=0)
        self.f1, self.f2, self.coords = torch.from_numpy(f1), torch.from_numpy(f2), torch.from_numpy(coords)
        self.wl = wl
        self.threads = torch.get_num_threads()

    def run(self):
        import warnings
        wl = self.wl
        t0 = time.perf_counter()
        with self.torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            blk = self.block(self.f1, self.f2, num_levels=wl["L"], radius=wl["r"])
            t1 = time.perf_counter()
            out = None
            for t in range(wl["iters"]):
                out = blk(self.coords[t])
        t2 = time.perf_counter()
        assert tuple(out.shape) == (wl["B"], wl["L"] * (2 * wl["r"] + 1), wl["H"], wl["W"])
        return t2 - t0, t1 - t0

    def sample(self, budget_s, min_runs=3, max_runs=50):
        self.run()
        times, ctor = [], []
        t_start = time.perf_counter()
        while len(times) < min_runs or (time.perf_counter() - t_start < budget_s and len(times) < max_runs):
            a, b = self.run()
            times.append(a)
            ctor.append(b)
        return times, ctor


def bench_config(wl):
    """`config` of the JSON line -- the SAME dict in the GPU arm and in the reference arm (the driver compares
    them); what is specific to an arm (input-set count, launch mode, CPU model) goes to `run`."""
    return {"workload": wl["desc"],
            "l2": "GPU arm: steps rotate over independent input sets whose combined footprint exceeds the 126 MB L2 "
                  "(count and bytes in run.l2); CPU arm: one input set in host memory"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the same
    config/metric -- the unmodified PyTorch CorrBlock1D when baseline/_ref is installed (kind
    "reference"), else the oracle's C port (kind "port").  Rank 0 only; all host threads."""
    if rank != 0:
        return
    kind, cores, ctor_ms = "reference", None, None
    try:
        ref = PytorchReferencePath(wl)
        cores = ref.threads
        for _ in range(max(args.warmup, 1)):
            ref.run()
        t0 = time.perf_counter()
        ctor = []
        for _ in range(args.steps):
            ctor.append(ref.run()[1])
        dt = time.perf_counter() - t0
        ctor_ms = 1e3 * statistics.median(ctor)
        what = "unmodified reference core/corr.py CorrBlock1D on PyTorch CPU (baseline/_ref)"
    except Exception as e:          # noqa: BLE001
        kind = "port"
        from oracle import corr_oracle as o
        o.build()
        o.use_all_host_threads()
        f1, f2, coords = synth_inputs(wl, seed=0)
        base = o.CorrPathBaseline(wl["B"], wl["C"], wl["H"], wl["W"], wl["W"], wl["L"], wl["r"], wl["iters"])
        for _ in range(max(args.warmup, 1)):
            base.run(f1, f2, coords)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            base.run(f1, f2, coords)
        dt = time.perf_counter() - t0
        cores = o.num_threads()
        what = f"C/OpenMP port of core/corr.py (oracle/corr_oracle.c); PyTorch reference unavailable: {e}"
    val = args.steps * wl["B"] / dt
    sample = f"{args.steps} full steps of {args.workload} (build + {wl['iters']} lookups each), {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(wl),
        "run": {"device": "host CPU", "cpu": cpu_model()},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "constructor_ms": ctor_ms},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ helpers
def bind_rank_to_cores(local_rank, local_world):
    """One process per GPU: give every rank its own slice of the host cores its GPU is attached to
    (NVML CPU affinity), so that N ranks staging pinned buffers do not share cores.  Returns a
    description for the JSON line."""
    try:
        avail = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    cores = avail
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(avail) // 64) + 1)
        aff = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        aff = [c for c in aff if c in avail]
        if aff:
            cores = aff
    except Exception:               # noqa: BLE001
        pass
    if local_world > 1 and len(cores) >= local_world:
        per = len(cores) // local_world
        mine = cores[local_rank * per:(local_rank + 1) * per]
    else:
        mine = cores
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return {"cores": f"{mine[0]}-{mine[-1]}" if mine else "", "n": len(mine), "gpu_affinity_n": len(cores)}


class Timer:
    """CUDA-event timing of graph replays on the current stream."""

    def __init__(self, torch, stream):
        self.torch, self.stream = torch, stream

    def __call__(self, fn, reps, warm=3):
        torch = self.torch
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for i in range(reps):
            fn(i)
        b.record(self.stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 1e3 / reps


def capture(torch, side, fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            keep = fn()
    return g, keep


def roof(bytes_, secs, hbm_peak, peak_src, traffic=None, written=None):
    """traffic: profiles/r02/traffic.json entry of this kernel instantiation (ncu --set full, one launch).
    written: bytes the kernel writes by design -- ncu's dram__bytes_write only counts what left L2 before
    the kernel ended, so the steady-state DRAM traffic of back-to-back launches is read + written."""
    ach = bytes_ / secs / 1e9
    r = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
         "traffic": (traffic["read"] + traffic["write"]) if traffic else None, "us_per_launch": secs * 1e6,
         "algorithmic_bytes": bytes_, "peak_source": peak_src}
    if traffic:
        r["traffic_detail"] = {"dram_bytes_read": traffic["read"], "dram_bytes_written_before_kernel_end": traffic["write"],
                               "ncu_duration_us": traffic.get("ncu_duration_us"),
                               "source": "profiles/r02/" + traffic.get("source", "traffic.json")}
        if written is not None:
            steady = traffic["read"] + written
            r["traffic_detail"]["steady_state_dram_bytes"] = steady
            r["traffic_detail"]["frac_of_peak_on_dram_bytes"] = steady / secs / 1e9 / hbm_peak
            r["traffic_detail"]["read_overfetch"] = traffic["read"] / max(bytes_ - written, 1)
    return r


def lookup_kernel_name(pyr_bytes):
    """The instantiation launch_fwd_px (csrc/lookup.cu) picks for a 4-level fp32 pyramid of this size."""
    if pyr_bytes > (256 << 20):
        return "lookup_fwd_px_kernel<float,float,4,4,true,true,true,5>"
    if pyr_bytes > (96 << 20):
        return "lookup_fwd_px_kernel<float,float,4,4,true,true,true,0>"
    return "lookup_fwd_px_kernel<float,float,4,4,false,true,false,0>"


# -------------------------------------------------------------------- BASELINE configs 2, 3, 4
def measure_config(ctx, name, B=None, H=None, reps=20):
    """One BASELINE config on this rank's shard (B pairs x H rows): device-resident inputs, the
    step as one CUDA graph, per-kernel-class graphs next to it.  Every lookup of a step writes its
    OWN output buffer (a consumer holds them; nothing is recycled through L2)."""
    torch, sb, dev, side, timer = ctx["torch"], ctx["sb"], ctx["dev"], ctx["side"], ctx["timer"]
    from smoe_stereo_b200.corr import (_batchable, _batched_lookup_backward, _build_backward, _build_into,
                                       _lookup)
    wl = dict(WORKLOADS[name])
    wl["B"] = B if B is not None else wl["B"]
    wl["H"] = H if H is not None else wl["H"]
    B, C, H, W, L, r, iters = (wl[k] for k in ("B", "C", "H", "W", "L", "r", "iters"))
    train = bool(wl.get("train"))
    px = B * H * W
    build_bytes, lookup_bytes, build_flops = algorithmic_bytes(wl)
    widths, _, pitch = sb._lib.pyramid_layout(W, L, sb._lib.DTYPE_F32)
    pyr_bytes = px * pitch * 4
    set_bytes = 2 * B * C * H * W * 4 + px * sum(w for i, w in enumerate(widths) if i % 2 == 0) * 4
    nsets = 1 if set_bytes > 3 * L2_BYTES else int(min(4, np.ceil(3 * L2_BYTES / set_bytes)))
    flags = sb._lib.BUILD_SKIP_ODD_LEVELS            # what CorrBlock1D passes for radius 4
    sets = []
    for s in range(nsets):
        f1, f2, coords = synth_inputs_gpu(torch, wl, B, H, seed=7000 + 100 * ctx["rank"] + s, dev=dev)
        pyr = sb.FusedPyramid(B, H, W, W, L, torch.float32, dev)
        gouts = None
        if train:
            g = torch.Generator(device=dev).manual_seed(9000 + s)
            gouts = [torch.randn((B, L * (2 * r + 1), H, W), device=dev, generator=g) for _ in range(iters)]
        sets.append(dict(f1=f1, f2=f2, coords=coords, pyr=pyr, gouts=gouts))
    torch.cuda.synchronize()

    def build(S):
        _build_into(S["pyr"], S["f1"], S["f2"], flags)

    def lookups(S):
        return [_lookup(S["pyr"], S["coords"][t], r, torch.float32) for t in range(iters)]

    def lookup_bwd(S):
        items = [_batchable(S["pyr"], S["coords"][t], S["gouts"][t], r) for t in range(iters)]
        return _batched_lookup_backward(S["pyr"], items, r, None)

    def step(S):
        build(S)
        outs = lookups(S)
        if not train:
            return outs
        g0 = lookup_bwd(S)
        return outs, g0, _build_backward(g0, S["f1"], S["f2"])

    t_wall0 = time.time()
    g_step = [capture(torch, side, lambda S=S: step(S)) for S in sets]
    g_build = [capture(torch, side, lambda S=S: build(S)) for S in sets]
    g_look = [capture(torch, side, lambda S=S: lookups(S)) for S in sets]
    t_step = timer(lambda i: g_step[i % nsets][0].replay(), reps)
    t_build = timer(lambda i: g_build[i % nsets][0].replay(), reps)
    t_look = timer(lambda i: g_look[i % nsets][0].replay(), reps) / iters
    res = {
        "workload": wl["desc"] if (B, H) == (WORKLOADS[name]["B"], WORKLOADS[name]["H"])
        else f"{wl['desc']} -- this rank's shard: B={B}, H={H}",
        "B_per_gpu": B, "H_per_gpu": H, "pyramid_mb": pyr_bytes / 1e6, "input_sets": nsets,
        "outputs": "every lookup writes its own buffer (no recycling through L2)",
        "launch": "one CUDA graph per step",
    }
    hbm, src = ctx["hbm_peak"], ctx["peak_src"]
    full_cfg = (B, H) == (WORKLOADS[name]["B"], WORKLOADS[name]["H"])
    traffic_of = ctx["traffic"] if full_cfg else (lambda key: None)      # the captures are of whole configs
    kname = lookup_kernel_name(pyr_bytes)
    out_bytes = px * L * (2 * r + 1) * 4
    res["roofline_lookup"] = roof(lookup_bytes, t_look, hbm, src, traffic_of(f"{kname}@{name}"), written=out_bytes)
    res["roofline_lookup"]["kernel"] = kname
    res["roofline_lookup"]["how"] = f"graph of {iters} lookups with {iters} distinct outputs / {iters}"
    moved = 2 * B * C * H * W * 4 + px * sum(w for i, w in enumerate(widths) if i % 2 == 0) * 4
    rb = roof(moved, t_build, hbm, src, traffic_of(f"corr_build_kernel<float,float,false,false>@{name}"),
              written=moved - 2 * B * C * H * W * 4)
    rb.update({"kernel": "corr_build_kernel<float,float,false,false>", "bytes": "moved by design: inputs + levels 0 and 2 "
               "(levels 1/3 are recomputed by the lookup, bit-exactly)", "survey_algorithmic_bytes": build_bytes,
               "frac_on_survey_bytes": build_bytes / t_build / 1e9 / hbm, "tensor_tflops": build_flops / t_build / 1e12,
               "tensor_frac_of_tf32_peak": build_flops / t_build / 1e12 / ctx["tf_peak"]})
    if W % 4:
        rb["note"] = f"W={W} violates TMA's 16-byte stride rule: the launch includes the pitch-padding copy of both feature maps"
    res["roofline_build"] = rb
    breakdown = {"build": t_build * 1e6, "lookup_per_iter": t_look * 1e6}
    if train:
        g_lb = [capture(torch, side, lambda S=S: lookup_bwd(S)) for S in sets]
        t_lb = timer(lambda i: g_lb[i % nsets][0].replay(), reps)
        g0s = [g[1] for g in g_lb]
        g_bb = [capture(torch, side, lambda S=S, g0=g0: _build_backward(g0, S["f1"], S["f2"])) for S, g0 in zip(sets, g0s)]
        t_bb = timer(lambda i: g_bb[i % nsets][0].replay(), reps)
        lb_bytes, bb_bytes = training_bytes(wl)
        res["roofline_lookup_bwd"] = roof(lb_bytes, t_lb, hbm, src, traffic_of(f"lookup_bwd_batched@{name}"),
                                          written=px * W * 4)
        res["roofline_lookup_bwd"].update({"kernel": ctx["bwd_kernel"], "bytes": "grad_out + coords of all "
                                           f"{iters} iterations read once + folded level-0 gradient written once",
                                           "survey_algorithmic_bytes": iters * px * 468})
        res["roofline_build_bwd"] = roof(bb_bytes, t_bb, hbm, src, traffic_of(f"corr_build_bwd_kernel@{name}"),
                                         written=2 * B * C * H * W * 4)
        res["roofline_build_bwd"].update({"kernel": "corr_build_bwd_kernel", "tensor_tflops": 2 * build_flops / t_bb / 1e12})
        breakdown.update({"lookup_backward_batched": t_lb * 1e6, "build_backward": t_bb * 1e6})
        del g_lb, g_bb
    t_wall1 = time.time()
    breakdown["sum_of_classes"] = (t_build + iters * t_look + (t_lb + t_bb if train else 0.0)) * 1e6
    res.update({"ms_per_step": t_step * 1e3, "pairs_per_s": B / t_step, "breakdown_us": breakdown,
                "gpu_launches_per_step": 1 + iters + (2 if train else 0) + (2 if W % 4 else 0) * (2 if train else 1)})
    if ctx["sampler"] is not None:
        res["clocks"] = ctx["sampler"].window(t_wall0, t_wall1)
    res["_t_step"] = t_step
    res["_sample_out"] = g_look[0][1][0]
    del g_step, g_build
    return res


def measure_fused_gather(ctx, name, world, rank):
    """Strong-scaling configs at N > 1: the optional output gather FUSED into the lookup kernel
    (smoe_corr_lookup_to through an NVSwitch multicast address; ShardedCorrBlock1D gather="fused")
    next to the shard-local lookup, both as CUDA graphs of 8 lookups, max over ranks."""
    import torch.distributed as dist
    torch, sb, dev, side, timer = ctx["torch"], ctx["sb"], ctx["dev"], ctx["side"], ctx["timer"]
    from smoe_stereo_b200.sharding import shard_bounds
    full = WORKLOADS[name]
    mode = "batch" if (full["B"] >= world and full["B"] % world == 0) else "rows"
    ext = full["B"] if mode == "batch" else full["H"]
    lo, hi = shard_bounds(ext, world, rank)
    B = hi - lo if mode == "batch" else full["B"]
    H = hi - lo if mode == "rows" else full["H"]
    wl = dict(full)
    f1, f2, coords = synth_inputs_gpu(torch, wl, B, H, seed=7300 + rank, dev=dev)
    blk = sb.ShardedCorrBlock1D(f1, f2, mode=mode, already_sharded=True)
    del f1, f2
    c = coords[0].contiguous()

    def graph_us(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn()
            torch.cuda.synchronize()
            dist.barrier()
            with torch.cuda.graph(g, stream=side):
                keep = [fn() for _ in range(8)]
        torch.cuda.synchronize()
        t = timer(lambda i: g.replay(), 10) / 8
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        del g, keep
        return float(tt.item())

    with torch.no_grad():
        t_local = graph_us(lambda: blk(c))
        t_fused = graph_us(lambda: blk(c, gather="fused"))
    gathered = full["B"] * full["L"] * (2 * full["r"] + 1) * full["H"] * full["W"] * 4
    return {"us_shard_lookup_only": t_local * 1e6, "us_fused_lookup_and_gather": t_fused * 1e6,
            "bytes_gathered": gathered, "gathered_gbs_per_rank": gathered / t_fused / 1e9,
            "what": "ONE kernel per lookup: every rank stores its shard of the result through an NVSwitch multicast "
                    "address (multimem.st, torch symmetric memory) + a symmetric-memory barrier; bit-identical to "
                    "lookup + NCCL all-gather (tests/test_gpu_fused_gather.py)"}


def measure_fp16(ctx, name, reps=20):
    """The reference's "reg_cuda" mode under its default mixed precision (core/SMoEStereo.py:217-218,
    core/corr.py:31-61): fp16 feature maps -> fp16 pyramid (tcgen05 kind::f16) -> fp16 lookups, what
    CorrBlockFast1D runs.  Whole config on this rank (N=1 only); same graphs as measure_config."""
    torch, sb, dev, side, timer = ctx["torch"], ctx["sb"], ctx["dev"], ctx["side"], ctx["timer"]
    from smoe_stereo_b200.corr import _build_into, _lookup
    wl = WORKLOADS[name]
    B, C, H, W, L, r, iters = (wl[k] for k in ("B", "C", "H", "W", "L", "r", "iters"))
    px = B * H * W
    f1, f2, coords = synth_inputs_gpu(torch, wl, B, H, seed=7100 + ctx["rank"], dev=dev)
    f1, f2 = f1.half(), f2.half()
    pyr = sb.FusedPyramid(B, H, W, W, L, torch.float16, dev)
    flags = sb._lib.BUILD_SKIP_ODD_LEVELS
    widths = pyr.widths
    g_build = capture(torch, side, lambda: _build_into(pyr, f1, f2, flags))
    g_look = capture(torch, side, lambda: [_lookup(pyr, coords[t], r, torch.float16) for t in range(iters)])
    t_build = timer(lambda i: g_build[0].replay(), reps)
    t_look = timer(lambda i: g_look[0].replay(), reps) / iters
    hbm, src = ctx["hbm_peak"], ctx["peak_src"]
    moved = 2 * B * C * H * W * 2 + px * sum(w for i, w in enumerate(widths) if i % 2 == 0) * 2
    look_bytes = px * (4 + L * (2 * r + 2) * 2 + L * (2 * r + 1) * 2)
    res = {"workload": wl["desc"] + " -- fp16 feature maps, fp16 pyramid, fp16 lookup outputs (CorrBlockFast1D / reg_cuda)",
           "pyramid_mb": pyr.buffer.numel() * 2 / 1e6,
           "roofline_build": roof(moved, t_build, hbm, src), "roofline_lookup": roof(look_bytes, t_look, hbm, src),
           "ms_per_step_sum": (t_build + iters * t_look) * 1e3,
           "breakdown_us": {"build": t_build * 1e6, "lookup_per_iter": t_look * 1e6}}
    res["roofline_build"]["kernel"] = "corr_build_kernel<__half,__half,false,false>"
    res["roofline_lookup"]["kernel"] = "lookup_fwd_px_kernel<__half,__half,4,4,...>"
    return res


def measure_cfg5(ctx, frames=3):
    """BASELINE config 5: the UNMODIFIED reference SMoEStereo (DAMv2 ViT-L + SMoE PEFT, random init,
    baseline/_ref) on a 540x960 frame padded to 544x960, 32 iterations, default mixed precision --
    once with the stock core/corr.py on the GPU and once after smoe_stereo_b200.install()."""
    torch, sb, dev = ctx["torch"], ctx["sb"], ctx["dev"]
    from baseline import ref_models as R
    if not R.available():
        return {"unavailable": "baseline/_ref/core not installed (tools/install_ref.sh)"}
    import warnings
    t_wall0 = time.time()
    args = R.make_args("reg", mixed_precision=True, model="smoe")
    model = R.build_model("smoe", args, seed=0, device=dev)
    left, right = R.synthetic_pair(1, 544, 960, max_disp=64.0, seed=5, device=dev)

    def run(n):
        out = None
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(n):
                out = model(left, right, iters=32, test_mode=True)["disp_preds"][-1]
        torch.cuda.synchronize()
        return out

    def timed():
        run(2)
        t0 = time.perf_counter()
        out = run(frames)
        return (time.perf_counter() - t0) / frames, out.float()

    sb.uninstall()
    t_stock, d_stock = timed()
    sb.install()
    try:
        t_ours, d_ours = timed()
    finally:
        sb.uninstall()
    sb.install(output_ring=2)          # opt-in: lookups served from a two-buffer ring
    try:
        t_ring, d_ring = timed()
    finally:
        sb.uninstall()
    res = {"workload": "SMoEStereo damv2/vitl + SMoE PEFT (unmodified reference model, random init), 544x960, B=1, "
                       "32 iterations, mixed_precision=True, corr_implementation=reg",
           "frames_per_s": 1.0 / t_ours, "frames_per_s_stock_corr_on_gpu": 1.0 / t_stock,
           "frames_per_s_install_output_ring_2": 1.0 / t_ring,
           "output_ring_equals_default": bool(torch.equal(d_ring, d_ours)),
           "ms_per_frame": 1e3 * t_ours, "ms_per_frame_stock_corr_on_gpu": 1e3 * t_stock,
           "epe_px_vs_stock": (d_ours - d_stock).abs().mean().item(), "frames_timed": frames,
           "how": "wall clock around model(left, right, iters=32, test_mode=True) + synchronize, eager PyTorch; "
                  "stock = reference core/corr.py CorrBlock1D on the same GPU; ours = after install()"}
    if ctx["sampler"] is not None:
        res["clocks"] = ctx["sampler"].window(t_wall0, time.time())
    res["_t"] = t_ours
    del model
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5",
                    help="extra BASELINE configs measured after the headline ('none' to skip)")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU baseline sampling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sets", type=int, default=0, help="independent input sets to rotate over")
    ap.add_argument("--reps", type=int, default=5, help="repetitions of the timed region (value_stats)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    extra = [] if args.configs.strip().lower() in ("", "none") else [c.strip() for c in args.configs.split(",")]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return

    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) are sent to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    binding = bind_rank_to_cores(local_rank, local_world) if world > 1 else None
    import torch
    import torch.distributed as dist
    import smoe_stereo_b200 as sb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def all_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, C, H, W, L, r, iters = (wl[k] for k in ("B", "C", "H", "W", "L", "r", "iters"))
    build_bytes, lookup_bytes, build_flops = algorithmic_bytes(wl)
    pitch = sb._lib.pyramid_layout(W, L, sb._lib.DTYPE_F32)[2]
    set_bytes = 2 * B * C * H * W * 4 + B * H * W * pitch * 4 + iters * B * 2 * H * W * 4
    nsets = args.sets or max(2, int(np.ceil(3 * L2_BYTES / set_bytes)) + 1)
    nsets = min(nsets, 8)

    # ---- resident inputs + one captured CUDA graph per input set
    sets = []
    host = []
    for s in range(nsets):
        f1, f2, coords = synth_inputs(wl, seed=1000 * rank + s)
        host.append((f1, f2, coords))
        sets.append((torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev),
                     torch.from_numpy(coords).to(dev)))
    torch.cuda.synchronize()

    # per-iteration coordinate tensors, as a model has them (slicing a stacked tensor costs ~2 us of host
    # time per call, two thirds of a lookup kernel)
    coord_lists = {id(s[2]): list(s[2].unbind(0)) for s in sets}

    def step_eager(f1, f2, coords, ring=0):
        blk = sb.CorrBlock1D(f1, f2, num_levels=L, radius=r)
        if ring:
            blk.reuse_outputs(ring)
        out = None
        for c in coord_lists[id(coords)]:
            out = blk(c)
        return blk, out

    side = torch.cuda.Stream(dev)
    graphs = []
    with torch.cuda.stream(side), torch.no_grad():
        for s in range(nsets):
            step_eager(*sets[s])                    # warm the allocator / module outside capture
        torch.cuda.synchronize()
        for s in range(nsets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                keep = step_eager(*sets[s])
            graphs.append((g, keep))
    torch.cuda.synchronize()

    # ---- value: K graph-replayed steps, device-timed; repeated args.reps times for the statistics
    stream = torch.cuda.current_stream(dev)
    timer = Timer(torch, stream)
    for i in range(args.warmup):
        graphs[i % nsets][0].replay()
    rep_values, dev_s, wall0, wall1 = [], None, None, None
    for rep in range(max(args.reps, 1)):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record(stream)
        for i in range(args.steps):
            graphs[i % nsets][0].replay()
        e1.record(stream)
        barrier()
        w1 = time.time()
        secs = all_max(e0.elapsed_time(e1) / 1e3)
        rep_values.append(world * args.steps * B / secs)
        if rep == 0:                                # the reported value is the FIRST timed region
            dev_s, wall0, wall1 = secs, w0, w1
    value = world * args.steps * B / dev_s
    value_stats = {"reps": len(rep_values), "min": min(rep_values), "max": max(rep_values),
                   "median": statistics.median(rep_values),
                   "stddev": statistics.pstdev(rep_values) if len(rep_values) > 1 else 0.0,
                   "note": "value is repetition 0; every repetition times exactly --steps graph replays"}
    clocks = sampler.window(wall0, wall1) if sampler else None

    # ---- the same steps issued round-robin on several CUDA streams: a B=1 step leaves most of the GPU
    # idle (single-wave, latency-bound lookups), independent pairs fill it.  NOT the headline (`value`
    # stays one stream, one pair at a time, comparable with round 1): reported next to it.
    concurrent = {}
    for ns in (2, 4):
        if ns > nsets:
            continue
        streams = [torch.cuda.Stream(dev) for _ in range(ns)]
        n_steps = max(args.steps // ns, 1) * ns

        def issue(n):
            for i in range(n):
                with torch.cuda.stream(streams[i % ns]):
                    graphs[i % nsets][0].replay()          # graph i % nsets only ever runs on stream i % ns
        for st_ in streams:
            st_.wait_stream(stream)
        issue(2 * ns)
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for st_ in streams:
            st_.wait_stream(stream)
        issue(n_steps)
        for st_ in streams:
            stream.wait_stream(st_)
        e1.record(stream)
        barrier()
        secs = all_max(e0.elapsed_time(e1) / 1e3)
        concurrent[f"{ns}_streams"] = {"pairs_per_s": world * n_steps * B / secs, "steps": n_steps,
                                       "us_per_step_amortised": 1e6 * secs / n_steps}
    if concurrent:
        concurrent["note"] = ("independent B=1 steps (one CUDA graph each) issued round-robin on 2 / 4 streams; the "
                              "pyramids of concurrent pairs no longer all fit L2; latency per pair is not improved")

    # ---- per-kernel durations (CUDA events on the launching stream, same rotation over sets)
    from smoe_stereo_b200.corr import _build_into, _lookup
    pyrs = [g[1][0]._state.pyr for g in graphs]

    def lookups_only(s):
        o = None
        for t in range(iters):      # the previous output is dropped, as in the GRU loop
            o = _lookup(pyrs[s], sets[s][2][t], r, torch.float32)
        return o

    lk_graphs = [capture(torch, side, lambda s=s: lookups_only(s)) for s in range(nsets)]
    # one graph holding a build of every input set: kernel time without per-graph launch overhead
    bflags = graphs[0][1][0]._state.build_flags      # what CorrBlock1D passes (levels 1/3 not written)
    bd_graph = capture(torch, side, lambda: [_build_into(pyrs[s], sets[s][0], sets[s][1], bflags) for s in range(nsets)])
    t_lookup_alone = timer(lambda i: lk_graphs[i % nsets][0].replay(), 50, warm=5) / iters
    # the same graph replayed on ONE input set: pyramid and coordinates L2-resident, i.e. where a real
    # update loop leaves them (the GRU has just written coords1); reported next to the cold-input numbers
    t_lookup_warm = timer(lambda i: lk_graphs[0][0].replay(), 50, warm=5) / iters
    t_build = timer(lambda i: bd_graph[0].replay(), 50, warm=5) / nsets
    # the lookup's duration INSIDE the timed step (pyramid just written by the build, L2-warm):
    # (step - build) / iters, from the same CUDA-event timing as `value`
    t_lookup = max((dev_s / args.steps - t_build) / iters, 1e-9)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:           # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    tf_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0     # tf32 = half the bf16 rate

    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r02", "traffic.json")))
    except Exception:           # noqa: BLE001
        pass

    def ncu_traffic(key):
        return traffic_tab.get(key)

    pyr_mb = B * H * W * pitch * 4 / 1e6
    big = pyr_mb * 1e6 > (96 << 20)
    kname = lookup_kernel_name(int(pyr_mb * 1e6))
    roofline = roof(lookup_bytes, t_lookup, hbm_peak, peak_src, ncu_traffic(f"{kname}@{args.workload}"))
    if roofline.get("traffic_detail"):
        roofline["traffic_detail"]["note"] = ("cold-cache single launch under ncu; inside the timed step the pyramid "
                                              "is L2-resident and only the coordinates and outputs touch DRAM")
    roofline["kernel"] = kname
    roofline["us_per_launch_standalone"] = t_lookup_alone * 1e6
    roofline["us_per_launch_l2_warm_inputs"] = t_lookup_warm * 1e6
    roofline["frac_l2_warm_inputs"] = lookup_bytes / t_lookup_warm / 1e9 / hbm_peak
    roofline["how"] = ("(event-timed step - event-timed build) / 32 lookups, i.e. the launches inside the timed "
                       "region; us_per_launch_standalone = lookups-only graphs over rotating (L2-cold) input sets; "
                       "us_per_launch_l2_warm_inputs = the same graph replayed on one set (pyramid AND coordinates in L2, "
                       "as after a GRU update), not the timed region")
    roofline["traffic_note"] = ("ncu cold-cache replay (profiles/r02); inside the timed step the pyramid is "
                                "L2-resident and DRAM traffic is ~0" if not big else "ncu capture, profiles/r02")
    roofline["note"] = ("pyramid of this workload (%.0f MB) fits the 126 MB L2: bandwidth is L2-assisted and the "
                        "kernel is launch/latency-bound at this size; see roofline_hbm for the HBM regime" % pyr_mb
                        if not big else "pyramid exceeds L2: HBM-bound")
    widths = level_widths(W, L)
    moved = 2 * B * C * H * W * 4 + B * H * W * sum(w for i, w in enumerate(widths)
                                                    if i % 2 == 0 or not (bflags & sb._lib.BUILD_SKIP_ODD_LEVELS)) * 4
    roofline_build = roof(moved, t_build, hbm_peak, peak_src,
                          ncu_traffic(f"corr_build_kernel<float,float,false,false>@{args.workload}"),
                          written=moved - 2 * B * C * H * W * 4)
    roofline_build["kernel"] = "corr_build_kernel<float,float,false,false>"
    roofline_build["bytes"] = ("bytes the kernel moves by design: both feature maps once + levels 0 and 2 (levels 1 and 3 "
                               "are recomputed by the lookup, bit-exactly); frac is on these")
    roofline_build["survey_algorithmic_bytes"] = build_bytes
    roofline_build["frac_on_survey_bytes"] = build_bytes / t_build / 1e9 / hbm_peak
    roofline_build["tensor_tflops"] = build_flops / t_build / 1e12
    roofline_build["tensor_frac_of_tf32_peak"] = build_flops / t_build / 1e12 / tf_peak

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region.
    # Software-pipelined like any host-fed serving loop: while pair i is being processed, pair i+1's
    # feature maps / coords are copied H2D on a copy stream and pair i's lookup results are
    # copied D2H on another (PCIe is full duplex).  Every step still moves all of its own bytes.
    pin = [(torch.from_numpy(h[0]).pin_memory(), torch.from_numpy(h[1]).pin_memory(),
            torch.from_numpy(h[2]).pin_memory()) for h in host[:2]]
    out_host = [torch.empty((iters, B, L * (2 * r + 1), H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    dbuf = [tuple(torch.empty_like(t) for t in sets[0]) for _ in range(2)]
    h2d_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_free = [torch.cuda.Event() for _ in range(2)]
    for e in in_free + out_free:
        e.record(stream)

    def issue_h2d(i):
        slot = i % 2
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(in_free[slot])
            for dst, src in zip(dbuf[slot], pin[slot]):
                dst.copy_(src, non_blocking=True)
            in_ready[slot].record(h2d_stream)

    # device-side staging for the lookup results of a pair: the caller owns these buffers (out=), so
    # the loop never touches the caching allocator and never blocks the host; reuse of a slot is
    # ordered on the device (compute stream waits for the D2H copies of pair i-2 out of that slot)
    out_dev = [torch.empty((iters, B, L * (2 * r + 1), H, W), dtype=torch.float32, device=dev) for _ in range(2)]

    def compute(i, with_kernels=True):
        slot = i % 2
        f1, f2, c = dbuf[slot]
        oh, od = out_host[slot], out_dev[slot]
        stream.wait_event(in_ready[slot])
        stream.wait_event(out_free[slot])            # D2H of pair i-2 has drained out_dev[slot] / out_host[slot]
        blk = sb.CorrBlock1D(f1, f2, num_levels=L, radius=r) if with_kernels else None
        for t in range(iters):
            if with_kernels:
                blk(c[t], out=od[t])
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev)
                oh[t].copy_(od[t], non_blocking=True)
        in_free[slot].record(stream)
        out_free[slot].record(d2h_stream)

    def run_e2e(n, with_kernels=True):
        issue_h2d(0)
        for i in range(n):
            if i + 1 < n:
                issue_h2d(i + 1)
            compute(i, with_kernels)
        stream.wait_stream(d2h_stream)

    e2e_steps = max(20, min(args.steps, 200))
    def timed_e2e(with_kernels):
        run_e2e(3, with_kernels)
        barrier()
        t0 = time.perf_counter()
        run_e2e(e2e_steps, with_kernels)
        barrier()
        return all_max(time.perf_counter() - t0)

    # the host <-> device path of a shared box is noisy (PCIe throughput moved between 220 and 315
    # pairs/s from one run to the next): three repetitions of exactly e2e_steps steps each, with
    # and without kernels interleaved; the reported value is the MEDIAN repetition, all are listed
    with torch.set_grad_enabled(os.environ.get("BENCH_E2E_GRAD", "0") == "1"):
        e2e_runs, copy_runs = [], []
        for _ in range(3):
            e2e_runs.append(timed_e2e(True))
            copy_runs.append(timed_e2e(False))     # the same bytes with no kernels at all
        e2e_s = statistics.median(e2e_runs)
        copy_s = statistics.median(copy_runs)
    h2d = 2 * B * C * H * W * 4 + iters * B * 2 * H * W * 4
    d2h = iters * B * L * (2 * r + 1) * H * W * 4
    e2e = {"value": world * e2e_steps * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "repetitions": {"pairs_per_s": [world * e2e_steps * B / t for t in e2e_runs],
                           "host_copy_ceiling_pairs_per_s": [world * e2e_steps * B / t for t in copy_runs],
                           "note": "value = median repetition; every repetition times exactly `steps` steps"},
           "pcie_gbs_per_rank": {"h2d": h2d * e2e_steps / e2e_s / 1e9, "d2h": d2h * e2e_steps / e2e_s / 1e9},
           "host_copy_ceiling": {"value": world * e2e_steps * B / copy_s, "unit": UNIT,
                                 "aggregate_gbs": world * (h2d + d2h) * e2e_steps / copy_s / 1e9,
                                 "how": "the same pinned H2D + D2H copies on the same streams with no kernels: the "
                                        "ceiling the host<->device path of this box sets at this rank count"},
           "frac_of_host_copy_ceiling": (world * e2e_steps * B / e2e_s) / (world * e2e_steps * B / copy_s),
           "host_binding": binding,
           "how": "CorrBlock1D(...) + 32 x __call__(coords, out=staging[t]) fed from pinned host buffers; all 32 lookup "
                  "results of every pair copied back to pinned host memory; H2D(i+1) / compute(i) / D2H(i) pipelined "
                  "on 3 streams, double-buffered, ordered by events on the device (the host never blocks)"}

    # ---- eager (non-graph) API throughput, inputs resident
    with torch.no_grad():
        t_eager = timer(lambda i: step_eager(*sets[i % nsets]), 50, warm=5)
        t_eager_ring = timer(lambda i: step_eager(*sets[i % nsets], ring=2), 50, warm=5)

    # ---- BASELINE configs 2-5 (N=1: whole configs; N>1: this rank's shard = strong scaling)
    from smoe_stereo_b200.sharding import gather_lookup, shard_bounds
    ctx = dict(torch=torch, sb=sb, dev=dev, side=side, timer=timer, rank=rank, hbm_peak=hbm_peak, peak_src=peak_src,
               tf_peak=tf_peak, sampler=sampler, traffic=ncu_traffic, bwd_kernel="lookup_bwd_tma_kernel<4>")
    del lk_graphs, bd_graph, graphs, pin, out_host, dbuf, out_dev
    torch.cuda.empty_cache()
    configs, strong = {}, {}
    for name in extra:
        try:
            if name == "cfg5":
                res = measure_cfg5(ctx)
                if "_t" in res:
                    t = all_max(res.pop("_t"))
                    res["frames_per_s_all_gpus"] = world / t
                    res["partitioning"] = f"{world} independent replica(s), one frame each"
                configs[name] = res
                continue
            full = WORKLOADS[name]
            if world == 1:
                res = measure_config(ctx, name)
            elif full["B"] >= world and full["B"] % world == 0:
                res = measure_config(ctx, name, B=full["B"] // world)
                res["sharding"] = f"batch: global B={full['B']} over {world} ranks"
            else:
                lo, hi = shard_bounds(full["H"], world, rank)
                res = measure_config(ctx, name, H=hi - lo)
                res["sharding"] = f"rows: H={full['H']} over {world} ranks ({hi - lo} on rank {rank})"
            t_step = all_max(res.pop("_t_step"))
            sample = res.pop("_sample_out")
            res["ms_per_step_max_over_ranks"] = t_step * 1e3
            res["pairs_per_s_all_gpus"] = full["B"] / t_step
            if world > 1:
                mode = "batch" if "batch" in res["sharding"] else "rows"
                full_extent = full["B"] if mode == "batch" else full["H"]
                gather_lookup(sample, mode, full_extent)            # warm NCCL
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(5):
                    gathered = gather_lookup(sample, mode, full_extent)
                b.record(stream)
                torch.cuda.synchronize()
                t_g = all_max(a.elapsed_time(b) / 1e3 / 5)
                res["gather_lookup"] = {"us": t_g * 1e6, "bytes_gathered": gathered.numel() * 4,
                                        "algbw_gbs": gathered.numel() * 4 / t_g / 1e9,
                                        "what": "NCCL all-gather of ONE lookup output to every rank (optional; only a "
                                                "non-sharded consumer needs it), timed apart from the step"}
                del gathered
                try:
                    res["fused_lookup_gather"] = measure_fused_gather(ctx, name, world, rank)
                except Exception as e:      # noqa: BLE001  (no multicast support, symmetric memory unavailable ...)
                    res["fused_lookup_gather"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
                strong[name] = {"n_gpus": world, "ms_per_step": t_step * 1e3, "pairs_per_s": full["B"] / t_step,
                                "sharding": res["sharding"]}
            configs[name] = res
            del sample
        except Exception as e:      # noqa: BLE001  (a failing extra config must not take the headline down)
            configs[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    if world == 1 and "cfg3" in extra:
        try:
            configs["cfg3_reg_cuda_fp16"] = measure_fp16(ctx, "cfg3")
        except Exception as e:      # noqa: BLE001
            configs["cfg3_reg_cuda_fp16"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    roofline_hbm = None
    if "cfg3" in configs and "roofline_lookup" in configs["cfg3"]:
        roofline_hbm = dict(configs["cfg3"]["roofline_lookup"])
        roofline_hbm["workload"] = configs["cfg3"]["workload"]
        roofline_hbm["note"] = ("the lookup kernel where the pyramid (%.0f MB on this rank) streams from HBM; "
                                "32 distinct output buffers per step" % configs["cfg3"]["pyramid_mb"])

    if sampler:
        sampler.stop()

    # ---- CPU baselines (rank 0, N=1 only): the reference itself, and the C port of it
    cpu = cpu_port = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = PytorchReferencePath(wl)
            times, ctor = ref.sample(args.cpu_budget)
            med = statistics.median(times)
            cpu = {"value": wl["B"] / med, "unit": UNIT, "cores": ref.threads, "kind": "reference",
                   "best": wl["B"] / min(times), "runs": len(times), "os_cpu_count": os.cpu_count(), "cpu": cpu_model(),
                   "constructor_ms": 1e3 * statistics.median(ctor), "lookups_ms": 1e3 * (med - statistics.median(ctor)),
                   "sample": f"{len(times)} full steps of {args.workload} (CorrBlock1D(...) + {iters} x __call__), unmodified "
                             f"reference core/corr.py on PyTorch CPU, torch.get_num_threads()={ref.threads}, median"}
        except Exception as e:      # noqa: BLE001
            cpu = None
            cpu_err = f"{type(e).__name__}: {e}"
        times, cores = cpu_port_timer(wl, args.cpu_budget / 2)
        cpu_port = {"value": wl["B"] / statistics.median(times), "unit": UNIT, "cores": cores, "kind": "port",
                    "best": wl["B"] / min(times), "runs": len(times),
                    "sample": f"{len(times)} full steps of {args.workload} (build + {iters} lookups), oracle C port "
                              f"(OpenMP, {cores} threads), median"}
        if cpu is None:
            cpu = dict(cpu_port)
            cpu["note"] = f"PyTorch reference unavailable ({cpu_err}); the C port is the baseline"

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (fp32 in/out, fp32 accumulate)",
            "data": "synthetic",
            # `config` is identical in both arms (the driver compares them); how this arm runs it is in `run`
            "config": bench_config(wl),
            "run": {"per_gpu_batch": B, "partitioning": f"batch x{world}, no collective",
                    "l2": f"steps rotate over {nsets} independent input sets ({nsets * set_bytes / 1e6:.0f} MB > 126 MB L2)",
                    "launch": "one CUDA graph per step (1 build + 32 lookup kernels)"},
            "value_stats": value_stats,
            "throughput_concurrent_streams": concurrent or None,
            "e2e": e2e, "gpu_launches": args.steps * (1 + iters),
            "roofline": roofline, "roofline_build": roofline_build, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu, "cpu_baseline_port": cpu_port, "clocks": clocks,
            "eager_api_pairs_per_s": B / t_eager,
            "eager_api_pairs_per_s_output_ring": B / t_eager_ring,
            "breakdown_us": {"build": t_build * 1e6, "lookup_per_iter_in_step": t_lookup * 1e6,
                             "lookup_per_iter_standalone": t_lookup_alone * 1e6,
                             "lookup_per_iter_l2_warm_inputs": t_lookup_warm * 1e6,
                             "graph_step": 1e6 * dev_s / args.steps},
            "configs": configs,
            "strong_scaling": strong or None,
            "wall_s": wall1 - wall0,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
