import sys
sys.path.insert(0, "/root/repo")
import torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _lookup
from bench import WORKLOADS, synth_inputs
dev = torch.device("cuda", 0)
side = torch.cuda.Stream(dev)
for cfg in ("cfg1", "cfg3"):
    wl = dict(WORKLOADS[cfg])
    if cfg == "cfg3":
        wl["B"] = 8
    f1, f2, coords = synth_inputs(wl, 0) if cfg == "cfg1" else (None, None, None)
    B, H, W = wl["B"], wl["H"], wl["W"]
    pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
    pyr.buffer.normal_()
    pyr.mark_missing(0b1010)
    if coords is None:
        g = torch.Generator(device="cpu").manual_seed(0)
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
        c = [(xs - torch.rand(B, 1, H, 1, generator=g) * 60 - torch.randn(B, 1, H, W, generator=g) * (1 + 0.5 * t)) for t in range(32)]
        m = [torch.rand(B, 1, H, W, generator=g) for _ in range(32)]
        c = [torch.where(mm < 0.015, -10.0 * torch.ones(()), torch.where(mm > 0.985, W + 10.0 * torch.ones(()), cc)) for cc, mm in zip(c, m)]
        coords_t = torch.stack(c).to(dev)
    else:
        coords_t = torch.from_numpy(coords[:, :, :1].copy()).to(dev)
    variants = {"bench coords (3% random OOB + left edge)": coords_t,
                "clamped to [24, W-24] (every window interior)": coords_t.clamp(24.0, W - 24.0)}
    for name, ct in variants.items():
        def lk():
            return [_lookup(pyr, ct[t], 4, torch.float32) for t in range(32)]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            lk(); torch.cuda.synchronize()
            with torch.cuda.graph(gr, stream=side):
                keep = lk()
        for _ in range(5): gr.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): gr.replay()
        b.record(); torch.cuda.synchronize()
        print(f"{cfg} {name}: {a.elapsed_time(b) * 1e3 / 50 / 32:6.2f} us per lookup", flush=True)
        del gr, keep
