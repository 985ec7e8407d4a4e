"""reg_cuda-under-autocast path: fp16 feature maps -> fp16 pyramid -> fp16 lookups (kind::f16)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _build_into, _lookup
from bench import WORKLOADS, synth_inputs
dev = torch.device("cuda", 0)
for name in (sys.argv[1:] or ["cfg1", "cfg3"]):
    wl = WORKLOADS[name]
    f1, f2, coords = synth_inputs(wl, 0)
    f1h, f2h = torch.from_numpy(f1).to(dev).half(), torch.from_numpy(f2).to(dev).half()
    coords = torch.from_numpy(coords).to(dev)
    blk = sb.CorrBlockFast1D(f1h, f2h)
    pyr = blk._state.pyr
    side = torch.cuda.Stream(dev)
    def graph_of(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn(); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                keep = fn()
        return g, keep
    def timeit(g, reps=100):
        for _ in range(5): g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / reps
    def lk():
        o = None
        for t in range(32):
            o = _lookup(pyr, coords[t], 4, torch.float16)
        return o
    px = wl["B"] * wl["H"] * wl["W"]
    tb = timeit(graph_of(lambda: _build_into(pyr, f1h, f2h))[0])
    tl = timeit(graph_of(lk)[0]) / 32
    bbytes = 2 * px * wl["C"] * 2 + px * sum(pyr.widths) * 2
    lbytes = px * (4 + 4 * 10 * 2 + 36 * 2)
    print(f"{name} fp16: build {tb:.1f} us ({bbytes/tb/1e3:.0f} GB/s algorithmic), lookup {tl:.2f} us/iter ({lbytes/tl/1e3:.0f} GB/s), "
          f"step(build+32) {tb + 32*tl:.0f} us")
