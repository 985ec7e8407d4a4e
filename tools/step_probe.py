"""Where does the bench step's time go?  One CUDA graph per input set holding CorrBlock1D(...) +
N lookups, replayed over 4 rotating sets (as bench.py), for N = 0, 1, 2, 8, 32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
from bench import WORKLOADS, synth_inputs
dev = torch.device("cuda", 0)
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg1"]
sets = []
for s in range(4):
    f1, f2, coords = synth_inputs(wl, seed=s)
    sets.append(tuple(torch.from_numpy(x).to(dev) for x in (f1, f2, coords)))
side = torch.cuda.Stream(dev)
for N in (0, 1, 2, 8, 32):
    def step(f1, f2, coords):
        blk = sb.CorrBlock1D(f1, f2)
        outs = [blk(coords[t]) for t in range(N)]
        return blk, outs
    graphs = []
    with torch.cuda.stream(side):
        for s in range(4): step(*sets[s])
        torch.cuda.synchronize()
        for s in range(4):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                keep = step(*sets[s])
            graphs.append((g, keep))
    for i in range(8): graphs[i % 4][0].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(400): graphs[i % 4][0].replay()
    b.record(); torch.cuda.synchronize()
    print(f"N={N:2d}: {a.elapsed_time(b) * 1e3 / 400:7.2f} us per step (all_levels={os.environ.get('SMOE_CORR_ALL_LEVELS')})")
    del graphs
