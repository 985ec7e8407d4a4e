"""Where does a B=1 lookup's ~4 us go?  Graph-node launch floor vs kernel time, warm vs cold L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _build_into, _lookup
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import WORKLOADS, synth_inputs

dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
wl = WORKLOADS[name]
f1, f2, coords = synth_inputs(wl, 0)
f1, f2, coords = (torch.from_numpy(x).to(dev) for x in (f1, f2, coords))
blk = sb.CorrBlock1D(f1, f2)
pyr = blk._state.pyr
side = torch.cuda.Stream(dev)

def graph_of(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            keep = fn()
    return g, keep

def timeit(g, reps=200, per=32):
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps / per

x = torch.zeros(32, device=dev)
g_empty, _ = graph_of(lambda: [x.add_(1.0) for _ in range(32)])
print("tiny torch kernel x32 in a graph : %.2f us per node" % timeit(g_empty))
def lk32(same=False):
    o = None
    for t in range(32):
        o = _lookup(pyr, coords[0 if same else t], 4, torch.float32)   # previous output is recycled
    return o
g_lk, _ = graph_of(lk32)
print("lookup x32 (same pyramid, warm L2): %.2f us per lookup" % timeit(g_lk))
g_lk1, _ = graph_of(lambda: lk32(True))
print("lookup x32 same coords           : %.2f us per lookup" % timeit(g_lk1))
out = torch.empty(wl["B"], 36, wl["H"], wl["W"], device=dev)
g_bd, _ = graph_of(lambda: _build_into(pyr, f1, f2))
print("build (inputs warm in L2 if they fit): %.2f us" % timeit(g_bd, per=1))
# cold: flush L2 between replays with a big memset (timed separately and subtracted)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def cold(g, reps=30):
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b) * 1e3
    return tot / reps
g_one, _ = graph_of(lambda: _lookup(pyr, coords[3], 4, torch.float32))
print("single lookup, L2 flushed (dirty)  : %.2f us" % cold(g_one))
print("single build,  L2 flushed (dirty)  : %.2f us" % cold(g_bd))
