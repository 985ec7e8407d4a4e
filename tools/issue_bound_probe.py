"""Is the level-pair lookup bound by instruction issue or by DRAM?  Same pixel count as cfg3
(B=8, 96x312 -> 239,616 pixels) over volumes of different widths: narrow ones are L2-resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
from smoe_stereo_b200.corr import _lookup
dev = torch.device("cuda", 0)
B, H, W1 = 8, 96, 312
side = torch.cuda.Stream(dev)
for W2 in (32, 64, 312):
    pyr = sb.FusedPyramid(B, H, W1, W2, 4, torch.float32, dev)
    pyr.buffer.normal_()
    g = torch.Generator(device="cpu").manual_seed(0)
    coords = [(torch.rand(B, 1, H, W1, generator=g) * (W2 - 1)).to(dev) for _ in range(8)]
    def lk():
        o = None
        for t in range(32):
            o = _lookup(pyr, coords[t % 8], 4, torch.float32)
        return o
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        lk(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=side):
            keep = lk()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): gr.replay()
    b.record(); torch.cuda.synchronize()
    print(f"W2={W2:4d} volume {pyr.buffer.numel()*4/1e6:7.1f} MB: {a.elapsed_time(b)*1e3/20/32:.2f} us per lookup "
          f"(env pair={os.environ.get('SMOE_CORR_FORCE_PAIR_LOOKUP')}, nopair={os.environ.get('SMOE_CORR_NO_PAIR_LOOKUP')})")
