"""Host-side cost of one eager CorrBlock1D.__call__ (the path an unmodified model takes)."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smoe_stereo_b200 as sb
dev = torch.device("cuda", 0)
f1 = torch.randn(1, 256, 135, 240, device=dev); f2 = torch.randn(1, 256, 135, 240, device=dev)
coords = torch.zeros(1, 2, 135, 240, device=dev); coords[:, 0] = torch.arange(240, device=dev).float() - 7.3
blk = sb.CorrBlock1D(f1, f2)
for _ in range(100): blk(coords)
torch.cuda.synchronize()
N = 3000
t0 = time.perf_counter()
for _ in range(N): out = blk(coords)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"eager lookup: host {1e6*(t1-t0)/N:.1f} us/call, incl. drain {1e6*(t2-t0)/N:.1f} us/call")
t0 = time.perf_counter()
for _ in range(300): b2 = sb.CorrBlock1D(f1, f2)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"eager build : host {1e6*(t1-t0)/300:.1f} us/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): out = blk(coords)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(s.getvalue()[:2500])
