import torch, time
dev = torch.device("cuda", 0)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
big_h = torch.empty(149_299_200 // 4, dtype=torch.float32).pin_memory()
big_d = torch.empty_like(big_h, device=dev)
print("D2H 149 MB one copy : %.1f GB/s" % (big_h.numel() * 4 / t(lambda: big_h.copy_(big_d, non_blocking=True)) / 1e9))
print("H2D 149 MB one copy : %.1f GB/s" % (big_h.numel() * 4 / t(lambda: big_d.copy_(big_h, non_blocking=True)) / 1e9))
ch = big_h.view(32, -1); cd = big_d.view(32, -1)
def many():
    for i in range(32): ch[i].copy_(cd[i], non_blocking=True)
print("D2H 32 x 4.67 MB    : %.1f GB/s" % (big_h.numel() * 4 / t(many) / 1e9))
s2 = torch.cuda.Stream(dev)
h2 = torch.empty(74_649_600 // 4, dtype=torch.float32).pin_memory(); d2 = torch.empty_like(h2, device=dev)
def both():
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    many()
dt = t(both)
print("bidirectional: D2H %.1f GB/s + H2D %.1f GB/s (step %.2f ms)" % (big_h.numel()*4/dt/1e9, h2.numel()*4/dt/1e9, dt*1e3))
