import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np
import smoe_stereo_b200 as sb
from smoe_stereo_b200 import _lib
from smoe_stereo_b200.corr import _batchable, _batched_lookup_backward, _stream_ptr
from gpu_util import make_coords
dev = torch.device("cuda", 0)
B, H, W, T = 2, 80, 184, 5
gen = torch.Generator(device="cpu").manual_seed(T)
coords = [torch.from_numpy(make_coords(B, H, W, W, 1, seed=100 + t)[0]).to(dev) for t in range(T)]
gouts = [torch.randn(B, 36, H, W, generator=gen).to(dev) for _ in range(T)]
pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev)
gp = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev, zero=True)
for c, g in zip(coords, gouts):
    _lib.check(_lib.lib().smoe_corr_lookup_bwd(gp.ref, c.data_ptr(), c.stride(0), 1.0, 4, g.data_ptr(), 0, 1, _stream_ptr(dev)), "b")
lv = [gp.level(i).clone() for i in range(4)]
_lib.check(_lib.lib().smoe_corr_fold_pyramid_grad(gp.ref, _stream_ptr(dev)), "fold")
g0 = _batched_lookup_backward(pyr, [_batchable(pyr, c, g, 4) for c, g in zip(coords, gouts)], 4, None)
a, b = g0.level(0), gp.level(0)
d = (a - b).abs()
print("max abs diff", d.max().item(), "max |b|", b.abs().max().item(), "num diff", int((d > 0).sum()), "of", d.numel())
idx = torch.nonzero(d > 0)[:8]
print(idx.tolist())
for i in idx[:4]:
    bb, hh, xx, jj = i.tolist()
    print(jj, a[bb, hh, xx, jj].item(), b[bb, hh, xx, jj].item(), [lv[k][bb, hh, xx, jj >> k].item() for k in range(4)])
print("---- per-T check")
for TT in (1, 2, 3):
    gp2 = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev, zero=True)
    for c, g in list(zip(coords, gouts))[:TT]:
        _lib.check(_lib.lib().smoe_corr_lookup_bwd(gp2.ref, c.data_ptr(), c.stride(0), 1.0, 4, g.data_ptr(), 0, 1, _stream_ptr(dev)), "b")
    _lib.check(_lib.lib().smoe_corr_fold_pyramid_grad(gp2.ref, _stream_ptr(dev)), "fold")
    g02 = _batched_lookup_backward(pyr, [_batchable(pyr, c, g, 4) for c, g in list(zip(coords, gouts))[:TT]], 4, None)
    d2 = (g02.level(0) - gp2.level(0)).abs()
    print("T", TT, "num diff", int((d2 > 0).sum()), "max", d2.max().item())
