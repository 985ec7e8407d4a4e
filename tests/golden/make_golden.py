"""Generate golden vectors from the UNMODIFIED reference implementation.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports the reference's own ``core/corr.py`` (CorrBlock1D and
PytorchAlternateCorrBlock1D, PyTorch CPU fp32), feeds it small seeded inputs and stores
inputs + outputs under tests/golden/*.npz.  The GPU box has no /root/reference, so the
fixtures are committed; this script is committed so they can be regenerated.

Stored per case:
  f1,f2           inputs (B,C,H,W1),(B,C,H,W2)
  coords          (T,B,2,H,W1) lookup coordinates (out-of-range, integer and fractional)
  vol             CorrBlock1D.corr output squeezed to (B,H,W1,W2)        core/corr.py:148-156
  lvl{i}          pyramid level i as (B,H,W1,W2_i)                         core/corr.py:119-125
  out             (T,B,L*(2r+1),H,W1) CorrBlock1D.__call__                 core/corr.py:127-146
  out_alt         (B,L*(2r+1),H,W1) PytorchAlternateCorrBlock1D (t=0)      core/corr.py:64-107
  gout            (T,B,L*(2r+1),H,W1) upstream gradient used for backward
  glvl{i}         d(sum_t <out_t,gout_t>)/d(level i)  (B,H,W1,W2_i)        autograd of grid_sample
  df1,df2         gradients w.r.t. the feature maps                        autograd of the whole path

convc1_<case>.npz (the consumer of the lookup), from the reference's
``core/update.py`` BasicMotionEncoder with seeded random weights:
  weight,bias     convc1 parameters (64, L*(2r+1)), (64,)                  core/update.py:70
  cor             (T,B,64,H,W1) = F.relu(convc1(out[t]))                   core/update.py:77
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("SMOE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
from core.corr import CorrBlock1D, PytorchAlternateCorrBlock1D  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (B, C, H, W1, W2, levels, radius, T, fp16_valued)
    "odd37": (2, 16, 3, 37, 37, 4, 4, 3, False),
    "c256_w40": (1, 256, 2, 40, 40, 4, 4, 2, False),
    "w1_ne_w2": (1, 32, 2, 24, 40, 4, 4, 2, False),
    "fp16_valued_w64": (1, 64, 2, 64, 64, 4, 4, 2, True),
    "r3_l3": (1, 8, 2, 20, 20, 3, 3, 2, False),
}


def make_inputs(name, B, C, H, W1, W2, L, r, T, fp16_valued):
    rs = np.random.RandomState(sum(map(ord, name)))
    f1 = rs.standard_normal((B, C, H, W1)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W2)).astype(np.float32)
    if fp16_valued:
        f1 = f1.astype(np.float16).astype(np.float32)
        f2 = f2.astype(np.float16).astype(np.float32)
    xs = np.arange(W1, dtype=np.float32)[None, None, :].repeat(B, 0).repeat(H, 1)
    ys = np.arange(H, dtype=np.float32)[None, :, None].repeat(B, 0).repeat(W1, 2)
    coords = np.zeros((T, B, 2, H, W1), dtype=np.float32)
    for t in range(T):
        disp = rs.uniform(-6.0, W2 * 0.6, size=(B, H, W1)).astype(np.float32)
        x = xs - disp
        # exact integers, half-integers, far out-of-range on both sides
        x[:, :, 0] = np.float32(3.0)
        x[:, :, 1] = np.float32(W2 - 1)
        x[:, :, 2] = np.float32(-0.5)
        x[:, :, 3] = np.float32(W2 + 7.25)
        x[:, :, 4] = np.float32(-13.75)
        x[:, :, 5] = np.float32(W2 - 0.25)
        coords[t, :, 0] = x
        coords[t, :, 1] = ys
    gout = rs.standard_normal((T, B, L * (2 * r + 1), H, W1)).astype(np.float32)
    return f1, f2, coords, gout


def run_case(name, spec):
    B, C, H, W1, W2, L, r, T, fp16_valued = spec
    f1, f2, coords, gout = make_inputs(name, *spec)
    tf1 = torch.from_numpy(f1).requires_grad_(True)
    tf2 = torch.from_numpy(f2).requires_grad_(True)
    blk = CorrBlock1D(tf1, tf2, num_levels=L, radius=r)
    data = {"f1": f1, "f2": f2, "coords": coords, "gout": gout,
            "meta": np.array([B, C, H, W1, W2, L, r, T], dtype=np.int64)}
    data["vol"] = CorrBlock1D.corr(tf1, tf2).detach().numpy().reshape(B, H, W1, W2)
    for i in range(L):
        data[f"lvl{i}"] = blk.corr_pyramid[i].detach().numpy().reshape(B, H, W1, -1)
    outs = []
    loss = 0.0
    for t in range(T):
        o = blk(torch.from_numpy(coords[t]))
        outs.append(o.detach().numpy())
        loss = loss + (o * torch.from_numpy(gout[t])).sum()
    data["out"] = np.stack(outs)
    grads = torch.autograd.grad(loss, [tf1, tf2] + [blk.corr_pyramid[i] for i in range(L)],
                                allow_unused=True)
    data["df1"] = grads[0].numpy()
    data["df2"] = grads[1].numpy()
    # NOTE: grad w.r.t. level i from autograd also contains the part flowing back from
    # the coarser levels through avg_pool2d; isolate the direct lookup part per level.
    for i in range(L):
        lv = blk.corr_pyramid[i].detach().clone().requires_grad_(True)
        sub = CorrBlock1D.__new__(CorrBlock1D)
        sub.num_levels, sub.radius = L, r
        sub.corr_pyramid = [p.detach() for p in blk.corr_pyramid]
        sub.corr_pyramid[i] = lv
        l2 = 0.0
        for t in range(T):
            l2 = l2 + (sub(torch.from_numpy(coords[t])) * torch.from_numpy(gout[t])).sum()
        (g,) = torch.autograd.grad(l2, [lv])
        data[f"glvl{i}"] = g.numpy().reshape(B, H, W1, -1)
    if W1 == W2:
        with torch.no_grad():
            alt = PytorchAlternateCorrBlock1D(torch.from_numpy(f1), torch.from_numpy(f2),
                                              num_levels=L, radius=r)
            data["out_alt"] = alt(torch.from_numpy(coords[0])).numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **data)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_convc1_case(name, spec):
    """Reference motion-encoder first layer on the stored lookup outputs of `name`."""
    from types import SimpleNamespace
    from core.update import BasicMotionEncoder
    B, C, H, W1, W2, L, r, T, _ = spec
    out = np.load(os.path.join(HERE, f"{name}.npz"))["out"]
    torch.manual_seed(1000 + sum(map(ord, name)))
    enc = BasicMotionEncoder(SimpleNamespace(corr_levels=L, corr_radius=r))
    with torch.no_grad():
        # default Conv2d init gives |w| <= 1/6; scale up so that the ReLU clips about half
        enc.convc1.weight.mul_(3.0)
        cor = np.stack([torch.relu(enc.convc1(torch.from_numpy(out[t]))).numpy() for t in range(T)])
    path = os.path.join(HERE, f"convc1_{name}.npz")
    np.savez_compressed(path, weight=enc.convc1.weight.detach().numpy().reshape(64, -1),
                        bias=enc.convc1.bias.detach().numpy(), cor=cor)
    print(f"convc1_{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(4)
    if "--convc1-only" not in sys.argv:
        for n, s in CASES.items():
            run_case(n, s)
    for n in ("odd37", "c256_w40", "fp16_valued_w64"):
        run_convc1_case(n, CASES[n])
