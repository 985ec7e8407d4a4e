"""GPU parity: lookup backward (accumulating fused kernel, dense corr_sampler.backward API),
pyramid-gradient fold, and the full autograd path of CorrBlock1D against the reference's own
autograd results (golden vectors) and the CPU oracle."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_max

pytestmark = pytest.mark.gpu

GRAD_TOL = 5e-5      # same arithmetic in fp32 (reference path goes through grid_sample backward)
E2E_GRAD_TOL = 1e-3  # includes the tf32 volume and cuBLAS einsum backward


@pytest.fixture(scope="module")
def sb():
    import smoe_stereo_b200 as m
    assert torch.cuda.is_available()
    return m


def _bwd_acc(sb, gp, coords, gout, r):
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _stream_ptr
    rc = _lib.lib().smoe_corr_lookup_bwd(gp.ref, coords.data_ptr(), coords.stride(0), 1.0, r,
                                         gout.data_ptr(), _lib.DTYPE_F32, _lib.BWD_ACCUMULATE,
                                         _stream_ptr(gp.device))
    _lib.check(rc, "lookup_bwd")


@pytest.mark.parametrize("name", golden_names())
def test_lookup_bwd_accumulate_and_fold_vs_golden(sb, name):
    from gpu_util import dev
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _stream_ptr
    g = load_golden(name)
    d = g["dims"]
    gp = sb.FusedPyramid(d["B"], d["H"], d["W1"], d["W2"], d["L"], torch.float32, dev(), zero=True)
    for t in range(d["T"]):
        _bwd_acc(sb, gp, torch.from_numpy(g["coords"][t]).to(dev()),
                 torch.from_numpy(g["gout"][t]).to(dev()), d["r"])
    for i in range(d["L"]):
        assert rel_max(gp.level(i).cpu().numpy(), g[f"glvl{i}"]) < GRAD_TOL
    # padding columns were never touched
    used = torch.zeros(gp.pitch, dtype=torch.bool)
    for i in range(d["L"]):
        used[gp.offsets[i]:gp.offsets[i] + gp.widths[i]] = True
    assert torch.all(gp.buffer[:, ~used.to(dev())] == 0)
    # fold onto level 0 == avg_pool2d backward chain
    from oracle import corr_oracle
    want = corr_oracle.numpy_fold_pyramid_grad([g[f"glvl{i}"] for i in range(d["L"])])
    _lib.check(_lib.lib().smoe_corr_fold_pyramid_grad(gp.ref, _stream_ptr(gp.device)), "fold")
    assert rel_max(gp.level(0).cpu().numpy(), want) < GRAD_TOL


@pytest.mark.parametrize("name", ["odd37", "c256_w40", "r3_l3"])
def test_corr_sampler_backward_api(sb, oracle, name):
    """Dense single-pass gradient == zeros_like + scatter (sampler_kernel.cu:138-166)."""
    from gpu_util import dev
    g = load_golden(name)
    d = g["dims"]
    rd = 2 * d["r"] + 1
    for i in range(d["L"]):
        vol = torch.from_numpy(np.ascontiguousarray(g[f"lvl{i}"])).to(dev())
        c = np.ascontiguousarray(g["coords"][0][:, :1] / np.float32(2 ** i))
        go = np.ascontiguousarray(g["gout"][0][:, i * rd:(i + 1) * rd])
        want = oracle.c_lookup_level_bwd(c, go, vol.shape[-1], d["r"])
        got, = sb.corr_sampler.backward(vol, torch.from_numpy(c).to(dev()),
                                        torch.from_numpy(go).to(dev()), d["r"])
        assert got.shape == vol.shape and got.dtype == vol.dtype
        assert rel_max(got.cpu().numpy(), want) < 2e-6
        # and through the CorrSampler autograd.Function (core/corr.py:17-29)
        v2 = vol.clone().requires_grad_(True)
        out = sb.CorrSampler.apply(v2, torch.from_numpy(c).to(dev()), d["r"])
        out.backward(torch.from_numpy(go).to(dev()))
        assert rel_max(v2.grad.cpu().numpy(), want) < 2e-6


@pytest.mark.parametrize("name", golden_names())
def test_corrblock_autograd_vs_reference(sb, name):
    """d loss / d fmap through build + T lookups equals the reference's autograd (golden df1/df2)."""
    from gpu_util import dev
    g = load_golden(name)
    d = g["dims"]
    f1 = torch.from_numpy(g["f1"]).to(dev()).requires_grad_(True)
    f2 = torch.from_numpy(g["f2"]).to(dev()).requires_grad_(True)
    blk = sb.CorrBlock1D(f1, f2, num_levels=d["L"], radius=d["r"])
    loss = 0.0
    for t in range(d["T"]):
        out = blk(torch.from_numpy(g["coords"][t]).to(dev()))
        assert out.requires_grad
        loss = loss + (out * torch.from_numpy(g["gout"][t]).to(dev())).sum()
    loss.backward()
    assert rel_max(f1.grad.cpu().numpy(), g["df1"]) < E2E_GRAD_TOL
    assert rel_max(f2.grad.cpu().numpy(), g["df2"]) < E2E_GRAD_TOL
    assert blk._state.grad_pyr is None     # persistent gradient pyramid released after use


def test_bwd_linearity_and_adjoint_full_size(sb):
    """Size-independent properties at BASELINE config-2 shape (B=8, 80x184): the backward is the
    exact adjoint of the forward (<out, gout> == <pyramid, grad_pyramid>) and is linear."""
    from gpu_util import dev, make_coords
    from smoe_stereo_b200.corr import _lookup
    B, H, W = 8, 80, 184
    gen = torch.Generator(device="cpu").manual_seed(1)
    pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev(), zero=True)
    pyr.level(0).copy_(torch.randn(B, H, W, W, generator=gen).to(dev()))
    for i in range(1, 4):        # a consistent pyramid (level i+1 = pair average), as the build writes it
        w = pyr.widths[i]
        prev = pyr.level(i - 1)
        pyr.level(i).copy_((prev[..., 0:2 * w:2] + prev[..., 1:2 * w:2]) * 0.5)
    coords = torch.from_numpy(make_coords(B, H, W, W, 1, seed=3)[0]).to(dev())
    gout = torch.randn(B, 36, H, W, generator=gen).to(dev())
    out = _lookup(pyr, coords, 4, torch.float32)
    gp = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev(), zero=True)
    _bwd_acc(sb, gp, coords, gout, 4)
    lhs = (out.double() * gout.double()).sum().item()
    rhs = sum((pyr.level(i).double() * gp.level(i).double()).sum().item() for i in range(4))
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-3
    gp2 = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev(), zero=True)
    _bwd_acc(sb, gp2, coords, 2 * gout, 4)
    _bwd_acc(sb, gp, coords, gout, 4)          # accumulate twice == once with 2x
    assert torch.equal(gp.buffer, gp2.buffer)


@pytest.mark.parametrize("shape", [
    (2, 256, 5, 184, 184),     # BASELINE config 2 width (training crop 320x736)
    (1, 256, 3, 240, 240),     # config 1 width
    (1, 100, 2, 72, 136),      # C not a multiple of 16, W1 != W2, K tails
    (1, 300, 2, 40, 312),      # C > 256 -> two N tiles; W2 > 256 -> three M tiles in GEMM 1
    (1, 64, 3, 37, 50),        # W % 4 != 0: pitch-padded operands (no torch.einsum fallback any more)
    (1, 48, 2, 30, 750),       # BASELINE config 4 width (W/4 = 750)
])
def test_native_build_backward_vs_oracle(sb, oracle, shape):
    """smoe_corr_build_bwd (tcgen05 tf32) == autograd of the einsum (core/corr.py:154-156)."""
    from gpu_util import dev
    from smoe_stereo_b200.corr import _build_backward
    B, C, H, W1, W2 = shape
    rs = np.random.RandomState(C + W1)
    f1 = rs.standard_normal((B, C, H, W1)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W2)).astype(np.float32)
    G = rs.standard_normal((B, H, W1, W2)).astype(np.float32)
    want1, want2 = oracle.c_corr_volume_bwd(f1, f2, G)
    gp = sb.FusedPyramid(B, H, W1, W2, 4, torch.float32, dev())
    gp.buffer.fill_(float("nan"))                    # padding / coarser levels must never be read
    gp.level(0).copy_(torch.from_numpy(G).to(dev()))
    df1, df2 = _build_backward(gp, torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()))
    e1, e2 = rel_max(df1.cpu().numpy(), want1), rel_max(df2.cpu().numpy(), want2)
    print(f"[build_bwd {shape}] rel_max d fmap1 {e1:.2e}, d fmap2 {e2:.2e}")
    assert e1 < E2E_GRAD_TOL and e2 < E2E_GRAD_TOL


@pytest.mark.parametrize("shape,T,gdtype", [((8, 80, 184), 22, torch.float32),     # BASELINE config 2, full size
                                            ((2, 80, 184), 5, torch.float32), ((1, 135, 240), 3, torch.float32),
                                            ((1, 6, 52), 50, torch.float32), ((2, 8, 40), 4, torch.float16),
                                            ((2, 5, 37), 3, torch.float32),     # H*W1 % 4 != 0, odd widths
                                            ((1, 3, 21), 2, torch.float16)])    # ragged 32-pixel tiles, fp16 grads
def test_batched_backward_equals_accumulate_then_fold(sb, shape, T, gdtype):
    """smoe_corr_lookup_bwd_batched (all iterations of a step in one pass, folded in shared memory)
    == T x smoe_corr_lookup_bwd(ACCUMULATE) + smoe_corr_fold_pyramid_grad, bit for bit."""
    from gpu_util import dev, make_coords
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _batchable, _batched_lookup_backward, _stream_ptr
    B, H, W = shape
    gen = torch.Generator(device="cpu").manual_seed(T)
    coords = [torch.from_numpy(make_coords(B, H, W, W, 1, seed=100 + t)[0]).to(dev()) for t in range(T)]
    gouts = [torch.randn(B, 36, H, W, generator=gen).to(dev()).to(gdtype) for _ in range(T)]
    pyr = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev())
    # per-call path
    gp = sb.FusedPyramid(B, H, W, W, 4, torch.float32, dev(), zero=True)
    for c, g in zip(coords, gouts):
        rc = _lib.lib().smoe_corr_lookup_bwd(gp.ref, c.data_ptr(), c.stride(0), 1.0, 4, g.data_ptr(),
                                             _lib.DTYPE_F16 if gdtype == torch.float16 else _lib.DTYPE_F32,
                                             _lib.BWD_ACCUMULATE, _stream_ptr(dev()))
        _lib.check(rc, "bwd")
    _lib.check(_lib.lib().smoe_corr_fold_pyramid_grad(gp.ref, _stream_ptr(dev())), "fold")
    # batched path (T = 50 exercises the chunking at 48 iterations and the accumulate flag)
    items = [_batchable(pyr, c, g, 4) for c, g in zip(coords, gouts)]
    assert all(it is not None for it in items)
    g0 = _batched_lookup_backward(pyr, items, 4, None)
    assert g0.num_levels == 1
    a, b = g0.level(0), gp.level(0)
    if T <= 48:
        assert torch.equal(a, b)
    else:                       # two chunks are folded separately: same math, different rounding
        assert rel_max(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
    # every batched kernel of the cross-check library gives the same bits as the product's:
    # [pixel][column] tile (lvl), [column][pixel] tile (cp), cp.async-staged first version
    from gpu_util import xcheck_lib
    for tuning in (dict(bwd_kernel=1), dict(bwd_kernel=2), dict(bwd_kernel=3), dict(bwd_kernel=4), dict(bwd_staged=1)):
        ges = 2 if gdtype == torch.float16 else 4
        if "bwd_staged" in tuning and ((H * W * ges) % 16 or (2 * H * W) % 4):
            continue            # the staged kernel needs 16-byte-multiple gradient / coordinate planes
        with xcheck_lib(**tuning):
            other = _batched_lookup_backward(pyr, items, 4, None)
        assert torch.equal(other.level(0), a), tuning


@pytest.mark.parametrize("levels,W", [(1, 40), (2, 37), (3, 50), (4, 9)])
def test_batched_kernels_agree_for_fewer_levels_and_accumulate(sb, levels, W):
    """1-4 levels, odd / tiny widths, and accumulate=1 on top of an existing level-0 gradient."""
    from gpu_util import dev, make_coords, xcheck_lib
    from smoe_stereo_b200.corr import _batchable, _batched_lookup_backward
    B, H, T = 2, 5, 3
    gen = torch.Generator(device="cpu").manual_seed(W)
    coords = [torch.from_numpy(make_coords(B, H, W, W, 1, seed=7 + t)[0]).to(dev()) for t in range(T)]
    gouts = [torch.randn(B, 9 * levels, H, W, generator=gen).to(dev()) for _ in range(T)]
    pyr = sb.FusedPyramid(B, H, W, W, levels, torch.float32, dev())
    items = [_batchable(pyr, c, g, 4) for c, g in zip(coords, gouts)]
    res = []
    for which in (1, 2, 3, 4):
        with xcheck_lib(bwd_kernel=which):
            g0 = _batched_lookup_backward(pyr, items[:2], 4, None)
            g0 = _batched_lookup_backward(pyr, items[2:], 4, g0)          # accumulate = 1
        res.append(g0.level(0).clone())
    assert all(torch.equal(res[0], r) for r in res[1:])
    assert torch.isfinite(res[0]).all() and res[0].abs().sum() > 0


def test_batched_and_per_call_paths_mix(sb):
    """A step whose lookups are partly batchable (radius 4 ...) and partly not still sums correctly:
    force the mix by marking every other lookup non-batchable."""
    from gpu_util import dev, make_coords
    import smoe_stereo_b200.corr as corr_mod
    B, C, H, W = 1, 32, 4, 48
    g = torch.Generator(device="cpu").manual_seed(0)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev())
    f2 = torch.randn(B, C, H, W, generator=g).to(dev())
    coords = [torch.from_numpy(make_coords(B, H, W, W, 1, seed=t)[0]).to(dev()) for t in range(4)]
    gouts = [torch.randn(B, 36, H, W, generator=g).to(dev()) for _ in range(4)]

    def run(flaky):
        a = f1.clone().requires_grad_(True)
        b = f2.clone().requires_grad_(True)
        blk = sb.CorrBlock1D(a, b)
        loss = sum((blk(c) * go).sum() for c, go in zip(coords, gouts))
        orig = corr_mod._batchable
        calls = {"n": 0}

        def sometimes(*args):
            calls["n"] += 1
            return None if (flaky and calls["n"] % 2) else orig(*args)
        corr_mod._batchable = sometimes
        try:
            loss.backward()
        finally:
            corr_mod._batchable = orig
        return a.grad, b.grad
    ga, gb = run(False)
    ma, mb = run(True)
    assert rel_max(ma.cpu().numpy(), ga.cpu().numpy()) < 1e-5
    assert rel_max(mb.cpu().numpy(), gb.cpu().numpy()) < 1e-5
