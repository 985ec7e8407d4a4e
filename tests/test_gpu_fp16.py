"""GPU parity for the fp16 ("reg_cuda" under autocast) path: fp16 feature maps -> fp16 pyramid ->
fp16 lookups (core/SMoEStereo.py:217-218, core/corr.py:37-51), checked against a numpy emulation
of the reference's half-precision rounding order."""
import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu


def rh(x):
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def emu_pyramid(f1h, f2h, L):
    C = f1h.shape[1]
    vol = np.einsum("bchk,bchw->bhkw", f1h.astype(np.float64), f2h.astype(np.float64))
    lv = [rh(rh(vol) / np.float32(np.sqrt(C)))]       # einsum output is fp16, then / sqrt(D)
    for _ in range(L - 1):
        w = lv[-1].shape[-1] // 2
        lv.append(rh((lv[-1][..., 0:2 * w:2] + lv[-1][..., 1:2 * w:2]) * np.float32(0.5)))
    return lv


def emu_lookup_level(vol, x, r):
    """sampler_kernel.cu:46-59 with scalar_t = c10::Half (fp32 op, round to half after each)."""
    B, H, W1, W2 = vol.shape
    fl = np.floor(x)
    dx = (x - fl).astype(np.float32)
    base = fl.astype(np.int64) - r
    w1, w0 = rh(np.float32(1) - dx), rh(dx)
    out = np.zeros((B, 2 * r + 1, H, W1), dtype=np.float32)

    def tap(idx):
        inb = (idx >= 0) & (idx < W2)
        v = np.take_along_axis(vol, np.clip(idx, 0, W2 - 1)[..., None], axis=-1)[..., 0]
        return np.where(inb, v, np.float32(0))
    for k in range(2 * r + 1):
        a = rh(tap(base + k) * w1)
        b = rh(tap(base + k + 1) * w0)
        out[:, k] = rh(a + b)
    return out


@pytest.fixture(scope="module")
def sb():
    import smoe_stereo_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.mark.parametrize("shape", [(1, 256, 6, 240, 240), (2, 64, 3, 72, 200), (1, 32, 2, 37, 37)])
def test_fast_block_fp16(sb, shape):
    from gpu_util import dev, make_coords
    B, C, H, W1, W2 = shape
    rs = np.random.RandomState(W1)
    f1 = rs.standard_normal((B, C, H, W1)).astype(np.float16)
    f2 = rs.standard_normal((B, C, H, W2)).astype(np.float16)
    blk = sb.CorrBlockFast1D(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()))
    got_lv = [lv.squeeze(3) for lv in blk.corr_pyramid]
    assert all(lv.dtype == torch.float16 for lv in got_lv)
    want_lv = emu_pyramid(f1.astype(np.float32), f2.astype(np.float32), 4)
    for i in range(4):
        g = got_lv[i].float().cpu().numpy()
        # one half-ulp of slack: the reference rounds the fp32-accumulated einsum to half first
        assert rel_max(g, want_lv[i]) < 1e-3
        frac_equal = float(np.mean(g == want_lv[i]))
        print(f"[fp16 build {shape}] level {i}: bit-equal fraction {frac_equal:.4f}")
        assert frac_equal > 0.98
    # pooling from the GPU's own rounded level is bit-exact
    for i in range(1, 4):
        prev = got_lv[i - 1].float().cpu().numpy()
        w = prev.shape[-1] // 2
        np.testing.assert_array_equal(got_lv[i].float().cpu().numpy(),
                                      rh((prev[..., 0:2 * w:2] + prev[..., 1:2 * w:2]) * np.float32(0.5)))
    coords = make_coords(B, H, W1, W2, 2, seed=W2)
    lv_np = [lv.float().cpu().numpy() for lv in got_lv]
    for t in range(2):
        out = blk(torch.from_numpy(coords[t]).to(dev()))
        assert out.dtype == torch.float16 and tuple(out.shape) == (B, 36, H, W1)
        want = np.concatenate([emu_lookup_level(lv_np[i], coords[t][:, 0] / np.float32(2 ** i), 4)
                               for i in range(4)], axis=1)
        np.testing.assert_array_equal(out.float().cpu().numpy(), want)


def test_corr_sampler_fp16_forward_backward(sb):
    from gpu_util import dev
    rs = np.random.RandomState(1)
    B, H, W1, W2, r = 2, 3, 20, 40, 4
    vol = rs.standard_normal((B, H, W1, W2)).astype(np.float16)
    x = rs.uniform(-8, W2 + 8, size=(B, 1, H, W1)).astype(np.float32)
    got, = sb.corr_sampler.forward(torch.from_numpy(vol).to(dev()), torch.from_numpy(x).to(dev()), r)
    np.testing.assert_array_equal(got.float().cpu().numpy(),
                                  emu_lookup_level(vol.astype(np.float32), x[:, 0], r))
    gout = rs.standard_normal((B, 2 * r + 1, H, W1)).astype(np.float16)
    gv, = sb.corr_sampler.backward(torch.from_numpy(vol).to(dev()), torch.from_numpy(x).to(dev()),
                                   torch.from_numpy(gout).to(dev()), r)
    assert gv.dtype == torch.float16 and gv.shape == (B, H, W1, W2)
    # emulate sampler_kernel.cu:92-102 in half arithmetic
    fl = np.floor(x[:, 0]); dx = (x[:, 0] - fl).astype(np.float32); base = fl.astype(np.int64) - r
    w0, w1 = rh(dx), rh(np.float32(1) - dx)
    want = np.zeros((B, H, W1, W2), dtype=np.float32)
    g32 = gout.astype(np.float32)
    for t in range(2 * r + 2):
        g = np.zeros((B, H, W1), dtype=np.float32)
        if t > 0:
            g = rh(g + rh(g32[:, t - 1] * w0))
        if t < 2 * r + 1:
            g = rh(g + rh(g32[:, t] * w1))
        idx = base + t
        inb = (idx >= 0) & (idx < W2)
        bi, hi, wi = np.nonzero(inb)
        want[bi, hi, wi, idx[inb]] = g[inb]
    np.testing.assert_array_equal(gv.float().cpu().numpy(), want)


def test_reg_block_accepts_fp16_and_returns_fp32(sb):
    from gpu_util import dev
    rs = np.random.RandomState(2)
    f1 = torch.from_numpy(rs.standard_normal((1, 64, 2, 64)).astype(np.float16)).to(dev())
    f2 = torch.from_numpy(rs.standard_normal((1, 64, 2, 64)).astype(np.float16)).to(dev())
    c = torch.zeros(1, 2, 2, 64, device=dev())
    c[:, 0] = torch.arange(64, device=dev()).float() - 3.3
    out = sb.CorrBlock1D(f1, f2)(c)
    assert out.dtype == torch.float32            # core/corr.py:146 .float()
    ref = sb.CorrBlock1D(f1.float(), f2.float())(c)
    assert rel_max(out.cpu().numpy(), ref.cpu().numpy()) < 2e-3
