"""GPU parity: lookup forward (fused multi-level kernel and the corr_sampler.forward API)
against the golden vectors of the reference and against the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_max

pytestmark = pytest.mark.gpu

LOOKUP_TOL = 5e-5   # reference goes through grid_sample's normalise round trip (utils.py:84-99)
ORACLE_TOL = 2e-6   # same formula, fp32, fma contraction only


@pytest.fixture(scope="module")
def sb():
    import smoe_stereo_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.mark.parametrize("name", golden_names())
def test_fused_lookup_vs_reference_golden(sb, name):
    from gpu_util import dev, fused_from_levels
    from smoe_stereo_b200.corr import _lookup
    g = load_golden(name)
    d = g["dims"]
    pyr = fused_from_levels([g[f"lvl{i}"] for i in range(d["L"])])
    for t in range(d["T"]):
        coords = torch.from_numpy(g["coords"][t]).to(dev())
        out = _lookup(pyr, coords, d["r"], torch.float32)
        assert out.shape == g["out"][t].shape and out.is_contiguous()
        assert rel_max(out.cpu().numpy(), g["out"][t]) < LOOKUP_TOL


@pytest.mark.parametrize("name", ["odd37", "c256_w40", "w1_ne_w2"])
def test_corr_sampler_forward_api(sb, oracle, name):
    """Planar single-level volumes, pre-scaled coords, as CorrBlockFast1D drives the op
    (core/corr.py:44-51).  W2=37 exercises the unaligned scalar kernel, 40 the vector one."""
    from gpu_util import dev
    g = load_golden(name)
    d = g["dims"]
    for i in range(d["L"]):
        vol = np.ascontiguousarray(g[f"lvl{i}"])
        for t in range(d["T"]):
            c = np.ascontiguousarray(g["coords"][t][:, :1] / np.float32(2 ** i))
            want = oracle.c_lookup([vol], c, d["r"])
            got, = sb.corr_sampler.forward(torch.from_numpy(vol).to(dev()),
                                           torch.from_numpy(c).to(dev()), d["r"])
            assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
            assert rel_max(got.cpu().numpy(), want) < ORACLE_TOL


def test_corr_sampler_errors(sb):
    from gpu_util import dev
    vol = torch.zeros(1, 2, 8, 8, device=dev())
    coords = torch.zeros(1, 1, 2, 8, device=dev())
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        sb.corr_sampler.forward(vol.cpu(), coords, 4)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        sb.corr_sampler.forward(vol.permute(0, 1, 3, 2), coords, 4)
    with pytest.raises(RuntimeError):
        sb.corr_sampler.forward(vol, coords[:, :, :1], 4)


def test_lookup_cfg1_full_size_vs_oracle(sb, oracle):
    """BASELINE config 1 pixel grid (B=1,H=135,W=240), 4 levels, radius 4, random pyramid."""
    from gpu_util import dev, fused_from_levels, make_coords
    from smoe_stereo_b200.corr import _lookup
    B, H, W = 1, 135, 240
    rs = np.random.RandomState(7)
    lv0 = rs.standard_normal((B, H, W, W)).astype(np.float32)
    levels = oracle.c_pyramid(lv0, 4)
    pyr = fused_from_levels(levels)
    coords = make_coords(B, H, W, W, 3, seed=11)
    for t in range(3):
        want = oracle.c_lookup(levels, np.ascontiguousarray(coords[t]), 4)
        got = _lookup(pyr, torch.from_numpy(coords[t]).to(dev()), 4, torch.float32)
        assert rel_max(got.cpu().numpy(), want) < ORACLE_TOL
    # single-channel coords are accepted too (reference op takes (B,1,H,W))
    got1 = _lookup(pyr, torch.from_numpy(coords[0][:, :1].copy()).to(dev()), 4, torch.float32)
    assert torch.equal(got1, _lookup(pyr, torch.from_numpy(coords[0]).to(dev()), 4, torch.float32))


def test_lookup_ragged_and_nonfinite(sb, oracle):
    """H*W not a multiple of the 32-pixel tile, B>1, NaN/inf/huge coordinates -> zeros."""
    from gpu_util import dev, fused_from_levels
    from smoe_stereo_b200.corr import _lookup
    B, H, W1, W2 = 3, 5, 13, 48
    rs = np.random.RandomState(3)
    levels = oracle.c_pyramid(rs.standard_normal((B, H, W1, W2)).astype(np.float32), 4)
    pyr = fused_from_levels(levels)
    coords = rs.uniform(-20, W2 + 20, size=(B, 2, H, W1)).astype(np.float32)
    coords[0, 0, 0, :4] = [np.nan, np.inf, -np.inf, 3e9]
    want = oracle.c_lookup(levels, coords, 4)
    got = _lookup(pyr, torch.from_numpy(coords).to(dev()), 4, torch.float32).cpu().numpy()
    assert np.all(np.isfinite(got))
    assert np.all(got[0, :, 0, :4] == 0)
    assert rel_max(got, want) < ORACLE_TOL


def test_lookup_empty(sb):
    from gpu_util import dev
    from smoe_stereo_b200.corr import _lookup
    pyr = sb.FusedPyramid(0, 4, 8, 16, 4, torch.float32, dev())
    out = _lookup(pyr, torch.zeros(0, 2, 4, 8, device=dev()), 4, torch.float32)
    assert tuple(out.shape) == (0, 36, 4, 8)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_pair_kernel_equals_window_kernel_bitwise(sb, oracle, dtype, monkeypatch):
    """Volumes beyond L2 take the level-pair kernel (two segments per pixel, coarse taps recomputed
    on the fly).  On a consistent pyramid it must agree BIT FOR BIT with the 4-window kernel; a
    260 MB volume (B=2, 120x260, fp32) crosses the size switch, the small one does not."""
    from gpu_util import dev, make_coords
    B, C, H, W = 2, 32, 120, 260
    g = torch.Generator(device="cpu").manual_seed(5)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev()).to(dtype)
    f2 = torch.randn(B, C, H, W, generator=g).to(dev()).to(dtype)
    coords = torch.from_numpy(make_coords(B, H, W, W, 1, seed=9)[0]).to(dev())
    cls = sb.CorrBlock1D if dtype == torch.float32 else sb.CorrBlockFast1D
    big = cls(f1, f2)
    assert big._state.pyr.buffer.numel() * big._state.pyr.buffer.element_size() > (96 << 20) or dtype == torch.float16
    out_big = big(coords)
    # same rows through the small-volume (4-window) path: one batch item, first 20 rows
    small = cls(f1[:1, :, :20].contiguous(), f2[:1, :, :20].contiguous())
    out_small = small(coords[:1, :, :20].contiguous())
    assert torch.equal(out_big[:1, :, :20], out_small)
    # and against the oracle for the fp32 case
    if dtype == torch.float32:
        lv = [l.squeeze(3)[:1, :20].cpu().numpy() for l in small.corr_pyramid]
        want = oracle.c_lookup([np.ascontiguousarray(x) for x in lv],
                               np.ascontiguousarray(coords[:1, :, :20].cpu().numpy()), 4)
        assert rel_max(out_small.cpu().numpy(), want) < ORACLE_TOL


@pytest.mark.parametrize("dtype,levels", [(torch.float32, 4), (torch.float16, 4), (torch.float32, 2),
                                          (torch.float32, 3), (torch.float16, 3), (torch.float32, 1)])
def test_thread_per_pixel_kernel_equals_lane_group_kernels_bitwise(sb, dtype, levels):
    """lookup_fwd_px_kernel (one thread per pixel, register shift network) against the four-window
    and level-pair lane-group kernels on the same pyramid: bit for bit, including out-of-range,
    integer and non-finite coordinates, odd widths at every level and rows shorter than a segment."""
    from gpu_util import dev, make_coords, xcheck_lib
    for (B, C, H, W1, W2) in [(2, 32, 9, 77, 77), (1, 16, 5, 40, 312), (1, 16, 3, 33, 22), (2, 16, 4, 50, 9)]:
        g = torch.Generator(device="cpu").manual_seed(W2)
        f1 = torch.randn(B, C, H, W1, generator=g).to(dev()).to(dtype)
        f2 = torch.randn(B, C, H, W2, generator=g).to(dev()).to(dtype)
        cls = sb.CorrBlock1D if dtype == torch.float32 else sb.CorrBlockFast1D
        import smoe_stereo_b200.corr as corr_mod
        old_skip = corr_mod._skip_odd_levels
        try:
            corr_mod._skip_odd_levels = True      # as for volumes beyond L2: levels 1/3 are not written
            blk = cls(f1, f2, num_levels=levels)
        finally:
            corr_mod._skip_odd_levels = old_skip
        coords = torch.from_numpy(make_coords(B, H, W1, W2, 3, seed=W1)).to(dev())
        coords[0, 0, 0, 0, :6] = torch.tensor([float("nan"), float("inf"), -float("inf"), 3e9, -0.0, W2 - 1.0])
        coords[1, 0, 0, 1, :4] = torch.tensor([-4.0, -4.5, W2 + 3.999, W2 + 4.0])
        product = [blk(coords[t]).clone() for t in range(3)]      # libsmoecorr.so, levels 1/3 never written
        outs = {}
        with xcheck_lib() as L:
            if levels > 1:
                # the block's build skipped levels 1/3: the lane-group kernels must refuse such a pyramid
                L.smoe_corr_debug_set_lookup_kernel(3)
                with pytest.raises(RuntimeError, match="were not built"):
                    blk(coords[0])
                L.smoe_corr_debug_set_lookup_kernel(-1)
            blk._state.pyr.materialize()             # complete the pyramid for the lane-group kernels
            for which, name in ((1, "px"), (2, "pair"), (3, "vec"), (4, "px256")):
                L.smoe_corr_debug_set_lookup_kernel(which)
                outs[name] = [blk(coords[t]).clone() for t in range(3)]
        for t in range(3):
            assert torch.equal(product[t], outs["vec"][t]), (B, C, H, W1, W2, t)   # product == first design
            assert torch.equal(outs["px"][t], outs["vec"][t])
            assert torch.equal(outs["pair"][t], outs["vec"][t])     # (pair falls through to vec for 1/3 levels)
            assert torch.equal(outs["px256"][t], outs["vec"][t])    # 256-bit loads (fp32 volumes; else == px)
            assert torch.isfinite(product[t]).all()


def test_kitti_cfg3_full_size_streaming_lookup_vs_oracle(sb, oracle):
    """BASELINE config 3 at full size (KITTI 384x1248: B=8, 96x312; 583 MB pyramid, far beyond L2):
    the streaming variant of the thread-per-pixel kernel (64-byte fills, one thread per level pair)
    against the oracle's lookup on the very pyramid the build wrote, three build rows against the
    oracle's volume, and the channels-last build against the NCHW build bit for bit."""
    from gpu_util import dev, make_coords
    from conftest import rel_max as rmax
    B, C, H, W = 8, 256, 96, 312
    g = torch.Generator(device="cpu").manual_seed(21)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev())
    f2 = torch.randn(B, C, H, W, generator=g).to(dev())
    blk = sb.CorrBlock1D(f1, f2)
    pyr = blk._state.pyr
    assert pyr.buffer.numel() * 4 > (512 << 20)
    for (b, h) in ((0, 0), (3, 50), (7, 95)):
        want = oracle.c_pyramid(oracle.c_corr_volume(f1[b:b + 1, :, h:h + 1].cpu().numpy().copy(),
                                                     f2[b:b + 1, :, h:h + 1].cpu().numpy().copy()), 4)
        for i in range(4):
            assert rmax(pyr.level(i)[b:b + 1, h:h + 1].cpu().numpy(), want[i]) < 1e-3
    levels = [np.ascontiguousarray(pyr.level(i).cpu().numpy()) for i in range(4)]
    coords = make_coords(B, H, W, W, 2, seed=17)
    for t in range(2):
        got = blk(torch.from_numpy(coords[t]).to(dev())).cpu().numpy()
        want = oracle.c_lookup(levels, np.ascontiguousarray(coords[t]), 4)
        assert rmax(got, want) < ORACLE_TOL
    cl = sb.CorrBlock1D(f1.contiguous(memory_format=torch.channels_last),
                        f2.contiguous(memory_format=torch.channels_last))
    for i in range(4):
        assert torch.equal(cl._state.pyr.level(i), pyr.level(i))
