"""GPU parity: tcgen05/TMA volume + fused pyramid build against the reference's golden vectors
and the CPU oracle.  Tolerance: north_star's <=1e-3 relative (max|d|/max|ref|) for the
tensor-core path; measured values are printed so the run log records the actual error."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu

VOL_TOL = 1e-3          # north_star tolerance for the tf32 tensor-core path
EXACT_TOL = 3e-6        # fp16-valued inputs are exact in tf32/f16 operands; fp32 accumulate only


@pytest.fixture(scope="module")
def sb():
    import smoe_stereo_b200 as m
    assert torch.cuda.is_available()
    return m


def _levels(blk):
    return [lv.squeeze(3).float().cpu().numpy() for lv in blk.corr_pyramid]


@pytest.mark.parametrize("name", golden_names())
def test_build_vs_reference_golden(sb, name):
    from gpu_util import dev
    g = load_golden(name)
    d = g["dims"]
    f1 = torch.from_numpy(g["f1"]).to(dev())
    f2 = torch.from_numpy(g["f2"]).to(dev())
    blk = sb.CorrBlock1D(f1, f2, num_levels=d["L"], radius=d["r"])
    lv = _levels(blk)
    tol = EXACT_TOL if name.startswith("fp16_valued") else VOL_TOL
    for i in range(d["L"]):
        assert lv[i].shape == g[f"lvl{i}"].shape
        e = rel_max(lv[i], g[f"lvl{i}"])
        print(f"[build {name}] level {i}: rel_max={e:.3e} rel_l2={rel_l2(lv[i], g[f'lvl{i}']):.3e}")
        assert e < tol
    # pooling is exact given level 0: (a+b)/2 in fp32, floor-halving of odd widths
    for i in range(1, d["L"]):
        w = lv[i].shape[-1]
        np.testing.assert_array_equal(lv[i], (lv[i - 1][..., 0:2 * w:2] + lv[i - 1][..., 1:2 * w:2]) * np.float32(0.5))
    # and the lookups on top of it agree with the reference end to end
    for t in range(d["T"]):
        out = blk(torch.from_numpy(g["coords"][t]).to(dev()))
        assert out.dtype == torch.float32
        assert rel_max(out.cpu().numpy(), g["out"][t]) < tol
    v = sb.CorrBlock1D.corr(f1, f2)
    assert tuple(v.shape) == (d["B"], d["H"], d["W1"], 1, d["W2"])
    assert rel_max(v.squeeze(3).cpu().numpy(), g["vol"]) < tol


@pytest.mark.parametrize("shape", [
    (1, 256, 135, 240, 240),    # BASELINE config 1 (540x960 at 1/4 res)
    (2, 256, 24, 184, 184),     # config 2 width (320x736), 2 M tiles with a 56-row tail
    (1, 256, 8, 312, 312),      # config 3 width (384x1248): 3 M tiles, 2 N tiles
    (1, 256, 3, 750, 750),      # config 4 width (W%4 != 0 -> padded staging), 6 M x 3 N tiles
    (1, 100, 2, 72, 520),       # C not a multiple of the K stage, W1 != W2, 3 N tiles
    (2, 64, 4, 37, 52),         # padded staging for fmap1 only
])
def test_build_shapes_vs_oracle(sb, oracle, shape):
    from gpu_util import dev
    B, C, H, W1, W2 = shape
    rs = np.random.RandomState(B * 1000 + W1)
    f1 = rs.standard_normal((B, C, H, W1)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W2)).astype(np.float32)
    want = oracle.c_pyramid(oracle.c_corr_volume(f1, f2), 4)
    blk = sb.CorrBlock1D(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()))
    lv = _levels(blk)
    for i in range(4):
        e = rel_max(lv[i], want[i])
        print(f"[build {shape}] level {i}: rel_max={e:.3e} rel_l2={rel_l2(lv[i], want[i]):.3e}")
        assert e < VOL_TOL


def test_build_fp16_valued_inputs_are_exact(sb, oracle):
    """Default mixed_precision=True feeds fp16-valued feature maps up-cast to fp32
    (core/SMoEStereo.py:168,213): tf32 operands then carry them exactly."""
    from gpu_util import dev
    rs = np.random.RandomState(5)
    f1 = rs.standard_normal((1, 256, 4, 240)).astype(np.float16).astype(np.float32)
    f2 = rs.standard_normal((1, 256, 4, 240)).astype(np.float16).astype(np.float32)
    want = oracle.numpy_corr_volume(f1, f2)
    blk = sb.CorrBlock1D(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()))
    assert rel_max(_levels(blk)[0], want) < EXACT_TOL


def test_tf32_modes_error_report(sb, oracle):
    """Record the error of both tf32 operand modes (TMA TFLOAT32 conversion vs MMA truncation)."""
    from gpu_util import dev
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _build_into
    rs = np.random.RandomState(9)
    f1 = rs.standard_normal((1, 256, 8, 240)).astype(np.float32)
    f2 = rs.standard_normal((1, 256, 8, 240)).astype(np.float32)
    want = oracle.numpy_corr_volume(f1, f2)
    t1, t2 = torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev())
    for flags, label in ((0, "tma_tfloat32"), (_lib.BUILD_TF32_TRUNCATE, "mma_truncate")):
        pyr = sb.FusedPyramid(1, 8, 240, 240, 4, torch.float32, dev())
        _build_into(pyr, t1, t2, flags)
        got = pyr.level(0).cpu().numpy()
        bias = float(np.sum(got.astype(np.float64) * want) / np.sum(want.astype(np.float64) ** 2) - 1)
        print(f"[tf32 {label}] rel_max={rel_max(got, want):.3e} rel_l2={rel_l2(got, want):.3e} "
              f"scale_bias={bias:+.3e}")
        assert rel_max(got, want) < VOL_TOL


def test_build_batch_independent_and_deterministic(sb):
    """Each (b,h) unit is independent: building a batch equals building its items one by one,
    bit for bit -- the property batch/row sharding relies on."""
    from gpu_util import dev
    g = torch.Generator(device="cpu").manual_seed(0)
    f1 = torch.randn(3, 64, 5, 96, generator=g).to(dev())
    f2 = torch.randn(3, 64, 5, 96, generator=g).to(dev())
    p_full, p_again = sb.CorrBlock1D(f1, f2)._state.pyr, sb.CorrBlock1D(f1, f2)._state.pyr
    for p_ in (p_full, p_again):
        p_.materialize()                    # levels 1/3 are skipped by the block's build
    full, again = p_full.buffer.clone(), p_again.buffer
    pyr = sb.CorrBlock1D(f1, f2)._state.pyr
    for b in range(3):
        one = sb.CorrBlock1D(f1[b:b + 1].contiguous(), f2[b:b + 1].contiguous())._state.pyr
        for i in range(4):
            assert torch.equal(one.level(i)[0], pyr.level(i)[b])
    for i in range(4):
        lvl = slice(pyr.offsets[i], pyr.offsets[i] + pyr.widths[i])
        assert torch.equal(full[:, lvl], again[:, lvl])
    # row shard == rows of the full build
    top = sb.CorrBlock1D(f1[:, :, :2].contiguous(), f2[:, :, :2].contiguous())._state.pyr
    for i in range(4):
        assert torch.equal(top.level(i), pyr.level(i)[:, :2])


def test_middlebury_full_resolution_single_gpu(sb, oracle):
    """BASELINE config 4: ~2000x3000 -> 500x750 at 1/4 res (W%4 != 0 -> padded staging, 6 M tiles x
    3 N tiles, 2.2 GB pyramid, >2^31-byte offsets).  Checked through size-independent properties:
    rows are independent (== a 2-row build of the same rows), pooling is exact, lookups on the big
    pyramid (level-pair kernel) equal lookups on the 2-row pyramid (4-window kernel), and those
    rows match the oracle."""
    from gpu_util import dev, make_coords
    B, C, H, W = 1, 256, 500, 750
    g = torch.Generator(device="cpu").manual_seed(11)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev())
    f2 = torch.randn(B, C, H, W, generator=g).to(dev())
    blk = sb.CorrBlock1D(f1, f2)
    pyr = blk._state.pyr
    assert pyr.buffer.numel() * 4 > 2 ** 31
    rows = [0, 249, 498]
    for r0 in rows:
        sub = sb.CorrBlock1D(f1[:, :, r0:r0 + 2].contiguous(), f2[:, :, r0:r0 + 2].contiguous())
        for i in range(4):
            assert torch.equal(sub._state.pyr.level(i), pyr.level(i)[:, r0:r0 + 2])
    lv = [pyr.level(i)[:, 498:500] for i in range(4)]
    for i in range(1, 4):
        w = lv[i].shape[-1]
        assert torch.equal(lv[i], (lv[i - 1][..., 0:2 * w:2] + lv[i - 1][..., 1:2 * w:2]) * 0.5)
    want = oracle.c_pyramid(oracle.c_corr_volume(f1[:, :, 498:500].cpu().numpy().copy(),
                                                 f2[:, :, 498:500].cpu().numpy().copy()), 4)
    for i in range(4):
        assert rel_max(lv[i].cpu().numpy(), want[i]) < VOL_TOL
    coords = torch.from_numpy(make_coords(B, H, W, W, 1, seed=4)[0]).to(dev())
    out = blk(coords)
    assert tuple(out.shape) == (1, 36, 500, 750) and torch.isfinite(out).all()
    sub = sb.CorrBlock1D(f1[:, :, 498:500].contiguous(), f2[:, :, 498:500].contiguous())
    assert torch.equal(out[:, :, 498:500], sub(coords[:, :, 498:500].contiguous()))


@pytest.mark.parametrize("shape", [
    (1, 256, 6, 240, 240),      # one 256-row pair tile, N tile padded 240 -> 256
    (2, 96, 5, 312, 312),       # two pair tiles (second CTA of the last pair idle), two N tiles
    (1, 64, 4, 37, 50),         # odd widths -> pitch-padded operands, partial tiles everywhere
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_build_cta_pair_kernel_is_bit_identical(sb, shape, dtype):
    """The CTA-pair kernel (tcgen05 cta_group::2; cross-check library only, Tuning::build_pair)
    accumulates the same products in the same order as the product's one-CTA kernel: every level
    must match bit for bit."""
    from gpu_util import dev, xcheck_lib
    B, C, H, W1, W2 = shape
    g = torch.Generator().manual_seed(11)
    f1 = torch.randn(B, C, H, W1, generator=g).to(dev(), dtype)
    f2 = torch.randn(B, C, H, W2, generator=g).to(dev(), dtype)
    cls = sb.CorrBlock1D if dtype == torch.float32 else sb.CorrBlockFast1D
    ref = [lv.clone() for lv in cls(f1, f2, num_levels=4, radius=4).corr_pyramid]
    with xcheck_lib(build_pair=1):
        got = [lv.clone() for lv in cls(f1, f2, num_levels=4, radius=4).corr_pyramid]
    for i, (a, b) in enumerate(zip(ref, got)):
        assert torch.equal(a, b), f"level {i} differs"


# ------------------------------------------------------------------ channels-last (K-major) inputs
@pytest.mark.parametrize("shape", [
    (1, 256, 135, 240, 240),    # BASELINE config 1
    (2, 256, 6, 184, 184),      # config 2 width: M tail of 56 rows
    (1, 256, 4, 312, 312),      # config 3 width: 2 N tiles
    (1, 256, 3, 750, 750),      # config 4 width: W%4 != 0 needs NO padding pass in this layout
    (1, 100, 2, 72, 520),       # C tail inside the last K stage (zero-filled by TMA), W1 != W2
    (2, 64, 4, 37, 52),         # narrow odd widths, boxes larger than the tensor
    (1, 8, 2, 16, 24),          # a single partial K stage
])
def test_build_channels_last_vs_oracle_and_nchw(sb, oracle, shape):
    """channels-last feature maps are read in place as K-major operands.
    fp32: same tf32 products, same fp32 accumulation order over C as the NCHW path -> the two
    layouts must agree (bit for bit in practice; 1e-6 asserted); both are held to the oracle at north_star's tolerance."""
    from gpu_util import dev
    B, C, H, W1, W2 = shape
    rs = np.random.RandomState(B * 1000 + W1 + 7)
    f1 = rs.standard_normal((B, C, H, W1)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W2)).astype(np.float32)
    want = oracle.c_pyramid(oracle.c_corr_volume(f1, f2), 4)
    t1, t2 = torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev())
    c1 = t1.contiguous(memory_format=torch.channels_last)
    c2 = t2.contiguous(memory_format=torch.channels_last)
    assert not c1.is_contiguous() and c1.stride(1) == 1
    blk = sb.CorrBlock1D(c1, c2)
    lv = _levels(blk)
    ref = _levels(sb.CorrBlock1D(t1, t2))
    for i in range(4):
        e = rel_max(lv[i], want[i])
        print(f"[build channels_last {shape}] level {i}: rel_max={e:.3e}")
        assert e < VOL_TOL
        print(f"   vs NCHW build: bit-identical={np.array_equal(lv[i], ref[i])} rel_max={rel_max(lv[i], ref[i]):.2e}")
        assert rel_max(lv[i], ref[i]) < 1e-6
    # mixed layouts: the NCHW map is converted, not misread
    mixed = _levels(sb.CorrBlock1D(c1, t2))
    np.testing.assert_array_equal(mixed[0], lv[0])


@pytest.mark.parametrize("fast", [False, True])
def test_build_channels_last_fp16(sb, fast):
    """fp16 channels-last feature maps (the autocast producer format of row f4, half the input
    bytes): kind::f16 K-major operands; same result as the NCHW fp16 build."""
    from gpu_util import dev
    g = torch.Generator(device="cpu").manual_seed(3)
    f1 = torch.randn(2, 256, 5, 184, generator=g).half().to(dev())
    f2 = torch.randn(2, 256, 5, 184, generator=g).half().to(dev())
    cls = sb.CorrBlockFast1D if fast else sb.CorrBlock1D
    a = cls(f1, f2)
    b = cls(f1.contiguous(memory_format=torch.channels_last), f2.contiguous(memory_format=torch.channels_last))
    assert a._state.pyr.dtype == b._state.pyr.dtype == torch.float16
    for la, lb in zip(a.corr_pyramid, b.corr_pyramid):
        print(f"[fp16 channels_last fast={fast}] bit-identical to NCHW: {torch.equal(la, lb)}")
        assert rel_max(lb.float().cpu().numpy(), la.float().cpu().numpy()) < 1e-3
    want = torch.einsum("bchw,bchv->bhwv", f1.float(), f2.float()) / 16.0
    assert rel_max(b.corr_pyramid[0].squeeze(3).float().cpu().numpy(), want.cpu().numpy()) < VOL_TOL


def test_build_channels_last_c_abi_rejects_unaligned_channels(sb):
    """C*sizeof(T) % 16 != 0 cannot be a TMA stride: the C ABI says so, the class converts to NCHW."""
    from gpu_util import dev
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _build_into
    g = torch.Generator(device="cpu").manual_seed(4)
    f1 = torch.randn(1, 6, 3, 40, generator=g).to(dev())
    f2 = torch.randn(1, 6, 3, 40, generator=g).to(dev())
    c1 = f1.contiguous(memory_format=torch.channels_last)
    c2 = f2.contiguous(memory_format=torch.channels_last)
    pyr = sb.FusedPyramid(1, 3, 40, 40, 4, torch.float32, dev())
    with pytest.raises(RuntimeError, match="channels-last"):
        _build_into(pyr, c1, c2)
    a = sb.CorrBlock1D(c1, c2)
    b = sb.CorrBlock1D(f1, f2)
    assert torch.equal(a.corr_pyramid[0], b.corr_pyramid[0])


def test_build_channels_last_backward(sb):
    """Training with channels-last feature maps: gradients equal those of the NCHW path."""
    from gpu_util import dev
    g = torch.Generator(device="cpu").manual_seed(6)
    f1 = torch.randn(1, 64, 4, 96, generator=g).to(dev())
    f2 = torch.randn(1, 64, 4, 96, generator=g).to(dev())
    coords = (torch.arange(96, dtype=torch.float32).view(1, 1, 1, 96).expand(1, 2, 4, 96)
              - torch.rand(1, 2, 4, 96, generator=g) * 20).to(dev())
    grads = []
    for cl in (False, True):
        a = f1.clone().requires_grad_(True)
        b = f2.clone().requires_grad_(True)
        x, y = (a, b)
        if cl:
            x = a.contiguous(memory_format=torch.channels_last)
            y = b.contiguous(memory_format=torch.channels_last)
        blk = sb.CorrBlock1D(x, y)
        (blk(coords).square().sum() + blk(coords + 0.3).sum()).backward()
        grads.append((a.grad.clone(), b.grad.clone()))
    for k in range(2):
        assert rel_max(grads[1][k].cpu().numpy(), grads[0][k].cpu().numpy()) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(1, 64, 6, 240, 240), (2, 32, 3, 77, 85), (1, 32, 2, 40, 9)])
def test_build_skip_odd_levels_equals_full_build(sb, shape, dtype):
    """SMOE_BUILD_SKIP_ODD_LEVELS: levels 0 and 2 are bit-identical to a full build, levels 1 and 3
    are left untouched by the kernel, and FusedPyramid.materialize() reproduces them bit for bit."""
    from gpu_util import dev
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _build_into
    B, C, H, W1, W2 = shape
    g = torch.Generator(device="cpu").manual_seed(W2)
    f1 = torch.randn(B, C, H, W1, generator=g).to(dev()).to(dtype)
    f2 = torch.randn(B, C, H, W2, generator=g).to(dev()).to(dtype)
    full = sb.FusedPyramid(B, H, W1, W2, 4, dtype, dev())
    _build_into(full, f1, f2)
    part = sb.FusedPyramid(B, H, W1, W2, 4, dtype, dev())
    part.buffer.fill_(12345.0)
    _build_into(part, f1, f2, _lib.BUILD_SKIP_ODD_LEVELS)
    assert part.c.levels_missing == 0b1010
    for i in (0, 2):
        assert torch.equal(part._level_view(i), full._level_view(i))
    for i in (1, 3):
        assert bool((part._level_view(i) == 12345.0).all())          # never written
    part.materialize()
    assert part.c.levels_missing == 0
    for i in range(4):
        assert torch.equal(part.level(i), full.level(i))
    # the blocks use the flag for radius 4 (unless SMOE_CORR_ALL_LEVELS=1) and build everything otherwise
    import smoe_stereo_b200.corr as corr_mod
    if corr_mod._skip_odd_levels is None:
        assert sb.CorrBlock1D(f1, f2)._state.pyr.c.levels_missing == 0b1010
    old = corr_mod._skip_odd_levels
    try:
        corr_mod._skip_odd_levels = False
        assert sb.CorrBlock1D(f1, f2)._state.pyr.c.levels_missing == 0
        corr_mod._skip_odd_levels = True
        assert sb.CorrBlock1D(f1, f2)._state.pyr.c.levels_missing == 0b1010
        assert sb.CorrBlock1D(f1, f2, radius=3)._state.pyr.c.levels_missing == 0
        assert sb.CorrBlock1D(f1, f2, num_levels=2)._state.pyr.c.levels_missing == 0b10
    finally:
        corr_mod._skip_odd_levels = old
