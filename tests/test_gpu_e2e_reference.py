"""GPU end-to-end parity through the UNMODIFIED reference models.

north_star tolerance 2: "<= 1e-2 px end-point error after 32 GRU iterations with identical
random-init weights".  The reference's pure-Python model tree travels to the GPU box as
baseline/_ref/core (tools/install_ref.sh; git-ignored, not gpurun-ignored).  Each test

  1. builds the reference model with seeded random init (core/raft_stereo.py:86 RAFTStereo,
     core/SMoEStereo.py:87 SMoEStereo -- DAMv2 ViT-L + SMoE PEFT with torch.load stubbed),
  2. runs 32 iterations with the STOCK corr implementation (core/corr.py) on the GPU in fp32
     (allow_tf32 off),
  3. calls smoe_stereo_b200.install() -- no reference file is edited -- and runs the very same
     model object again, now on libsmoecorr.so,
  4. asserts the end-point error between the two final disparity maps.

"reg_cuda" needs the reference's CUDA sampler for the stock arm (oracle/_ref, built by
oracle/build_ref_sampler.py); without it the stock arm of that case falls back to "reg".
"""
import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_models as R  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="baseline/_ref/core not installed (tools/install_ref.sh)")]

EPE_TOL = 1e-2      # px, north_star


def _fp32_everywhere():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False


def _disp(model, left, right, iters):
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(left, right, iters=iters, test_mode=True)["disp_preds"][-1]
    torch.cuda.synchronize()
    return out.float()


def _loaded_native_lib():
    return any("libsmoecorr.so" in line for line in open("/proc/self/maps"))


def _stock_vs_installed(model, impl, left, right, iters=32, stock_impl=None):
    import smoe_stereo_b200 as sb
    import core.corr as ref_corr
    sb.uninstall()
    assert ref_corr.CorrBlock1D.__module__ == "core.corr"
    model.args.corr_implementation = stock_impl or impl
    stock = _disp(model, left, right, iters)
    stock2 = _disp(model, left, right, iters)                  # run-to-run noise floor of the stock arm
    replaced = sb.install()
    import smoe_stereo_b200.corr as our_corr
    calls = {"n": 0, "dtypes": set()}
    real_lookup = our_corr._lookup

    def counting_lookup(pyr, coords, radius, out_dtype, *a, **k):
        calls["n"] += 1
        calls["dtypes"].add((pyr.dtype, out_dtype))
        return real_lookup(pyr, coords, radius, out_dtype, *a, **k)

    our_corr._lookup = counting_lookup
    try:
        assert any(r.endswith(".CorrBlock1D") for r in replaced), replaced
        import core.raft_stereo as rs
        assert rs.CorrBlock1D is sb.CorrBlock1D and rs.CorrBlockFast1D is sb.CorrBlockFast1D
        model.args.corr_implementation = impl
        ours = _disp(model, left, right, iters)
    finally:
        our_corr._lookup = real_lookup
        sb.uninstall()
    assert _loaded_native_lib(), "libsmoecorr.so was not loaded: the drop-in did not run the CUDA path"
    assert calls["n"] >= iters, f"the installed arm made {calls['n']} lookups through this package, expected >= {iters}"
    d = (stock - ours).abs()
    return dict(epe=d.mean().item(), epe_max=d.max().item(), floor=(stock - stock2).abs().mean().item(),
                disp=stock.abs().mean().item(), lookups=calls["n"], dtypes=calls["dtypes"])


@pytest.fixture(scope="module")
def raft_fp32():
    _fp32_everywhere()
    dev = torch.device("cuda", 0)
    model = R.build_model("raft", R.make_args("reg", mixed_precision=False), seed=0, device=dev)
    left, right = R.synthetic_pair(1, 320, 736, max_disp=40.0, seed=1, device=dev)
    return model, left, right


@pytest.mark.parametrize("impl", ["reg", "alt", "reg_cuda"])
def test_raftstereo_32_iterations_epe(raft_fp32, impl):
    """core/raft_stereo.py:177-207 with identical random-init weights, fp32, 320x736 (the cfg2 crop)."""
    model, left, right = raft_fp32
    stock_impl = impl
    if impl == "reg_cuda" and not R.ref_sampler_available():
        stock_impl = "reg"
    r = _stock_vs_installed(model, impl, left, right, 32, stock_impl)
    print(f"\n[e2e RAFTStereo {impl} vs stock {stock_impl}] mean|disp| {r['disp']:.3f} px  EPE {r['epe']:.3e} px "
          f"(max {r['epe_max']:.3e}; stock run-to-run {r['floor']:.1e})")
    assert r["disp"] > 0.05, "model output is ~0 everywhere; the comparison would be vacuous"
    assert r["epe"] <= EPE_TOL


def test_raftstereo_batch2_mixed_precision_reg(raft_fp32):
    """The default pipeline (mixed_precision=True, 'reg'): fp16-valued features up-cast to fp32
    (core/raft_stereo.py:179) -- the tensor-core build is exact on them."""
    model, _, _ = raft_fp32
    dev = torch.device("cuda", 0)
    left, right = R.synthetic_pair(2, 256, 512, max_disp=32.0, seed=2, device=dev)
    model.args.mixed_precision = True
    try:
        r = _stock_vs_installed(model, "reg", left, right, 32)
    finally:
        model.args.mixed_precision = False
    print(f"\n[e2e RAFTStereo reg, autocast fp16 GRU, B=2] mean|disp| {r['disp']:.3f} px  EPE {r['epe']:.3e} px "
          f"(max {r['epe_max']:.3e}; stock run-to-run {r['floor']:.1e})")
    # fp16 GRU arithmetic amplifies 1-ulp differences of the lookup; bound = max(north_star, 3x the
    # stock arm's own run-to-run floor)
    assert r["epe"] <= max(EPE_TOL, 3 * r["floor"])


@pytest.fixture(scope="module")
def smoe_fp32():
    _fp32_everywhere()
    dev = torch.device("cuda", 0)
    model = R.build_model("smoe", R.make_args("reg", mixed_precision=False, model="smoe"), seed=0, device=dev)
    left, right = R.synthetic_pair(1, 256, 512, max_disp=32.0, seed=3, device=dev)
    return model, left, right


@pytest.mark.parametrize("impl", ["reg", "reg_cuda"])
def test_smoestereo_vitl_32_iterations_epe(smoe_fp32, impl):
    """core/SMoEStereo.py:211-248: DAMv2 ViT-L + SMoE PEFT, random init, 32 iterations."""
    model, left, right = smoe_fp32
    stock_impl = impl if (impl != "reg_cuda" or R.ref_sampler_available()) else "reg"
    r = _stock_vs_installed(model, impl, left, right, 32, stock_impl)
    print(f"\n[e2e SMoEStereo vitl {impl} vs stock {stock_impl}] mean|disp| {r['disp']:.3f} px  EPE {r['epe']:.3e} px "
          f"(max {r['epe_max']:.3e}; stock run-to-run {r['floor']:.1e})")
    assert r["epe"] <= EPE_TOL


def test_install_with_output_ring_gives_identical_disparities(raft_fp32):
    """install(output_ring=2): the update loop of the reference consumes `corr` inside the iteration
    (core/raft_stereo.py:205-214), so serving lookups from a two-buffer ring changes nothing."""
    import smoe_stereo_b200 as sb
    model, left, right = raft_fp32
    model.args.corr_implementation = "reg"
    sb.uninstall()
    sb.install()
    try:
        a = _disp(model, left, right, 12)
    finally:
        sb.uninstall()
    sb.install(output_ring=2)
    try:
        import smoe_stereo_b200.corr as c
        assert c._default_output_ring == 2
        b = _disp(model, left, right, 12)
    finally:
        sb.uninstall()
    assert c._default_output_ring == 0
    assert torch.equal(a, b)


def test_channels_last_producer_feeds_the_build_in_place(raft_fp32):
    """SURVEY.md 8f row f4 with a REAL producer: the reference model run in
    memory_format=torch.channels_last makes cuDNN emit (B,H,W,C) feature maps; the drop-in hands
    them to the build kernel as they are (K-major operands, no .contiguous() copy) and the
    32-iteration disparity stays within the north_star tolerance of the stock NCHW run."""
    import smoe_stereo_b200 as sb
    import smoe_stereo_b200.corr as c
    model, left, right = raft_fp32
    model.args.corr_implementation = "reg"
    sb.uninstall()
    stock = _disp(model, left, right, 32)
    seen = []
    real_build = c._build_into

    def spy(pyr, f1, f2, flags=0):
        seen.append((c._is_channels_last(f1), c._is_channels_last(f2), f1.data_ptr()))
        return real_build(pyr, f1, f2, flags)

    model.to(memory_format=torch.channels_last)
    c._build_into = spy
    sb.install()
    try:
        ours = _disp(model, left.contiguous(memory_format=torch.channels_last),
                     right.contiguous(memory_format=torch.channels_last), 32)
    finally:
        sb.uninstall()
        c._build_into = real_build
        model.to(memory_format=torch.contiguous_format)
    assert seen and all(a and b for a, b, _ in seen), f"the producer did not emit channels-last features: {seen}"
    epe = (stock - ours).abs().mean().item()
    print(f"\n[e2e RAFTStereo channels-last producer] EPE vs stock NCHW {epe:.3e} px")
    assert epe <= EPE_TOL


def test_raftstereo_reg_cuda_under_autocast_fp16_path(raft_fp32):
    """The mode that actually runs under the reference's defaults when reg_cuda is selected:
    mixed_precision=True keeps the feature maps in fp16 (core/raft_stereo.py:183-184), so the volume, the
    pyramid and every lookup are fp16 -- stock arm = the reference's own CUDA sampler (oracle/_ref),
    ours = tcgen05 kind::f16 build + fp16 lookup kernels.  fp16 GRU arithmetic amplifies 1-ulp
    differences of the volume; bound = max(north_star, 3x the stock arm's run-to-run floor)."""
    if not R.ref_sampler_available():
        pytest.skip("oracle/_ref/corr_sampler_ref.so not built (oracle/build_ref_sampler.py)")
    model, left, right = raft_fp32
    model.args.mixed_precision = True
    try:
        r = _stock_vs_installed(model, "reg_cuda", left, right, 32)
    finally:
        model.args.mixed_precision = False
    print(f"\n[e2e RAFTStereo reg_cuda, autocast: fp16 volume + fp16 lookups] mean|disp| {r['disp']:.3f} px  "
          f"EPE {r['epe']:.3e} px (max {r['epe_max']:.3e}; stock run-to-run {r['floor']:.1e})")
    assert r["disp"] > 0.05
    assert r["dtypes"] == {(torch.float16, torch.float16)}, r["dtypes"]      # fp16 pyramid, fp16 lookups
    assert r["epe"] <= max(EPE_TOL, 3 * r["floor"])
