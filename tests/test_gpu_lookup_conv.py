"""GPU parity of the fused lookup + 1x1 convolution (+bias, +ReLU) kernel
against the reference's motion-encoder first layer (golden, core/update.py:70,77) and the oracle.
Tolerance: north_star's <=1e-3 relative (max|d|/max|ref|) for tensor-core paths; the products run
in tf32 (round to nearest), accumulation in fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import convc1_names, load_convc1, load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def sb():
    import smoe_stereo_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.mark.parametrize("name", convc1_names())
def test_fused_lookup_conv_vs_reference_golden(sb, name):
    from gpu_util import dev
    g, c = load_golden(name), load_convc1(name)
    d = g["dims"]
    blk = sb.CorrBlock1D(torch.from_numpy(g["f1"]).to(dev()), torch.from_numpy(g["f2"]).to(dev()),
                         num_levels=d["L"], radius=d["r"])
    w = torch.from_numpy(c["weight"]).to(dev())
    b = torch.from_numpy(c["bias"]).to(dev())
    for t in range(d["T"]):
        coords = torch.from_numpy(g["coords"][t]).to(dev())
        got = blk.lookup_conv(coords, w.view(64, -1, 1, 1), b, relu=True)
        assert got.shape == (d["B"], 64, d["H"], d["W1"]) and got.dtype == torch.float32
        assert got.is_contiguous()
        e = rel_max(got.cpu().numpy(), c["cor"][t])
        print(f"[lookup_conv {name} t={t}] rel_max={e:.3e} rel_l2={rel_l2(got.cpu().numpy(), c['cor'][t]):.3e}")
        assert e < TOL
        assert (got >= 0).all()


@pytest.mark.parametrize("shape", [
    (1, 64, 135, 240, 4, 64),      # BASELINE config 1 plane, the reference's 36 -> 64 layer
    (2, 32, 9, 77, 4, 64),         # ragged plane (693 px: partial last 128-pixel tile), odd width
    (1, 32, 5, 40, 3, 32),         # 3 levels -> K = 27, 32 output channels
    (1, 32, 5, 40, 1, 16),         # 1 level  -> K = 9 (one padded MMA k-step pair), 16 channels
    (1, 32, 4, 48, 2, 256),        # widest supported layer
])
@pytest.mark.parametrize("relu,use_bias", [(True, True), (False, False)])
def test_fused_lookup_conv_vs_oracle(sb, oracle, shape, relu, use_bias):
    from gpu_util import dev
    B, C, H, W, L, cout = shape
    g = torch.Generator().manual_seed(5)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.randn(B, C, H, W, generator=g)
    coords = torch.zeros(B, 2, H, W)
    coords[:, 0] = torch.arange(W, dtype=torch.float32)[None, None, :] - torch.rand(B, H, W, generator=g) * (W * 0.7) + 3.0
    coords[:, 0, :, 0] = -20.0           # fully out of range
    coords[:, 0, :, 1] = float(W - 1)    # exact last column
    weight = torch.randn(cout, L * 9, generator=g) * 0.3
    bias = torch.randn(cout, generator=g) if use_bias else None
    blk = sb.CorrBlock1D(f1.to(dev()), f2.to(dev()), num_levels=L, radius=4)
    got = blk.lookup_conv(coords.to(dev()), weight.to(dev()), bias.to(dev()) if use_bias else None, relu=relu)
    # oracle on the SAME pyramid (isolates the fused op from the tf32 volume build)
    levels = [lv.squeeze(3).cpu().numpy() for lv in blk.corr_pyramid]
    corr = oracle.numpy_lookup(levels, coords.numpy(), 4)
    want = oracle.numpy_conv1x1(corr, weight.numpy(), bias.numpy() if use_bias else None, relu=relu)
    e = rel_max(got.cpu().numpy(), want)
    print(f"[lookup_conv oracle {shape} relu={relu}] rel_max={e:.3e}")
    assert got.shape == want.shape
    assert e < TOL
    # and against the unfused composition of this library's own lookup + torch's fp32 conv
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(blk(coords.to(dev())), weight.to(dev()).view(cout, -1, 1, 1), bias.to(dev()) if use_bias else None)
    ref = F.relu(ref) if relu else ref
    assert rel_max(got.cpu().numpy(), ref.cpu().numpy()) < TOL


def test_fused_lookup_conv_fp16_volume_and_half_output(sb):
    """reg_cuda flavour: fp16 features -> fp16 pyramid -> fp16 lookup arithmetic (as the reference's
    c10::Half sampler) -> convolution; fp16 output as autocast would produce."""
    from gpu_util import dev
    g = torch.Generator().manual_seed(9)
    B, C, H, W = 1, 64, 16, 96
    f1 = torch.randn(B, C, H, W, generator=g).half().to(dev())
    f2 = torch.randn(B, C, H, W, generator=g).half().to(dev())
    coords = torch.zeros(B, 2, H, W)
    coords[:, 0] = torch.arange(W, dtype=torch.float32)[None, None, :] - torch.rand(B, H, W, generator=g) * 40
    coords = coords.to(dev())
    w = (torch.randn(64, 36, generator=g) * 0.3).to(dev())
    b = torch.randn(64, generator=g).to(dev())
    blk = sb.CorrBlockFast1D(f1, f2, num_levels=4, radius=4)
    corr = blk(coords)
    assert corr.dtype == torch.float16
    want = F.relu(F.conv2d(corr.float(), w.view(64, 36, 1, 1), b))
    got32 = blk.lookup_conv(coords, w, b, out_dtype=torch.float32)
    got16 = blk.lookup_conv(coords, w, b)                      # default: the lookup dtype (fp16)
    assert got32.dtype == torch.float32 and got16.dtype == torch.float16
    # the lookup values are fp16 -> exact in tf32; only the weights are rounded
    assert rel_max(got32.cpu().numpy(), want.cpu().numpy()) < TOL
    assert rel_max(got16.float().cpu().numpy(), want.cpu().numpy()) < 2e-3   # + fp16 output rounding


def test_lookup_conv_unsupported_shapes_compose_and_grads_flow(sb):
    """radius 3 / 24 output channels are outside the fused kernel: the method composes the lookup
    kernel with F.conv2d; so does a call that needs gradients."""
    from gpu_util import dev
    g = torch.Generator().manual_seed(3)
    f1 = torch.randn(1, 16, 4, 32, generator=g).to(dev())
    f2 = torch.randn(1, 16, 4, 32, generator=g).to(dev())
    coords = torch.zeros(1, 2, 4, 32)
    coords[:, 0] = torch.arange(32, dtype=torch.float32)[None, None, :] - 2.5
    coords = coords.to(dev())
    blk = sb.CorrBlock1D(f1, f2, num_levels=3, radius=3)
    w = torch.randn(24, 21, generator=g).to(dev())
    got = blk.lookup_conv(coords, w, None, relu=True)
    want = F.relu(F.conv2d(blk(coords), w.view(24, 21, 1, 1)))
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    blk4 = sb.CorrBlock1D(f1, f2, num_levels=4, radius=4)
    w4 = torch.randn(64, 36, generator=g).to(dev()).requires_grad_(True)
    y = blk4.lookup_conv(coords, w4, None)
    y.sum().backward()
    assert w4.grad is not None and w4.grad.abs().sum() > 0
    with pytest.raises(RuntimeError):
        blk4.lookup_conv(coords, torch.randn(64, 35).to(dev()))


def test_lookup_conv_c_abi_rejects_bad_arguments(sb):
    from gpu_util import dev
    import ctypes
    L = sb._lib.lib()
    f = torch.randn(1, 16, 4, 32).to(dev())
    blk = sb.CorrBlock1D(f, f, num_levels=4, radius=4)
    p = blk._state.pyr
    coords = torch.zeros(1, 2, 4, 32, device=dev())
    w = torch.zeros(64, 36, device=dev())
    out = torch.empty(1, 64, 4, 32, device=dev())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = lambda radius, cout: (p.ref, coords.data_ptr(), 2 * 4 * 32, 1.0, radius, 0, w.data_ptr(), None,
                                 cout, 1, out.data_ptr(), 0, stream)
    assert L.smoe_corr_lookup_conv1x1(*args(3, 64)) != 0 and "radius" in sb._lib.last_error()
    assert L.smoe_corr_lookup_conv1x1(*args(4, 24)) != 0 and "output channels" in sb._lib.last_error()
    assert L.smoe_corr_lookup_conv1x1(*args(4, 64)) == 0
