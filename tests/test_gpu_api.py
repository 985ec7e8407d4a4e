"""GPU: behaviour of the reference-facing classes beyond the golden cases -- level counts, radii,
the "alt" alias, fp16 training path, streams, non-contiguous coords, the drop-in installer."""
import sys
import types

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_max

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import smoe_stereo_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.mark.parametrize("levels,radius", [(1, 4), (2, 4), (3, 4), (4, 2), (4, 5), (2, 7)])
def test_levels_and_radii_vs_oracle(sb, oracle, levels, radius):
    """num_levels 1..4 use the templated vector kernel (radius 4) or the generic one (others)."""
    from gpu_util import dev, make_coords
    B, C, H, W = 2, 64, 4, 96
    rs = np.random.RandomState(levels * 10 + radius)
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float16).astype(np.float32)   # exact in tf32
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float16).astype(np.float32)
    want_lv = oracle.c_pyramid(oracle.numpy_corr_volume(f1, f2), levels)
    blk = sb.CorrBlock1D(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()),
                         num_levels=levels, radius=radius)
    assert blk.num_levels == levels and blk.radius == radius and len(blk.corr_pyramid) == levels
    coords = make_coords(B, H, W, W, 2, seed=5)
    for t in range(2):
        out = blk(torch.from_numpy(coords[t]).to(dev()))
        assert tuple(out.shape) == (B, levels * (2 * radius + 1), H, W)
        want = oracle.c_lookup(want_lv, np.ascontiguousarray(coords[t]), radius)
        assert rel_max(out.cpu().numpy(), want) < 5e-6


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "w1_ne_w2"])
def test_alt_block_matches_reference_alt_output(sb, name):
    from gpu_util import dev
    g = load_golden(name)
    d = g["dims"]
    alt = sb.PytorchAlternateCorrBlock1D(torch.from_numpy(g["f1"]).to(dev()),
                                         torch.from_numpy(g["f2"]).to(dev()),
                                         num_levels=d["L"], radius=d["r"])
    out = alt(torch.from_numpy(g["coords"][0]).to(dev()))
    assert out.dtype == torch.float32
    assert rel_max(out.cpu().numpy(), g["out_alt"]) < 1e-3      # core/corr.py:64-107
    assert alt.fmap1 is not None and alt.fmap2 is not None


def test_fast_block_fp16_training_path(sb):
    """reg_cuda under autocast: fp16 features, fp16 lookups, gradients flow back in fp16."""
    from gpu_util import dev, make_coords
    B, C, H, W = 1, 64, 4, 64
    g = torch.Generator(device="cpu").manual_seed(3)
    f1 = torch.randn(B, C, H, W, generator=g).half().to(dev()).requires_grad_(True)
    f2 = torch.randn(B, C, H, W, generator=g).half().to(dev()).requires_grad_(True)
    r1 = f1.detach().float().requires_grad_(True)
    r2 = f2.detach().float().requires_grad_(True)
    coords = torch.from_numpy(make_coords(B, H, W, W, 1, seed=2)[0]).to(dev())
    gout = torch.randn(B, 36, H, W, generator=g).to(dev())
    out16 = sb.CorrBlockFast1D(f1, f2)(coords)
    assert out16.dtype == torch.float16 and out16.requires_grad
    (out16.float() * gout).sum().backward()
    out32 = sb.CorrBlock1D(r1, r2)(coords)
    (out32 * gout).sum().backward()
    assert f1.grad.dtype == torch.float16
    assert rel_max(out16.float().detach().cpu().numpy(), out32.detach().cpu().numpy()) < 3e-3
    assert rel_max(f1.grad.float().cpu().numpy(), r1.grad.cpu().numpy()) < 5e-3
    assert rel_max(f2.grad.float().cpu().numpy(), r2.grad.cpu().numpy()) < 5e-3


def test_no_grad_and_detached_inputs_skip_autograd(sb):
    from gpu_util import dev
    f = torch.randn(1, 32, 2, 32, device=dev(), requires_grad=True)
    coords = torch.zeros(1, 2, 2, 32, device=dev())
    with torch.no_grad():
        out = sb.CorrBlock1D(f, f)(coords)
    assert not out.requires_grad
    out = sb.CorrBlock1D(f.detach(), f.detach())(coords)
    assert not out.requires_grad


def test_side_stream_and_noncontiguous_coords(sb, oracle):
    from gpu_util import dev
    B, C, H, W = 2, 32, 3, 40
    rs = np.random.RandomState(0)
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float16).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float16).astype(np.float32)
    x = (np.arange(W, dtype=np.float32) - rs.uniform(0, 9, size=(B, H, W))).astype(np.float32)
    levels = oracle.c_pyramid(oracle.numpy_corr_volume(f1, f2), 4)
    want = oracle.c_lookup(levels, np.ascontiguousarray(x[:, None]), 4)
    big = torch.zeros(B, 5, H, W, device=dev())
    big[:, 2] = torch.from_numpy(x).to(dev())
    s = torch.cuda.Stream(dev())
    s.wait_stream(torch.cuda.current_stream(dev()))
    with torch.cuda.stream(s):
        blk = sb.CorrBlock1D(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()))
        out_slice = blk(big[:, 2:4])                 # batch stride 5*H*W, channel 0 of the view = x
        out_perm = blk(big.permute(0, 1, 3, 2)[:, 2:3].permute(0, 1, 3, 2))
        out_f64 = blk(big[:, 2:3].double())
    s.synchronize()
    for o in (out_slice, out_perm, out_f64):
        assert rel_max(o.cpu().numpy(), want) < 5e-6


def test_install_serves_reference_style_callers(sb):
    """A module written the way core/corr.py:17-51 uses the extension (bare `import corr_sampler`,
    `corr, = corr_sampler.forward(...)`) runs on this implementation after install()."""
    from gpu_util import dev
    fake = types.ModuleType("core.corr")
    exec("import torch\n"
         "def lookup(vol, coords, r):\n"
         "    import corr_sampler\n"
         "    corr, = corr_sampler.forward(vol, coords, r)\n"
         "    return corr\n", fake.__dict__)
    fake.CorrBlock1D = object
    sys.modules["core.corr"] = fake
    try:
        replaced = sb.install()
        assert "core.corr.CorrBlock1D" in replaced and fake.CorrBlock1D is sb.CorrBlock1D
        vol = torch.randn(1, 2, 16, 32, device=dev())
        coords = torch.rand(1, 1, 2, 16, device=dev()) * 30
        got = fake.lookup(vol, coords, 4)
        want, = sb.corr_sampler.forward(vol, coords, 4)
        assert torch.equal(got, want)
    finally:
        sb.uninstall()
        sys.modules.pop("core.corr", None)


def test_sharded_block_single_process_equals_full(sb):
    from gpu_util import dev
    f1 = torch.randn(2, 32, 6, 48, device=dev())
    f2 = torch.randn(2, 32, 6, 48, device=dev())
    coords = torch.zeros(2, 2, 6, 48, device=dev())
    coords[:, 0] = torch.arange(48, device=dev()).float() - 3.7
    full = sb.CorrBlock1D(f1, f2)(coords)
    for mode in ("batch", "rows"):
        blk = sb.ShardedCorrBlock1D(f1, f2, mode=mode)
        assert torch.equal(blk(coords, gather=True), full)
    # manual 2-way row shard, as two ranks would hold it: bit-identical to the full result
    parts = []
    for r in range(2):
        lo, hi = sb.shard_bounds(6, 2, r)
        b = sb.CorrBlock1D(f1[:, :, lo:hi].contiguous(), f2[:, :, lo:hi].contiguous())
        parts.append(b(coords[:, :, lo:hi].contiguous()))
    assert torch.equal(torch.cat(parts, dim=2), full)


@pytest.mark.parametrize("shape,levels,radius", [((2, 32, 6, 48), 4, 4), ((1, 32, 3, 37), 4, 4), ((1, 32, 4, 40), 3, 3)])
def test_fused_lookup_update_equals_loop_glue(sb, shape, levels, radius):
    """lookup_update == the reference loop body around the lookup (core/SMoEStereo.py:234-248),
    bit for bit: coords1 + (dx, 0), lookup at the new x, flow = coords1 - coords0."""
    from gpu_util import dev
    B, C, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(W)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev())
    f2 = torch.randn(B, C, H, W, generator=g).to(dev())
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords0 = torch.stack([xs, ys], 0).float()[None].repeat(B, 1, 1, 1).to(dev())
    coords1 = (coords0 - torch.rand(B, 2, H, W, generator=g).to(dev()) * 9).contiguous()
    delta = torch.randn(B, 2, H, W, generator=g).to(dev())
    blk = sb.CorrBlock1D(f1, f2, num_levels=levels, radius=radius)
    with torch.no_grad():
        d = delta.clone()
        d[:, 1] = 0
        want_c1 = coords1 + d
        want_corr = blk(want_c1)
        want_flow = want_c1 - coords0
        corr, c1, flow = blk.lookup_update(coords1, delta, coords0)
        assert torch.equal(c1, want_c1) and torch.equal(flow, want_flow) and torch.equal(corr, want_corr)
        corr2, c2, flow2 = blk.lookup_update(coords1)            # no update, no flow
        assert flow2 is None and torch.equal(c2, coords1) and torch.equal(corr2, blk(coords1))
    # training: gradients reach delta_flow (x only) and the feature maps
    f1g = f1.clone().requires_grad_(True)
    dg = delta.clone().requires_grad_(True)
    blk_g = sb.CorrBlock1D(f1g, f2, num_levels=levels, radius=radius)
    corr, c1, flow = blk_g.lookup_update(coords1, dg, coords0)
    (corr.sum() + (flow * 2.0).sum()).backward()
    assert torch.all(dg.grad[:, 0] == 2.0) and torch.all(dg.grad[:, 1] == 0.0)
    assert f1g.grad is not None and f1g.grad.abs().sum() > 0


def test_second_device_and_threads_like_dataparallel(sb):
    """nn.DataParallel drives the path from one host thread per GPU (train_stereo_sceneflow.py:133):
    tensors on cuda:1 while cuda:0 is current, and concurrent calls from two threads."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import threading
    torch.cuda.set_device(0)
    g = torch.Generator(device="cpu").manual_seed(0)
    f1 = torch.randn(1, 64, 4, 64, generator=g)
    f2 = torch.randn(1, 64, 4, 64, generator=g)
    coords = torch.zeros(1, 2, 4, 64)
    coords[:, 0] = torch.arange(64).float() - 5.5
    ref = sb.CorrBlock1D(f1.cuda(0), f2.cuda(0))(coords.cuda(0)).cpu()
    out1 = sb.CorrBlock1D(f1.cuda(1), f2.cuda(1))(coords.cuda(1))      # current device is still 0
    assert out1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(out1.cpu(), ref)
    results = {}

    def worker(d):
        with torch.cuda.device(d):
            for _ in range(20):
                results[d] = sb.CorrBlock1D(f1.cuda(d), f2.cuda(d))(coords.cuda(d)).cpu()
    th = [threading.Thread(target=worker, args=(d,)) for d in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert torch.equal(results[0], ref) and torch.equal(results[1], ref)


def test_optin_fp16_volume_storage(sb, oracle):
    """set_default_volume_dtype(float16): fp32 feature maps, tensor-core result rounded once to fp16
    for storage (opt-in).  Still inside north_star's 1e-3; lookups stay fp32; training works."""
    from gpu_util import dev, make_coords
    B, C, H, W = 1, 256, 8, 240
    rs = np.random.RandomState(3)
    f1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    want_lv = oracle.c_pyramid(oracle.c_corr_volume(f1, f2), 4)
    coords = make_coords(B, H, W, W, 1, seed=1)[0]
    want = oracle.c_lookup(want_lv, np.ascontiguousarray(coords), 4)
    sb.set_default_volume_dtype(torch.float16)
    try:
        t1 = torch.from_numpy(f1).to(dev()).requires_grad_(True)
        t2 = torch.from_numpy(f2).to(dev())
        blk = sb.CorrBlock1D(t1, t2)
        assert blk._state.pyr.dtype == torch.float16
        for i, lv in enumerate(blk.corr_pyramid):
            assert rel_max(lv.squeeze(3).float().detach().cpu().numpy(), want_lv[i]) < 1e-3
        out = blk(torch.from_numpy(coords).to(dev()))
        assert out.dtype == torch.float32
        e = rel_max(out.detach().cpu().numpy(), want)
        print(f"[fp16 storage] lookup rel_max {e:.2e}")
        assert e < 1e-3
        out.square().sum().backward()
        assert torch.isfinite(t1.grad).all() and t1.grad.abs().sum() > 0
        # the fast block keeps the reference rule (volume dtype == feature dtype)
        assert sb.CorrBlockFast1D(t1.detach(), t2)._state.pyr.dtype == torch.float32
    finally:
        sb.set_default_volume_dtype(None)
    assert sb.CorrBlock1D(t1.detach(), t2)._state.pyr.dtype == torch.float32


def test_call_with_caller_owned_output_buffer(sb):
    """`blk(coords, out=buf)`: same bits as the allocating call, written in place; wrong shapes /
    dtypes raise; refused while autograd records the lookup."""
    from gpu_util import dev, make_coords
    B, C, H, W = 2, 64, 4, 96
    g = torch.Generator(device="cpu").manual_seed(11)
    f1 = torch.randn((B, C, H, W), generator=g).to(dev())
    f2 = torch.randn((B, C, H, W), generator=g).to(dev())
    coords = torch.from_numpy(make_coords(B, H, W, W, 2, seed=2)).to(dev())
    blk = sb.CorrBlock1D(f1, f2, num_levels=4, radius=4)
    staging = torch.full((2, B, 36, H, W), float("nan"), device=dev())
    with torch.no_grad():
        for t in range(2):
            want = blk(coords[t])
            got = blk(coords[t], out=staging[t])
            assert got.data_ptr() == staging[t].data_ptr()
            assert torch.equal(got, want)
        with pytest.raises(RuntimeError):
            blk(coords[0], out=staging[0, :, :35])
        with pytest.raises(RuntimeError):
            blk(coords[0], out=staging[0].half())
        with pytest.raises(RuntimeError):
            blk(coords[0], out=staging[0].cpu())
    tr = sb.CorrBlock1D(f1.clone().requires_grad_(True), f2, num_levels=4, radius=4)
    with pytest.raises(RuntimeError):
        tr(coords[0], out=staging[0])


def test_gru_loop_captured_in_a_cuda_graph(sb):
    """INTEGRATION.md "CUDA graphs": the whole update loop (build + N x [coords update, lookup]) is
    captured ONCE and replayed on new inputs copied into the static buffers -- no host work per
    lookup at all.  Results equal the eager loop bit for bit."""
    from gpu_util import dev
    B, C, H, W, iters = 1, 64, 6, 96, 5
    g = torch.Generator(device="cpu").manual_seed(4)
    static_f1 = torch.empty((B, C, H, W), device=dev())
    static_f2 = torch.empty((B, C, H, W), device=dev())
    static_c0 = torch.empty((B, 2, H, W), device=dev())
    delta = [torch.randn((B, 2, H, W), generator=g).to(dev()) * 0.5 for _ in range(iters)]

    def loop():
        blk = sb.CorrBlock1D(static_f1, static_f2, num_levels=4, radius=4)
        coords = static_c0
        outs = []
        for t in range(iters):
            coords = coords + delta[t]                 # stands in for the GRU's delta_flow
            outs.append(blk(coords))
        return blk, outs

    def fill(seed):
        gg = torch.Generator(device="cpu").manual_seed(seed)
        static_f1.copy_(torch.randn((B, C, H, W), generator=gg))
        static_f2.copy_(torch.randn((B, C, H, W), generator=gg))
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
        ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(B, 1, H, W)
        static_c0.copy_(torch.cat([xs - 3.25 * seed, ys], dim=1))

    side = torch.cuda.Stream(dev())
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad():
        fill(1)
        with torch.cuda.stream(side):
            loop()                                       # warm-up outside capture (allocator, module load)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                keep = loop()                            # the block and its outputs live in the graph's pool
        for seed in (2, 3):
            fill(seed)
            graph.replay()
            torch.cuda.synchronize()
            got = [o.clone() for o in keep[1]]
            _, want = loop()
            torch.cuda.synchronize()
            for a, b in zip(got, want):
                assert torch.equal(a, b)


def test_concurrent_host_threads_and_streams_on_one_device(sb):
    """Re-entrancy on ONE GPU (what the driver's single-GPU box can run): four host threads, each with
    its own CUDA stream, build blocks of different shapes, look up and back-propagate concurrently;
    every thread must get exactly what it gets when it runs alone (no global mutable state in the
    library, thread-local error strings, kernels on the caller's current stream)."""
    import threading
    from gpu_util import dev, make_coords
    shapes = [(1, 64, 4, 64), (2, 32, 5, 96), (1, 128, 3, 40), (2, 64, 6, 37)]
    cases = []
    for i, (B, C, H, W) in enumerate(shapes):
        g = torch.Generator(device="cpu").manual_seed(100 + i)
        f1 = torch.randn(B, C, H, W, generator=g).to(dev())
        f2 = torch.randn(B, C, H, W, generator=g).to(dev())
        coords = torch.from_numpy(make_coords(B, H, W, W, 3, seed=i)).to(dev())
        cases.append((f1, f2, coords))

    def run(case):
        f1, f2, coords = case
        a = f1.clone().requires_grad_(True)
        blk = sb.CorrBlock1D(a, f2, num_levels=4, radius=4)
        outs = [blk(coords[t]) for t in range(3)]
        sum(o.square().sum() for o in outs).backward()
        return [o.detach().clone() for o in outs], a.grad.clone()

    alone = [run(c) for c in cases]
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(i):
        try:
            s = torch.cuda.Stream(dev())
            with torch.cuda.device(dev()), torch.cuda.stream(s):
                for _ in range(10):
                    results[i] = run(cases[i])
                s.synchronize()
        except Exception as e:          # noqa: BLE001
            errors.append((i, repr(e)))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(cases))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    for i, (outs, grad) in enumerate(alone):
        for a, b in zip(outs, results[i][0]):
            assert torch.equal(a, b), f"thread {i}: lookup differs from the single-threaded run"
        assert torch.equal(grad, results[i][1]), f"thread {i}: gradient differs from the single-threaded run"
    # an error raised in one thread leaves the other threads' error state alone
    with pytest.raises(RuntimeError):
        sb.CorrBlock1D(cases[0][0], cases[1][1])        # mismatched shapes
