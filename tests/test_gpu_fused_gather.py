"""smoe_corr_lookup_to: the lookup written through a strided view of a larger (gathered) output, and --
on a box with >= 2 GPUs -- through an NVSwitch multicast address (lookup + all-gather in one kernel)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flags", [0, 2])          # 0: 16-byte stores where alignment allows; 2 = SMOE_OUT_SCALAR
@pytest.mark.parametrize("mode", ["batch", "rows"])
@pytest.mark.parametrize("shape", [(2, 64, 6, 96), (3, 32, 5, 37), (1, 32, 9, 750), (2, 32, 8, 100), (1, 256, 135, 240)])
def test_lookup_into_a_strided_view_equals_the_plain_lookup(mode, shape, flags):
    """Single GPU: every "rank's" shard is computed in turn and written straight into its place inside
    the full (B,36,H,W) tensor (batch shards: an offset; row shards: an offset and a larger channel
    stride).  The assembled tensor equals the unsharded lookup bit for bit."""
    import smoe_stereo_b200 as sb
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _stream_ptr
    from smoe_stereo_b200.sharding import shard_bounds
    from gpu_util import dev, make_coords
    B, C, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(B * 100 + W)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev())
    f2 = torch.randn(B, C, H, W, generator=g).to(dev())
    coords = torch.from_numpy(make_coords(B, H, W, W, 1, seed=3)[0]).to(dev())
    want = sb.CorrBlock1D(f1, f2)(coords)
    full = torch.full((B, 36, H, W), float("nan"), device=dev())
    world = 2 if (B if mode == "batch" else H) >= 2 else 1
    plane = H * W
    for rank in range(world):
        lo, hi = shard_bounds(B if mode == "batch" else H, world, rank)
        sl = (slice(lo, hi),) if mode == "batch" else (slice(None), slice(None), slice(lo, hi))
        blk = sb.CorrBlock1D(f1[sl].contiguous(), f2[sl].contiguous())
        c = coords[sl].contiguous()
        p = blk._state.pyr
        off = lo * 36 * plane if mode == "batch" else lo * W
        rc = _lib.lib().smoe_corr_lookup_to(p.ref, c.data_ptr(), c.stride(0), 1.0, 4, full.data_ptr() + 4 * off,
                                            36 * plane, plane, _lib.DTYPE_F32, flags, _stream_ptr(dev()))
        _lib.check(rc, "smoe_corr_lookup_to")
    assert torch.equal(full, want)


def test_lookup_to_rejects_what_it_does_not_cover():
    import smoe_stereo_b200 as sb
    from smoe_stereo_b200 import _lib
    from smoe_stereo_b200.corr import _stream_ptr
    from gpu_util import dev
    f = torch.randn(1, 32, 4, 64, device=dev())
    blk = sb.CorrBlock1D(f, f, num_levels=4, radius=3)
    p = blk._state.pyr
    out = torch.empty(1, 28, 4, 64, device=dev())
    c = torch.zeros(1, 2, 4, 64, device=dev())
    rc = _lib.lib().smoe_corr_lookup_to(p.ref, c.data_ptr(), c.stride(0), 1.0, 3, out.data_ptr(), 28 * 256, 256,
                                        _lib.DTYPE_F32, 0, _stream_ptr(dev()))
    assert rc == _lib.ERR_UNSUPPORTED
    blk4 = sb.CorrBlock1D(f, f)
    rc = _lib.lib().smoe_corr_lookup_to(blk4._state.pyr.ref, c.data_ptr(), c.stride(0), 1.0, 4, out.data_ptr(), 100, 256,
                                        _lib.DTYPE_F32, 0, _stream_ptr(dev()))
    assert rc == _lib.ERR_INVALID_ARG and b"strides" in _lib.lib().smoe_corr_last_error()


def test_fused_multicast_gather_on_two_gpus():
    """2 ranks (NCCL + torch symmetric memory): ShardedCorrBlock1D(...)(coords, gather="fused") returns the
    full lookup on both ranks, bit-identical to lookup + NCCL all-gather, for batch and row sharding."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tools", "fused_gather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "fused gather ok" in r.stdout
