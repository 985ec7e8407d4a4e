"""CPU, world_size 2 over gloo: the N>1 host path (shard -> rank-local lookups -> optional
all-gather) reproduces the unsharded result exactly.  The rank-local block is a CPU stand-in
built on the oracle so that only the sharding/collective plumbing is under test."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleBlock:
    """CPU stand-in with CorrBlock1D's interface (test infrastructure)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        from oracle import corr_oracle as o
        self.o, self.r = o, radius
        vol = o.numpy_corr_volume(fmap1.numpy(), fmap2.numpy())
        self.levels = o.numpy_pyramid(vol, num_levels)

    def __call__(self, coords):
        return torch.from_numpy(self.o.numpy_lookup(self.levels, coords.numpy(), self.r))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import smoe_stereo_b200 as sb
        g = torch.Generator().manual_seed(0)
        B, C, H, W = 3, 8, 5, 24
        f1 = torch.randn(B, C, H, W, generator=g)
        f2 = torch.randn(B, C, H, W, generator=g)
        coords = torch.zeros(B, 2, H, W)
        coords[:, 0] = torch.arange(W).float() - torch.rand(B, H, W, generator=g) * 10
        full = OracleBlock(f1, f2)(coords)
        blk = sb.ShardedCorrBlock1D(f1, f2, mode=mode, block_cls=OracleBlock)
        local = blk(coords)
        d = 0 if mode == "batch" else 2
        lo, hi = sb.shard_bounds(full.shape[d], world, rank)
        assert torch.equal(local, full.narrow(d, lo, hi - lo))
        gathered = blk(coords, gather=True)
        assert torch.equal(gathered, full)
        # already-sharded construction (each rank only ever holds its shard)
        blk2 = sb.ShardedCorrBlock1D(f1.narrow(d, lo, hi - lo), f2.narrow(d, lo, hi - lo), mode=mode,
                                     already_sharded=True, block_cls=OracleBlock)
        assert blk2.full_extent == full.shape[d]
        assert torch.equal(blk2(coords.narrow(d, lo, hi - lo).contiguous(), gather=True), full)
        out_q.put((rank, "ok"))
    except Exception as e:                     # noqa: BLE001
        out_q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _run(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_batch_shard_world2():
    _run("batch")       # B=3 over 2 ranks: uneven shards 2+1


def test_row_shard_world2():
    _run("rows")        # H=5 over 2 ranks: 3+2 rows, zero halo
