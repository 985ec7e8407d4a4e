"""Multi-GPU plumbing for the correlation path: one process per GPU (torchrun), no collective
inside the path.

corr[b,h,:,:] depends only on fmap1[b,:,h,:] and fmap2[b,:,h,:], and a lookup at pixel (b,y,x)
touches only pixel row (b,y,x) (core/corr.py:154, sampler/sampler_kernel.cu:50), so the path
shards by batch (dim 0) or by image rows (dim 2) with zero halo.  Each rank builds and keeps its
own pyramid shard; lookups and backward are rank-local.  The only optional communication is an
all-gather of lookup outputs for a consumer that is not sharded the same way (the reference's
equivalent is nn.DataParallel's implicit gather, train_stereo_sceneflow.py:133).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_DIM = {"batch": 0, "rows": 2}


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of range(n); the first n % world_size ranks get one more
    (500 rows on 8 ranks -> 63,63,63,63,62,62,62,62)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world_size: int) -> list[int]:
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0]
            for r in range(world_size)]


def take_shard(t: torch.Tensor, mode: str, world_size: int, rank: int) -> torch.Tensor:
    """This rank's slice of a (B,C,H,W) tensor along the batch or row axis."""
    d = _DIM[mode]
    lo, hi = shard_bounds(t.shape[d], world_size, rank)
    return t.narrow(d, lo, hi - lo)


def gather_lookup(local: torch.Tensor, mode: str, full_extent: int, group=None) -> torch.Tensor:
    """All-gather per-rank lookup outputs (B_r,36,H_r,W) into the full (B,36,H,W) tensor.
    Shards may be uneven; each rank's block is padded to the largest shard for the collective."""
    d = _DIM[mode]
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(full_extent, world)
    if local.shape[d] != sizes[dist.get_rank(group)]:
        raise RuntimeError(f"local shard extent {local.shape[d]} != expected {sizes[dist.get_rank(group)]}")
    mx = max(sizes)
    moved = local.movedim(d, 0).contiguous()
    if moved.shape[0] < mx:
        pad = moved.new_zeros((mx - moved.shape[0],) + tuple(moved.shape[1:]))
        moved = torch.cat([moved, pad], dim=0)
    out = moved.new_empty((world * mx,) + tuple(moved.shape[1:]))
    dist.all_gather_into_tensor(out, moved, group=group)
    parts = [out[r * mx:r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0).movedim(0, d).contiguous()


class ShardedCorrBlock1D:
    """CorrBlock1D over this rank's batch- or row-shard of the feature maps.

    fmap1/fmap2 are the FULL tensors unless ``already_sharded``; coords passed to __call__ follow
    the same convention.  ``__call__(coords, gather=True)`` returns the full lookup on every rank.
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, mode="batch", group=None,
                 already_sharded=False, block_cls=None):
        if mode not in _DIM:
            raise ValueError("mode must be 'batch' or 'rows'")
        if block_cls is None:
            from .corr import CorrBlock1D as block_cls
        self.mode, self.group, self.already_sharded = mode, group, already_sharded
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        d = _DIM[mode]
        if already_sharded:
            ext = torch.tensor([fmap1.shape[d]], device=fmap1.device)
            if self.world > 1:
                dist.all_reduce(ext, group=group)
            self.full_extent = int(ext.item())
        else:
            self.full_extent = fmap1.shape[d]
            fmap1 = take_shard(fmap1, mode, self.world, self.rank)
            fmap2 = take_shard(fmap2, mode, self.world, self.rank)
        self.block = block_cls(fmap1.contiguous(), fmap2.contiguous(), num_levels=num_levels,
                               radius=radius)
        self.num_levels, self.radius = num_levels, radius

    def __call__(self, coords, gather=False):
        if not self.already_sharded:
            coords = take_shard(coords, self.mode, self.world, self.rank).contiguous()
        out = self.block(coords)
        if gather and self.world > 1:
            out = gather_lookup(out, self.mode, self.full_extent, self.group)
        return out
