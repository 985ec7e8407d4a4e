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


class FusedGather:
    """Lookup + all-gather in ONE kernel (smoe_corr_lookup_to): every rank's lookup kernel stores its
    shard of the result through an NVSwitch MULTICAST address (`multimem.st`), and the switch replicates
    each value into the gathered buffer of every GPU of the group -- no NCCL kernel, no second pass
    over the result.  The buffers are torch symmetric-memory allocations (one (B,36,H,W) fp32 tensor
    per GPU, a ring of `depth` of them); a symmetric-memory barrier after the lookup makes the peers'
    stores visible.  Raises if the group has no multicast mapping; use gather=True (NCCL) there."""

    def __init__(self, B, C, H, W, device, group=None, depth=2):
        import torch.distributed._symmetric_memory as symm
        self.shape = (B, C, H, W)
        self.depth = depth
        self.buf = symm.empty((depth, B, C, H, W), dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.mc_ptr = int(self.hdl.multicast_ptr or 0)
        if not self.mc_ptr:
            raise RuntimeError("the process group's GPUs have no multicast (NVLS) mapping for symmetric memory")
        self.pos = 0

    def lookup(self, block, coords, mode, lo):
        """Run `block`'s lookup for this rank's shard (which starts at index `lo` of the sharded axis)
        and return the FULL (B,36,H,W) result, valid on every rank after the call."""
        from . import _lib
        from .corr import _coords_plane, _on_device, _stream_ptr
        B, C, H, W = self.shape
        p = block._state.pyr
        if torch.is_grad_enabled() and block._token is not None:
            raise RuntimeError("gather='fused' writes through a multicast address outside autograd: use it under "
                               "torch.no_grad() / for inference, or gather=True while training")
        if p.dtype is not torch.float32 and not block._float_output:
            raise RuntimeError("gather='fused' produces fp32 outputs (CorrBlock1D); use gather=True for CorrBlockFast1D")
        coords, bstride = _coords_plane(coords, p.B, p.H, p.W1)
        slot = self.pos
        self.pos = (self.pos + 1) % self.depth
        plane = H * W
        off = slot * B * C * plane + (lo * C * plane if mode == "batch" else lo * W)
        # (a slot is reused after `depth` further lookups: the barrier of the previous use has long passed)
        with _on_device(p.device):
            rc = _lib.lib().smoe_corr_lookup_to(p.ref, coords.data_ptr(), bstride, 1.0, block.radius,
                                                self.mc_ptr + 4 * off, C * plane, plane, _lib.DTYPE_F32,
                                                _lib.OUT_MULTIMEM, _stream_ptr(p.device))
        _lib.check(rc, "smoe_corr_lookup_to")
        self.hdl.barrier(channel=0)
        return self.buf[slot]


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
        self._fused = None

    def __call__(self, coords, gather=False):
        """gather=False: this rank's shard of the lookup.  gather=True: the full lookup on every rank via
        lookup + NCCL all-gather.  gather="fused": the full lookup on every rank from ONE kernel whose
        stores go through an NVSwitch multicast address (FusedGather; fp32 "reg" blocks, radius 4, no
        autograd); the returned tensor is a ring slot, valid until two further fused lookups."""
        if not self.already_sharded:
            coords = take_shard(coords, self.mode, self.world, self.rank).contiguous()
        if gather == "fused" and self.world > 1:
            if self._fused is None:
                p = self.block._state.pyr
                B = self.full_extent if self.mode == "batch" else p.B
                H = self.full_extent if self.mode == "rows" else p.H
                self._fused = FusedGather(B, p.num_levels * (2 * self.radius + 1), H, p.W1, p.device, self.group)
            lo = shard_bounds(self.full_extent, self.world, self.rank)[0]
            return self._fused.lookup(self.block, coords, self.mode, lo)
        out = self.block(coords)
        if gather and self.world > 1:
            out = gather_lookup(out, self.mode, self.full_extent, self.group)
        return out
