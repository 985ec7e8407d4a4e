"""torchrun target (>= 2 ranks): lookup + all-gather fused into one kernel through an NVSwitch multicast
address (ShardedCorrBlock1D gather="fused") against lookup + NCCL all-gather: equality and timing.

    python -m torch.distributed.run --nproc-per-node N tools/fused_gather_check.py [cfg2|cfg3|cfg4 ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import smoe_stereo_b200 as sb  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def inputs(B, C, H, W, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)           # the same full tensors on every rank
    f1 = torch.randn(B, C, H, W, generator=g).to(dev)
    f2 = torch.randn(B, C, H, W, generator=g).to(dev)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    c = torch.zeros(B, 2, H, W)
    c[:, :1] = xs - torch.rand(B, 1, H, 1, generator=g) * 40 - torch.randn(B, 1, H, W, generator=g) * 2
    return f1, f2, c.to(dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e3 / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


with torch.no_grad():
    # correctness on small shapes, both shardings, uneven shards
    for mode, (B, C, H, W) in (("batch", (4, 64, 6, 96)), ("rows", (2, 32, 7, 120)), ("batch", (3, 32, 4, 37))):
        f1, f2, c = inputs(B, C, H, W, seed=5)
        blk = sb.ShardedCorrBlock1D(f1, f2, mode=mode)
        want = sb.CorrBlock1D(f1, f2)(c)
        nccl = blk(c, gather=True)
        for it in range(3):                                        # ring slots are reused
            fused = blk(c, gather="fused")
            torch.cuda.synchronize()
            assert torch.equal(fused, want), f"rank {rank} {mode} {B, C, H, W}: fused gather differs (iteration {it})"
        assert torch.equal(nccl, want)
    # training blocks are refused (the multicast stores bypass autograd); every rank raises, nobody hangs
    f1, f2, c = inputs(2, 32, 4, 64, seed=9)
    with torch.enable_grad():
        tr = sb.ShardedCorrBlock1D(f1.requires_grad_(True), f2, mode="batch")
        try:
            tr(c, gather="fused")
            raise AssertionError("gather='fused' accepted an autograd-recording block")
        except RuntimeError as e:
            assert "autograd" in str(e)
    if rank == 0:
        print("fused gather ok: bit-identical to the unsharded lookup and to lookup + NCCL all-gather", flush=True)
    # timing at BASELINE shapes
    from bench import WORKLOADS
    for arg in [a for a in sys.argv[1:] if a.split(":")[0] in WORKLOADS or a[0].isdigit()]:
        name = arg.split(":")[0]
        if name[0].isdigit():                            # BxHxW[:mode] -- an ad-hoc shape
            B, H, W = (int(v) for v in name.split("x"))
        else:
            wl = WORKLOADS[name]
            B, H, W = wl["B"], wl["H"], wl["W"]
        mode = "batch" if B >= world and B % world == 0 else "rows"
        if ":" in arg:
            mode = arg.split(":")[1]                     # e.g. cfg3:rows forces the row split
        f1, f2, c = inputs(B, 256, H, W, seed=7)
        blk = sb.ShardedCorrBlock1D(f1, f2, mode=mode)
        del f1, f2
        t_local = timed(lambda: blk(c))
        t_nccl = timed(lambda: blk(c, gather=True))
        t_fused = timed(lambda: blk(c, gather="fused"))
        # the same inside CUDA graphs of 8 lookups (no host launch cost), when the pieces can be captured
        def graph_time(fn):
            try:
                side = torch.cuda.Stream(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    fn(); torch.cuda.synchronize(); dist.barrier()
                    with torch.cuda.graph(g, stream=side):
                        keep = [fn() for _ in range(8)]
                torch.cuda.synchronize()
                return timed(g.replay, n=10) / 8
            except Exception as e:          # noqa: BLE001
                return float("nan")
        g_local = graph_time(lambda: blk(c))
        g_fused = graph_time(lambda: blk(c, gather="fused"))
        if rank == 0:
            mb = B * 36 * H * W * 4 / 1e6
            print(f"{name} ({mode}-sharded over {world} ranks, {mb:.1f} MB gathered per rank): shard lookup only "
                  f"{t_local:7.1f} us | lookup + NCCL all-gather {t_nccl:7.1f} us | fused multicast lookup "
                  f"{t_fused:7.1f} us   [in a CUDA graph: shard lookup {g_local:6.1f} us, fused {g_fused:6.1f} us = "
                  f"{mb / g_fused * 1e3:.0f} GB/s gathered per rank]", flush=True)
        del blk
        torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
