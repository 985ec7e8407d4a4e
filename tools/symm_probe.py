"""Feasibility: torch symmetric memory on this box -- peer pointers, NVSwitch multicast (NVLS) support."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty((1024,), dtype=torch.float32, device=dev)
t.fill_(float(rank))
hdl = symm.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: backend={symm.get_backend(dev)} world={hdl.world_size} buffer_ptrs={[hex(p) for p in hdl.buffer_ptrs]} "
      f"multicast_support={hdl.has_multicast_support} multicast_ptr={hex(hdl.multicast_ptr) if hdl.multicast_ptr else None}", flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(f"rank {rank}: peer buffer value {peer[0].item()} (expected {(rank + 1) % world})", flush=True)
hdl.barrier()
dist.destroy_process_group()
