"""Probe of torch symmetric memory on this box (run under torchrun): peer pointers, multicast pointer, barrier."""
import os
import sys
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
hdl.barrier()
dist.destroy_process_group()
