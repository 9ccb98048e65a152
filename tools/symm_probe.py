"""Feasibility probe: torch symmetric memory across the ranks of one node (peer pointers for a custom all-reduce)."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import data_parallel as dp
rank, local_rank, world = dp.init_from_env()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(4096, dtype=torch.float64, device=dev)
    t.fill_(rank + 1)
    h = symm.rendezvous(t, dist.group.WORLD)
    torch.cuda.synchronize()
    dist.barrier()
    print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs][:2],
          "multicast", hex(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else None, flush=True)
    peer = h.get_buffer((rank + 1) % world, (4096,), torch.float64)
    print(rank, "peer value", float(peer[0]), flush=True)
    dist.barrier()
    print("SYMM OK")
except Exception as e:
    print(rank, "SYMM FAILED", type(e).__name__, e, flush=True)
