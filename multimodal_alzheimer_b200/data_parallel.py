"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The sharded step equals the single-process full-batch step through three exchanges:
BatchNorm statistic sums and the loss normaliser are all-reduced inside the autograd Functions
(multimodal_alzheimer_b200/autograd.py); this module sums the parameter gradients in a few large buckets.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment. Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def shard_bounds(global_batch, rank, world):
    """rank r owns samples [r*B/N, (r+1)*B/N) (SURVEY.md §8e 'Partitioning')."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


class GradientBuckets:
    """Sum parameter gradients across ranks in ~bucket_mb buckets, last-produced gradients first (the order the
    backward pass finishes them), using flat fp32 staging buffers."""

    def __init__(self, params, bucket_mb=64):
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_mb * (1 << 20):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)

    def all_reduce(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        works = []
        for bucket in self.buckets:
            grads = [p.grad for p in bucket if p.grad is not None]
            if not grads:
                continue
            flat = torch._utils._flatten_dense_tensors(grads)
            works.append((dist.all_reduce(flat, async_op=True), flat, grads))
        for work, flat, grads in works:
            work.wait()
            for g, synced in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(synced)
