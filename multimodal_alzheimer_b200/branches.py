"""Run the two independent encoder branches of a fusion model (PET trunk, MRI trunk) on two CUDA streams.

The branches share nothing until the feature concat, so their kernels may overlap: the HBM-bound BatchNorm / pooling
kernels of one branch run under the tensor-bound convolutions of the other, wave tails are filled, and (data
parallel) one branch computes while the other waits for its BatchNorm statistic all-reduce.  autograd replays every
backward node on the stream of its forward, so the backward pass overlaps the same way.  Each branch gets its own
NCCL communicator so that its reductions do not queue behind the other branch's.  Works eagerly and under CUDA-graph
capture (fork/join through stream waits).  ADNI_PARALLEL_BRANCHES=0 disables it.
"""
import os

import torch
import torch.distributed as dist

from . import autograd as A

_ENABLED = os.environ.get("ADNI_PARALLEL_BRANCHES", "1") != "0"
_side_streams = {}
_groups = None


def _branch_groups():
    global _groups
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return (None, None)
    if _groups is None:  # collective: every rank reaches this on its first fusion forward
        _groups = (dist.new_group(), dist.new_group())
    return _groups


def run_two(f_a, f_b, *inputs_a):
    """Returns (f_a(), f_b()) with f_a enqueued on a side stream. inputs_a: tensors f_a reads (for the allocator)."""
    ref = next((t for t in inputs_a if isinstance(t, torch.Tensor)), None)
    if not _ENABLED or ref is None or not ref.is_cuda:
        return f_a(), f_b()
    cur = torch.cuda.current_stream()
    key = (ref.device.index, cur.cuda_stream)
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device=ref.device)
    ga, gb = _branch_groups()
    side.wait_stream(cur)
    with torch.cuda.stream(side), A.dp_group(ga):
        for t in inputs_a:
            if isinstance(t, torch.Tensor):
                t.record_stream(side)
        out_a = f_a()
    with A.dp_group(gb):
        out_b = f_b()
    cur.wait_stream(side)
    if isinstance(out_a, torch.Tensor):
        out_a.record_stream(cur)
    return out_a, out_b
