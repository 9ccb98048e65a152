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


class PeerReducer:
    """One-shot NVLink all-reduce for the small fp64 vectors of synchronised BatchNorm and the loss normaliser
    (csrc/peer_reduce.cu).  PyTorch symmetric memory provides the peer-mapped buffers (plumbing); the exchange itself
    is one CTA of our own kernel per call.  Construction is collective (all ranks, same order) and must happen outside
    CUDA-graph capture; one instance per concurrent stream of calls (fusion models: one per encoder branch)."""

    MAX_N = 4096  # doubles per call: 2 x C for C <= 2048

    def __init__(self, device):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        nbytes = int(_lib.load().adni_peer_buffer_bytes(self.world, self.MAX_N))
        self.buf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.counter = torch.zeros(1, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()  # every buffer is zeroed before any rank stores into a peer

    def usable(self, t):
        return t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 0 < t.numel() <= self.MAX_N

    def all_reduce_(self, t):
        from . import _lib
        _lib.call("adni_peer_allreduce_f64", _lib.ptr(t), t.numel(), _lib.ptr(self.peers), _lib.ptr(self.counter),
                  self.rank, self.world, self.MAX_N, _lib.stream_ptr())
        return t


_PEER_REDUCERS = {}
# On by default (ADNI_PEER_REDUCE=0 selects NCCL).  Measured on 2 x B200: bit-identical to NCCL on 200 test vectors,
# 2252-2264 volumes/s against 2106 through NCCL (+7 %), 4 of 4 runs clean.  (Two earlier failures next to it were the
# barrier-phase aliasing bug of the dual-issuer conv kernel, fixed in conv_halo.cu.)  The spinning exchange CTA can
# delay, not block, a persistent conv grid: conv CTAs own disjoint work ranges, so the CTA that finds the SM taken
# runs on the first SM another CTA of its grid vacates.
_PEER_DISABLED = [os.environ.get("ADNI_PEER_REDUCE", "1") == "0"]


def peer_reducer(channel, device):
    """The PeerReducer of a channel (None = default group, else one of the per-branch groups), created on first use;
    None when symmetric memory is unavailable (callers fall back to NCCL)."""
    if _PEER_DISABLED[0] or not torch.cuda.is_available():
        return None
    key = (id(channel) if channel is not None else 0, device.index)
    red = _PEER_REDUCERS.get(key)
    if red is None:
        if torch.cuda.is_current_stream_capturing():
            return None  # collective construction cannot be captured: this call goes through NCCL
        try:
            red = _PEER_REDUCERS[key] = PeerReducer(device)
        except Exception as e:  # noqa: BLE001 - optional fast path
            _PEER_DISABLED[0] = True
            if dist.get_rank() == 0:
                print(f"[adni_b200] peer all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)
            return None
    return red


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
