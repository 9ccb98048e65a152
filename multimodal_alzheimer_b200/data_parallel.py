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
# ADNI_PEER_REDUCE: "1" (default) = the one-shot NVLink kernel, and if symmetric memory cannot be set up on this box the
# exchange goes through NCCL with a warning on EVERY rank's stderr and `exchange_status()` says so (bench.py prints
# it in its JSON line - never a silent switch); "require" = raise instead; "0" = NCCL for everything.
# Measured on 2 x B200: bit-identical to NCCL on 200 test vectors, +7 % volumes/s.  The spinning exchange CTA can
# delay, not block, a persistent conv grid: conv CTAs own disjoint work ranges, so the CTA that finds the SM taken
# runs on the first SM another CTA of its grid vacates.
_PEER_MODE = os.environ.get("ADNI_PEER_REDUCE", "1")
_PEER_DISABLED = [_PEER_MODE == "0"]
_PEER_STATUS = {"mode": "nccl (ADNI_PEER_REDUCE=0)" if _PEER_MODE == "0" else "unused", "error": None}


def exchange_status():
    """Which path the sync-BN / loss-normaliser sums take: 'peer' (csrc/peer_reduce.cu), 'nccl (...)' with the reason,
    or 'unused' (single rank so far)."""
    return dict(_PEER_STATUS)


def peer_reducer(channel, device):
    """The PeerReducer of a channel (None = default group, else one of the per-branch groups), created on first use;
    None when disabled or when symmetric memory is unavailable (callers use NCCL; see ADNI_PEER_REDUCE above)."""
    if _PEER_DISABLED[0] or not torch.cuda.is_available():
        return None
    key = (id(channel) if channel is not None else 0, device.index)
    red = _PEER_REDUCERS.get(key)
    if red is None:
        if torch.cuda.is_current_stream_capturing():
            return None  # collective construction cannot be captured: this call goes through NCCL
        try:
            red = _PEER_REDUCERS[key] = PeerReducer(device)
            _PEER_STATUS["mode"] = "peer"
        except Exception as e:  # noqa: BLE001 - reported, never silent
            if _PEER_MODE == "require":
                raise
            _PEER_DISABLED[0] = True
            _PEER_STATUS.update(mode=f"nccl (peer all-reduce unavailable: {type(e).__name__})", error=str(e))
            import sys
            print(f"[adni_b200 rank {dist.get_rank()}] WARNING: one-shot NVLink all-reduce unavailable "
                  f"({type(e).__name__}: {e}); sync-BN sums go through NCCL", file=sys.stderr, flush=True)
            return None
    return red


# Gradient slots: parameter storage pointer -> (view into a flat bucket buffer, weakref to the parameter, bucket object).
# The conv autograd Functions ask `grad_slot(w)` where to WRITE a weight gradient: with data-parallel buckets active the
# wgrad layout kernel stores straight into the bucket (autograd then adopts that view as `.grad`), so that the
# all-reduce needs neither a flatten nor an unflatten copy of the 265 MB of conv gradients.
_GRAD_SLOTS = {}
_GRAD_SLOTS_ON = os.environ.get("ADNI_GRAD_SLOTS", "1") != "0"   # 0: gradients are gathered into the buckets by copy (A/B)


def grad_slot(w):
    """The bucket view a gradient of parameter storage `w` should be written into, or None (no buckets, the
    parameter already holds a gradient that autograd would accumulate into, or the slot was already handed out in
    this backward pass - a weight used twice)."""
    entry = _GRAD_SLOTS.get(w.data_ptr() if isinstance(w, torch.Tensor) else w) if _GRAD_SLOTS_ON else None
    if entry is None:
        return None
    view, pref, owner = entry
    param = pref()
    if param is None or param.grad is not None or id(view) in owner._handed_out:
        return None
    owner._handed_out.add(id(view))
    return view.view(param.shape)      # a fresh tensor object: autograd adopts it (use count 1) instead of cloning


class _PeerArena:
    """A flat fp32 gradient arena in symmetric memory plus the flag buffer / counters of csrc/peer_reduce.cu's two-shot
    NVLink all-reduce (adni_peer_allreduce_f32).  Construction is collective (all ranks, same order)."""

    def __init__(self, numel, device):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.flat = symm.empty(numel, dtype=torch.float32, device=device)
        self.flat.zero_()
        self.flags = symm.empty(int(lib.adni_peer_grad_flag_bytes()) // 8, dtype=torch.int64, device=device)
        self.flags.zero_()
        self._h_data = symm.rendezvous(self.flat, dist.group.WORLD)
        self._h_flags = symm.rendezvous(self.flags, dist.group.WORLD)
        self.data_peers = torch.tensor([int(p) for p in self._h_data.buffer_ptrs], dtype=torch.int64, device=device)
        self.flag_peers = torch.tensor([int(p) for p in self._h_flags.buffer_ptrs], dtype=torch.int64, device=device)
        self.counter = torch.zeros(1, dtype=torch.int64, device=device)
        self.arrive = torch.zeros(1, dtype=torch.int32, device=device)
        # CTAs of 256 threads: two per SM keep (world x 4) float4 loads per thread in flight on every SM (measured on
        # 2 x B200: 16 / 32 / 64 CTAs = 9.64 / 8.90 / 8.60 ms per step at 4 pairs per rank - the kernel is latency-bound)
        self.ctas = int(os.environ.get("ADNI_PEER_GRAD_CTAS", str(2 * torch.cuda.get_device_properties(device).multi_processor_count)))
        torch.cuda.synchronize(device)
        dist.barrier()  # every buffer is zeroed before any rank touches a peer

    def all_reduce_range(self, start, end):
        from . import _lib
        _lib.call("adni_peer_allreduce_f32", start, end - start, _lib.ptr(self.data_peers), _lib.ptr(self.flag_peers),
                  _lib.ptr(self.counter), _lib.ptr(self.arrive), self.rank, self.world, self.ctas, _lib.stream_ptr())


# ADNI_PEER_GRADS: "1" (default) = gradient buckets live in symmetric memory and are summed by our own two-shot NVLink
# kernel; if symmetric memory cannot be set up the exchange goes through NCCL with a warning on every rank's stderr
# and `gradient_exchange_status()` says so; "require" = raise instead; "0" = NCCL.
_PEER_GRADS_MODE = os.environ.get("ADNI_PEER_GRADS", "1")
_GRAD_STATUS = {"mode": "unused", "error": None}


def gradient_exchange_status():
    return dict(_GRAD_STATUS)


class GradientBuckets:
    """Sum parameter gradients across ranks in ~bucket_mb buckets, last-produced gradients first (the order the
    backward pass finishes them).  All gradients of a (device, dtype) group share one persistent flat ARENA (in
    symmetric memory when the NVLink kernel is used); a bucket is a contiguous range of it.  A gradient that already
    lives in its slot (written there by the wgrad kernels through `grad_slot`) is reduced in place, the others are
    gathered with one multi-tensor copy, and after the all-reduce `p.grad` is REBOUND to the slot view - there is no
    copy back."""

    ALIGN = 64   # elements: 256-byte slot alignment (the Adam kernel reads gradients as float4)

    def __init__(self, params, bucket_mb=64):
        self.params = [p for p in params if p.requires_grad]
        self._active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.buckets, self._views, self._ranges, self._arenas = [], [], [], []
        self._handed_out = set()
        groups = {}
        for p in reversed(self.params):
            groups.setdefault((p.device, p.dtype), []).append(p)
        for (device, dtype), plist in groups.items():
            offs, total = [], 0
            for p in plist:
                offs.append(total)
                total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            ai = len(self._arenas)
            self._arenas.append(self._make_arena(total, dtype, device) if self._active else None)
            limit = bucket_mb * (1 << 20) / plist[0].element_size()
            cur, start = [], 0
            for i, (p, o) in enumerate(zip(plist, offs)):
                cur.append((p, o))
                end = offs[i + 1] if i + 1 < len(plist) else total
                if end - start >= limit or i + 1 == len(plist):
                    self.buckets.append([q for q, _ in cur])
                    self._ranges.append((ai, start, end))
                    flat = self._arenas[ai]["flat"] if self._arenas[ai] else None
                    self._views.append([flat[oo:oo + q.numel()] for q, oo in cur] if flat is not None else None)
                    cur, start = [], end
        if self._active:
            import weakref
            for bucket, views in zip(self.buckets, self._views):
                for p, v in zip(bucket, views):
                    _GRAD_SLOTS[p.data_ptr()] = (v, weakref.ref(p), self)

    @staticmethod
    def _make_arena(numel, dtype, device):
        if device.type == "cuda" and dtype == torch.float32 and _PEER_GRADS_MODE != "0" and not _GRAD_STATUS.get("disabled"):
            try:
                peer = _PeerArena(numel, device)
                _GRAD_STATUS["mode"] = "peer (two-shot NVLink kernel)"
                return {"flat": peer.flat, "peer": peer}
            except Exception as e:  # noqa: BLE001 - reported, never silent
                if _PEER_GRADS_MODE == "require":
                    raise
                _GRAD_STATUS.update(mode=f"nccl (peer gradient all-reduce unavailable: {type(e).__name__})", error=str(e),
                                    disabled=True)
                import sys
                print(f"[adni_b200 rank {dist.get_rank()}] WARNING: NVLink gradient all-reduce unavailable "
                      f"({type(e).__name__}: {e}); gradient buckets go through NCCL", file=sys.stderr, flush=True)
        if _GRAD_STATUS["mode"] == "unused":
            _GRAD_STATUS["mode"] = "nccl" if device.type == "cuda" else str(dist.get_backend())
        return {"flat": torch.zeros(numel, dtype=dtype, device=device), "peer": None}

    def _gather(self, bi):
        """Bring the bucket's gradients into its slots; returns [(param, slot view)] of those with a gradient."""
        have, src, dst = [], [], []
        for p, v in zip(self.buckets[bi], self._views[bi]):
            if p.grad is None:
                continue
            have.append((p, v))
            if p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad.reshape(-1))
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        return have

    def _reduce(self, bi):
        """Start the all-reduce of bucket bi (its gradients are gathered); returns a work handle or None (stream-ordered)."""
        ai, start, end = self._ranges[bi]
        arena = self._arenas[ai]
        if arena["peer"] is not None:
            arena["peer"].all_reduce_range(start, end)
            return None
        return dist.all_reduce(arena["flat"][start:end], async_op=True)

    @staticmethod
    def _rebind(have):
        for p, v in have:
            if p.grad.data_ptr() != v.data_ptr():
                p.grad = v.view(p.shape)

    def all_reduce(self):
        if not self._active:
            return
        works = []
        for bi in range(len(self.buckets)):
            have = self._gather(bi)
            if have:
                works.append((self._reduce(bi), have))
        for work, have in works:
            if work is not None:
                work.wait()
            self._rebind(have)
        self._handed_out.clear()


class _EventWork:
    """work.wait() for a stream-ordered exchange issued on a side stream: the current stream waits for its event."""

    def __init__(self, event, device):
        self.event, self.device = event, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_event(self.event)


class OverlappedGradientBuckets(GradientBuckets):
    """The same buckets, but a bucket's exchange starts as soon as the backward pass has produced its last gradient
    (`register_post_accumulate_grad_hook`), so that the 265 MB of gradient traffic of the two-encoder model overlaps
    the remaining dgrad / wgrad kernels instead of following them.  bench.py's choice: +1.3 % at 8 GPUs with the NVLink
    kernel on a side stream; SLOWER with NCCL, whose CTAs take whole SMs away from the persistent conv grids.
    `all_reduce()` after `backward()` then only launches what is still pending (parameters without a gradient), joins
    the exchanges and rebinds `p.grad` to the bucket slots.

    Stream safety: a hook runs on the stream its gradient was produced on (the two encoder branches use two streams);
    every hook records an event, and the stream that launches a bucket first waits for the events of all of the
    bucket's gradients.  Collective order: buckets complete in the autograd engine's (deterministic, rank-independent)
    execution order.  One backward pass per optimizer step (no gradient accumulation across backward calls)."""

    def __init__(self, params, bucket_mb=64):
        super().__init__(params, bucket_mb)
        self._bucket_of = {}
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._bucket_of[id(p)] = bi
        self._side = {}
        self._reset()
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params] if self._active else []

    def _reset(self):
        self._seen = [0] * len(self.buckets)
        self._events = [[] for _ in self.buckets]
        self._launched = [None] * len(self.buckets)

    def _launch(self, bi):
        if not any(p.grad is not None for p in self.buckets[bi]):
            self._launched[bi] = False
            return
        dev = self.buckets[bi][0].device
        ai = self._ranges[bi][0]
        if dev.type == "cuda" and self._arenas[ai]["peer"] is not None:
            # NVLink kernel: stream-ordered, so it runs on a SIDE stream - the backward kernels of the producing
            # streams go on while the bucket crosses the links (the kernel has no shared memory and co-resides with
            # the persistent conv grids); all_reduce() joins the side stream
            side = self._side.get(dev.index)
            if side is None:
                side = self._side[dev.index] = torch.cuda.Stream(device=dev)
            for ev in self._events[bi]:
                side.wait_event(ev)
            with torch.cuda.stream(side):
                have = self._gather(bi)
                self._reduce(bi)
                done = torch.cuda.Event()
                done.record(side)
            self._launched[bi] = (_EventWork(done, dev), have)
            return
        if dev.type == "cuda":
            cur = torch.cuda.current_stream(dev)
            for ev in self._events[bi]:
                cur.wait_event(ev)
        have = self._gather(bi)
        self._launched[bi] = (self._reduce(bi), have)

    def _on_grad(self, p):
        bi = self._bucket_of[id(p)]
        if self._launched[bi] is not None:
            raise RuntimeError("OverlappedGradientBuckets: a second backward pass before all_reduce() "
                               "(gradient accumulation is not supported)")
        if p.grad is not None and p.grad.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(p.grad.device))
            self._events[bi].append(ev)
        self._seen[bi] += 1
        if self._seen[bi] == len(self.buckets[bi]):
            self._launch(bi)

    def all_reduce(self):
        if not self._active:
            return
        for bi in range(len(self.buckets)):
            if self._launched[bi] is None:
                self._launch(bi)            # parameters of this bucket that took no part in the backward pass
        for item in self._launched:
            if not item:
                continue
            work, have = item
            if work is not None:
                work.wait()
            self._rebind(have)
        self._handed_out.clear()
        self._reset()

    def remove_hooks(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def make_gradient_buckets(params, bucket_mb=64, overlap=None):
    """GradientBuckets, or the variant that starts a bucket's exchange from the backward pass (one backward pass per
    optimizer step: no gradient accumulation across backward calls).  `overlap=None`: ADNI_OVERLAP_GRADS decides
    (default off); ADNI_OVERLAP_GRADS=0 / 1 overrides the argument.  Measured on 8 x B200 (profiles/r02_bench_dp8*.json):
    NCCL after backward 9.22 ms per step, NVLink kernel after backward 9.15, NVLink kernel on a side stream during
    backward 9.04."""
    env = os.environ.get("ADNI_OVERLAP_GRADS")
    if env is not None:
        overlap = env == "1"
    if overlap:
        return OverlappedGradientBuckets(params, bucket_mb)
    return GradientBuckets(params, bucket_mb)
