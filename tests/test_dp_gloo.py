"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, bucketed gradient all-reduce, and the
statistic / loss-normaliser exchange that makes N shards equal the single-process full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from multimodal_alzheimer_b200 import data_parallel as dp
    r, lr, w = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _bucket_fn(rank, world):
    from multimodal_alzheimer_b200 import data_parallel as dp
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 1000, 3, 70000, 11)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    params[2].grad = None  # a frozen / unused parameter must be skipped, not crash
    b = dp.GradientBuckets(params, bucket_mb=0.2)
    assert len(b.buckets) >= 2
    b.all_reduce()
    return [None if p.grad is None else float(p.grad[0]) for p in params]


def _slot_fn(rank, world):
    """Gradient slots: a gradient produced INSIDE its bucket slot (what the wgrad layout kernel does through
    data_parallel.grad_slot) is adopted by autograd as `.grad`, reduced in place and never copied; gradients produced
    elsewhere are gathered; after all_reduce() every `.grad` is a view of the bucket's flat buffer; a parameter that
    already holds a gradient, or a slot handed out twice in one backward pass, falls back to ordinary allocation."""
    from multimodal_alzheimer_b200 import data_parallel as dp

    class SlotGrad(torch.autograd.Function):      # stands in for the conv Functions: writes dW where grad_slot says
        @staticmethod
        def forward(ctx, x, w):
            ctx.ptr, ctx.shape = w.data_ptr(), w.shape
            ctx.save_for_backward(x)
            return (x * w).sum()

        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            slot = dp.grad_slot(ctx.ptr)
            dw = x * g
            if slot is not None:
                slot.copy_(dw)
                dw = slot
            return None, dw

    torch.manual_seed(0)
    w_in = torch.nn.Parameter(torch.ones(1000))
    w_out = torch.nn.Parameter(torch.ones(333))
    w_twice = torch.nn.Parameter(torch.ones(50))
    params = [w_in, w_out, w_twice]
    b = dp.GradientBuckets(params, bucket_mb=0.002)
    res = {}
    for step in range(2):
        x = torch.full((1000,), float(rank + 1 + step))
        loss = SlotGrad.apply(x, w_in) + (w_out * 2.0 * (rank + 1)).sum() + SlotGrad.apply(x[:50], w_twice) \
            + SlotGrad.apply(x[:50], w_twice)
        loss.backward()
        flat_of = {id(p): (bi, v) for bi in range(len(b.buckets)) for p, v in zip(b.buckets[bi], b._views[bi])}
        res[f"in_place_before_{step}"] = w_in.grad.data_ptr() == flat_of[id(w_in)][1].data_ptr()
        b.all_reduce()
        res[f"views_{step}"] = all(p.grad.data_ptr() == flat_of[id(p)][1].data_ptr() for p in params)
        res[f"vals_{step}"] = [float(w_in.grad[0]), float(w_out.grad[0]), float(w_twice.grad[0])]
        for p in params:
            p.grad = None
    # a parameter that still holds a gradient gets no slot (autograd would accumulate into it)
    w_in.grad = torch.zeros_like(w_in)
    res["slot_with_grad"] = dp.grad_slot(w_in) is None
    return res


def test_gradient_slots_reduce_in_place():
    out = _run(_slot_fn)
    for rank in range(2):
        r = out[rank]
        for step in range(2):
            tot = sum(k + 1 + step for k in range(2))
            assert r[f"in_place_before_{step}"], "autograd did not adopt the slot view as .grad"
            assert r[f"views_{step}"]
            assert r[f"vals_{step}"] == [float(tot), 2.0 * 3, 2.0 * tot]
        assert r["slot_with_grad"]


def test_gradient_buckets_sum_across_ranks():
    out = _run(_bucket_fn)
    tot = sum(r + 1 for r in range(2))
    for rank in range(2):
        assert out[rank] == [tot * 1.0, tot * 2.0, None, tot * 4.0, tot * 5.0]


def _overlap_fn(rank, world):
    """OverlappedGradientBuckets == GradientBuckets on a small MLP: hooks fire during backward, one parameter is
    unused (its bucket is completed by all_reduce()), two steps in a row (state reset), double backward rejected."""
    from multimodal_alzheimer_b200 import data_parallel as dp
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(300, 200), torch.nn.ReLU(), torch.nn.Linear(200, 100), torch.nn.ReLU(),
                              torch.nn.Linear(100, 3))
    unused = torch.nn.Parameter(torch.zeros(7))
    params = list(net.parameters()) + [unused]
    ob = dp.OverlappedGradientBuckets(params, bucket_mb=0.05)
    assert len(ob.buckets) >= 2
    ok = True
    for step in range(2):
        g = torch.Generator().manual_seed(10 * step + rank)
        x = torch.randn(5, 300, generator=g)
        for p in params:
            p.grad = None
        net(x).square().sum().backward()
        launched_early = sum(1 for it in ob._launched if it)         # before all_reduce(): complete buckets are in flight
        local = [None if p.grad is None else p.grad.clone() for p in params]
        # reference: what every rank's gradient sums to
        gathered = [None] * world
        dist.all_gather_object(gathered, [None if t is None else t.tolist() for t in local])
        ob.all_reduce()
        for i, p in enumerate(params):
            if local[i] is None:
                ok = ok and p.grad is None
                continue
            want = sum(torch.tensor(gathered[r][i]) for r in range(world))
            ok = ok and torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7)
        ok = ok and launched_early >= 1 and launched_early < len(ob.buckets)   # the bucket holding `unused` waits
    net(torch.randn(2, 300)).sum().backward()
    try:
        net(torch.randn(2, 300)).sum().backward()
        ok = False
    except RuntimeError:
        pass
    ob.remove_hooks()
    return bool(ok)


def test_overlapped_gradient_buckets_equal_plain_buckets():
    out = _run(_overlap_fn)
    assert all(out[r] is True for r in range(2)), out


def _syncbn_fn(rank, world):
    """The reductions the autograd Functions perform (sum x, sum x^2 | loss numerator, normaliser) reproduce the
    full-batch BatchNorm statistics and weighted-CE loss when every rank holds a shard."""
    from multimodal_alzheimer_b200 import autograd as A
    from multimodal_alzheimer_b200 import data_parallel as dp
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 5, generator=g, dtype=torch.float64)          # full batch (identical on every rank)
    lo, hi = dp.shard_bounds(8, rank, world)
    xs = x[lo:hi]
    stats = torch.stack([xs.sum(0), (xs * xs).sum(0)])
    A._allreduce_(stats)
    count = xs.shape[0] * A._world()
    mean = stats[0] / count
    var = stats[1] / count - mean * mean
    ok_bn = torch.allclose(mean, x.mean(0)) and torch.allclose(var, x.var(0, unbiased=False))
    # weighted CE: numerator and sum of target weights are both global sums
    w = torch.tensor([0.4651162790697675, 0.6712473572938689, 0.8636363636363636], dtype=torch.float64)
    z = torch.randn(8, 3, generator=g, dtype=torch.float64)
    t = torch.randint(0, 3, (8,), generator=g)
    nll = -torch.log_softmax(z[lo:hi], 1).gather(1, t[lo:hi, None]).squeeze(1)
    partial = torch.stack([(w[t[lo:hi]] * nll).sum(), w[t[lo:hi]].sum()])
    A._allreduce_(partial)
    ref = torch.nn.functional.cross_entropy(z, t, weight=w)
    return bool(ok_bn), abs(float(partial[0] / partial[1]) - float(ref)) < 1e-14


def test_sharded_statistics_equal_full_batch():
    out = _run(_syncbn_fn)
    assert all(out[r] == (True, True) for r in range(2))


def test_shard_bounds():
    from multimodal_alzheimer_b200 import data_parallel as dp
    assert [dp.shard_bounds(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    assert dp.shard_bounds(16, 0, 1) == (0, 16)
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)
