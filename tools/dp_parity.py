"""N-rank sharded training step vs the CPU ORACLE on the full batch (SURVEY.md 8(e) 'Parity test').

Run under torchrun with 2+ ranks (tests/test_gpu_multirank.py launches it when the box has >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_parity.py [--workload pet_mri_fusion_r18] [--volume 64] [--per-rank 2] [--depth 18]

Every rank runs the CUDA product on its shard of ONE global batch (sync-BN statistics, global loss normaliser and
gradient buckets exchanged as in training).  Rank 0 runs the oracle (plain torch.nn, fp32, CPU) on the whole batch and
compares: loss, the gathered logits, EVERY parameter gradient after the bucket all-reduce, and every BatchNorm running
statistic - at the module tolerances of tests/test_gpu_models.py (logits / loss 2e-2, gradients / statistics 3e-2
rel-L2, each OR within 2x of the error of PyTorch's own bf16-autocast run of the oracle on the same batch).  It also
checks the one-shot NVLink all-reduce against NCCL and that all ranks end with bit-identical gradients.
"""
import argparse
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_alzheimer_b200 import data_parallel as dp  # noqa: E402
from multimodal_alzheimer_b200 import workloads as W  # noqa: E402
from tests._util import rel_l2  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="pet_mri_fusion_r18")
ap.add_argument("--volume", type=int, default=64)
ap.add_argument("--per-rank", type=int, default=2)
ap.add_argument("--depth", type=int, default=None)
args = ap.parse_args()

rank, local_rank, world = dp.init_from_env()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# --- the one-shot NVLink all-reduce against NCCL, on its own (200 calls of varying length on one channel) ---------
red = dp.peer_reducer(None, dev)
if red is not None:
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    worst = 0.0
    for it in range(200):
        n = [2, 128, 1024, 4096, 250][it % 5]
        v = torch.randn(n, generator=g, dtype=torch.float64).to(dev) * (10.0 ** (it % 7))
        ref = v.clone()
        dist.all_reduce(ref)
        out = red.all_reduce_(v.clone())
        worst = max(worst, float((out - ref).abs().max() / ref.abs().max()))
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree bitwise"
    torch.cuda.synchronize()
    if rank == 0:
        print(f"peer all-reduce vs NCCL: worst rel diff {worst:.2e} over 200 calls, ranks bit-identical")
    assert worst < 1e-14

# --- the two-shot NVLink gradient all-reduce against NCCL, on its own (ragged ranges of one symmetric arena) --------
if dp._PEER_GRADS_MODE != "0":
    arena = dp._PeerArena(3_000_000, dev)
    g = torch.Generator(device="cpu").manual_seed(200 + rank)
    worst = 0.0
    for it, (start, n) in enumerate([(0, 4), (64, 2_000_000), (0, 3_000_000), (1_000_000, 1_234_568), (128, 64), (4, 36)] * 4):
        v = torch.randn(n, generator=g, dtype=torch.float32).to(dev) * (10.0 ** (it % 5))
        arena.flat.zero_()
        arena.flat[start:start + n].copy_(v)
        guard = arena.flat.clone()
        ref = v.clone()
        dist.all_reduce(ref)
        dist.barrier()
        torch.cuda.synchronize()
        arena.all_reduce_range(start, start + n)
        out = arena.flat[start:start + n].clone()
        worst = max(worst, float((out - ref).abs().max() / ref.abs().max()))
        guard[start:start + n] = out
        assert torch.equal(guard, arena.flat), "the kernel wrote outside its range"
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree bitwise"
    torch.cuda.synchronize()
    if rank == 0:
        print(f"peer gradient all-reduce vs NCCL: worst rel diff {worst:.2e} over 24 ranges, ranks bit-identical")
    assert worst < 1e-5
    del arena

w = W.WORKLOADS[args.workload]
volume = (args.volume,) * 3
B = args.per_rank * world
lo, hi = dp.shard_bounds(B, rank, world)

# identical weights everywhere: the oracle's random init (seed 15), loaded into the product through state_dict
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tests.test_gpu_baseline_sizes import _autocast_step, _oracle_ns, oracle_batch  # noqa: E402

ns = _oracle_ns()
oracle = W.build_model(ns, args.workload, depth=args.depth)
product = W.build_model(W.product_namespace(), args.workload, depth=args.depth)
product.load_state_dict(copy.deepcopy(oracle.state_dict()), strict=True)
product.to(dev).train()

raw = W.synth_batch(0, B, volume, w["modalities"])          # the whole global batch on every rank (CPU), sliced below
raw["label"][0], raw["label"][-1] = 0, 2
shard = {k: v[lo:hi].to(dev) for k, v in raw.items()}
params = [p for p in product.parameters() if p.requires_grad]
buckets = dp.make_gradient_buckets(params)   # before backward: the wgrad kernels write into the bucket slots (dp.grad_slot)
out = product.general_step(W.normalized_batch_gpu(shard), 0, "train")
out["loss"].backward()
in_slots = sum(1 for bi in range(len(buckets.buckets)) for p, v in zip(buckets.buckets[bi], buckets._views[bi])
               if p.grad is not None and p.grad.data_ptr() == v.data_ptr())
buckets.all_reduce()
torch.cuda.synchronize()
if rank == 0:
    print(f"gradient slots: {in_slots} of {len(params)} gradients were produced inside their bucket slot")

# all ranks must hold bit-identical gradients after the exchange
flat = torch.cat([p.grad.flatten() for p in params if p.grad is not None])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
identical = all(torch.equal(gathered[0], t) for t in gathered)
logits = [torch.zeros_like(out["outputs"].detach()) for _ in range(world)]
dist.all_gather(logits, out["outputs"].detach().contiguous())
dist.barrier()

ok = True
if rank == 0:
    ob = oracle_batch(raw, ns.tab_key)
    bracket, out_a = _autocast_step(oracle, ob, dev)
    out_o = oracle.general_step(ob, 0, "train")
    out_o["loss"].backward()
    lg = torch.cat(logits).cpu()
    e_logits = rel_l2(lg, out_o["outputs"].detach())
    a_logits = rel_l2(out_a["outputs"].detach().cpu(), out_o["outputs"].detach())
    e_loss = abs(float(out["loss"].detach()) - float(out_o["loss"].detach()))
    a_loss = abs(float(out_a["loss"].detach()) - float(out_o["loss"].detach()))
    print(f"{args.workload} depth={args.depth or w['depth']} {volume} global batch {B} over {world} ranks")
    print(f"loss sharded {float(out['loss'].detach()):.9f} oracle {float(out_o['loss'].detach()):.9f} (|diff| {e_loss:.2e}, "
          f"autocast {a_loss:.2e})")
    print(f"logits rel-L2 {e_logits:.3e} (autocast {a_logits:.3e})")
    ok &= e_logits <= max(2e-2, 2 * a_logits) and e_loss <= max(2e-2, 2 * a_loss)
    po, pa = dict(oracle.named_parameters()), dict(bracket.named_parameters())
    rows = []
    for n, p in product.named_parameters():
        q = po[n]
        if q.grad is None or float(q.grad.norm()) < 1e-12:
            continue
        e = rel_l2(p.grad.detach().cpu(), q.grad)
        ea = rel_l2(pa[n].grad.detach().float().cpu(), q.grad) if pa[n].grad is not None else 0.0
        rows.append((e - 2 * ea, e, ea, n))
    rows.sort(reverse=True)
    for _, e, ea, n in rows[:6]:
        print(f"  grad rel-L2 {e:.3e} (autocast {ea:.3e}) {n}")
    bad = [(e, ea, n) for _, e, ea, n in rows if e > max(3e-2, 2 * ea)]
    print(f"gradients: {len(rows)} tensors, worst {max(r[1] for r in rows):.3e}, outside tolerance: {len(bad)}")
    ok &= not bad
    bo, ba = dict(oracle.named_buffers()), dict(bracket.named_buffers())
    worst_rs = 0.0
    for n, b in product.named_buffers():
        if n.endswith("running_mean") or n.endswith("running_var"):
            e = rel_l2(b.detach().cpu(), bo[n])
            ea = rel_l2(ba[n].detach().float().cpu(), bo[n])
            worst_rs = max(worst_rs, e)
            if e > max(3e-2, 2 * ea):
                print(f"  running statistic {n}: {e:.3e} (autocast {ea:.3e})")
                ok = False
    print(f"running statistics worst rel-L2 {worst_rs:.3e}")
    print(f"gradients bit-identical on all ranks: {identical}")
    ok &= identical
    print("DP ORACLE PARITY", "OK" if ok else "FAILED")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
