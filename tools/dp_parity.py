"""N-rank sharded step == 1-rank full-batch step (sync-BN statistics, global loss normaliser, gradient sum).
Run under torchrun with 2+ ranks; rank 0 also runs the full batch alone (with the process group temporarily
bypassed) and compares loss / logits / gradients / running statistics."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import autograd as A  # noqa: E402
from multimodal_alzheimer_b200 import data_parallel as dp  # noqa: E402
from tests._models import build_pair, synthetic_batch  # noqa: E402
from tests._util import rel_l2  # noqa: E402

rank, local_rank, world = dp.init_from_env()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
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
elif rank == 0:
    print("peer all-reduce unavailable: NCCL path")
B = 4 * world
batch = synthetic_batch(B, (48, 48, 48), 3, modalities=("mri", "pet1451"))
_, model = build_pair("anat_pet_2resnet", depth=10, fl_gamma=None)   # weighted CE: exercises the global normaliser
model.to(dev).train()
lo, hi = dp.shard_bounds(B, rank, world)
shard = {k: v[lo:hi].to(dev) for k, v in batch.items()}
out = model.general_step(shard, 0, "train")
out["loss"].backward()
params = [p for p in model.parameters() if p.requires_grad]
dp.GradientBuckets(params).all_reduce()
torch.cuda.synchronize()
sharded = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
sharded_loss = float(out["loss"].detach())
sharded_rm = model.model_mri.model.layer4[0].bn2.running_mean.detach().clone()
logits = [torch.zeros_like(out["outputs"].detach()) for _ in range(world)]
dist.all_gather(logits, out["outputs"].detach().contiguous())
dist.barrier()
if rank == 0:
    # single-process reference on the full batch: disable the collectives inside the Functions
    A._world = lambda: 1
    A._allreduce_ = lambda t: t
    _, ref = build_pair("anat_pet_2resnet", depth=10, fl_gamma=None)
    ref.to(dev).train()
    full = {k: v.to(dev) for k, v in batch.items()}
    o = ref.general_step(full, 0, "train")
    o["loss"].backward()
    torch.cuda.synchronize()
    print("loss sharded %.9f full %.9f" % (sharded_loss, float(o["loss"].detach())))
    print("logits rel-L2 %.3e" % rel_l2(torch.cat(logits), o["outputs"].detach()))
    worst = max((rel_l2(sharded[n], p.grad), n) for n, p in ref.named_parameters() if p.grad is not None)
    print("worst gradient rel-L2 %.3e (%s)" % worst)
    print("running_mean rel-L2 %.3e" % rel_l2(sharded_rm, ref.model_mri.model.layer4[0].bn2.running_mean))
    ok = abs(sharded_loss - float(o["loss"].detach())) < 1e-3 and worst[0] < 5e-2
    print("DP PARITY", "OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
