"""Per-parameter gradient comparison: product vs fp32 oracle, and torch bf16-autocast oracle vs fp32 oracle."""
import copy
import sys

import torch

sys.path.insert(0, ".")
from tests._models import build_pair, oracle_step, product_step, synthetic_batch  # noqa: E402
from tests._util import rel_l2  # noqa: E402


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


kind = sys.argv[1] if len(sys.argv) > 1 else "anat"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
oracle, product = build_pair(kind, depth=depth)
mods = ("mri", "pet1451", "tabular")
batch = synthetic_batch(B, (S, S, S), 3, modalities=mods)
ac = copy.deepcopy(oracle).to(dev)  # identical weights
out_o = oracle_step(oracle, batch)
out_p = product_step(product, batch, dev)
# autocast bf16 oracle on GPU
ac.train()
b = {k: v.to(dev) for k, v in batch.items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    yh = ac(b["mri"].unsqueeze(1).float())
loss = ac.criterion(yh.double(), b["label"])
loss.backward()
print("loss oracle %.6f product %.6f autocast %.6f" % (float(out_o["loss"]), float(out_p["loss"]), float(loss)))
print("logits rel: product %.3e autocast %.3e" % (rel_l2(out_p["outputs"].cpu(), out_o["outputs"]),
                                                    rel_l2(yh.double().cpu(), out_o["outputs"])))
po = dict(oracle.named_parameters())
pa = dict(ac.named_parameters())
print(f"{'param':45s} {'cos_prod':>9s} {'rel_prod':>9s} {'cos_ac':>9s} {'rel_ac':>9s}")
for n, p in product.named_parameters():
    r = po[n].grad
    if r is None or p.grad is None:
        continue
    g = p.grad.cpu()
    a = pa[n].grad.float().cpu()
    print(f"{n:45s} {cos(g, r):9.5f} {rel_l2(g, r):9.2e} {cos(a, r):9.5f} {rel_l2(a, r):9.2e}")
