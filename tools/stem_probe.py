"""Time the plane-resident stem fprop under the ADNI_STEM_DEBUG diagnostic bits, and stem wgrad."""
import os
import sys
import torch
sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import kernels as K

dev = torch.device("cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


N, S = 32, 128
x = torch.randn((N, S, S, S, 1), device=dev).to(torch.bfloat16)
w = torch.randn((64, 1, 7, 7, 7), device=dev) * 0.05
x8 = K.stem_expand(x)
y, st = K.stem_fprop(x8, tuple(x.shape), w)
dy = torch.randn_like(y)
for dbg, name in ((0, "normal"), (1, "no BN sums"), (2, "no stores"), (3, "no sums, no stores"), (4, "no MMA"), (7, "skeleton")):
    os.environ["ADNI_STEM_DEBUG"] = str(dbg)
    t = timeit(lambda: K.stem_fprop(x8, tuple(x.shape), w))
    print(f"stem fprop debug={dbg} {name:20s}: {t:.3f} ms", flush=True)
os.environ["ADNI_STEM_DEBUG"] = "0"
print(f"stem wgrad: {timeit(lambda: K.stem_wgrad(x8, dy, tuple(x.shape))):.3f} ms")
for planes in ("0",):
    os.environ["ADNI_STEM_PLANES"] = planes
    print(f"planes={planes}: fprop {timeit(lambda: K.stem_fprop(x8, tuple(x.shape), w)):.3f} ms, wgrad {timeit(lambda: K.stem_wgrad(x8, dy, tuple(x.shape))):.3f} ms")
