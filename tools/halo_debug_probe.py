"""Time the halo conv engine under the ADNI_HALO_DEBUG diagnostic bits (see conv_halo.cu)."""
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


for (N, S, C) in ((32, 32, 64), (32, 16, 128)):
    x = torch.randn((N, S, S, S, C), device=dev).to(torch.bfloat16)
    w = torch.randn((C, C, 3, 3, 3), device=dev) * 0.05
    oti, ito = K.weights_to_kernel_layout(w)
    flops = 2.0 * N * S ** 3 * C * C * 27
    for dbg, name in ((0, "normal"), (16, "rotated taps"), (1, "no MMA"), (2, "no weight TMA"), (4, "no plane TMA"),
                      (8, "no epilogue stores"), (6, "no TMA at all"), (7, "barriers only"), (18, "rot + no weight TMA")):
        os.environ["ADNI_HALO_DEBUG"] = str(dbg)
        t = timeit(lambda: K.conv3d_fprop(x, oti, None, 3, 1, 1, 1, stats=False))
        ts = timeit(lambda: K.conv3d_fprop(x, oti, None, 3, 1, 1, 1, stats=True))
        print(f"N{N} {S}^3 C{C} debug={dbg:2d} {name:22s}: {t:.3f} ms ({flops / t / 1e9:.0f} TF/s)  with stats {ts:.3f} ms", flush=True)
