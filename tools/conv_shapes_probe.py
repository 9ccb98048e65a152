"""Per-shape timing of the ResNet-18 conv kernels (fprop / dgrad / wgrad), halo engine off and on."""
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


N = 32
for (S, C, dil) in ((32, 64, 1), (16, 128, 1), (16, 256, 2), (16, 512, 4)):
    x = torch.randn((N, S, S, S, C), device=dev).to(torch.bfloat16)
    dy = torch.randn((N, S, S, S, C), device=dev).to(torch.bfloat16)
    w = torch.randn((C, C, 3, 3, 3), device=dev) * 0.05
    oti, ito = K.weights_to_kernel_layout(w)
    flops = 2.0 * N * S ** 3 * C * C * 27
    for halo in ("0", "1"):
        if halo == "1" and C > 128:
            continue
        os.environ["ADNI_HALO"] = halo
        tf = timeit(lambda: K.conv3d_fprop(x, oti, None, 3, 1, dil, dil, stats=True))
        td = timeit(lambda: K.conv3d_dgrad(dy, ito, tuple(x.shape), 3, 1, dil, dil))
        tw = timeit(lambda: K.conv3d_wgrad(x, dy, 3, 1, dil, dil))
        print(f"N{N} {S}^3 C{C} dil{dil} halo={halo}: fprop {tf:.3f} ms {flops / tf / 1e9:5.0f} TF/s | dgrad {td:.3f} ms "
              f"{flops / td / 1e9:5.0f} TF/s | wgrad {tw:.3f} ms {flops / tw / 1e9:5.0f} TF/s", flush=True)
