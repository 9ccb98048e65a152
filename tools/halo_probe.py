"""Halo-resident conv engine probe: parity vs the tap-per-box engine + torch, and timing, for one
(ADNI_HALO_PITCH, ADNI_HALO_BASE) variant given in the environment."""
import os
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import kernels as K
from tests._util import to_ncdhw_f32, to_ndhwc_bf16

dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def run(N, D, H, W, C, check_torch):
    g = torch.Generator().manual_seed(3)
    x = torch.randn((N, C, D, H, W), generator=g).to(dev)
    w = (torch.randn((C, C, 3, 3, 3), generator=g) / (C * 27) ** 0.5).to(dev)
    xb = to_ndhwc_bf16(x)
    oti, ito = K.weights_to_kernel_layout(w)
    dy = to_ndhwc_bf16(torch.randn((N, C, D, H, W), generator=g).to(dev))
    add = to_ndhwc_bf16(torch.randn((N, C, D, H, W), generator=g).to(dev))
    os.environ["ADNI_HALO"] = "0"
    y0, st0 = K.conv3d_fprop(xb, oti, None, 3, 1, 1, 1, stats=True)
    dx0 = K.conv3d_dgrad(dy, ito, tuple(xb.shape), 3, 1, 1, 1, addend=add)
    torch.cuda.synchronize()
    os.environ["ADNI_HALO"] = "1"
    y1, st1 = K.conv3d_fprop(xb, oti, None, 3, 1, 1, 1, stats=True)
    dx1 = K.conv3d_dgrad(dy, ito, tuple(xb.shape), 3, 1, 1, 1, addend=add)
    torch.cuda.synchronize()
    msg = f"shape N{N} {D}x{H}x{W} C{C}: fprop rel {rel(y1, y0):.2e} stats rel {rel(st1[1], st0[1]):.2e} dgrad rel {rel(dx1, dx0):.2e}"
    if check_torch:
        ref = F.conv3d(to_ncdhw_f32(xb), w.to(torch.bfloat16).float(), None, 1, 1, 1)
        msg += f" | vs torch: halo {rel(to_ncdhw_f32(y1), ref):.2e} base {rel(to_ncdhw_f32(y0), ref):.2e}"
    print(msg, flush=True)
    return xb, oti, ito, dy


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


print("variant pitch", os.environ.get("ADNI_HALO_PITCH"), "base", os.environ.get("ADNI_HALO_BASE"), flush=True)
run(1, 16, 16, 16, 64, True)
run(2, 9, 23, 28, 64, True)
run(1, 12, 14, 12, 128, True)
for (N, S, C) in ((32, 32, 64), (32, 16, 128)):
    xb, oti, ito, dy = run(N, S, S, S, C, False)
    flops = 2.0 * N * S ** 3 * C * C * 27
    for halo in ("0", "1"):
        os.environ["ADNI_HALO"] = halo
        tf = timeit(lambda: K.conv3d_fprop(xb, oti, None, 3, 1, 1, 1, stats=True))
        td = timeit(lambda: K.conv3d_dgrad(dy, ito, tuple(xb.shape), 3, 1, 1, 1))
        print(f"  N{N} {S}^3 C{C} halo={halo}: fprop {tf:.3f} ms {flops / tf / 1e9:.0f} TF/s | dgrad {td:.3f} ms {flops / td / 1e9:.0f} TF/s", flush=True)
