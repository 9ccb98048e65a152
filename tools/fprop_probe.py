"""Time igemm fprop for the dominant shapes under the diagnostic modes of ADNI_DEBUG_MODE."""
import os, sys, subprocess
import torch
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from multimodal_alzheimer_b200 import kernels as K
    dev = torch.device("cuda:0")
    N = 32
    for D, Cin, Cout, dil in [(32, 64, 64, 1), (16, 128, 128, 1), (16, 256, 256, 2), (16, 512, 512, 4)]:
        x = torch.randn((N, D, D, D, Cin), device=dev).to(torch.bfloat16)
        w = torch.randn((Cout, Cin, 3, 3, 3), device=dev) * 0.05
        oti, ito = K.weights_to_kernel_layout(w)
        for _ in range(2):
            y, st = K.conv3d_fprop(x, oti, None, 3, 1, dil, dil, stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y, st = K.conv3d_fprop(x, oti, None, 3, 1, dil, dil, stats=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2 * N * D ** 3 * Cout * Cin * 27
        print(f"  {Cin:4d}->{Cout:4d} @{D}^3 dil{dil}: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TF(alg)")
else:
    for mode, name in [(0, "normal"), (1, "no MMA issue (TMA + barriers only)"), (2, "no TMA (MMA on stale smem)"),
                       (3, "no epilogue stores")]:
        print(f"mode {mode}: {name}", flush=True)
        env = dict(os.environ, ADNI_DEBUG_MODE=str(mode))
        subprocess.run([sys.executable, __file__, "child"], env=env, timeout=300)
