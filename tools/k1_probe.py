"""Time the tap-per-box engine on the ResNet-50 1x1x1 conv shapes under the diagnostic modes of ADNI_DEBUG_MODE
(1 = no MMA issue, 2 = no TMA loads, 3 = no epilogue stores): which part of the kernel is the tile rate made of?"""
import os, sys, subprocess
import torch
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from multimodal_alzheimer_b200 import kernels as K
    dev = torch.device("cuda:0")
    N = 8
    for (D, H, W, Cin, Cout) in [(40, 48, 40, 64, 256), (40, 48, 40, 256, 64), (20, 24, 20, 256, 1024), (20, 24, 20, 1024, 256),
                                 (20, 24, 20, 512, 2048)]:
        x = torch.randn((N, D, H, W, Cin), device=dev).to(torch.bfloat16)
        w = torch.randn((Cout, Cin, 1, 1, 1), device=dev) * 0.05
        oti, ito = K.weights_to_kernel_layout(w)
        for stats in (True, False):
            for _ in range(2):
                y, st = K.conv3d_fprop(x, oti, None, 1, 1, 0, 1, stats=stats)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y, st = K.conv3d_fprop(x, oti, None, 1, 1, 0, 1, stats=stats)
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / 10
            mb = (x.numel() + y.numel()) * 2 / 1e6
            print(f"  {Cin:4d}->{Cout:4d} @{D}x{H}x{W} stats={int(stats)}: {us:7.1f} us  {mb / us * 1e-3:6.2f} TB/s of HBM-algorithmic bytes", flush=True)
else:
    for mode, name in [(0, "normal"), (1, "no MMA issue (TMA + barriers + epilogue)"), (2, "no TMA (MMA on stale smem)"),
                       (3, "no epilogue stores (generic epilogue path)")]:
        print(f"mode {mode}: {name}", flush=True)
        env = dict(os.environ, ADNI_DEBUG_MODE=str(mode))
        subprocess.run([sys.executable, __file__, "child"], env=env, timeout=300)
