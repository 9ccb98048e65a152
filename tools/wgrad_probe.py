import os, sys, subprocess
import torch
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from multimodal_alzheimer_b200 import kernels as K
    dev = torch.device("cuda:0")
    N = 32
    for D, Cin, Cout, dil in [(32, 64, 64, 1), (16, 128, 128, 1), (16, 256, 256, 2), (16, 512, 512, 4)]:
        x = torch.randn((N, D, D, D, Cin), device=dev).to(torch.bfloat16)
        dy = torch.randn((N, D, D, D, Cout), device=dev).to(torch.bfloat16)
        for _ in range(2):
            K.conv3d_wgrad(x, dy, 3, 1, dil, dil)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            K.conv3d_wgrad(x, dy, 3, 1, dil, dil)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2 * N * D ** 3 * Cout * Cin * 27
        print(f"  {Cin:4d}->{Cout:4d} @{D}^3 dil{dil}: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TF(alg)")
else:
    for mt in (0, 1, 2, 4):
        print(f"ADNI_WGRAD_MT={mt}", flush=True)
        env = dict(os.environ, ADNI_WGRAD_MT=str(mt))
        subprocess.run([sys.executable, __file__, "child"], env=env, timeout=300)
