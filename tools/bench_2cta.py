"""A/B micro-benchmark of the tap-per-box conv kernel (1 CTA per tile) against the experimental CTA-pair engine
(csrc/conv_igemm_2cta.cu): run once plain and once with ADNI_IGEMM_2CTA=1 (the switch is read once per process).
Prints one JSON line per run: ms per launch, algorithmic TFLOP/s and an output checksum per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_alzheimer_b200 import kernels as K  # noqa: E402

SHAPES = {  # N, D, H, W, Cin, Cout, k, stride, pad, dil (ResNet-18 at 128^3, 32 volumes)
    "layer4_512_512_d4": (32, 16, 16, 16, 512, 512, 3, 1, 4, 4),
    "layer3_256_256_d2": (32, 16, 16, 16, 256, 256, 3, 1, 2, 2),
}


def main():
    dev = torch.device("cuda:0")
    out = {"two_cta": os.environ.get("ADNI_IGEMM_2CTA", "0")}
    g = torch.Generator(device=dev).manual_seed(1)
    for name, (N, D, H, W, Cin, Cout, k, s, p, d) in SHAPES.items():
        x = torch.randn((N, D, H, W, Cin), device=dev, generator=g).to(torch.bfloat16)
        w = (torch.randn((Cout, Cin, k, k, k), device=dev, generator=g) / (Cin * k ** 3) ** 0.5)
        oti, ito = K.weights_to_kernel_layout(w)
        flops = 2.0 * N * D * H * W * Cout * Cin * k ** 3
        for mode in ("fprop", "dgrad"):
            fn = (lambda: K.conv3d_fprop(x, oti, None, k, s, p, d, stats=True, engine=1)[0]) if mode == "fprop" else \
                 (lambda: K.conv3d_dgrad(x, ito, (N, D, H, W, Cin), k, s, p, d, engine=1))
            for _ in range(3):
                y = fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                y = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            out[f"{name}_{mode}"] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                                     "checksum": float(y.double().sum()), "abs": float(y.double().abs().sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
