"""Run the dominant conv shapes of the ResNet-18 128^3 step once each (for ncu captures)."""
import sys

import torch

sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SHAPES = [  # D, Cin, Cout, k, stride, pad, dil
    (32, 64, 64, 3, 1, 1, 1),      # layer1
    (16, 128, 128, 3, 1, 1, 1),    # layer2
    (16, 256, 256, 3, 1, 2, 2),    # layer3
    (16, 512, 512, 3, 1, 4, 4),    # layer4
]


def run_once():
    for D, Cin, Cout, k, s, p, d in SHAPES:
        x = torch.randn((N, D, D, D, Cin), device=dev).to(torch.bfloat16)
        w = torch.randn((Cout, Cin, k, k, k), device=dev) * 0.05
        oti, ito = K.weights_to_kernel_layout(w)
        y, st = K.conv3d_fprop(x, oti, None, k, s, p, d, stats=True)
        dy = torch.randn_like(y)
        dx = K.conv3d_dgrad(dy, ito, tuple(x.shape), k, s, p, d)
        dw, _ = K.conv3d_wgrad(x, dy, k, s, p, d)
    xs = torch.randn((N, 128, 128, 128, 1), device=dev).to(torch.bfloat16)
    ws = torch.randn((64, 1, 7, 7, 7), device=dev) * 0.05
    x8 = K.stem_expand(xs)
    ys, st = K.stem_fprop(x8, tuple(xs.shape), ws)
    g = K.stem_wgrad(x8, torch.randn_like(ys), tuple(xs.shape))
    torch.cuda.synchronize()


run_once()
run_once()
print("ok")
