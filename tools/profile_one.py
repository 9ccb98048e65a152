import sys
import torch
sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import kernels as K
dev = torch.device("cuda:0")
N, D, Cin, Cout, dil = 32, 32, 64, 64, 1
x = torch.randn((N, D, D, D, Cin), device=dev).to(torch.bfloat16)
w = torch.randn((Cout, Cin, 3, 3, 3), device=dev) * 0.05
oti, ito = K.weights_to_kernel_layout(w)
for _ in range(3):
    y, st = K.conv3d_fprop(x, oti, None, 3, 1, dil, dil, stats=True)
torch.cuda.synchronize()
print("ok")
