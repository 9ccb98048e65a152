import sys
import torch
sys.path.insert(0, ".")
from multimodal_alzheimer_b200 import kernels as K
dev = torch.device("cuda:0")
N, S = 32, 128
x = torch.randn((N, S, S, S, 1), device=dev).to(torch.bfloat16)
w = torch.randn((64, 1, 7, 7, 7), device=dev) * 0.05
x8 = K.stem_expand(x)
for _ in range(3):
    y, st = K.stem_fprop(x8, tuple(x.shape), w)
torch.cuda.synchronize()
print("ok")
