"""One cfg2-shaped dual backward (sgg_f.cu) for ncu: python tools/prof_dual.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import functional as F

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
torch.manual_seed(0)
mx, my, k = 4096, 50257, 1024
x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
y = (torch.randn(my, k, device=dev) * 0.02).to(torch.bfloat16)
row = (torch.full((mx,), 11.0, device=dev), torch.randn(mx, device=dev),
       torch.randint(0, my, (mx,), device=dev, dtype=torch.int32))
for _ in range(iters):
    ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, out_x_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("ok", ox.float().abs().mean().item(), oy.abs().mean().item())
