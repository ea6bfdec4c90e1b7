"""One NT-Xent backward (both softmax terms, sgg_f.cu) on one rank's slice of cfg3 for ncu: python tools/prof_ntxent_bwd.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import functional as F

dev = "cuda"
torch.manual_seed(0)
ra, rb = 4096, 32768
a = torch.nn.functional.normalize(torch.randn(ra, 512, device=dev), dim=-1).bfloat16()
b = torch.nn.functional.normalize(torch.randn(rb, 512, device=dev), dim=-1).bfloat16()
lr, dg, lc = F.ntxent_fwd(a, b, 2.0, 0, bounded=True)
one = torch.ones(1, device=dev)
for _ in range(2):
    da, db = F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 0.5 / rb)
torch.cuda.synchronize()
print("ok", da.abs().mean().item(), db.abs().mean().item())
