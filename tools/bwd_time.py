"""Time the dual backward kernel on the three shapes that matter: cfg2 DPO (4096 x 50257 x 1024, row term) and the NT-Xent
backward (both terms) at 4096 x 32768 x 512 (one rank of cfg3) and 16384^2 x 512.  Minimum over round-robin repeats."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import functional as F

dev = "cuda"
torch.manual_seed(0)
x = (torch.randn(4096, 1024, device=dev) * 0.5).to(torch.bfloat16)
y = (torch.randn(50257, 1024, device=dev) * 0.02).to(torch.bfloat16)
lse, _ = F.gemm_lse(x, y, 1.0)
row = (lse, torch.randn(4096, device=dev), torch.randint(0, 50257, (4096,), device=dev, dtype=torch.int32))
cases = {"cfg2 dual 4096x50257x1024": lambda: F.softmax_grad_gemm_dual(x, y, 1.0, row=row, out_x_dtype=torch.bfloat16)}
one = torch.ones(1, device=dev)
for ra, rb in ((4096, 32768), (16384, 16384)):
    a = torch.nn.functional.normalize(torch.randn(ra, 512, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(rb, 512, device=dev), dim=-1).bfloat16()
    lr, dg, lc = F.ntxent_fwd(a, b, 2.0, 0, bounded=True)
    cases[f"ntxent bwd {ra}x{rb}x512"] = (lambda a=a, b=b, lr=lr, lc=lc: F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 0.5 / rb))
    cases[f"ntxent bwd {ra}x{rb}x512 one exponential"] = (
        lambda a=a, b=b, lr=lr, lc=lc: F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 0.5 / rb, bounded=True))
best = {}
for rnd in range(3):
    for name, fn in cases.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best[name] = min(best.get(name, 1e9), e0.elapsed_time(e1) / 8)
for name, ms in best.items():
    print(f"{name}: {ms * 1e3:.1f} us", flush=True)
