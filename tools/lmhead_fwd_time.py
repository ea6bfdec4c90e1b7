"""Time the LM-head forward of cfg2 (gemm_lse: row LSE + target gather, 4096 x 50257 x 1024)."""
import sys, torch
sys.path.insert(0, ".")
from preference_guided_image_captioning_alignment_b200 import functional as F

# the LM-head forward of cfg2 (row LSE + target gather), for regressions of the shared kernel
x = (torch.randn(4096, 1024, device="cuda") * 0.5).to(torch.bfloat16)
y = (torch.randn(50257, 1024, device="cuda") * 0.02).to(torch.bfloat16)
lab = torch.randint(0, 50257, (4096,), device="cuda", dtype=torch.int32)
for _ in range(3):
    F.gemm_lse(x, y, 1.0, lab, 0)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F.gemm_lse(x, y, 1.0, lab, 0)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print("cfg2 forward 4096x50257x1024", round(best * 1e3, 1), "us", flush=True)
