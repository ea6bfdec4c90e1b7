"""A/B timing of the dual backward kernel on the cfg2 shape: role splits x drain variants.
    python tools/dual_ab.py [mx my k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import _lib
from preference_guided_image_captioning_alignment_b200 import functional as F

mx, my, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 50257, 1024)
dev = "cuda"
torch.manual_seed(0)
x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
y = (torch.randn(my, k, device=dev) * 0.02).to(torch.bfloat16)
lse, _ = F.gemm_lse(x, y, 1.0)
row = (lse, torch.randn(mx, device=dev), torch.randint(0, my, (mx,), device=dev, dtype=torch.int32))


def set_plan(plan):
    r2, c2 = (int(v) for v in plan.split(",")) if plan else (0, 0)
    _lib.set_option("sggf_plan_r2", r2)
    _lib.set_option("sggf_plan_c2", c2)


def timed(iters=10):
    for _ in range(3):
        ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, out_x_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, out_x_dtype=torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, ox, oy


plans = sys.argv[4].split(";") if len(sys.argv) > 4 else ["", "8,11", "16,8", "16,9", "16,10", "16,11"]
ref = None
for plan in plans:
    for drain in (0, 1):
        set_plan(plan or None)
        _lib.set_option("sggf_direct_drain", drain)
        try:
            ms, ox, oy = timed()
        except Exception as e:
            print(f"plan {plan or 'auto':6s} direct_drain {drain}: {e}")
            continue
        if ref is None:
            ref = (ox.float(), oy)
        ex = ((ox.float() - ref[0]).norm() / ref[0].norm()).item()
        ey = ((oy - ref[1]).norm() / ref[1].norm()).item()
        flops = 4.0 * mx * my * k
        print(f"plan {plan or 'auto':6s} direct_drain {drain}: {ms:.4f} ms  {flops / ms / 1e9:7.1f} TF/s algorithmic   "
              f"rel diff vs first: dX {ex:.2e} dY {ey:.2e}", flush=True)
