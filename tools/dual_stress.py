"""Randomised stress of the dual softmax-gradient GEMM: random shapes, terms, role splits and output dtypes, each checked
against the two single-product launches (and repeated for determinism).  python tools/dual_stress.py [cases] [seed]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import _lib
from preference_guided_image_captioning_alignment_b200 import functional as F


def _set_plan(plan):
    """"R2,C2" pins the dual kernel's role split (options sggf_plan_r2 / sggf_plan_c2); None hands it back to the planner."""
    r2, c2 = (int(v) for v in plan.split(",")) if plan else (0, 0)
    _lib.set_option("sggf_plan_r2", r2)
    _lib.set_option("sggf_plan_c2", c2)


dev = "cuda"
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


worst = 0.0
for i in range(cases):
    k = rng.choice([512, 512, 1024, 1024, 1536, 2048])
    S = k // 512
    mx = rng.choice([rng.randint(1, 300), rng.randint(300, 3000), 128 * rng.randint(1, 20), 128 * rng.randint(1, 20) + 1])
    my = rng.choice([rng.randint(1, 300), rng.randint(300, 6000), 128 * rng.randint(1, 40), 256 * rng.randint(1, 20) + 129])
    mode = rng.choice(["row", "row", "col", "both"])
    plan = None
    if rng.random() < 0.6:
        r2 = rng.randint(1, max(1, min(12, (70 // S) // 2)))
        c2 = rng.randint(1, max(1, min(12, (70 // S) // 2)))
        plan = f"{r2},{c2}"
        _set_plan(plan)
    else:
        _set_plan(None)
    torch.manual_seed(i)
    x = (torch.randn(mx, k, device=dev) * 0.3).to(torch.bfloat16)
    y = (torch.randn(my, k, device=dev) * 0.3).to(torch.bfloat16)
    row = col = None
    if mode in ("row", "both"):
        row = (F.gemm_lse(x, y, 1.0)[0], torch.randn(mx, device=dev), torch.randint(-1, my, (mx,), device=dev, dtype=torch.int32))
    if mode in ("col", "both"):
        col = (F.gemm_lse(y, x, 1.0)[0], torch.randn(my, device=dev), torch.randint(-1, mx, (my,), device=dev, dtype=torch.int32))
    ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
    ox2, oy2 = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
    sx = F.softmax_grad_gemm(x, y, 1.0, row=row, col=col)
    sy = F.softmax_grad_gemm(y, x, 1.0, row=col, col=row)
    torch.cuda.synchronize()
    ex, ey = rel(ox, sx), rel(oy, sy)
    det = torch.equal(ox, ox2) and torch.equal(oy, oy2)
    worst = max(worst, ex, ey)
    ok = ex < 2e-3 and ey < 2e-3 and det
    print(f"{i:3d} {mx}x{my}x{k} {mode:4s} plan={plan}: out_x {ex:.1e} out_y {ey:.1e} deterministic={det} {'ok' if ok else 'FAIL'}",
          flush=True)
    if not ok:
        sys.exit(1)
print(f"all {cases} cases ok; worst relative difference to the single-product launches {worst:.1e}")
