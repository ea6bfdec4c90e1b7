"""Bring-up check of the dual softmax-gradient GEMM (sgg_f.cu): small parity cases, then cfg2 timing.
python tools/dual_check.py [small|cfg2|all]"""
import os
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
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def ref(x, y, row, col):
    xf, yf = x.float(), y.float()
    z = xf @ yf.t()
    g = torch.zeros_like(z)
    if row is not None:
        lse, coef, tgt = row
        oh = (tgt[:, None].long() == torch.arange(y.shape[0], device=dev)[None, :]).float()
        g += coef[:, None] * (torch.exp(z - lse[:, None]) - oh)
    if col is not None:
        lse, coef, tgt = col
        oh = (tgt[None, :].long() == torch.arange(x.shape[0], device=dev)[:, None]).float()
        g += coef[None, :] * (torch.exp(z - lse[None, :]) - oh)
    return g @ yf, g.t() @ xf


def case(mx, my, k, mode, plan=None):
    if plan:
        _set_plan(plan)
    else:
        _set_plan(None)
    torch.manual_seed(mx + my)
    x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
    y = (torch.randn(my, k, device=dev) * 0.2).to(torch.bfloat16)
    row = col = None
    if mode in ("row", "both"):
        row = (F.gemm_lse(x, y, 1.0)[0], torch.randn(mx, device=dev),
               torch.randint(0, my, (mx,), device=dev, dtype=torch.int32))
    if mode in ("col", "both"):
        col = (F.gemm_lse(y, x, 1.0)[0], torch.randn(my, device=dev),
               torch.randint(0, mx, (my,), device=dev, dtype=torch.int32))
    ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
    torch.cuda.synchronize()
    ex, ey = ref(x, y, row, col)
    print(f"dual {mx}x{my}x{k} {mode} plan={plan}: rel out_x {rel(ox, ex):.2e} out_y {rel(oy, ey):.2e}", flush=True)


if what in ("small", "all"):
    case(128, 128, 512, "row")
    case(128, 256, 512, "row")
    case(256, 128, 512, "row")
    case(300, 1000, 512, "both")
    case(200, 333, 1024, "row")
    case(128 * 5 + 7, 128 * 9 + 1, 1024, "row", "2,4")
    case(128 * 7, 128 * 6, 512, "both", "3,2")
    case(2048, 5003, 1024, "row")

if what in ("cfg2", "all"):
    _set_plan(None)
    B, T, d, V = 16, 128, 1024, 50257
    g = torch.Generator().manual_seed(1234)
    W = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    H = torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev)
    y = torch.randint(0, V, (2 * B, T), generator=g).to(dev)
    m = torch.ones(2 * B, T, dtype=torch.long, device=dev)
    seq, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(H, W, y, m, False)
    gseq = torch.randn(2 * B, device=dev)
    _lib.set_option("sgg_fused", 0)
    dh0, dw0 = F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
    _lib.set_option("sgg_fused", 1)
    plans = [None] + [p for p in os.environ.get("DUAL_PLANS", "32,22;32,20;32,21").split(";") if p]
    for plan in plans:
        if plan:
            _set_plan(plan)
        dh1, dw1 = F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
        e0.record()
        for _ in range(10):
            F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
        e1.record()
        torch.cuda.synchronize()
        print(f"cfg2 dual plan={plan}: {e0.elapsed_time(e1) / 10:.3f} ms/bwd; vs split: dH rel {rel(dh1.float(), dh0.float()):.2e} "
              f"dW rel {rel(dw1, dw0):.2e}", flush=True)
    _lib.set_option("sgg_fused", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
    e1.record()
    torch.cuda.synchronize()
    print(f"cfg2 split: {e0.elapsed_time(e1) / 10:.3f} ms/bwd")

if what in ("bf16dw",):
    _set_plan(None)
    _lib.set_option("sgg_fused", 1)
    B, T, d, V = 16, 128, 1024, 50257
    g = torch.Generator().manual_seed(1234)
    W = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    H = torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev)
    y = torch.randint(0, V, (2 * B, T), generator=g).to(dev)
    m = torch.ones(2 * B, T, dtype=torch.long, device=dev)
    seq, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(H, W, y, m, False)
    gseq = torch.randn(2 * B, device=dev)
    for dt in (torch.float32, torch.bfloat16):
        for _ in range(3):
            F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False, dweight_dtype=dt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False, dweight_dtype=dt)
        e1.record()
        torch.cuda.synchronize()
        print(f"cfg2 dual dW {dt}: {e0.elapsed_time(e1) / 10:.3f} ms/bwd")
