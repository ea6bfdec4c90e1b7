"""Role splits of the dual backward kernel on the NT-Xent backward (both softmax terms, k = 512): round-robin minimum.
    python tools/ntxent_plan_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import _lib
from preference_guided_image_captioning_alignment_b200 import functional as F

dev = "cuda"
torch.manual_seed(0)
one = torch.ones(1, device=dev)
shapes = {(4096, 32768): ["", "16,22", "16,20", "16,24", "16,14,2", "16,16,2", "16,18,2", "16,20,2", "16,22,2", "16,12,3"],
          (16384, 16384): ["", "22,22", "16,24", "16,22", "16,20", "16,26", "13,24", "13,26", "11,26", "22,24", "22,20"]}
bounded = True
for (ra, rb), plans in shapes.items():
    a = torch.nn.functional.normalize(torch.randn(ra, 512, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(rb, 512, device=dev), dim=-1).bfloat16()
    lr, dg, lc = F.ntxent_fwd(a, b, 2.0, 0, bounded=True)
    best = {}
    for rnd in range(3):
        for plan in plans:
            v = [int(t) for t in plan.split(",")] if plan else [0, 0]
            _lib.set_option("sggf_plan_r2", v[0])
            _lib.set_option("sggf_plan_c2", v[1])
            _lib.set_option("sggf_col_groups", v[2] if len(v) > 2 else 1)
            for _ in range(2):
                F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 0.5 / rb, bounded=bounded)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(6):
                F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 0.5 / rb, bounded=bounded)
            e1.record()
            torch.cuda.synchronize()
            best[plan] = min(best.get(plan, 1e9), e0.elapsed_time(e1) / 6)
    print(ra, rb, {k or "auto": round(v * 1e3, 1) for k, v in best.items()}, flush=True)
