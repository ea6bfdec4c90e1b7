"""NT-Xent backward: dual kernel (one launch) vs one launch per product, on the shapes of cfg1 / cfg3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import _lib
from preference_guided_image_captioning_alignment_b200 import functional as F

dev = "cuda"
one = torch.ones((), device=dev)
for ra, rb in ((64, 64), (4096, 4096), (4096, 32768), (16384, 16384)):
    a = torch.nn.functional.normalize(torch.randn(ra, 512, device=dev), dim=-1).to(torch.bfloat16)
    b = torch.nn.functional.normalize(torch.randn(rb, 512, device=dev), dim=-1).to(torch.bfloat16)
    lr, dg, lc = F.ntxent_fwd(a, b, 2.0)
    res = {}
    for mode in ("1", "0"):
        _lib.set_option("sgg_fused", int(mode))
        for _ in range(3):
            out = F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 1.0 / (2 * ra))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = F.ntxent_bwd(a, b, 2.0, 0, lr, lc, one, 1.0 / (2 * ra))
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 10, out)
    da1, db1 = res["1"][1]
    da0, db0 = res["0"][1]
    rel = lambda x, y: ((x.double() - y.double()).norm() / y.double().norm()).item()
    print(f"ntxent bwd {ra}x{rb}x512: dual {res['1'][0] * 1e3:.1f} us, split {res['0'][0] * 1e3:.1f} us; "
          f"da rel {rel(da1, da0):.1e} db rel {rel(db1, db0):.1e}", flush=True)
