"""Time the NT-Xent forward, two-pass against one-pass (unit-norm rows), on the sizes of BASELINE cfg1/cfg3 and B=16384."""
import sys, torch
sys.path.insert(0, ".")
from preference_guided_image_captioning_alignment_b200 import functional as F

for ra, rb in ((4096, 4096), (4096, 32768), (16384, 16384)):
    a = torch.nn.functional.normalize(torch.randn(ra, 512, device="cuda"), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(rb, 512, device="cuda"), dim=-1).bfloat16()
    for bounded in (False, True):
        for _ in range(3):
            F.ntxent_fwd(a, b, 2.0, 0, bounded=bounded)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            F.ntxent_fwd(a, b, 2.0, 0, bounded=bounded)
        e1.record()
        torch.cuda.synchronize()
        print(ra, rb, "one-pass" if bounded else "two-pass", round(e0.elapsed_time(e1) / 10 * 1e3, 1), "us", flush=True)

