"""Small-batch NT-Xent: one-launch kernel vs the general path (fwd + bwd), timings for B = 8 / 64 / 128."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import preference_guided_image_captioning_alignment_b200 as pg
from preference_guided_image_captioning_alignment_b200 import functional as F
from preference_guided_image_captioning_alignment_b200 import ops

dev = "cuda"
for B, D in ((8, 512), (64, 512), (128, 512)):
    a = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1).requires_grad_(True)
    b = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1).requires_grad_(True)
    ab, bb = a.detach().to(torch.bfloat16), b.detach().to(torch.bfloat16)

    def timed(fn, iters=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    def small_kernel():
        return F.ntxent_small(ab, bb, 2.0, True)

    def module_small():
        a.grad = b.grad = None
        pg.ContrastiveLoss(temperature=0.5)(a, b).backward()

    def module_general():
        a.grad = b.grad = None
        ops.ntxent(a, b, 2.0, True)[0].backward()

    gstep = pg.GraphedContrastiveStep(pg.ContrastiveLoss(temperature=0.5), a.detach(), b.detach())

    def module_graphed():
        return gstep.replay()

    l1 = small_kernel()
    module_small()
    ga = a.grad.clone()
    module_general()
    rel = ((ga - a.grad).norm() / a.grad.norm()).item()
    print(f"B={B} D={D}: kernel alone {timed(small_kernel):.1f} us; module fwd+bwd small {timed(module_small):.1f} us, "
          f"general {timed(module_general):.1f} us, graphed {timed(module_graphed):.1f} us; grad rel diff small vs general {rel:.1e}; loss {l1[0].item():.5f}", flush=True)
