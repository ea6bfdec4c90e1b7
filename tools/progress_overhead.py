"""Does publishing progress cost the dual backward kernel anything?  cfg2 shapes on ONE GPU, no all-reduce kernel beside
it: lmhead_logprob_bwd (plain) against lmhead_logprob_bwd_progress (per-segment release counters at system scope)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import functional as F

dev = "cuda"
torch.manual_seed(0)
B, T, d, V = 32, 128, 1024, 50257
H = torch.randn(B, T, d, device=dev).to(torch.bfloat16)
W = (torch.randn(V, d, device=dev) * 0.02).to(torch.bfloat16)
y = torch.randint(0, V, (B, T), device=dev)
m = torch.ones(B, T, dtype=torch.long, device=dev)
gs = torch.randn(B, device=dev)
_, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(H, W, y, m, False)
dw = torch.empty(V, d, device=dev)
prog = torch.zeros(64, dtype=torch.int32, device=dev)
pairs = (V + 255) // 256
rows_per_seg = ((pairs + 7) // 8) * 256
cases = {"plain": lambda: F.lmhead_logprob_bwd(H, W, rl, rw, lse, gs, False),
         "progress": lambda: F.lmhead_logprob_bwd_progress(H, W, rl, rw, lse, gs, dw, prog, rows_per_seg, False)}
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
print({k: round(v * 1e3, 1) for k, v in best.items()}, "us")
