"""Per-GPU slice of BASELINE config 4 (seq 512, 32 pairs per GPU at 8 GPUs): DPO head step timing on one B200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import _lib
from preference_guided_image_captioning_alignment_b200 import functional as F

dev = "cuda"
B, T, d, V = int(os.environ.get("PAIRS", 32)), 512, 1024, 50257
g = torch.Generator().manual_seed(1234)
W = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
Wr = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
H = torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev)
Hr = torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev)
y = torch.randint(0, V, (2 * B, T), generator=g).to(dev)
m = torch.ones(2 * B, T, dtype=torch.long, device=dev)
one = torch.ones((), device=dev)


def step(ev=None):
    def mark(i):
        if ev is not None:
            ev[i].record()
    mark(0)
    seq_p, lse_p, _, rl, rw, _ = F.lmhead_logprob_fwd(H, W, y, m, False)
    mark(1)
    seq_r = F.lmhead_logprob_fwd(Hr, Wr, y, m, False)[0]
    mark(2)
    loss, metrics, dpc = F.dpo_loss_fwd(seq_p[:B], seq_p[B:], seq_r[:B], seq_r[B:], 0.1, 0.0, B)
    gseq = F.dpo_grad_seq(dpc, one)
    dh, dw = F.lmhead_logprob_bwd(H, W, rl, rw, lse_p, gseq, False)
    mark(3)
    return loss


for mode in ("1", "0"):
    _lib.set_option("sgg_fused", int(mode))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(5)]
    for k in range(5):
        step(ev[k])
    torch.cuda.synchronize()
    ph = [sum(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(5)) / 5 for i in range(3)]
    tot = sum(ph)
    rows = 2 * B * T
    flops = 16.0 * B * (T - 1) * d * V
    print(f"cfg4 slice ({B} pairs, T={T}) fused={mode}: step {tot:.2f} ms (fwd {ph[0]:.2f}, ref {ph[1]:.2f}, loss+bwd {ph[2]:.2f}); "
          f"{B * (T - 1) / tot * 1e3:.0f} pair-tokens/s, {flops / tot / 1e9:.0f} algorithmic TFLOP/s", flush=True)
