"""One cfg2 step of the DPO head (policy fwd, reference fwd, scalar, dW, dH) for ncu: python tools/prof_step.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from preference_guided_image_captioning_alignment_b200 import functional as F

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
B, T, d, V = 16, 128, 1024, 50257
g = torch.Generator().manual_seed(1234)
W = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
Wr = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
H = torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev)
Hr = torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev)
y = torch.randint(0, V, (2 * B, T), generator=g).to(dev)
m = torch.ones(2 * B, T, dtype=torch.long, device=dev)
one = torch.ones((), device=dev)
for _ in range(iters):
    seq_p, lse_p, _, rl, rw, _ = F.lmhead_logprob_fwd(H, W, y, m, False)
    seq_r = F.lmhead_logprob_fwd(Hr, Wr, y, m, False)[0]
    loss, metrics, dpc = F.dpo_loss_fwd(seq_p[:B], seq_p[B:], seq_r[:B], seq_r[B:], 0.1, 0.0, B)
    gseq = F.dpo_grad_seq(dpc, one)
    dh, dw = F.lmhead_logprob_bwd(H, W, rl, rw, lse_p, gseq, False)
torch.cuda.synchronize()
print("loss", loss.item(), "dh", dh.float().abs().mean().item(), "dw", dw.abs().mean().item())
