"""The reference loss heads restated with the same torch ops the reference calls, so that a box without
/root/reference can still run (and time) "the reference's CPU path".

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function names the reference lines it follows;
"pkg/" = src/preference_guided_image_captioning_alignment/.  tests/test_oracle.py checks every function here
against fixtures produced by the real reference (tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def scaled_similarity(v, t, temperature, normalize, clamp_tau, min_temp=0.1, max_temp=2.0):
    # pkg/models/components.py:73-81 (normalize + clamp) / pkg/models/model.py:988 (neither)
    if normalize:
        v = F.normalize(v, p=2, dim=-1)
        t = F.normalize(t, p=2, dim=-1)
    tau = torch.as_tensor(temperature, dtype=v.dtype, device=v.device)
    if clamp_tau:
        tau = torch.clamp(tau, min_temp, max_temp)
    return torch.matmul(v, t.T) / tau


def ntxent_components(v, t, temperature=0.5, reduction="mean"):
    # pkg/models/components.py:117-145
    sim = scaled_similarity(v, t, temperature, normalize=True, clamp_tau=True)
    target = torch.arange(v.size(0), device=v.device)
    fwd = F.cross_entropy(sim, target, reduction=reduction)
    bwd = F.cross_entropy(sim.T, target, reduction=reduction)
    return (fwd + bwd) / 2.0


def ntxent_trainer(img, txt, temperature=0.07):
    # pkg/models/model.py:970-1000
    sim = torch.matmul(img, txt.t()) / temperature
    target = torch.arange(img.size(0), device=img.device)
    return (F.cross_entropy(sim, target) + F.cross_entropy(sim.t(), target)) / 2


def _shifted_token_logprobs(logits, labels):
    # pkg/models/components.py:339-352 == pkg/models/model.py:1069-1079
    lp = F.log_softmax(logits[:, :-1, :].contiguous(), dim=-1)
    tgt = labels[:, 1:].contiguous()
    return torch.gather(lp, dim=-1, index=tgt.unsqueeze(-1)).squeeze(-1)


def sequence_logprobs_sum(logits, labels, attention_mask=None):
    # compute_sequence_logprobs, pkg/models/components.py:321-362
    tok = _shifted_token_logprobs(logits, labels)
    if attention_mask is None:
        m = torch.ones_like(labels[:, 1:])
    else:
        m = attention_mask[:, 1:].contiguous()
    return (tok * m).sum(dim=1)


def sequence_logprobs_mean(logits, labels, mask):
    # PreferenceLoss._compute_log_probs, pkg/models/model.py:1052-1085
    tok = _shifted_token_logprobs(logits, labels)
    m = mask[..., 1:].contiguous()
    return (tok * m).sum(dim=-1) / m.sum(dim=-1)


def preference_loss_trainer(pref_logits, rej_logits, pref_labels, rej_labels, pref_mask, rej_mask, beta=0.1):
    # PreferenceLoss.forward, pkg/models/model.py:1016-1050
    gap = sequence_logprobs_mean(pref_logits, pref_labels, pref_mask) - sequence_logprobs_mean(
        rej_logits, rej_labels, rej_mask)
    return -F.logsigmoid(beta * gap).mean()


def dpo_components(pc, pr, rc=None, rr=None, beta=0.1, reference_free=False, label_smoothing=0.0):
    # DPOPreferenceLoss.forward, pkg/models/components.py:192-249
    pol = pc - pr
    ref = torch.zeros_like(pol) if (reference_free or rc is None) else rc - rr
    x = beta * (pol - ref)
    if label_smoothing > 0:
        loss = F.binary_cross_entropy_with_logits(x, (1.0 - label_smoothing) * torch.ones_like(x), reduction="mean")
    else:
        loss = -F.logsigmoid(x).mean()
    with torch.no_grad():
        margin = (pol - ref).mean()
        acc = (pol > ref).float().mean()
    metrics = {
        "dpo_loss": loss.item(),
        "reward_margin": margin.item(),
        "reward_accuracy": acc.item(),
        "policy_chosen_logprob": pc.mean().item(),
        "policy_rejected_logprob": pr.mean().item(),
    }
    return loss, metrics


def lm_head(hidden, weight):
    # GPT2LMHeadModel.lm_head = nn.Linear(n_embd, vocab, bias=False): transformers modeling_gpt2.py:651,706
    return F.linear(hidden, weight)


def dpo_head_step(hc, hr, W, yc, yr, mc, mr, ref_hc=None, ref_hr=None, ref_W=None, beta=0.1):
    """One Stage-2 head evaluation the way a user of the reference's components would write it: policy
    LM head + sequence log-probs (with grad), frozen-reference LM head + log-probs (no grad), DPO loss,
    backward.  This is the unit bench.py times as the CPU baseline."""
    pc = sequence_logprobs_sum(lm_head(hc, W), yc, mc)
    pr = sequence_logprobs_sum(lm_head(hr, W), yr, mr)
    rc = rr = None
    if ref_W is not None:
        with torch.no_grad():
            rc = sequence_logprobs_sum(lm_head(ref_hc, ref_W), yc, mc)
            rr = sequence_logprobs_sum(lm_head(ref_hr, ref_W), yr, mr)
    loss, metrics = dpo_components(pc, pr, rc, rr, beta=beta)
    loss.backward()
    return loss.detach(), metrics
