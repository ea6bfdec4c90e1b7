"""Mirror of the reference's stand-alone loss components (pkg/models/components.py): same class / function
names, constructor arguments, forward signatures and return types, computed by the fused sm_100a kernels.

    TemperatureScaledSimilarity   components.py:24-83
    ContrastiveLoss               components.py:86-145   (normalises inside, clamps tau to [0.1, 2.0])
    DPOPreferenceLoss             components.py:148-249  -> (loss, metrics dict)
    compute_sequence_logprobs     components.py:321-362  (masked SUM of target log-probs)

plus the hidden-state-level entry points that avoid materialising logits altogether
(`lmhead_sequence_logprobs`, `FusedDPOHead`).
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .losses import LazyLogits

_METRIC_KEYS = ("dpo_loss", "reward_margin", "reward_accuracy", "policy_chosen_logprob", "policy_rejected_logprob")


class TemperatureScaledSimilarity(nn.Module):
    """normalize(v) · normalize(t)ᵀ / clamp(τ, min_temp, max_temp) as a dense (B, B) matrix.

    The fused ContrastiveLoss below never forms this matrix; this module exists for callers that want it
    (retrieval-style scoring).  τ is a buffer, or a Parameter when `learnable` — like the reference, so it
    shows up in state_dict() under the same key."""

    def __init__(self, temperature: float = 0.5, learnable: bool = False, min_temp: float = 0.1,
                 max_temp: float = 2.0):
        super().__init__()
        self.min_temp = min_temp
        self.max_temp = max_temp
        if learnable:
            self.temperature = nn.Parameter(torch.tensor(temperature))
        else:
            self.register_buffer("temperature", torch.tensor(temperature))

    def _tau_host(self) -> float:
        """τ as a Python float without a device->host read per call: the value is re-read only when the tensor was
        written to (its `_version` moves on load_state_dict / an optimiser step of a learnable τ) or replaced."""
        t = self.temperature
        key = (id(t), t._version, t.device)
        cached = self.__dict__.get("_tau_cache")
        if cached is None or cached[0] != key:
            cached = (key, float(t.detach()))
            self.__dict__["_tau_cache"] = cached
        return cached[1]

    def effective_temperature(self) -> float:
        return float(min(max(self._tau_host(), self.min_temp), self.max_temp))

    def forward(self, vision_embeds: torch.Tensor, text_embeds: torch.Tensor) -> torch.Tensor:
        return _DenseSimilarity.apply(vision_embeds, text_embeds, self.temperature, self._tau_host(), self.min_temp,
                                      self.max_temp)


def _gemm_nn(a: torch.Tensor, b: torch.Tensor, scale: float) -> torch.Tensor:
    """scale * a @ b on the library's tcgen05 GEMM (`pgica_similarity` = the gemm_lse main loop with a plain store
    epilogue): a (m, k) and b (k, n) float; operands go in as bf16, k is zero-padded to a multiple of 8, fp32 out."""
    from . import functional as F
    k = a.shape[1]
    bt = b.t()
    if k % 8:
        pad = 8 - k % 8
        a = torch.nn.functional.pad(a, (0, pad))
        bt = torch.nn.functional.pad(bt, (0, pad))
    return F.similarity(F.as_bf16(a.contiguous()), F.as_bf16(bt.contiguous()), scale)


class _DenseSimilarity(torch.autograd.Function):
    """S = normalize(v) normalize(t)^T / clamp(tau) as a dense matrix, differentiable like the reference's module
    (components.py:61-83).  Forward: `pgica_rownorm_fwd` + `pgica_similarity`.  Backward of a DENSE upstream gradient:
    dS t_hat and dS^T v_hat on the same tcgen05 GEMM (`_gemm_nn`; there is no softmax to fuse and nothing to
    recompute), then `pgica_rownorm_bwd`; d tau = -sum(dS * S) / tau inside the clamp range.  Training goes through
    ContrastiveLoss, which never forms S; this path exists for callers that score with the matrix."""

    @staticmethod
    def forward(ctx, v, t, temperature, tau, min_temp, max_temp):
        from . import functional as F
        tau_eff = float(min(max(tau, min_temp), max_temp))
        vn, vinv = ops.l2_normalize(v.detach(), 1e-12)
        tn, tinv = ops.l2_normalize(t.detach(), 1e-12)
        S = F.similarity(vn, tn, 1.0 / tau_eff)
        ctx.save_for_backward(v, t, vn, tn, vinv, tinv, S)
        ctx.tau_eff = tau_eff
        ctx.tau_inside = min_temp < tau < max_temp  # clamp passes the gradient only strictly inside its range
        ctx.temp_dtype = temperature.dtype
        return S

    @staticmethod
    def backward(ctx, dS):
        v, t, vn, tn, vinv, tinv, S = ctx.saved_tensors
        dS = dS.float().contiguous()
        dv = dt = dtau = None
        if ctx.needs_input_grad[0]:
            dvn = _gemm_nn(dS, tn, 1.0 / ctx.tau_eff)
            dv = ops.l2_normalize_bwd(v.detach(), vinv, dvn.contiguous()).to(v.dtype)
        if ctx.needs_input_grad[1]:
            dtn = _gemm_nn(dS.t(), vn, 1.0 / ctx.tau_eff)
            dt = ops.l2_normalize_bwd(t.detach(), tinv, dtn.contiguous()).to(t.dtype)
        if ctx.needs_input_grad[2]:
            dtau = (-(dS * S).sum() / ctx.tau_eff if ctx.tau_inside else torch.zeros((), device=dS.device))
            dtau = dtau.to(ctx.temp_dtype)
        return dv, dt, dtau, None, None, None


class ContrastiveLoss(nn.Module):
    """NT-Xent / InfoNCE, components flavour (components.py:86-145): L2-normalise both inputs (eps 1e-12),
    clamp τ to [0.1, 2.0], symmetric cross-entropy with reduction 'mean' or 'sum', halved."""

    def __init__(self, temperature: float = 0.5, reduction: str = "mean"):
        super().__init__()
        self.similarity = TemperatureScaledSimilarity(temperature=temperature)
        self.reduction = reduction

    def forward(self, vision_embeds: torch.Tensor, text_embeds: torch.Tensor) -> torch.Tensor:
        if self.reduction not in ("mean", "sum"):
            # F.cross_entropy would raise on an unknown reduction string as well
            raise ValueError(f"{self.reduction} is not a valid value for reduction")
        out = ops.ntxent_cosine(vision_embeds, text_embeds, 1.0 / self.similarity.effective_temperature(),
                                self.reduction == "mean", 1e-12)
        return out[0]


class DPOPreferenceLoss(nn.Module):
    """Direct Preference Optimisation loss on per-sequence log-probs (components.py:148-249)."""

    def __init__(self, beta: float = 0.1, reference_free: bool = False, label_smoothing: float = 0.0):
        super().__init__()
        self.beta = beta
        self.reference_free = reference_free
        self.label_smoothing = label_smoothing

    def forward(self, policy_chosen_logprobs: torch.Tensor, policy_rejected_logprobs: torch.Tensor,
                reference_chosen_logprobs: Optional[torch.Tensor] = None,
                reference_rejected_logprobs: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, dict]:
        loss, metrics = self.forward_tensors(policy_chosen_logprobs, policy_rejected_logprobs,
                                             reference_chosen_logprobs, reference_rejected_logprobs)
        vals = metrics.tolist()  # ONE device->host read (the reference does five .item() calls, :241-247)
        return loss, dict(zip(_METRIC_KEYS, vals))

    def forward_tensors(self, pc, pr, rc=None, rr=None, n_global: Optional[int] = None):
        """Same computation, metrics left on the device as a (5,) tensor (no host sync)."""
        if self.reference_free or rc is None:
            rc = rr = None
        loss, metrics, _ = ops.dpo_loss(pc, pr, rc, rr, float(self.beta), float(self.label_smoothing),
                                        int(n_global or pc.numel()))
        return loss, metrics.detach()  # metrics are computed under no_grad in the reference (components.py:234)


def compute_sequence_logprobs(logits, labels: torch.Tensor,
                              attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, T, V) logits (or LazyLogits), (B, T) labels, optional (B, T) mask -> (B,) masked sum of
    log p(labels[t+1] | position t)  (components.py:321-362)."""
    if isinstance(logits, LazyLogits):
        return logits.seq_logprobs(labels, attention_mask, False)
    return ops.logits_seq_logprob(logits, labels, attention_mask, False)[0]


def lmhead_sequence_logprobs(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                             attention_mask: Optional[torch.Tensor] = None,
                             length_normalize: bool = False) -> torch.Tensor:
    """compute_sequence_logprobs(lm_head(hidden), labels, mask) without the logits: hidden (B, T, d), tied
    LM-head / wte weight (V, d).  Differentiable w.r.t. hidden and weight."""
    return ops.lmhead_seq_logprob(hidden, weight, labels, attention_mask, length_normalize)[0]


class FusedDPOHead(nn.Module):
    """The whole Stage-2 head at the hidden-state level: policy log-probs (with grad) for chosen and rejected,
    frozen-reference log-probs (no grad), DPO loss and metrics.  Equivalent to

        DPOPreferenceLoss(beta, ...)(csl(lm_head(hc)), csl(lm_head(hr)), csl(ref_lm_head(rhc)), csl(ref_lm_head(rhr)))

    with csl = compute_sequence_logprobs, and logits never materialised."""

    def __init__(self, beta: float = 0.1, reference_free: bool = False, label_smoothing: float = 0.0,
                 length_normalize: bool = False):
        super().__init__()
        self.loss = DPOPreferenceLoss(beta, reference_free, label_smoothing)
        self.length_normalize = length_normalize

    def forward_stacked(self, hidden, weight, labels, mask=None, ref_hidden=None, ref_weight=None,
                        n_global: Optional[int] = None):
        """Same head on pre-stacked inputs: rows [0, B) are the chosen captions, rows [B, 2B) the rejected ones
        (the layout a concatenated policy forward produces) — no concatenation copy."""
        ln = self.length_normalize
        B = hidden.shape[0] // 2
        seq = lmhead_sequence_logprobs(hidden, weight, labels, mask, ln)
        rc = rr = None
        if ref_weight is not None and not self.loss.reference_free:
            with torch.no_grad():
                rseq = lmhead_sequence_logprobs(ref_hidden, ref_weight, labels, mask, ln)
            rc, rr = rseq[:B], rseq[B:]
        return self.loss.forward_tensors(seq[:B], seq[B:], rc, rr, n_global)

    def forward(self, hidden_chosen, hidden_rejected, weight, labels_chosen, labels_rejected, mask_chosen=None,
                mask_rejected=None, ref_hidden_chosen=None, ref_hidden_rejected=None, ref_weight=None,
                n_global: Optional[int] = None):
        ln = self.length_normalize
        B = hidden_chosen.shape[0]
        same_shape = hidden_chosen.shape == hidden_rejected.shape and (mask_chosen is None) == (mask_rejected is None)
        if same_shape:
            # one launch over chosen ++ rejected rows: twice the row blocks per wave, one dW accumulation
            hidden = torch.cat([hidden_chosen, hidden_rejected], 0)
            labels = torch.cat([labels_chosen, labels_rejected], 0)
            mask = None if mask_chosen is None else torch.cat([mask_chosen, mask_rejected], 0)
            seq = lmhead_sequence_logprobs(hidden, weight, labels, mask, ln)
            pc, pr = seq[:B], seq[B:]
        else:
            pc = lmhead_sequence_logprobs(hidden_chosen, weight, labels_chosen, mask_chosen, ln)
            pr = lmhead_sequence_logprobs(hidden_rejected, weight, labels_rejected, mask_rejected, ln)
        rc = rr = None
        if ref_weight is not None and not self.loss.reference_free:
            with torch.no_grad():
                if same_shape and ref_hidden_chosen.shape == ref_hidden_rejected.shape:
                    rseq = lmhead_sequence_logprobs(torch.cat([ref_hidden_chosen, ref_hidden_rejected], 0), ref_weight,
                                                    labels, mask, ln)
                    rc, rr = rseq[:B], rseq[B:]
                else:
                    rc = lmhead_sequence_logprobs(ref_hidden_chosen, ref_weight, labels_chosen, mask_chosen, ln)
                    rr = lmhead_sequence_logprobs(ref_hidden_rejected, ref_weight, labels_rejected, mask_rejected, ln)
        return self.loss.forward_tensors(pc, pr, rc, rr, n_global)


class NaNSafeGradientNorm(nn.Module):
    """Gradient clipping with the non-finite check folded in (components.py:252-318), one multi-tensor pass on the GPU:
    `forward(parameters) -> (total_norm, is_finite)`; gradients are clipped in place to `max_norm` when the norm is
    finite, left untouched (and `is_finite` False) otherwise.  Replaces, in one call, the per-parameter
    `torch.isfinite(p.grad).all()` scan and `clip_grad_norm_` of the trainer (trainer.py:494-515, 619-628).
    Only the L2 norm (the reference's default and the trainer's choice) is implemented on the device."""

    def __init__(self, max_norm: float = 1.0, norm_type: float = 2.0, error_if_nonfinite: bool = False):
        super().__init__()
        if float(norm_type) != 2.0:
            raise ValueError("NaNSafeGradientNorm: only norm_type=2.0 is implemented by the fused kernel")
        self.max_norm = max_norm
        self.norm_type = norm_type
        self.error_if_nonfinite = error_if_nonfinite

    def forward(self, parameters) -> Tuple[torch.Tensor, bool]:
        from . import functional as F
        grads = [p.grad for p in parameters if p.grad is not None]
        if len(grads) == 0:
            return torch.tensor(0.0), True
        native = all(g.is_contiguous() and g.dtype in (torch.float32, torch.bfloat16) for g in grads)
        if native:
            stats = F.grad_norm_clip(grads, self.max_norm, clip=True)
        else:
            # gradient layouts the multi-tensor kernel does not take in place (non-contiguous views, fp16 / fp64): the
            # norm is taken over contiguous fp32 copies of those, and the clip coefficient — still on the device, no
            # host round trip — is applied to every gradient in place
            views = [g if (g.is_contiguous() and g.dtype in (torch.float32, torch.bfloat16)) else
                     g.detach().float().contiguous() for g in grads]
            stats = F.grad_norm_clip(views, self.max_norm, clip=False)
            scale = torch.where((stats[2] != 0) & (stats[1] < 1), stats[1], torch.ones_like(stats[1]))
            for g in grads:
                g.mul_(scale.to(g.dtype))
        total_norm = stats[0]
        is_finite = bool(stats[2].item() != 0.0)  # the one host read the reference makes too (components.py:307)
        if not is_finite and self.error_if_nonfinite:
            raise RuntimeError("Non-finite gradient norm detected")
        return total_norm, is_finite
