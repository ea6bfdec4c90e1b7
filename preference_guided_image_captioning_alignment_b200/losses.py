"""Trainer-facing loss modules: same names, constructor arguments and forward signatures as
pkg/models/model.py:957-1085 of the reference (the classes PreferenceGuidedTrainer instantiates at
pkg/training/trainer.py:204-209), backed by the fused sm_100a kernels.

`PreferenceLoss.forward` accepts either dense (B, T, V) logits — exactly the reference signature, served by
the streaming kernels — or `LazyLogits` handles (hidden states + LM-head weight, produced by
install.LazyLMHead) in which case the LM-head GEMM, log-softmax, gather and masked mean run fused and the
logits never exist.
"""
import logging
from typing import Optional

import torch
import torch.nn as nn

from . import ops


class LazyLogits:
    """What `lm_head(hidden)` WOULD be: logits = hidden @ weight.T, kept symbolic.

    Carries the (B, T, d) hidden states and the (V, d) LM-head weight (tied to wte in GPT-2,
    transformers modeling_gpt2.py:646,706).  `.materialize()` is the escape hatch for code that really needs
    the tensor (it is what the reference computes)."""

    def __init__(self, hidden: torch.Tensor, weight: torch.Tensor):
        self.hidden = hidden
        self.weight = weight
        self._cache = {}

    @property
    def shape(self):
        return torch.Size((*self.hidden.shape[:-1], self.weight.shape[0]))

    @property
    def device(self):
        return self.hidden.device

    @property
    def dtype(self):
        return self.hidden.dtype

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return self.hidden.dim()

    def materialize(self) -> torch.Tensor:
        return torch.nn.functional.linear(self.hidden, self.weight)

    def seq_logprobs(self, labels, mask, length_normalize):
        return ops.lmhead_seq_logprob(self.hidden, self.weight, labels, mask, length_normalize)[0]


class ContrastiveLoss(nn.Module):
    """Drop-in for pkg/models/model.py:957-1000: symmetric cross-entropy over I·Tᵀ/τ on inputs that are used
    AS GIVEN (the model normalises them, model.py:826-829); temperature is not clamped."""

    def __init__(self, temperature: float = 0.07) -> None:
        super().__init__()
        self.temperature = temperature
        self.logger = logging.getLogger(__name__)

    def forward(self, image_embeddings: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
        loss, _, _ = ops.ntxent_auto(image_embeddings, text_embeddings, 1.0 / float(self.temperature), True)
        return loss.to(image_embeddings.dtype) if image_embeddings.dtype == torch.float64 else loss


class PreferenceLoss(nn.Module):
    """Drop-in for pkg/models/model.py:1003-1085: length-normalised sequence log-probs of the preferred and the
    rejected caption, loss = -logsigmoid(beta * (lp_w - lp_l)).mean(); no reference model."""

    def __init__(self, beta: float = 0.1) -> None:
        super().__init__()
        self.beta = beta
        self.logger = logging.getLogger(__name__)

    def forward(self, preferred_logits, rejected_logits, preferred_labels: torch.Tensor,
                rejected_labels: torch.Tensor, preferred_mask: torch.Tensor,
                rejected_mask: torch.Tensor) -> torch.Tensor:
        lw = self._compute_log_probs(preferred_logits, preferred_labels, preferred_mask)
        ll = self._compute_log_probs(rejected_logits, rejected_labels, rejected_mask)
        loss, _, _ = ops.dpo_loss(lw, ll, None, None, float(self.beta), 0.0, lw.numel())
        return loss

    def _compute_log_probs(self, logits, labels: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        """(B, T, V) logits or LazyLogits, (B, T) labels, (B, T) mask -> (B,) mean log-prob per scored token
        (pkg/models/model.py:1052-1085; NaN for a sequence whose shifted mask is all zero, like the reference)."""
        if isinstance(logits, LazyLogits):
            return logits.seq_logprobs(labels, mask, True)
        return ops.logits_seq_logprob(logits, labels, mask, True)[0]
