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

    def float(self):
        return self

    def materialize(self) -> torch.Tensor:
        return torch.nn.functional.linear(self.hidden, self.weight)

    def seq_logprobs(self, labels, mask, length_normalize):
        return ops.lmhead_seq_logprob_compact([self.hidden], self.weight, [labels], [mask], length_normalize)[0]


class DeferredLoss(torch.Tensor):
    """A 0-dim loss that is computed the first time anybody looks at it.

    `CaptionDecoder.forward` passes `labels` to GPT-2, so HF computes the causal-LM cross-entropy on every forward
    (modeling_gpt2.py:709-716) and the reference model returns it as `generation_loss` (pkg/models/model.py:849-851) —
    but the Stage-2 trainer never reads it (pkg/training/trainer.py:575-603): on the reference that costs a full
    log-softmax over (B, T, V) per forward; here it would cost an LM-head GEMM over ALL rows, pads included, that the
    compacted preference loss does not need.  The fused LM head therefore hands HF's `loss_function` result back as
    this placeholder; any torch function applied to it (`.item()`, `.backward()`, arithmetic, printing, ...) first
    evaluates the real loss — with the autograd graph it would have had — and continues on that tensor."""

    @staticmethod
    def __new__(cls, thunk, device):
        t = torch.Tensor._make_subclass(cls, torch.empty((), device=device))
        t._thunk = thunk
        t._value = None
        t._grad_mode = torch.is_grad_enabled()
        return t

    def materialize(self) -> torch.Tensor:
        if self._value is None:
            with torch.set_grad_enabled(self._grad_mode):
                self._value = self._thunk()
            self._thunk = None
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        from torch.utils._pytree import tree_map
        real = lambda x: x.materialize() if isinstance(x, DeferredLoss) else x
        with torch._C.DisableTorchFunctionSubclass():
            return func(*tree_map(real, args), **tree_map(real, kwargs or {}))


class ContrastiveLoss(nn.Module):
    """Drop-in for pkg/models/model.py:957-1000: symmetric cross-entropy over I·Tᵀ/τ on inputs that are used
    AS GIVEN (the model normalises them, model.py:826-829); temperature is not clamped.

    The constructor is the reference's.  Setting the attribute `assume_normalized = True` afterwards promises
    unit-norm rows, which lets large batches take the one-pass forward / one-exponential backward; embeddings that
    come out of this package's fused LayerNorm + L2-normalise (prologue.fuse_projection_tail) carry that promise
    themselves."""

    assume_normalized = False

    def __init__(self, temperature: float = 0.07) -> None:
        super().__init__()
        self.temperature = temperature
        self.logger = logging.getLogger(__name__)

    def forward(self, image_embeddings: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
        unit = self.assume_normalized or (getattr(image_embeddings, "_pgica_unit_norm", False)
                                          and getattr(text_embeddings, "_pgica_unit_norm", False))
        loss, _, _ = ops.ntxent_auto(image_embeddings, text_embeddings, 1.0 / float(self.temperature), True, bool(unit))
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
        if (isinstance(preferred_logits, LazyLogits) and isinstance(rejected_logits, LazyLogits)
                and preferred_logits.weight is rejected_logits.weight):
            # both captions in ONE pass over the LM head: scored rows of the two sets gathered into one matrix, one
            # forward GEMM, one dual backward launch (dW accumulated once), logits never formed
            lw, ll = ops.lmhead_seq_logprob_compact([preferred_logits.hidden, rejected_logits.hidden],
                                                    preferred_logits.weight, [preferred_labels, rejected_labels],
                                                    [preferred_mask, rejected_mask], True)
        else:
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
