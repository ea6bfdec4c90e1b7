"""Swap the fused heads into the reference package without editing it.

    import preference_guided_image_captioning_alignment_b200 as pgica
    pgica.install()                      # before the trainer is constructed
    # ... scripts/train.py / PreferenceGuidedTrainer run unchanged ...

What is rebound:
  * `ContrastiveLoss`, `PreferenceLoss` in pkg.models.model, pkg.models and pkg.training.trainer — the names
    PreferenceGuidedTrainer.__init__ resolves at pkg/training/trainer.py:204-209.
  * `ContrastiveLoss`, `DPOPreferenceLoss`, `TemperatureScaledSimilarity`, `compute_sequence_logprobs` in
    pkg.models.components.
  * with fuse_lm_head=True, every CaptionDecoder built afterwards gets its GPT-2 `lm_head` wrapped so that the
    training forward (pkg/models/model.py:604-610) returns a LazyLogits handle instead of the (B, T, V) tensor, and
    the HF causal-LM loss that forward computes from `labels` (modeling_gpt2.py:709-716) is taken from the same
    fused kernel.  Generation (model.py:621-678) keeps dense logits.
`python -m preference_guided_image_captioning_alignment_b200.install script.py [args...]` runs a script
(e.g. the reference's scripts/train.py) with the swap applied.
"""
import importlib
import runpy
import sys
import threading

import torch
import torch.nn as nn

from . import components, losses, ops

REFERENCE_PACKAGE = "preference_guided_image_captioning_alignment"
_originals = {}
_lazy = threading.local()


class LazyLMHead(nn.Module):
    """Wraps the original bias-free nn.Linear (whose weight stays the Parameter tied to wte)."""

    def __init__(self, linear: nn.Linear):
        super().__init__()
        if getattr(linear, "bias", None) is not None:
            raise ValueError("fused LM head expects a bias-free lm_head (GPT-2)")
        self.linear = linear

    @property
    def weight(self):
        return self.linear.weight

    def forward(self, hidden):
        if getattr(_lazy, "on", False) and hidden.is_cuda and hidden.dim() == 3 and hidden.shape[1] > 1:
            return losses.LazyLogits(hidden, self.linear.weight)
        return self.linear(hidden)


def lazy_causal_lm_loss(logits, labels, vocab_size=None, num_items_in_batch=None, ignore_index=-100, **kwargs):
    """Stand-in for transformers' ForCausalLMLoss (loss_utils.py:45-67): mean over shifted positions whose label
    is not ignore_index of -log p; LazyLogits go through the fused kernel, tensors through the stock path."""
    if not isinstance(logits, losses.LazyLogits):
        from transformers.loss.loss_utils import ForCausalLMLoss
        return ForCausalLMLoss(logits, labels, vocab_size, num_items_in_batch=num_items_in_batch,
                               ignore_index=ignore_index, **kwargs)
    valid = (labels != ignore_index)
    safe = torch.where(valid, labels, torch.zeros_like(labels))
    seq = ops.lmhead_seq_logprob(logits.hidden, logits.weight, safe, valid, False)[0]
    denom = valid[:, 1:].sum() if num_items_in_batch is None else num_items_in_batch
    return -seq.sum() / denom


def fuse_decoder(decoder):
    """Wrap decoder.lm_model.lm_head / loss_function in place (idempotent).  `decoder` is a CaptionDecoder."""
    lm = decoder.lm_model
    head = lm.get_output_embeddings() if hasattr(lm, "get_output_embeddings") else lm.lm_head
    if not isinstance(head, LazyLMHead):
        lm.lm_head = LazyLMHead(head)
    lm.loss_function = lazy_causal_lm_loss
    return decoder


def _wrap_caption_decoder(cls):
    if getattr(cls, "_pgica_wrapped", False):
        return
    orig_init, orig_forward = cls.__init__, cls.forward

    def __init__(self, *a, **k):
        orig_init(self, *a, **k)
        fuse_decoder(self)

    def forward(self, vision_features, input_ids=None, *a, **k):
        prev = getattr(_lazy, "on", False)
        _lazy.on = input_ids is not None
        try:
            return orig_forward(self, vision_features, input_ids, *a, **k)
        finally:
            _lazy.on = prev

    cls.__init__, cls.forward = __init__, forward
    cls._pgica_wrapped = True
    _originals[(cls, "__init__")] = orig_init
    _originals[(cls, "forward")] = orig_forward


def _rebind(module, name, value):
    if hasattr(module, name):
        _originals.setdefault((module, name), getattr(module, name))
        setattr(module, name, value)


def install(package: str = REFERENCE_PACKAGE, fuse_lm_head: bool = True):
    """Rebind the reference's loss names to the fused modules.  Returns the list of (module, name) rebound."""
    done = []
    model_mod = importlib.import_module(package + ".models.model")
    targets = [model_mod]
    for sub in (".models", ".training.trainer", ".training", ""):
        try:
            targets.append(importlib.import_module(package + sub))
        except Exception:  # trainer needs accelerate/mlflow; rebinding what is importable is enough
            pass
    for mod in targets:
        for name, value in (("ContrastiveLoss", losses.ContrastiveLoss), ("PreferenceLoss", losses.PreferenceLoss)):
            if hasattr(mod, name):
                _rebind(mod, name, value)
                done.append((mod.__name__, name))
    try:
        comp_mod = importlib.import_module(package + ".models.components")
        for name in ("ContrastiveLoss", "DPOPreferenceLoss", "TemperatureScaledSimilarity", "compute_sequence_logprobs"):
            _rebind(comp_mod, name, getattr(components, name))
            done.append((comp_mod.__name__, name))
    except Exception:
        pass
    if fuse_lm_head and hasattr(model_mod, "CaptionDecoder"):
        _wrap_caption_decoder(model_mod.CaptionDecoder)
        done.append((model_mod.__name__, "CaptionDecoder.lm_head"))
    return done


def uninstall():
    for (owner, name), value in list(_originals.items()):
        setattr(owner, name, value)
        if name == "forward":
            owner._pgica_wrapped = False
    _originals.clear()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m preference_guided_image_captioning_alignment_b200.install script.py [args]")
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
