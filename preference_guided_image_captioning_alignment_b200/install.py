"""Swap the fused heads into the reference package without editing it.

    import preference_guided_image_captioning_alignment_b200 as pgica
    pgica.install()                      # before the trainer is constructed
    # ... scripts/train.py / PreferenceGuidedTrainer run unchanged ...

What is rebound:
  * `ContrastiveLoss`, `PreferenceLoss` in pkg.models.model, pkg.models and pkg.training.trainer — the names
    PreferenceGuidedTrainer.__init__ resolves at pkg/training/trainer.py:204-209.
  * `ContrastiveLoss`, `DPOPreferenceLoss`, `TemperatureScaledSimilarity`, `compute_sequence_logprobs` in
    pkg.models.components.
  * `PreferenceGuidedCaptioningModel.compute_similarity` (pkg/models/model.py:925-954) -> scoring.model_compute_similarity.
  * with fuse_lm_head=True, every CaptionDecoder built afterwards (and every one passed to `fuse_decoder`) gets the
    `forward` of its GPT-2 `lm_head` Linear rebound on the instance, so that the training forward
    (pkg/models/model.py:604-610) returns a LazyLogits handle instead of the (B, T, V) tensor; the HF causal-LM loss
    that forward computes from `labels` (modeling_gpt2.py:709-716) becomes a deferred value taken from the same fused
    kernel if anybody reads it.  The module tree and the state_dict keys do not change (checkpoints stay
    interchangeable, pkg/training/trainer.py:795,845); under LoRA the GPT2LMHeadModel inside the PeftModel is the one
    patched.  Generation (model.py:621-678) keeps dense logits.
`python -m preference_guided_image_captioning_alignment_b200.install script.py [args...]` runs a script
(e.g. the reference's scripts/train.py) with the swap applied.
"""
import importlib
import runpy
import sys
import threading
import types
import weakref

import torch
import torch.nn as nn

from . import components, losses, ops, prologue, scoring

REFERENCE_PACKAGE = "preference_guided_image_captioning_alignment"
_originals = {}
_lazy = threading.local()
_patched = []  # instance-level patches (weak references), undone by uninstall()
_MISSING = object()


def _lazy_linear_forward(self, hidden):
    """Replacement `forward` of the GPT-2 `lm_head` Linear (bound on the INSTANCE, so the module tree, the
    state_dict keys and HF's weight tying are untouched): inside a training / validation forward of CaptionDecoder the
    (B, T, d) hidden states are handed on as a LazyLogits handle; everywhere else (generation, CPU, 2-D inputs) the
    Linear runs as usual."""
    if getattr(_lazy, "on", False) and hidden.dim() == 3 and hidden.shape[1] > 1 and self.bias is None \
            and (hidden.is_cuda or getattr(_lazy, "any_device", False)):  # any_device: host-logic tests only
        return losses.LazyLogits(hidden, self.weight)
    return nn.functional.linear(hidden, self.weight, self.bias)


def lazy_causal_lm_loss(logits, labels, vocab_size=None, num_items_in_batch=None, ignore_index=-100, **kwargs):
    """Stand-in for transformers' ForCausalLMLoss (loss_utils.py:45-67): mean over shifted positions whose label
    is not ignore_index of -log p.  Tensors go through the stock path; for LazyLogits the value is DEFERRED
    (losses.DeferredLoss): the Stage-2 trainer never reads `generation_loss`, and computing it means an LM-head
    pass over every position, pads included.  When somebody does read it, it comes from the fused kernel."""
    if not isinstance(logits, losses.LazyLogits):
        from transformers.loss.loss_utils import ForCausalLMLoss
        return ForCausalLMLoss(logits, labels, vocab_size, num_items_in_batch=num_items_in_batch,
                               ignore_index=ignore_index, **kwargs)
    return losses.DeferredLoss(lambda: causal_lm_loss_now(logits, labels, num_items_in_batch, ignore_index),
                               logits.device)


def causal_lm_loss_now(logits, labels, num_items_in_batch=None, ignore_index=-100):
    """The HF causal-LM loss of a LazyLogits handle, evaluated immediately (fused kernel, no logits)."""
    valid = (labels != ignore_index)
    safe = torch.where(valid, labels, torch.zeros_like(labels))
    seq = ops.lmhead_seq_logprob(logits.hidden, logits.weight, safe, valid, False)[0]
    denom = valid[:, 1:].sum() if num_items_in_batch is None else num_items_in_batch
    return -seq.sum() / denom


def resolve_causal_lm(lm):
    """The module that owns `lm_head` and calls `self.loss_function`: `lm` itself, or — when LoRA is configured and
    CaptionDecoder.lm_model is a PeftModel (pkg/models/model.py:538-558) — the GPT2LMHeadModel inside the wrappers."""
    seen = set()
    while id(lm) not in seen:
        seen.add(id(lm))
        if hasattr(lm, "get_base_model") and callable(lm.get_base_model):
            inner = lm.get_base_model()
        elif hasattr(lm, "base_model") and hasattr(lm.base_model, "model") and lm.base_model is not lm:
            inner = lm.base_model.model
        else:
            break
        if inner is None or inner is lm:
            break
        lm = inner
    return lm


def fuse_model(model, cross_attention: bool = True, projection_tails: bool = True):
    """Everything install() can fuse on an already-built PreferenceGuidedCaptioningModel: the lazy LM head, the one-key
    cross-attention and the LayerNorm + L2-normalise tail of both projection heads (row 3).  All as
    instance-level forward patches: the module tree and every state_dict key stay as they are."""
    fuse_decoder(model.caption_decoder, cross_attention=cross_attention)
    if projection_tails:
        for enc in (getattr(model, "vision_encoder", None), getattr(model, "text_encoder", None)):
            if enc is not None and hasattr(enc, "projection"):
                prologue.fuse_projection_tail(enc)
    return model


def unfuse_model(model):
    unfuse_decoder(model.caption_decoder)
    for enc in (getattr(model, "vision_encoder", None), getattr(model, "text_encoder", None)):
        if enc is not None and hasattr(enc, "projection"):
            prologue.unfuse_projection_tail(enc)
    return model


def fuse_decoder(decoder, cross_attention: bool = False):
    """Make `decoder.lm_model`'s LM head lazy in place (idempotent).  `decoder` is a CaptionDecoder.  Nothing is added
    to or renamed in the module tree: checkpoints written before and after install() are interchangeable."""
    if cross_attention and hasattr(decoder, "cross_attention") and hasattr(decoder, "attention_norm"):
        prologue.fuse_cross_attention(decoder)
    lm = resolve_causal_lm(decoder.lm_model)
    head = lm.get_output_embeddings() if hasattr(lm, "get_output_embeddings") else lm.lm_head
    if head is None or not isinstance(head, nn.Linear):
        raise TypeError(f"fused LM head expects an nn.Linear output embedding, found {type(head).__name__}")
    if getattr(head, "bias", None) is not None:
        raise ValueError("fused LM head expects a bias-free lm_head (GPT-2)")
    if "forward" not in head.__dict__:
        head.forward = types.MethodType(_lazy_linear_forward, head)
        _patched.append((weakref.ref(head), "forward"))
    if getattr(lm, "_loss_function", None) is not lazy_causal_lm_loss:
        _patched.append((weakref.ref(lm), "_loss_function", getattr(lm, "__dict__", {}).get("_loss_function", _MISSING)))
        lm.loss_function = lazy_causal_lm_loss
    return decoder


def unfuse_decoder(decoder):
    if hasattr(decoder, "cross_attention") and hasattr(decoder, "attention_norm"):
        prologue.unfuse_cross_attention(decoder)
    lm = resolve_causal_lm(decoder.lm_model)
    head = lm.get_output_embeddings() if hasattr(lm, "get_output_embeddings") else lm.lm_head
    head.__dict__.pop("forward", None)
    if getattr(lm, "_loss_function", None) is lazy_causal_lm_loss:
        del lm._loss_function
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
    if hasattr(model_mod, "PreferenceGuidedCaptioningModel"):
        cls = model_mod.PreferenceGuidedCaptioningModel
        if hasattr(cls, "compute_similarity"):
            _originals.setdefault((cls, "compute_similarity"), cls.compute_similarity)
            cls.compute_similarity = scoring.model_compute_similarity
            done.append((model_mod.__name__, "PreferenceGuidedCaptioningModel.compute_similarity"))
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
    for rec in _patched:
        obj = rec[0]()
        if obj is None:
            continue
        if rec[1] == "forward":
            obj.__dict__.pop("forward", None)
        elif rec[1] == "_loss_function" and obj.__dict__.get("_loss_function") is lazy_causal_lm_loss:
            if rec[2] is _MISSING:
                del obj.__dict__["_loss_function"]
            else:
                obj.__dict__["_loss_function"] = rec[2]
    _patched.clear()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m preference_guided_image_captioning_alignment_b200.install script.py [args]")
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
