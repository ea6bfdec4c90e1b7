"""install() against REAL models: a real transformers GPT2LMHeadModel (always available) and, where the reference
package can be imported (`/root/reference/src` in the build container, the git-ignored `baseline/_ref` install on the
GPU box), the reference's own PreferenceGuidedCaptioningModel and trainer module.

CPU tests cover the host logic (module tree / state_dict unchanged, PEFT resolution, LazyLogits + DeferredLoss flow
with an oracle-backed stand-in for the kernels); `-m gpu` tests run the same Stage-2 micro-step patched and unpatched
on the B200 and compare loss, gradients and generation_loss (BASELINE config 5).
"""
import importlib
import math

import pytest
import torch
import torch.nn as nn

import preference_guided_image_captioning_alignment_b200 as pg
from oracle import ref_model
from oracle import torch_port as tp
from preference_guided_image_captioning_alignment_b200 import losses, ops

inst = importlib.import_module("preference_guided_image_captioning_alignment_b200.install")

needs_reference = pytest.mark.skipif(not ref_model.available(), reason="reference package not importable here")


def tiny_lm(n_embd=32, vocab=101, n_layer=1):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(0)
    return GPT2LMHeadModel(GPT2Config(n_embd=n_embd, n_layer=n_layer, n_head=2, vocab_size=vocab, n_positions=32))


class Decoder(nn.Module):
    """Shape of pkg/models/model.py:477-616 as far as install.py cares: `.lm_model` and a forward that calls it."""

    def __init__(self, lm):
        super().__init__()
        self.lm_model = lm

    def forward(self, vision_features, input_ids=None, attention_mask=None, labels=None):
        emb = self.lm_model.transformer.wte(input_ids).float() + vision_features[:, None, :]
        return self.lm_model(inputs_embeds=emb, attention_mask=attention_mask, labels=labels, return_dict=True)


class FakePeft(nn.Module):
    """What peft.get_peft_model returns, structurally: wrapper.base_model.model is the GPT2LMHeadModel and attribute
    access falls through to it (model.py:555)."""

    def __init__(self, lm):
        super().__init__()
        self.base_model = nn.Module()
        self.base_model.model = lm

    def get_base_model(self):
        return self.base_model.model

    def forward(self, *a, **k):
        return self.base_model.model(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model.model, name)


@pytest.fixture
def oracle_kernels(monkeypatch):
    """Stand-ins (tests only) for the two ops PreferenceLoss / the deferred causal-LM loss call, written with the
    torch ops of the reference: lets the host logic of the swap run on CPU."""

    def compact(hiddens, weight, labels, masks, length_normalize):
        fn = tp.sequence_logprobs_mean if length_normalize else tp.sequence_logprobs_sum
        return [fn(tp.lm_head(h, weight), y, m if m is not None else torch.ones_like(y))
                for h, y, m in zip(hiddens, labels, masks)]

    def seq_logprob(hidden, weight, labels, mask, length_normalize):
        fn = tp.sequence_logprobs_mean if length_normalize else tp.sequence_logprobs_sum
        return (fn(tp.lm_head(hidden, weight), labels, mask.long() if mask is not None else None),)

    def dpo(lw, ll, rc, rr, beta, ls, n):
        return -torch.nn.functional.logsigmoid(beta * (lw - ll)).sum() / n, None, None

    monkeypatch.setattr(ops, "lmhead_seq_logprob_compact", compact)
    monkeypatch.setattr(ops, "lmhead_seq_logprob", seq_logprob)
    monkeypatch.setattr(ops, "dpo_loss", dpo)
    inst._lazy.any_device = True
    yield
    inst._lazy.any_device = False


# ------------------------------------------------------------------------------------------------ CPU: host logic
@pytest.mark.parametrize("peft", [False, True])
def test_fuse_keeps_module_tree_and_checkpoints(peft):
    """fusing must not rename `lm_head.weight` — a checkpoint written by the reference trainer
    (trainer.py:795) loads into a fused model and vice versa, strict, and the weight stays tied to wte."""
    lm = tiny_lm()
    dec = Decoder(FakePeft(lm) if peft else lm)
    before = {k: v.clone() for k, v in dec.state_dict().items()}
    tree_before = [n for n, _ in dec.named_modules()]
    inst.fuse_decoder(dec)
    inst.fuse_decoder(dec)  # idempotent
    assert [n for n, _ in dec.named_modules()] == tree_before
    assert list(dec.state_dict().keys()) == list(before.keys())
    assert any(k.endswith("lm_head.weight") for k in before)
    assert lm.lm_head.weight is lm.transformer.wte.weight
    assert type(lm.lm_head) is nn.Linear and "forward" in lm.lm_head.__dict__
    assert lm.loss_function is inst.lazy_causal_lm_loss  # the GPT2LMHeadModel inside any wrapper, not the wrapper
    fresh = Decoder(FakePeft(tiny_lm()) if peft else tiny_lm())
    fresh.load_state_dict(dec.state_dict(), strict=True)        # fused -> plain
    dec.load_state_dict(before, strict=True)                    # plain -> fused
    # outside a decoder forward (generation, direct calls) the head is the ordinary Linear
    h = torch.randn(2, 5, 32)
    torch.testing.assert_close(lm.lm_head(h), h @ lm.lm_head.weight.T)
    inst.unfuse_decoder(dec)
    assert "forward" not in lm.lm_head.__dict__ and lm.loss_function is not inst.lazy_causal_lm_loss


def test_lazy_flow_through_a_real_gpt2(oracle_kernels):
    """CaptionDecoder-style forward of a real GPT2LMHeadModel with the head fused: logits come back as a LazyLogits
    handle, HF's loss as a DeferredLoss that nobody has to pay for, and PreferenceLoss on the handles equals the
    reference arithmetic on the dense logits — forward and gradients."""
    lm = tiny_lm().eval()   # no dropout: both runs see the same activations
    dec = Decoder(lm)
    g = torch.Generator().manual_seed(1)
    B, T, V = 3, 12, 101
    ids_w, ids_l = torch.randint(0, V, (B, T), generator=g), torch.randint(0, V, (B, T), generator=g)
    lens = torch.tensor([12, 5, 8])
    m = (torch.arange(T)[None] < lens[:, None]).long()
    vis = torch.randn(B, 32, generator=g)

    def step(fused):
        lm.zero_grad()
        if fused:
            inst.fuse_decoder(dec)
            inst._lazy.on = True
        try:
            ow, ol = dec(vis, ids_w, m, ids_w), dec(vis, ids_l, m, ids_l)
        finally:
            inst._lazy.on = False
        return ow, ol

    ow, ol = step(False)
    ref = tp.preference_loss_trainer(ow.logits, ol.logits, ids_w, ids_l, m, m, 0.1)
    ref.backward()
    ref_grad = lm.transformer.wte.weight.grad.clone()
    ref_gen = ow.loss.detach().clone()

    ow, ol = step(True)
    assert isinstance(ow.logits, losses.LazyLogits) and ow.logits.shape == (B, T, V)
    assert isinstance(ow.loss, losses.DeferredLoss) and ow.loss._value is None
    loss = pg.PreferenceLoss(0.1)(ow.logits, ol.logits, ids_w, ids_l, m, m)
    assert ow.loss._value is None                      # generation_loss still unevaluated: nobody read it
    loss.backward()
    torch.testing.assert_close(loss, ref)
    torch.testing.assert_close(lm.transformer.wte.weight.grad, ref_grad)
    assert math.isclose(ow.loss.item(), ref_gen.item(), rel_tol=1e-5)   # ... and correct when somebody does
    torch.testing.assert_close(ow.logits.materialize(), tp.lm_head(ow.logits.hidden, lm.lm_head.weight))
    inst.unfuse_decoder(dec)


@needs_reference
def test_install_on_the_real_reference_package(oracle_kernels):
    """install() against the real package: the names the trainer resolves (trainer.py:76-80, 204-209) are rebound,
    CaptionDecoder is wrapped, a model built afterwards has a lazy head and an unchanged state_dict, and the Stage-2
    micro-step of trainer.py:575-603 gives the unpatched loss / gradient through the swapped-in modules."""
    import importlib
    mm = ref_model.load_package()
    pkg = "preference_guided_image_captioning_alignment"
    trainer_mod = importlib.import_module(pkg + ".training.trainer")
    small = dict(n_embd=32, n_layer=1, n_head=2, clip_layers=1)
    plain = ref_model.build_model(**small)
    ref_pl = mm.PreferenceLoss(0.1)
    try:
        done = pg.install()
        assert trainer_mod.PreferenceLoss is losses.PreferenceLoss and trainer_mod.ContrastiveLoss is losses.ContrastiveLoss
        assert mm.PreferenceLoss is losses.PreferenceLoss and (mm.__name__, "CaptionDecoder.lm_head") in done
        fused = ref_model.build_model(**small)                     # same seed: same weights
        assert list(fused.state_dict().keys()) == list(plain.state_dict().keys())
        fused.load_state_dict(plain.state_dict(), strict=True)
        lm = fused.caption_decoder.lm_model
        assert "forward" in lm.lm_head.__dict__ and lm.lm_head.weight is lm.transformer.wte.weight
        batch = ref_model.stage2_batch(2, seq_len=16, vocab=50257, seed=3, min_len=4, max_len=9)
        plain.eval(), fused.eval()                                 # no dropout: the two steps see identical activations
        l0, p0, _ = ref_model.stage2_micro_step(plain, ref_pl, batch)
        l0.backward()
        l1, p1, _ = ref_model.stage2_micro_step(fused, trainer_mod.PreferenceLoss(0.1), batch)
        assert isinstance(p1["logits"], losses.LazyLogits) and isinstance(p1["generation_loss"], losses.DeferredLoss)
        l1.backward()
        torch.testing.assert_close(l1, l0, rtol=1e-5, atol=1e-6)
        g0 = plain.caption_decoder.lm_model.transformer.wte.weight.grad
        g1 = fused.caption_decoder.lm_model.transformer.wte.weight.grad
        torch.testing.assert_close(g1, g0, rtol=1e-4, atol=1e-7)
        assert math.isclose(float(p1["generation_loss"]), float(p0["generation_loss"]), rel_tol=1e-5)
    finally:
        pg.uninstall()
    assert trainer_mod.PreferenceLoss is ref_pl.__class__ and mm.CaptionDecoder.forward.__name__ == "forward"
    assert not getattr(mm.CaptionDecoder, "_pgica_wrapped", False)


# ------------------------------------------------------------------------------------------------ GPU: the real swap
def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


@pytest.mark.gpu
def test_preference_loss_lazy_logits_real_gpt2(cuda_device):
    """PreferenceLoss.forward on LazyLogits produced by a real GPT2LMHeadModel (d=512, real vocabulary 50260, ragged
    right-padded masks): fused, compacted path vs the reference arithmetic on dense fp32 logits (TF32 off)."""
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    lm = GPT2LMHeadModel(GPT2Config(n_embd=512, n_layer=2, n_head=8, vocab_size=50260, n_positions=128)).to(cuda_device)
    lm.eval()
    dec = Decoder(lm)
    B, T, V = 4, 128, 50260
    g = torch.Generator().manual_seed(5)
    ids = [torch.randint(0, 50257, (B, T), generator=g).to(cuda_device) for _ in range(2)]
    lens = torch.randint(10, 21, (2, B), generator=g)
    masks = [(torch.arange(T)[None] < lens[i][:, None]).long().to(cuda_device) for i in range(2)]
    ids = [torch.where(m.bool(), y, torch.full_like(y, 50257)) for y, m in zip(ids, masks)]   # [PAD] = 50257
    vis = torch.randn(B, 512, generator=g).to(cuda_device) * 0.1

    lm.zero_grad()
    ow, ol = dec(vis, ids[0], masks[0], ids[0]), dec(vis, ids[1], masks[1], ids[1])
    ref = tp.preference_loss_trainer(ow.logits, ol.logits, ids[0], ids[1], masks[0], masks[1], 0.1)
    ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in lm.named_parameters() if p.grad is not None}
    ref_gen = ow.loss.item()

    inst.fuse_decoder(dec)
    lm.zero_grad()
    inst._lazy.on = True
    try:
        ow, ol = dec(vis, ids[0], masks[0], ids[0]), dec(vis, ids[1], masks[1], ids[1])
    finally:
        inst._lazy.on = False
    assert isinstance(ow.logits, losses.LazyLogits)
    loss = pg.PreferenceLoss(0.1)(ow.logits, ol.logits, ids[0], ids[1], masks[0], masks[1])
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    for n, gref in ref_grads.items():
        assert _rel(lm.get_parameter(n).grad, gref) < 1e-2, n
    assert abs(ow.loss.item() - ref_gen) <= 1e-4 * abs(ref_gen)      # the deferred HF loss, from the fused kernel
    inst.unfuse_decoder(dec)


@pytest.mark.gpu
@needs_reference
def test_cfg5_stage2_step_patched_vs_unpatched(cuda_device):
    """BASELINE config 5 on the B200: the reference's random-init 867 M model (CLIP ViT-B/32 + GPT-2 Medium decoder),
    one Stage-2 micro-step as the trainer issues it (trainer.py:575-616), unpatched and with install() applied,
    identical seeds.  Loss within 1e-4 relative, global gradient norm and the tied wte/lm_head gradient within 1e-2,
    generation_loss within 1e-4 (fp32 hidden states and weights are rounded to bf16 operands inside the fused head)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    mm = ref_model.load_package()
    model = ref_model.build_model().to(cuda_device)
    model.train()
    batch = ref_model.stage2_batch(8, seq_len=128, vocab=50257, seed=11, device=cuda_device)

    def run(pl):
        model.zero_grad(set_to_none=True)
        torch.manual_seed(123)                                     # same dropout masks in both runs
        loss, pref, _ = ref_model.stage2_micro_step(model, pl, batch)
        loss.backward()
        gen = float(pref["generation_loss"])
        grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
        gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
        wte = model.caption_decoder.lm_model.transformer.wte.weight.grad.clone()
        return loss.item(), gen, gnorm, wte, pref

    l0, gen0, n0, w0, _ = run(mm.PreferenceLoss(0.1))
    try:
        pg.install()
        inst.fuse_decoder(model.caption_decoder)                   # the model was built before install()
        l1, gen1, n1, w1, pref = run(mm.PreferenceLoss(0.1))       # the rebound name: losses.PreferenceLoss
        assert isinstance(pref["logits"], losses.LazyLogits)
    finally:
        inst.unfuse_decoder(model.caption_decoder)
        pg.uninstall()
    assert abs(l1 - l0) <= 1e-4 * abs(l0), (l1, l0)
    assert abs(n1 - n0) <= 1e-2 * n0, (n1, n0)
    assert _rel(w1, w0) < 1e-2
    assert abs(gen1 - gen0) <= 1e-4 * abs(gen0), (gen1, gen0)

    # Everything install() can fuse (SURVEY 8(f) rows 3 and 4 on top of the LM head): the one-key cross-attention and
    # the LayerNorm + normalise tails draw their own dropout masks, so the comparison runs with dropout off (eval mode,
    # gradients on) — the Stage-2 step and a Stage-1 contrastive step.
    model.eval()
    l0, gen0, n0, w0, _ = run(mm.PreferenceLoss(0.1))
    cap = {"ids": batch["preferred_ids"], "mask": batch["preferred_mask"]}

    def stage1(cl):
        model.zero_grad(set_to_none=True)
        out = model(images=batch["image"], caption_ids=cap["ids"], caption_mask=cap["mask"], mode="contrastive")
        loss = cl(out["image_embeddings"], out["text_embeddings"])
        loss.backward()
        g = model.vision_encoder.projection[0].weight.grad.clone()
        return loss.item(), g, out

    c0, gc0, _ = stage1(mm.ContrastiveLoss(0.07))
    try:
        pg.install()
        pg.fuse_model(model)
        l1, gen1, n1, w1, pref = run(mm.PreferenceLoss(0.1))
        c1, gc1, out1 = stage1(mm.ContrastiveLoss(0.07))
        assert isinstance(pref["logits"], losses.LazyLogits)
        assert "forward" in model.caption_decoder.cross_attention.__dict__
        assert "forward" in model.vision_encoder.projection[-1].__dict__
    finally:
        pg.unfuse_model(model)
        pg.uninstall()
    assert abs(l1 - l0) <= 1e-4 * abs(l0), (l1, l0)
    assert abs(n1 - n0) <= 1e-2 * n0 and _rel(w1, w0) < 1e-2
    assert abs(c1 - c0) <= 1e-4 * abs(c0), (c1, c0)
    assert _rel(gc1, gc0) < 1e-2
    assert "forward" not in model.caption_decoder.cross_attention.__dict__
