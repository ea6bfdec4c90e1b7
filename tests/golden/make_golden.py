"""Mint golden vectors from the REAL reference modules (only runnable where /root/reference exists).

    python tests/golden/make_golden.py

Every fixture stores the exact inputs (float inputs are rounded to bf16-representable values first, so the
bf16 tensor-core path and the fp32 reference see identical numbers) and the reference's outputs computed in
float64 (loss, intermediate log-probs, gradients from the reference's own autograd).
Reference code paths exercised (relative to the reference root, pkg/ = src/preference_guided_image_captioning_alignment/):
  pkg/models/components.py:86-145   ContrastiveLoss (normalise + clamp, mean|sum)
  pkg/models/model.py:957-1000      ContrastiveLoss (trainer variant)
  pkg/models/components.py:321-362  compute_sequence_logprobs
  pkg/models/model.py:1003-1085     PreferenceLoss / _compute_log_probs
  pkg/models/components.py:148-249  DPOPreferenceLoss
  pkg/models/components.py:252-318  NaNSafeGradientNorm
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

torch.set_default_dtype(torch.float64)


def bf16_round(t):
    return t.float().to(torch.bfloat16).double()


def np64(t):
    return t.detach().cpu().numpy().astype(np.float64)


def bf16_bits(t):
    """bf16-representable float tensor -> uint16 bit patterns (half the fixture size; exact)."""
    return t.float().to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def make_ntxent(comp, TrainerCL):
    out = {}
    for seed, (B, D) in [(1234, (64, 512)), (1, (16, 64)), (2, (33, 128))]:
        g = torch.Generator().manual_seed(seed)
        v = bf16_round(torch.randn(B, D, generator=g, dtype=torch.float32))
        t = bf16_round(v.float() + 0.5 * torch.randn(B, D, generator=g, dtype=torch.float32))
        key = f"s{seed}"
        big = B * D > 4096  # cfg-1 sized case: inputs as bf16 bits, gradients only for the default variant, fp32
        out[key + "_v_bf16"], out[key + "_t_bf16"] = bf16_bits(v), bf16_bits(t)
        for tau in (0.5, 0.07):
            for red in ("mean", "sum"):
                vv, tt = v.clone().requires_grad_(True), t.clone().requires_grad_(True)
                loss = comp.ContrastiveLoss(temperature=tau, reduction=red).double()(vv, tt)
                loss.backward()
                k = f"{key}_comp_tau{tau}_{red}"
                out[k + "_loss"] = np64(loss)
                if not big or (tau == 0.5 and red == "mean"):
                    out[k + "_dv"], out[k + "_dt"] = np64(vv.grad).astype(np.float32), np64(tt.grad).astype(np.float32)
        # trainer variant expects pre-normalised inputs (model.py:826-829), here normalised then bf16-rounded
        vn, tn = bf16_round(F.normalize(v, dim=-1)), bf16_round(F.normalize(t, dim=-1))
        out[key + "_vn_bf16"], out[key + "_tn_bf16"] = bf16_bits(vn), bf16_bits(tn)
        for tau in (0.5, 0.07):
            vv, tt = vn.clone().requires_grad_(True), tn.clone().requires_grad_(True)
            loss = TrainerCL(temperature=tau)(vv, tt)
            loss.backward()
            k = f"{key}_trainer_tau{tau}"
            out[k + "_loss"] = np64(loss)
            if not big or tau == 0.5:
                out[k + "_dv"], out[k + "_dt"] = np64(vv.grad).astype(np.float32), np64(tt.grad).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "ntxent.npz"), **out)


def make_seq_logprobs(comp, TrainerPL):
    out = {}
    g = torch.Generator().manual_seed(7)
    B, T, V = 3, 10, 50
    logits = bf16_round(torch.randn(B, T, V, generator=g, dtype=torch.float32) * 2)
    labels = torch.randint(0, V, (B, T), generator=g)
    lens = torch.tensor([10, 6, 8])
    mask_i = (torch.arange(T)[None, :] < lens[:, None]).long()
    mask_f = mask_i.double() * torch.tensor([1.0, 0.5, 1.0])[:, None]  # float masks are multiplied in
    out.update(logits=np64(logits), labels=labels.numpy(), mask_i=mask_i.numpy(), mask_f=np64(mask_f))
    for name, m in (("none", None), ("i64", mask_i), ("f64", mask_f)):
        lg = logits.clone().requires_grad_(True)
        s = comp.compute_sequence_logprobs(lg, labels, m)
        gs = torch.arange(1, B + 1).double()
        (s * gs).sum().backward()
        out[f"sum_{name}"], out[f"sum_{name}_dlogits"] = np64(s), np64(lg.grad)
    pl = TrainerPL(beta=0.1)
    for name, m in (("i64", mask_i), ("f64", mask_f)):
        lg = logits.clone().requires_grad_(True)
        s = pl._compute_log_probs(lg, labels, m)
        (s * torch.arange(1, B + 1).double()).sum().backward()
        out[f"mean_{name}"], out[f"mean_{name}_dlogits"] = np64(s), np64(lg.grad)
    # trainer-variant loss on two logits tensors
    logits2 = bf16_round(torch.randn(B, T, V, generator=g, dtype=torch.float32) * 2)
    labels2 = torch.randint(0, V, (B, T), generator=g)
    mask2 = (torch.arange(T)[None, :] < torch.tensor([7, 10, 5])[:, None]).long()
    a, b = logits.clone().requires_grad_(True), logits2.clone().requires_grad_(True)
    loss = pl(a, b, labels, labels2, mask_i, mask2)
    loss.backward()
    out.update(logits2=np64(logits2), labels2=labels2.numpy(), mask2=mask2.numpy(), pref_loss=np64(loss),
               pref_dlogits=np64(a.grad), pref_dlogits2=np64(b.grad))
    np.savez_compressed(os.path.join(HERE, "seq_logprobs.npz"), **out)


def make_dpo(comp):
    out = {}
    g = torch.Generator().manual_seed(11)
    n = 16
    pc, pr, rc, rr = [bf16_round(torch.randn(n, generator=g, dtype=torch.float32) * 30 - 600) for _ in range(4)]
    out.update(pc=np64(pc), pr=np64(pr), rc=np64(rc), rr=np64(rr))
    for tag, kw, use_ref in (("std", dict(beta=0.1), True), ("ls", dict(beta=0.1, label_smoothing=0.1), True),
                             ("free", dict(beta=0.1, reference_free=True), True), ("noref", dict(beta=0.25), False)):
        xs = [t.clone().requires_grad_(True) for t in (pc, pr, rc, rr)]
        mod = comp.DPOPreferenceLoss(**kw)
        loss, metrics = mod(xs[0], xs[1], xs[2], xs[3]) if use_ref else mod(xs[0], xs[1])
        loss.backward()
        out[tag + "_loss"] = np64(loss)
        out[tag + "_metrics"] = np.array([metrics[k] for k in ("dpo_loss", "reward_margin", "reward_accuracy",
                                                               "policy_chosen_logprob", "policy_rejected_logprob")])
        for nm, x in zip(("dpc", "dpr", "drc", "drr"), xs):
            out[f"{tag}_{nm}"] = np64(x.grad) if x.grad is not None else np.zeros(n)
    np.savez_compressed(os.path.join(HERE, "dpo_loss.npz"), **out)


def make_dpo_head(comp):
    """Hidden-state level composite: lm_head (F.linear, bias-free, modeling_gpt2.py:706) -> components."""
    out = {}
    g = torch.Generator().manual_seed(1234)
    B, T, d, V = 4, 12, 64, 300
    r = lambda *s, sc=1.0: bf16_round(torch.randn(*s, generator=g, dtype=torch.float32) * sc)
    W, Wr = r(V, d, sc=0.2), r(V, d, sc=0.2)
    hc, hr, rhc, rhr = r(B, T, d), r(B, T, d), r(B, T, d), r(B, T, d)
    yc, yr = torch.randint(0, V, (B, T), generator=g), torch.randint(0, V, (B, T), generator=g)
    lc, lr = torch.randint(T // 2, T + 1, (B,), generator=g), torch.randint(T // 2, T + 1, (B,), generator=g)
    mc = (torch.arange(T)[None, :] < lc[:, None]).long()
    mr = (torch.arange(T)[None, :] < lr[:, None]).long()
    out.update(W=np64(W), Wr=np64(Wr), hc=np64(hc), hr=np64(hr), rhc=np64(rhc), rhr=np64(rhr), yc=yc.numpy(),
               yr=yr.numpy(), mc=mc.numpy(), mr=mr.numpy())
    Wg, hcg, hrg = W.clone().requires_grad_(True), hc.clone().requires_grad_(True), hr.clone().requires_grad_(True)
    pc = comp.compute_sequence_logprobs(F.linear(hcg, Wg), yc, mc)
    pr = comp.compute_sequence_logprobs(F.linear(hrg, Wg), yr, mr)
    with torch.no_grad():
        rc = comp.compute_sequence_logprobs(F.linear(rhc, Wr), yc, mc)
        rr = comp.compute_sequence_logprobs(F.linear(rhr, Wr), yr, mr)
    loss, metrics = comp.DPOPreferenceLoss(beta=0.1)(pc, pr, rc, rr)
    loss.backward()
    out.update(pc=np64(pc), pr=np64(pr), rc=np64(rc), rr=np64(rr), loss=np64(loss), dW=np64(Wg.grad),
               dhc=np64(hcg.grad), dhr=np64(hrg.grad),
               metrics=np.array([metrics[k] for k in ("dpo_loss", "reward_margin", "reward_accuracy",
                                                      "policy_chosen_logprob", "policy_rejected_logprob")]))
    np.savez_compressed(os.path.join(HERE, "dpo_head.npz"), **out)


def make_grad_clip(comp):
    """pkg/models/components.py:252-318 NaNSafeGradientNorm on a handful of 'parameters' of odd sizes (fp32)."""
    out = {}
    g = torch.Generator().manual_seed(77)
    shapes = [(3, 5), (1000,), (65, 33), (9001,), (8, 8, 8)]
    for tag, scale, max_norm in (("clip", 1.0, 1.0), ("noclip", 1e-3, 1.0), ("big", 30.0, 0.5)):
        params = []
        for i, sh in enumerate(shapes):
            p = torch.nn.Parameter(torch.zeros(*sh, dtype=torch.float32))
            p.grad = (torch.randn(*sh, generator=g, dtype=torch.float32) * scale)
            out[f"{tag}_g{i}"] = p.grad.numpy().copy()
            params.append(p)
        total, finite = comp.NaNSafeGradientNorm(max_norm=max_norm)(params)
        out[f"{tag}_max_norm"] = np.float64(max_norm)
        out[f"{tag}_total"] = np64(total)
        out[f"{tag}_finite"] = np.int64(bool(finite))
        for i, p in enumerate(params):
            out[f"{tag}_c{i}"] = p.grad.numpy().copy()
    # a non-finite gradient: reported, nothing clipped
    params = []
    for i, sh in enumerate(shapes[:2]):
        p = torch.nn.Parameter(torch.zeros(*sh, dtype=torch.float32))
        p.grad = torch.randn(*sh, generator=g, dtype=torch.float32) * 5
        params.append(p)
    params[1].grad[17] = float("nan")
    for i, p in enumerate(params):
        out[f"nan_g{i}"] = p.grad.numpy().copy()
    total, finite = comp.NaNSafeGradientNorm(max_norm=1.0)(params)
    out["nan_finite"] = np.int64(bool(finite))
    for i, p in enumerate(params):
        out[f"nan_c{i}"] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "grad_clip.npz"), **out)


def make_fp32_inputs(comp, TrainerCL, TrainerPL):
    """Inputs that are NOT bf16-representable: fp32 unit vectors as the model hands them to the trainer's
    ContrastiveLoss (model.py:826-829 -> 970-1000), fp32 embeddings for the components flavour, and an fp32 hidden /
    weight pair for the LM head + PreferenceLoss._compute_log_probs.  The reference runs in float64 on the fp32 values."""
    out = {}
    for seed, (B, D) in [(5, (32, 512)), (6, (40, 128))]:
        g = torch.Generator().manual_seed(seed)
        v = torch.randn(B, D, generator=g, dtype=torch.float32)
        t = (v + 0.5 * torch.randn(B, D, generator=g, dtype=torch.float32))
        vn, tn = F.normalize(v, dim=-1), F.normalize(t, dim=-1)           # fp32 values, not bf16-representable
        key = f"s{seed}"
        out[key + "_v"], out[key + "_t"] = v.numpy(), t.numpy()
        out[key + "_vn"], out[key + "_tn"] = vn.numpy(), tn.numpy()
        for tau in (0.5, 0.07):
            vv, tt = vn.double().requires_grad_(True), tn.double().requires_grad_(True)
            loss = TrainerCL(temperature=tau)(vv, tt)
            loss.backward()
            k = f"{key}_trainer_tau{tau}"
            out[k + "_loss"] = np64(loss)
            out[k + "_dv"], out[k + "_dt"] = np64(vv.grad).astype(np.float32), np64(tt.grad).astype(np.float32)
            vv, tt = v.double().requires_grad_(True), t.double().requires_grad_(True)
            loss = comp.ContrastiveLoss(temperature=tau).double()(vv, tt)
            loss.backward()
            k = f"{key}_comp_tau{tau}"
            out[k + "_loss"] = np64(loss)
            out[k + "_dv"], out[k + "_dt"] = np64(vv.grad).astype(np.float32), np64(tt.grad).astype(np.float32)
    # LM head on fp32 hidden / weight (what install() sees): trainer-variant mean log-probs + PreferenceLoss
    g = torch.Generator().manual_seed(8)
    B, T, d, V = 3, 12, 64, 257
    f32 = dict(generator=g, dtype=torch.float32)
    W = torch.randn(V, d, **f32) * 0.05
    hw, hl = torch.randn(B, T, d, **f32), torch.randn(B, T, d, **f32)
    yw, yl = torch.randint(0, V, (B, T), generator=g), torch.randint(0, V, (B, T), generator=g)
    lens = torch.tensor([[12, 7, 4], [9, 12, 5]])
    mw = (torch.arange(T)[None] < lens[0][:, None]).long()
    ml = (torch.arange(T)[None] < lens[1][:, None]).long()
    Wd, hwd, hld = (x.double().requires_grad_(True) for x in (W, hw, hl))
    pl = TrainerPL(beta=0.1)
    loss = pl(F.linear(hwd, Wd), F.linear(hld, Wd), yw, yl, mw, ml)
    loss.backward()
    out.update(lm_W=W.numpy(), lm_hw=hw.numpy(), lm_hl=hl.numpy(), lm_yw=yw.numpy(), lm_yl=yl.numpy(), lm_mw=mw.numpy(),
               lm_ml=ml.numpy(), lm_loss=np64(loss), lm_dW=np64(Wd.grad).astype(np.float32),
               lm_dhw=np64(hwd.grad).astype(np.float32), lm_dhl=np64(hld.grad).astype(np.float32),
               lm_lpw=np64(pl._compute_log_probs(F.linear(hwd, Wd), yw, mw)))
    np.savez_compressed(os.path.join(HERE, "fp32_inputs.npz"), **out)


def main():
    if not ref_loader.available():
        raise SystemExit("reference not found at " + ref_loader.REFERENCE_ROOT)
    comp = ref_loader.load_components()
    TrainerCL, TrainerPL = ref_loader.load_model_losses()
    make_ntxent(comp, TrainerCL)
    make_seq_logprobs(comp, TrainerPL)
    make_dpo(comp)
    make_dpo_head(comp)
    make_grad_clip(comp)
    make_fp32_inputs(comp, TrainerCL, TrainerPL)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
