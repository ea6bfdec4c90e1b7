"""CPU tests: the oracle restatements (oracle/closed_form.py, oracle/torch_port.py) against golden vectors
minted from the REAL reference (tests/golden/make_golden.py), plus the closed-form known answers of
SURVEY.md §8c.  These pin the oracle; the GPU parity tests then compare the CUDA path against it."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_loader
from oracle import torch_port as tp

METRIC_KEYS = ("dpo_loss", "reward_margin", "reward_accuracy", "policy_chosen_logprob", "policy_rejected_logprob")


def bits_to_f64(u16):
    return torch.from_numpy(u16.astype(np.int16)).view(torch.bfloat16).double().numpy()


@pytest.fixture(scope="module")
def ntx(golden_dir):
    return np.load(os.path.join(golden_dir, "ntxent.npz"))


@pytest.fixture(scope="module")
def seqg(golden_dir):
    return np.load(os.path.join(golden_dir, "seq_logprobs.npz"))


@pytest.fixture(scope="module")
def dpog(golden_dir):
    return np.load(os.path.join(golden_dir, "dpo_loss.npz"))


@pytest.fixture(scope="module")
def headg(golden_dir):
    return np.load(os.path.join(golden_dir, "dpo_head.npz"))


# ------------------------------------------------------------------------------------------ NT-Xent
@pytest.mark.parametrize("seed", [1234, 1, 2])
@pytest.mark.parametrize("tau", [0.5, 0.07])
@pytest.mark.parametrize("red", ["mean", "sum"])
def test_ntxent_components_closed_form(ntx, seed, tau, red):
    v, t = bits_to_f64(ntx[f"s{seed}_v_bf16"]), bits_to_f64(ntx[f"s{seed}_t_bf16"])
    out = cf.ntxent(v, t, tau, normalize=True, clamp_tau=True, reduction=red)
    k = f"s{seed}_comp_tau{tau}_{red}"
    assert out["loss"] == pytest.approx(float(ntx[k + "_loss"]), rel=1e-12)
    if k + "_dv" in ntx:
        np.testing.assert_allclose(out["dx"], ntx[k + "_dv"], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(out["dy"], ntx[k + "_dt"], rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("seed", [1234, 1, 2])
@pytest.mark.parametrize("tau", [0.5, 0.07])
def test_ntxent_trainer_closed_form(ntx, seed, tau):
    v, t = bits_to_f64(ntx[f"s{seed}_vn_bf16"]), bits_to_f64(ntx[f"s{seed}_tn_bf16"])
    out = cf.ntxent(v, t, tau, normalize=False, clamp_tau=False)
    k = f"s{seed}_trainer_tau{tau}"
    assert out["loss"] == pytest.approx(float(ntx[k + "_loss"]), rel=1e-12)
    if k + "_dv" in ntx:
        np.testing.assert_allclose(out["dx"], ntx[k + "_dv"], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(out["dy"], ntx[k + "_dt"], rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("seed", [1, 2])
def test_ntxent_torch_port(ntx, seed):
    v = torch.from_numpy(bits_to_f64(ntx[f"s{seed}_v_bf16"]))
    t = torch.from_numpy(bits_to_f64(ntx[f"s{seed}_t_bf16"]))
    for tau in (0.5, 0.07):
        for red in ("mean", "sum"):
            got = tp.ntxent_components(v, t, tau, red)
            assert got.item() == pytest.approx(float(ntx[f"s{seed}_comp_tau{tau}_{red}_loss"]), rel=1e-12)
    vn = torch.from_numpy(bits_to_f64(ntx[f"s{seed}_vn_bf16"]))
    tn = torch.from_numpy(bits_to_f64(ntx[f"s{seed}_tn_bf16"]))
    for tau in (0.5, 0.07):
        assert tp.ntxent_trainer(vn, tn, tau).item() == pytest.approx(float(ntx[f"s{seed}_trainer_tau{tau}_loss"]),
                                                                      rel=1e-12)


def test_ntxent_known_answers():
    rng = np.random.default_rng(0)
    B, D = 8, 16
    # KA3: identical rows -> ln B
    x = np.tile(rng.standard_normal((1, D)), (B, 1))
    assert cf.ntxent(x, x, 0.5, normalize=True, clamp_tau=True)["loss"] == pytest.approx(math.log(B), rel=1e-12)
    # KA4: orthonormal one-hot rows, a == b -> ln(1 + (B-1) e^{-1/tau})
    e = np.eye(B, D)
    for tau in (0.5, 0.2):
        assert cf.ntxent(e, e, tau)["loss"] == pytest.approx(math.log(1 + (B - 1) * math.exp(-1 / tau)), rel=1e-12)
    # KA5: aligned embeddings, low temperature gives the lower loss (reference tests/test_model.py:452-466)
    u, _ = cf.l2_normalize(rng.standard_normal((4, 256)))
    assert cf.ntxent(u, u, 0.01)["loss"] < cf.ntxent(u, u, 1.0)["loss"]
    # KA6: components variant clamps tau=0.07 to 0.1
    v, t = rng.standard_normal((B, D)), rng.standard_normal((B, D))
    assert cf.ntxent(v, t, 0.07, True, True)["loss"] == pytest.approx(cf.ntxent(v, t, 0.1, True, True)["loss"],
                                                                      rel=1e-14)


def test_ntxent_row_slices_sum_to_global():
    """The multi-GPU decomposition: row-slice gradients (dx final, dy partial) add up to the global ones."""
    rng = np.random.default_rng(3)
    B, D, W = 12, 8, 3
    x, _ = cf.l2_normalize(rng.standard_normal((B, D)))
    y, _ = cf.l2_normalize(rng.standard_normal((B, D)))
    full = cf.ntxent(x, y, 0.5)
    b = B // W
    S = full["sim"]
    lse_c = full["lse_col"]
    dy = np.zeros_like(y)
    for r in range(W):
        rows = slice(r * b, (r + 1) * b)
        Sr = S[rows]
        lr = np.log(np.exp(Sr).sum(1))
        onehot = np.zeros_like(Sr)
        onehot[np.arange(b), np.arange(b) + r * b] = 1
        dS = (np.exp(Sr - lr[:, None]) + np.exp(Sr - lse_c[None, :]) - 2 * onehot) / (2 * B)
        np.testing.assert_allclose(dS @ y / 0.5, full["dx"][rows], rtol=1e-10, atol=1e-14)
        dy += dS.T @ x[rows] / 0.5
    np.testing.assert_allclose(dy, full["dy"], rtol=1e-10, atol=1e-14)


# ------------------------------------------------------------------------------------------ sequence log-probs
@pytest.mark.parametrize("name", ["none", "i64", "f64"])
def test_sequence_logprobs_sum(seqg, name):
    mask = {"none": None, "i64": seqg["mask_i"], "f64": seqg["mask_f"]}[name]
    got = cf.sequence_logprobs(seqg["logits"], seqg["labels"], mask, length_normalize=False)
    np.testing.assert_allclose(got, seqg[f"sum_{name}"], rtol=1e-12)
    lg = torch.from_numpy(seqg["logits"]).requires_grad_(True)
    m = None if mask is None else torch.from_numpy(mask)
    s = tp.sequence_logprobs_sum(lg, torch.from_numpy(seqg["labels"]), m)
    np.testing.assert_allclose(s.detach().numpy(), seqg[f"sum_{name}"], rtol=1e-12)
    (s * torch.arange(1, 4).double()).sum().backward()
    np.testing.assert_allclose(lg.grad.numpy(), seqg[f"sum_{name}_dlogits"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("name", ["i64", "f64"])
def test_sequence_logprobs_mean(seqg, name):
    mask = {"i64": seqg["mask_i"], "f64": seqg["mask_f"]}[name]
    got = cf.sequence_logprobs(seqg["logits"], seqg["labels"], mask, length_normalize=True)
    np.testing.assert_allclose(got, seqg[f"mean_{name}"], rtol=1e-12)
    s = tp.sequence_logprobs_mean(torch.from_numpy(seqg["logits"]), torch.from_numpy(seqg["labels"]),
                                  torch.from_numpy(mask))
    np.testing.assert_allclose(s.numpy(), seqg[f"mean_{name}"], rtol=1e-12)


def test_preference_loss_trainer(seqg):
    loss, _, _ = cf.preference_loss_from_logits(seqg["logits"], seqg["logits2"], seqg["labels"], seqg["labels2"],
                                                seqg["mask_i"], seqg["mask2"], beta=0.1)
    assert loss == pytest.approx(float(seqg["pref_loss"]), rel=1e-12)
    t = lambda k: torch.from_numpy(seqg[k])
    got = tp.preference_loss_trainer(t("logits"), t("logits2"), t("labels"), t("labels2"), t("mask_i"), t("mask2"), 0.1)
    assert got.item() == pytest.approx(float(seqg["pref_loss"]), rel=1e-12)


# ------------------------------------------------------------------------------------------ DPO scalar head
@pytest.mark.parametrize("tag,kw,use_ref", [("std", dict(beta=0.1), True),
                                            ("ls", dict(beta=0.1, label_smoothing=0.1), True),
                                            ("free", dict(beta=0.1, reference_free=True), True),
                                            ("noref", dict(beta=0.25), False)])
def test_dpo_loss(dpog, tag, kw, use_ref):
    a = [dpog[k] for k in ("pc", "pr", "rc", "rr")]
    out = cf.dpo_loss(a[0], a[1], a[2] if use_ref else None, a[3] if use_ref else None, **kw)
    assert out["loss"] == pytest.approx(float(dpog[tag + "_loss"]), rel=1e-12)
    np.testing.assert_allclose([out["metrics"][k] for k in METRIC_KEYS], dpog[tag + "_metrics"], rtol=1e-12)
    np.testing.assert_allclose(out["d_pc"], dpog[tag + "_dpc"], rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(out["d_pr"], dpog[tag + "_dpr"], rtol=1e-10, atol=1e-16)
    if out["d_rc"] is not None:
        np.testing.assert_allclose(out["d_rc"], dpog[tag + "_drc"], rtol=1e-10, atol=1e-16)
        np.testing.assert_allclose(out["d_rr"], dpog[tag + "_drr"], rtol=1e-10, atol=1e-16)
    else:
        assert not dpog[tag + "_drc"].any()
    ts = [torch.from_numpy(x) for x in a]
    loss, metrics = tp.dpo_components(ts[0], ts[1], ts[2] if use_ref else None, ts[3] if use_ref else None, **kw)
    assert loss.item() == pytest.approx(float(dpog[tag + "_loss"]), rel=1e-12)
    np.testing.assert_allclose([metrics[k] for k in METRIC_KEYS], dpog[tag + "_metrics"], rtol=1e-12)


def test_dpo_known_answers():
    rng = np.random.default_rng(5)
    B, T, d, V = 3, 7, 16, 40
    h = rng.standard_normal((2, B, T, d))
    W = rng.standard_normal((V, d)) * 0.3
    y = rng.integers(0, V, (2, B, T))
    # KA1: policy == reference -> ln 2, margin 0, accuracy 0 (strict >)
    out = cf.dpo_head(h[0], h[1], W, y[0], y[1], ref=dict(hc=h[0], hr=h[1], W=W))
    assert out["loss"] == pytest.approx(math.log(2), rel=1e-14)
    assert out["metrics"]["reward_margin"] == 0 and out["metrics"]["reward_accuracy"] == 0
    # KA2: W = 0 -> every token log-prob is -ln V
    z = cf.lmhead_sequence_logprobs(h[0], np.zeros_like(W), y[0])
    np.testing.assert_allclose(z["seq_logp"], -(T - 1) * math.log(V), rtol=1e-14)
    zl = cf.lmhead_sequence_logprobs(h[0], np.zeros_like(W), y[0], length_normalize=True)
    np.testing.assert_allclose(zl["seq_logp"], -math.log(V), rtol=1e-14)
    # KA7: trainer variant == reference-free DPO on length-normalised log-probs
    m = (np.arange(T)[None, :] < np.array([7, 4, 5])[:, None]).astype(np.int64)
    lw = cf.sequence_logprobs(h[0] @ W.T, y[0], m, True)
    ll = cf.sequence_logprobs(h[1] @ W.T, y[1], m, True)
    a, _, _ = cf.preference_loss_from_logits(h[0] @ W.T, h[1] @ W.T, y[0], y[1], m, m, beta=0.1)
    assert a == cf.dpo_loss(lw, ll, beta=0.1, reference_free=True)["loss"]
    # KA8: label smoothing identity
    x = rng.standard_normal(6) * 3
    ls = 0.2
    want = np.mean(-(1 - ls) * np.log(1 / (1 + np.exp(-x))) - ls * np.log(1 / (1 + np.exp(x))))
    assert cf.dpo_loss(x / 0.1, np.zeros(6), beta=0.1, label_smoothing=ls)["loss"] == pytest.approx(want, rel=1e-12)
    # KA9: all-ones mask == no mask; an all-zero mask row gives 0 (sum) / NaN (mean)
    np.testing.assert_array_equal(cf.sequence_logprobs(h[0] @ W.T, y[0], np.ones((B, T))),
                                  cf.sequence_logprobs(h[0] @ W.T, y[0], None))
    m0 = m.copy()
    m0[1] = 0
    assert cf.sequence_logprobs(h[0] @ W.T, y[0], m0)[1] == 0
    assert np.isnan(cf.sequence_logprobs(h[0] @ W.T, y[0], m0, True)[1])


def test_dpo_head_composite(headg):
    g = headg
    out = cf.dpo_head(g["hc"], g["hr"], g["W"], g["yc"], g["yr"], g["mc"], g["mr"],
                      ref=dict(hc=g["rhc"], hr=g["rhr"], W=g["Wr"]), beta=0.1)
    assert out["loss"] == pytest.approx(float(g["loss"]), rel=1e-12)
    for k in ("pc", "pr", "rc", "rr"):
        np.testing.assert_allclose(out[k], g[k], rtol=1e-12)
    np.testing.assert_allclose([out["metrics"][k] for k in METRIC_KEYS], g["metrics"], rtol=1e-10)
    np.testing.assert_allclose(out["dW"], g["dW"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(out["dhc"], g["dhc"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(out["dhr"], g["dhr"], rtol=1e-9, atol=1e-15)
    # torch port of the same step (what bench.py times as the CPU baseline)
    t = lambda k: torch.from_numpy(g[k])
    W = t("W").clone().requires_grad_(True)
    hc, hr = t("hc").clone().requires_grad_(True), t("hr").clone().requires_grad_(True)
    loss, metrics = tp.dpo_head_step(hc, hr, W, t("yc"), t("yr"), t("mc"), t("mr"), t("rhc"), t("rhr"), t("Wr"), 0.1)
    assert loss.item() == pytest.approx(float(g["loss"]), rel=1e-12)
    np.testing.assert_allclose(W.grad.numpy(), g["dW"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(hc.grad.numpy(), g["dhc"], rtol=1e-9, atol=1e-15)


# ------------------------------------------------------------------------------------------ live reference
@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_against_live_reference():
    comp = ref_loader.load_components()
    CL, PL = ref_loader.load_model_losses()
    g = torch.Generator().manual_seed(99)
    v, t = torch.randn(10, 32, generator=g, dtype=torch.float64), torch.randn(10, 32, generator=g, dtype=torch.float64)
    # the reference stores tau as a float32 tensor (components.py:56-59), so 0.3 is 0.30000001192...: compare at
    # a float32-exact temperature for the tight check and at 0.3 with a float32-epsilon tolerance
    assert comp.ContrastiveLoss(0.25).double()(v, t).item() == pytest.approx(
        cf.ntxent(v.numpy(), t.numpy(), 0.25, True, True)["loss"], rel=1e-12)
    assert comp.ContrastiveLoss(0.3).double()(v, t).item() == pytest.approx(
        cf.ntxent(v.numpy(), t.numpy(), 0.3, True, True)["loss"], rel=1e-7)
    vn, tn = torch.nn.functional.normalize(v, dim=-1), torch.nn.functional.normalize(t, dim=-1)
    assert CL(0.07)(vn, tn).item() == pytest.approx(cf.ntxent(vn.numpy(), tn.numpy(), 0.07)["loss"], rel=1e-12)
    logits = torch.randn(2, 6, 30, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 30, (2, 6), generator=g)
    mask = torch.ones(2, 6, dtype=torch.float64)
    np.testing.assert_allclose(PL(0.1)._compute_log_probs(logits, labels, mask).numpy(),
                               cf.sequence_logprobs(logits.numpy(), labels.numpy(), mask.numpy(), True), rtol=1e-12)


# ================================================================================ SURVEY 8(f) row 2: grad norm + clip
@pytest.mark.parametrize("tag", ["clip", "noclip", "big", "nan"])
def test_grad_norm_clip_closed_form(golden_dir, tag):
    """oracle.closed_form.grad_norm_clip against the real NaNSafeGradientNorm (components.py:252-318) outputs."""
    g = np.load(os.path.join(golden_dir, "grad_clip.npz"))
    n = 2 if tag == "nan" else 5
    grads = [g[f"{tag}_g{i}"] for i in range(n)]
    max_norm = 1.0 if tag == "nan" else float(g[f"{tag}_max_norm"])
    o = cf.grad_norm_clip(grads, max_norm)
    assert o["is_finite"] == bool(g[f"{tag}_finite"])
    if tag != "nan":
        assert abs(o["total_norm"] - float(g[f"{tag}_total"])) <= 1e-6 * float(g[f"{tag}_total"])  # reference is fp32
        assert (o["clip_coef"] < 1.0) == (tag != "noclip")
    for i in range(n):
        np.testing.assert_allclose(o["clipped"][i], g[f"{tag}_c{i}"], rtol=2e-6, atol=0, equal_nan=True)


def test_fp32_input_fixture_against_oracle(golden_dir):
    """tests/golden/fp32_inputs.npz (inputs that are not bf16-representable, minted from the real reference in
    float64): both restatements reproduce it."""
    g = np.load(os.path.join(golden_dir, "fp32_inputs.npz"))
    for seed in (5, 6):
        for tau in (0.5, 0.07):
            vn, tn = g[f"s{seed}_vn"].astype(np.float64), g[f"s{seed}_tn"].astype(np.float64)
            o = cf.ntxent(vn, tn, tau, normalize=False, clamp_tau=False)
            assert abs(o["loss"] - float(g[f"s{seed}_trainer_tau{tau}_loss"])) < 1e-12
            np.testing.assert_allclose(o["dx"], g[f"s{seed}_trainer_tau{tau}_dv"], rtol=2e-6, atol=1e-9)
            v, t = g[f"s{seed}_v"].astype(np.float64), g[f"s{seed}_t"].astype(np.float64)
            o = cf.ntxent(v, t, tau, normalize=True, clamp_tau=True)
            assert abs(o["loss"] - float(g[f"s{seed}_comp_tau{tau}_loss"])) < 1e-12
            np.testing.assert_allclose(o["dy"], g[f"s{seed}_comp_tau{tau}_dt"], rtol=2e-6, atol=1e-9)
    import torch
    from oracle import torch_port as tp
    W, hw, hl = (torch.from_numpy(g[k]).double() for k in ("lm_W", "lm_hw", "lm_hl"))
    yw, yl, mw, ml = (torch.from_numpy(g[k]) for k in ("lm_yw", "lm_yl", "lm_mw", "lm_ml"))
    loss = tp.preference_loss_trainer(tp.lm_head(hw, W), tp.lm_head(hl, W), yw, yl, mw, ml, 0.1)
    assert abs(loss.item() - float(g["lm_loss"])) < 1e-12
