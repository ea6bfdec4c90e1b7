"""GPU parity tests (`-m gpu`, B200): the CUDA path — module -> torch.library op -> C ABI -> kernels — against
(1) golden vectors minted from the real reference, (2) the float64 oracle on seeded inputs at sizes it finishes
in seconds, (3) size-independent properties at BASELINE.json's full sizes.

Tolerances are the ones BASELINE.json states: bf16 inputs with fp32 accumulation -> loss within 1e-4 relative,
gradients within 1e-2 relative (||delta|| / ||ref||), index / mask plumbing bit-exact."""
import math
import os
import types

import numpy as np
import pytest
import torch

from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-2
METRIC_KEYS = ("dpo_loss", "reward_margin", "reward_accuracy", "policy_chosen_logprob", "policy_rejected_logprob")


def bits_to_f32(u16, dev):
    return torch.from_numpy(u16.astype(np.int16)).view(torch.bfloat16).float().to(dev)


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def bf16r(t):
    return t.to(torch.bfloat16).float()


def loss_close(got, ref, logit_scale=1.0):
    """1e-4 relative (BASELINE.json) plus an absolute term of a few fp32 ulps on the logit scale: with well-aligned
    pairs and a small temperature the loss is a tiny difference of O(1/tau) terms (lse - diag), where even the
    fp32 reference carries ~1e-6 * |logit| of rounding noise."""
    return abs(got - ref) <= LOSS_RTOL * abs(ref) + 2e-6 * max(1.0, logit_scale)


def set_plan(plan):
    """"R2,C2" pins the dual kernel's role split through the option API; None hands it back to the planner."""
    from preference_guided_image_captioning_alignment_b200 import _lib
    r2, c2 = (int(v) for v in plan.split(",")) if plan else (0, 0)
    _lib.set_option("sggf_plan_r2", r2)
    _lib.set_option("sggf_plan_c2", c2)


@pytest.fixture(autouse=True)
def _reset_library_options():
    yield
    try:
        from preference_guided_image_captioning_alignment_b200 import _lib
        set_plan(None)
        _lib.set_option("sgg_fused", 1)
        _lib.set_option("sgg_cluster", 0)
    except Exception:
        pass


@pytest.fixture(scope="module")
def pg(cuda_device):
    import preference_guided_image_captioning_alignment_b200 as pkg
    from preference_guided_image_captioning_alignment_b200 import _lib
    lib = _lib.load()
    assert lib.pgica_device_check() == 0, lib.pgica_last_error()
    return pkg


# ================================================================================================ NT-Xent
@pytest.mark.parametrize("seed", [1234, 1, 2])
@pytest.mark.parametrize("tau,red", [(0.5, "mean"), (0.5, "sum"), (0.07, "mean")])
def test_ntxent_components_golden(pg, cuda_device, golden_dir, seed, tau, red):
    from preference_guided_image_captioning_alignment_b200 import components
    g = np.load(os.path.join(golden_dir, "ntxent.npz"))
    v = bits_to_f32(g[f"s{seed}_v_bf16"], cuda_device).requires_grad_(True)
    t = bits_to_f32(g[f"s{seed}_t_bf16"], cuda_device).requires_grad_(True)
    loss = components.ContrastiveLoss(temperature=tau, reduction=red)(v, t)
    k = f"s{seed}_comp_tau{tau}_{red}"
    assert loss.dim() == 0
    assert loss_close(loss.item(), float(g[k + "_loss"]), 1.0 / max(tau, 0.1)), (loss.item(), float(g[k + "_loss"]))
    loss.backward()
    if k + "_dv" in g:
        assert rel(v.grad, g[k + "_dv"]) < GRAD_RTOL
        assert rel(t.grad, g[k + "_dt"]) < GRAD_RTOL


@pytest.mark.parametrize("seed", [1234, 1, 2])
@pytest.mark.parametrize("tau", [0.5, 0.07])
def test_ntxent_trainer_golden(pg, cuda_device, golden_dir, seed, tau):
    g = np.load(os.path.join(golden_dir, "ntxent.npz"))
    v = bits_to_f32(g[f"s{seed}_vn_bf16"], cuda_device).requires_grad_(True)
    t = bits_to_f32(g[f"s{seed}_tn_bf16"], cuda_device).requires_grad_(True)
    loss = pg.ContrastiveLoss(temperature=tau)(v, t)
    k = f"s{seed}_trainer_tau{tau}"
    assert loss_close(loss.item(), float(g[k + "_loss"]), 1.0 / tau), (loss.item(), float(g[k + "_loss"]))
    loss.backward()
    if k + "_dv" in g:
        assert rel(v.grad, g[k + "_dv"]) < GRAD_RTOL
        assert rel(t.grad, g[k + "_dt"]) < GRAD_RTOL


def test_ntxent_bf16_inputs_and_no_grad(pg, cuda_device):
    torch.manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(200, 512, device=cuda_device), dim=-1).to(torch.bfloat16)
    b = torch.nn.functional.normalize(torch.randn(200, 512, device=cuda_device), dim=-1).to(torch.bfloat16)
    ref = cf.ntxent(a.double().cpu().numpy(), b.double().cpu().numpy(), 0.2)
    with torch.no_grad():
        l0 = pg.ContrastiveLoss(0.2)(a, b)
    assert l0.item() == pytest.approx(ref["loss"], rel=LOSS_RTOL)
    a.requires_grad_(True)
    b.requires_grad_(True)
    loss = pg.ContrastiveLoss(0.2)(a, b)
    loss.backward()
    assert a.grad.dtype == torch.bfloat16
    assert rel(a.grad.float(), ref["dx"]) < GRAD_RTOL and rel(b.grad.float(), ref["dy"]) < GRAD_RTOL


def test_ntxent_known_answers(pg, cuda_device):
    from preference_guided_image_captioning_alignment_b200 import components
    dev = cuda_device
    B, D = 48, 64
    x = torch.randn(1, D, device=dev).repeat(B, 1)
    assert components.ContrastiveLoss(0.5)(x, x).item() == pytest.approx(math.log(B), rel=LOSS_RTOL)      # KA3
    e = torch.eye(B, D, device=dev)
    for tau in (0.5, 0.2):                                                                                # KA4
        assert pg.ContrastiveLoss(tau)(e, e).item() == pytest.approx(math.log(1 + (B - 1) * math.exp(-1 / tau)),
                                                                     rel=LOSS_RTOL)
    u = torch.nn.functional.normalize(torch.randn(4, 256, device=dev), dim=-1)                            # KA5
    lo, hi = pg.ContrastiveLoss(0.01)(u, u).item(), pg.ContrastiveLoss(1.0)(u, u).item()
    assert lo < hi and lo >= 0.0
    v, t = torch.randn(B, D, device=dev), torch.randn(B, D, device=dev)                                   # KA6
    assert components.ContrastiveLoss(0.07)(v, t).item() == components.ContrastiveLoss(0.1)(v, t).item()


def test_reference_test_suite_properties(pg, cuda_device):
    """The reference's own TestLossFunctions (tests/test_model.py:383-499) re-run against the fused modules."""
    dev = cuda_device
    ie = torch.nn.functional.normalize(torch.randn(4, 256, device=dev), dim=-1)
    te = torch.nn.functional.normalize(torch.randn(4, 256, device=dev), dim=-1)
    loss = pg.ContrastiveLoss(temperature=0.07)(ie, te)
    assert isinstance(loss, torch.Tensor) and loss.dim() == 0 and loss.item() >= 0.0
    lf = pg.PreferenceLoss(beta=0.1)
    pl, rl = torch.randn(2, 20, 1000, device=dev), torch.randn(2, 20, 1000, device=dev)
    py, ry = torch.randint(0, 1000, (2, 20), device=dev), torch.randint(0, 1000, (2, 20), device=dev)
    m = torch.ones(2, 20, device=dev)
    loss = lf(preferred_logits=pl, rejected_logits=rl, preferred_labels=py, rejected_labels=ry, preferred_mask=m,
              rejected_mask=m)
    assert loss.dim() == 0 and loss.item() >= 0.0
    lp = lf._compute_log_probs(torch.randn(2, 10, 100, device=dev), torch.randint(0, 100, (2, 10), device=dev),
                               torch.ones(2, 10, device=dev))
    assert lp.shape == (2,)
    a = torch.randn(2, 256, device=dev, requires_grad=True)
    b = torch.randn(2, 256, device=dev, requires_grad=True)
    pg.ContrastiveLoss()(torch.nn.functional.normalize(a, dim=-1), torch.nn.functional.normalize(b, dim=-1)).backward()
    assert a.grad is not None and b.grad is not None
    pl = torch.randn(2, 10, 100, device=dev, requires_grad=True)
    rl = torch.randn(2, 10, 100, device=dev, requires_grad=True)
    y = torch.randint(0, 100, (2, 10), device=dev)
    m = torch.ones(2, 10, device=dev)
    pg.PreferenceLoss()(pl, rl, y, y, m, m).backward()
    assert pl.grad is not None and rl.grad is not None


def test_ntxent_rank_emulation(pg, cuda_device):
    """All ranks' shards through the same kernels on one device == the global batch."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    W, nb, D, tau = 4, 96, 512, 0.5
    B = W * nb
    torch.manual_seed(1)
    a = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1).to(torch.bfloat16)
    b = torch.nn.functional.normalize(a.float() + 0.3 * torch.randn(B, D, device=dev), dim=-1).to(torch.bfloat16)
    ref = cf.ntxent(a.double().cpu().numpy(), b.double().cpu().numpy(), tau)
    fw = [F.ntxent_fwd(a[r * nb:(r + 1) * nb].contiguous(), b, 1 / tau, r * nb) for r in range(W)]
    lse_col = F.lse_combine(torch.stack([f[2] for f in fw]))
    assert rel(lse_col, ref["lse_col"]) < 1e-5
    loss = sum(F.ntxent_loss(f[0], f[1], lse_col[r * nb:(r + 1) * nb].contiguous(), 1.0 / B).item()
               for r, f in enumerate(fw))
    assert loss == pytest.approx(ref["loss"], rel=LOSS_RTOL)
    one = torch.ones((), device=dev)
    db = torch.zeros(B, D, device=dev)
    for r, f in enumerate(fw):
        da, dbp = F.ntxent_bwd(a[r * nb:(r + 1) * nb].contiguous(), b, 1 / tau, r * nb, f[0], lse_col, one, 1 / (2 * B))
        assert rel(da, ref["dx"][r * nb:(r + 1) * nb]) < GRAD_RTOL
        db += dbp
    assert rel(db, ref["dy"]) < GRAD_RTOL


# ================================================================================================ logits path
@pytest.mark.parametrize("name", ["none", "i64", "f64"])
def test_sequence_logprobs_golden(pg, cuda_device, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "seq_logprobs.npz"))
    dev = cuda_device
    logits = torch.tensor(g["logits"], dtype=torch.float32, device=dev, requires_grad=True)
    labels = torch.tensor(g["labels"], device=dev)
    mask = {"none": None, "i64": torch.tensor(g["mask_i"], device=dev),
            "f64": torch.tensor(g["mask_f"], dtype=torch.float32, device=dev)}[name]
    s = pg.compute_sequence_logprobs(logits, labels, mask)
    np.testing.assert_allclose(s.detach().cpu().numpy(), g[f"sum_{name}"], rtol=2e-6)
    (s * torch.arange(1, 4, device=dev).float()).sum().backward()
    assert rel(logits.grad, g[f"sum_{name}_dlogits"]) < 1e-5
    if name != "none":
        lg2 = torch.tensor(g["logits"], dtype=torch.float32, device=dev, requires_grad=True)
        s2 = pg.PreferenceLoss(0.1)._compute_log_probs(lg2, labels, mask)
        np.testing.assert_allclose(s2.detach().cpu().numpy(), g[f"mean_{name}"], rtol=2e-6)
        (s2 * torch.arange(1, 4, device=dev).float()).sum().backward()
        assert rel(lg2.grad, g[f"mean_{name}_dlogits"]) < 1e-5


def test_preference_loss_golden(pg, cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "seq_logprobs.npz"))
    dev = cuda_device
    a = torch.tensor(g["logits"], dtype=torch.float32, device=dev, requires_grad=True)
    b = torch.tensor(g["logits2"], dtype=torch.float32, device=dev, requires_grad=True)
    t = lambda k: torch.tensor(g[k], device=dev)
    loss = pg.PreferenceLoss(beta=0.1)(a, b, t("labels"), t("labels2"), t("mask_i"), t("mask2"))
    assert loss.item() == pytest.approx(float(g["pref_loss"]), rel=1e-5)
    loss.backward()
    assert rel(a.grad, g["pref_dlogits"]) < 1e-4 and rel(b.grad, g["pref_dlogits2"]) < 1e-4


@pytest.mark.parametrize("tag,kw,use_ref", [("std", dict(beta=0.1), True),
                                            ("ls", dict(beta=0.1, label_smoothing=0.1), True),
                                            ("free", dict(beta=0.1, reference_free=True), True),
                                            ("noref", dict(beta=0.25), False)])
def test_dpo_loss_golden(pg, cuda_device, golden_dir, tag, kw, use_ref):
    g = np.load(os.path.join(golden_dir, "dpo_loss.npz"))
    xs = [torch.tensor(g[k], dtype=torch.float32, device=cuda_device, requires_grad=True)
          for k in ("pc", "pr", "rc", "rr")]
    mod = pg.DPOPreferenceLoss(**kw)
    loss, metrics = mod(*xs) if use_ref else mod(xs[0], xs[1])
    assert loss.item() == pytest.approx(float(g[tag + "_loss"]), rel=1e-5)
    assert tuple(metrics.keys()) == METRIC_KEYS
    np.testing.assert_allclose([metrics[k] for k in METRIC_KEYS], g[tag + "_metrics"], rtol=1e-5, atol=1e-6)
    loss.backward()
    assert rel(xs[0].grad, g[tag + "_dpc"]) < 1e-5 and rel(xs[1].grad, g[tag + "_dpr"]) < 1e-5
    if use_ref and not kw.get("reference_free"):
        assert rel(xs[2].grad, g[tag + "_drc"]) < 1e-5 and rel(xs[3].grad, g[tag + "_drr"]) < 1e-5
    else:
        assert xs[2].grad is None


# ================================================================================================ fused LM head
def test_dpo_head_golden(pg, cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "dpo_head.npz"))
    dev = cuda_device
    f = lambda k, rg=False: torch.tensor(g[k], dtype=torch.float32, device=dev, requires_grad=rg)
    i = lambda k: torch.tensor(g[k], device=dev)
    W, hc, hr = f("W", True), f("hc", True), f("hr", True)
    head = pg.FusedDPOHead(beta=0.1)
    loss, metrics = head(hc, hr, W, i("yc"), i("yr"), i("mc"), i("mr"), f("rhc"), f("rhr"), f("Wr"))
    assert loss.item() == pytest.approx(float(g["loss"]), rel=LOSS_RTOL)
    np.testing.assert_allclose(metrics.cpu().numpy(), g["metrics"], rtol=1e-4, atol=1e-4)
    assert not metrics.requires_grad
    loss.backward()
    assert rel(W.grad, g["dW"]) < GRAD_RTOL
    assert rel(hc.grad, g["dhc"]) < GRAD_RTOL and rel(hr.grad, g["dhr"]) < GRAD_RTOL
    # per-sequence log-probs, sum and mean flavours
    for ln in (False, True):
        got = pg.lmhead_sequence_logprobs(hc.detach(), W.detach(), i("yc"), i("mc"), ln)
        want = cf.lmhead_sequence_logprobs(g["hc"], g["W"], g["yc"], g["mc"], ln)["seq_logp"]
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5)


def _cfg2_inputs(dev, B, T=128, d=1024, V=50257, seed=1234):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    W = (torch.randn(V, d, generator=gen) * 0.02).to(torch.bfloat16)
    h = torch.randn(2 * B, T, d, generator=gen).to(torch.bfloat16)
    y = torch.randint(0, V, (2 * B, T), generator=gen)
    lens = torch.randint(T // 2, T + 1, (2 * B,), generator=gen)
    m = (torch.arange(T)[None, :] < lens[:, None]).long()
    return W.to(dev), h.to(dev), y.to(dev), m.to(dev)


def test_lmhead_full_vocab_vs_oracle(pg, cuda_device):
    """GPT-2 Medium head shape (d=1024, V=50257, T=128) on 2 pairs: everything against the float64 oracle."""
    dev = cuda_device
    B = 2
    W, h, y, m = _cfg2_inputs(dev, B)
    Wg, hg = W.clone().requires_grad_(True), h.clone().requires_grad_(True)
    seq = pg.lmhead_sequence_logprobs(hg, Wg, y, m)
    gseq = torch.tensor([0.3, -1.1, 0.7, 2.0], device=dev)
    (seq * gseq).sum().backward()
    o = cf.lmhead_sequence_logprobs(h.double().cpu().numpy(), W.double().cpu().numpy(), y.cpu().numpy(),
                                    m.cpu().numpy(), False, grad_seq=gseq.double().cpu().numpy())
    np.testing.assert_allclose(seq.detach().cpu().numpy(), o["seq_logp"], rtol=LOSS_RTOL)
    assert rel(hg.grad.float(), o["dhidden"]) < GRAD_RTOL
    assert rel(Wg.grad.float(), o["dweight"]) < GRAD_RTOL
    assert torch.count_nonzero(hg.grad[:, -1]).item() == 0  # nothing is scored at the last position


def test_plumbing_bit_exact(pg, cuda_device):
    """KA10: the index / mask plumbing equals labels[:, 1:], mask[:, 1:] exactly; ids >= V at masked slots are inert."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    nseq, T, V = 5, 17, 50257
    gen = torch.Generator().manual_seed(3)
    labels = torch.randint(0, V, (nseq, T), generator=gen)
    lens = torch.tensor([17, 9, 1, 12, 2])
    mask = (torch.arange(T)[None, :] < lens[:, None]).long()
    labels[mask == 0] = 50257  # pad id, outside the V=50257 vocabulary
    for mk in (mask, mask.float(), mask.bool(), mask.int()):
        rl, rw = F.prep_rows(labels.to(dev), mk.to(dev), V)
        rl, rw = rl.view(nseq, T).cpu(), rw.view(nseq, T).cpu()
        want_l = torch.where(mask[:, 1:] == 1, labels[:, 1:], torch.full_like(labels[:, 1:], -1))
        assert torch.equal(rl[:, :-1].long(), want_l)
        assert torch.equal(rw[:, :-1], mask[:, 1:].float())
        assert (rl[:, -1] == -1).all() and (rw[:, -1] == 0).all()
    rl, rw = F.prep_rows(labels.clamp(max=V - 1).to(dev), None, V)
    assert torch.equal(rw.view(nseq, T)[:, :-1].cpu(), torch.ones(nseq, T - 1))


def test_lmhead_edge_cases(pg, cuda_device):
    dev = cuda_device
    torch.manual_seed(5)
    # ragged shapes: rows not a multiple of 128, V smaller than a tile, d not a multiple of 64
    for (nseq, T, d, V) in [(1, 2, 64, 50), (3, 7, 72, 300), (2, 130, 256, 1000)]:
        h = torch.randn(nseq, T, d, device=dev).to(torch.bfloat16).requires_grad_(True)
        W = (torch.randn(V, d, device=dev) * 0.3).to(torch.bfloat16).requires_grad_(True)
        y = torch.randint(0, V, (nseq, T), device=dev)
        seq = pg.lmhead_sequence_logprobs(h, W, y, None, True)
        seq.sum().backward()
        o = cf.lmhead_sequence_logprobs(h.detach().double().cpu().numpy(), W.detach().double().cpu().numpy(),
                                        y.cpu().numpy(), None, True, grad_seq=np.ones(nseq))
        np.testing.assert_allclose(seq.detach().cpu().numpy(), o["seq_logp"], rtol=LOSS_RTOL)
        assert rel(h.grad.float(), o["dhidden"]) < GRAD_RTOL and rel(W.grad.float(), o["dweight"]) < GRAD_RTOL
    # all-zero mask row: 0 in sum mode, NaN in length-normalised mode (like the reference)
    h = torch.randn(2, 6, 64, device=dev).to(torch.bfloat16)
    W = torch.randn(40, 64, device=dev).to(torch.bfloat16)
    y = torch.randint(0, 40, (2, 6), device=dev)
    m = torch.ones(2, 6, dtype=torch.long, device=dev)
    m[1] = 0
    assert pg.lmhead_sequence_logprobs(h, W, y, m, False)[1].item() == 0.0
    assert math.isnan(pg.lmhead_sequence_logprobs(h, W, y, m, True)[1].item())
    # NaN in -> NaN out (the trainer's NaN-skip logic relies on it, trainer.py:482,607)
    hn = h.clone()
    hn[0, 2, 5] = float("nan")
    assert math.isnan(pg.lmhead_sequence_logprobs(hn, W, y, None, False)[0].item())
    # CPU tensors: loud failure, no fallback
    with pytest.raises(Exception):
        pg.lmhead_sequence_logprobs(h.cpu(), W.cpu(), y.cpu())


def test_cfg2_properties_full_size(pg, cuda_device):
    """BASELINE config 2 (B=16, T=128, d=1024, V=50257): size-independent properties."""
    dev = cuda_device
    B = 16
    W, h, y, m = _cfg2_inputs(dev, B)
    head = pg.FusedDPOHead(beta=0.1)
    # KA1: policy == reference -> ln 2, margin 0, accuracy 0
    loss, met = head(h[:B], h[B:], W, y[:B], y[B:], m[:B], m[B:], h[:B], h[B:], W)
    assert loss.item() == pytest.approx(math.log(2), rel=1e-6)
    assert met[1].item() == 0.0 and met[2].item() == 0.0
    # KA2: W = 0 -> every token log-prob is -ln V
    seq = pg.lmhead_sequence_logprobs(h, torch.zeros_like(W), y, m, True)
    np.testing.assert_allclose(seq.cpu().numpy(), -math.log(50257), rtol=1e-6)
    # KA9: all-ones mask == no mask, bit for bit
    ones = torch.ones_like(m)
    assert torch.equal(pg.lmhead_sequence_logprobs(h, W, y, ones), pg.lmhead_sequence_logprobs(h, W, y, None))
    # determinism and linearity of the backward in the upstream gradient
    Wg = W.clone().requires_grad_(True)
    hg = h.clone().requires_grad_(True)
    seq = pg.lmhead_sequence_logprobs(hg, Wg, y, m)
    g1 = torch.autograd.grad(seq.sum(), (hg, Wg), retain_graph=True)
    g1b = torch.autograd.grad(seq.sum(), (hg, Wg), retain_graph=True)
    g2 = torch.autograd.grad((2.0 * seq).sum(), (hg, Wg))
    assert torch.equal(g1[0], g1b[0]) and torch.equal(g1[1], g1b[1])
    assert rel(g2[1].float(), 2.0 * g1[1].float()) < 1e-2
    # masked rows receive exactly zero gradient
    assert torch.count_nonzero(g1[0][(m[:, 1:] == 0).nonzero(as_tuple=True)]).item() == 0
    # sequence sums against the oracle on a slice of the batch
    sl = [0, B, 2 * B - 1]
    o = cf.lmhead_sequence_logprobs(h[sl].double().cpu().numpy(), W.double().cpu().numpy(), y[sl].cpu().numpy(),
                                    m[sl].cpu().numpy())
    np.testing.assert_allclose(seq.detach()[sl].cpu().numpy(), o["seq_logp"], rtol=LOSS_RTOL)


def test_ntxent_cfg1_and_large(pg, cuda_device):
    """cfg1 (B=64, D=512, tau=0.5) against the oracle; a 4096-row slice of cfg3 via symmetry properties."""
    dev = cuda_device
    gen = torch.Generator().manual_seed(1234)
    v = torch.randn(64, 512, generator=gen).to(torch.bfloat16).float()
    t = torch.randn(64, 512, generator=gen).to(torch.bfloat16).float()
    from preference_guided_image_captioning_alignment_b200 import components
    vg, tg = v.to(dev).requires_grad_(True), t.to(dev).requires_grad_(True)
    loss = components.ContrastiveLoss(0.5)(vg, tg)
    loss.backward()
    o = cf.ntxent(v.double().numpy(), t.double().numpy(), 0.5, True, True)
    assert loss.item() == pytest.approx(o["loss"], rel=LOSS_RTOL)
    assert rel(vg.grad, o["dx"]) < GRAD_RTOL and rel(tg.grad, o["dy"]) < GRAD_RTOL
    # symmetry: swapping the two sides swaps the gradients and keeps the loss
    a = torch.nn.functional.normalize(torch.randn(4096, 512, device=dev), dim=-1).requires_grad_(True)
    b = torch.nn.functional.normalize(torch.randn(4096, 512, device=dev), dim=-1).requires_grad_(True)
    l1 = pg.ContrastiveLoss(0.5)(a, b)
    ga, gb = torch.autograd.grad(l1, (a, b))
    l2 = pg.ContrastiveLoss(0.5)(b, a)
    hb, ha = torch.autograd.grad(l2, (b, a))
    assert l1.item() == pytest.approx(l2.item(), rel=1e-6)
    assert rel(ga, ha) < 1e-3 and rel(gb, hb) < 1e-3
    o = cf.ntxent(a.detach().to(torch.bfloat16).double().cpu().numpy(),
                  b.detach().to(torch.bfloat16).double().cpu().numpy(), 0.5)
    assert l1.item() == pytest.approx(o["loss"], rel=LOSS_RTOL)
    assert rel(ga, o["dx"]) < GRAD_RTOL


def test_similarity_matrix(pg, cuda_device):
    dev = cuda_device
    v, t = torch.randn(70, 96, device=dev), torch.randn(50, 96, device=dev)
    sim = pg.TemperatureScaledSimilarity(temperature=0.07)(v, t)  # clamped to 0.1
    vn = torch.nn.functional.normalize(v, dim=-1).to(torch.bfloat16).double()
    tn = torch.nn.functional.normalize(t, dim=-1).to(torch.bfloat16).double()
    assert sim.shape == (70, 50)
    assert rel(sim, vn @ tn.T / 0.1) < 1e-5
    assert "temperature" in pg.TemperatureScaledSimilarity().state_dict()


def test_similarity_matrix_is_differentiable(pg, cuda_device):
    """a1: the dense similarity module is differentiable w.r.t. both inputs and a learnable temperature, like the
    reference's (components.py:45-83); checked against fp64 autograd of the same formula with a random upstream
    gradient."""
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    v0, t0 = torch.randn(70, 96, generator=g), torch.randn(50, 96, generator=g)
    up = torch.randn(70, 50, generator=g)
    mod = pg.TemperatureScaledSimilarity(temperature=0.5, learnable=True).to(dev)
    v, t = v0.to(dev).requires_grad_(True), t0.to(dev).requires_grad_(True)
    (mod(v, t) * up.to(dev)).sum().backward()
    vd, td = v0.double().requires_grad_(True), t0.double().requires_grad_(True)
    tau = torch.tensor(0.5, dtype=torch.float64, requires_grad=True)
    S = torch.nn.functional.normalize(vd, dim=-1) @ torch.nn.functional.normalize(td, dim=-1).T / tau.clamp(0.1, 2.0)
    (S * up.double()).sum().backward()
    assert rel(v.grad, vd.grad) < GRAD_RTOL and rel(t.grad, td.grad) < GRAD_RTOL
    assert abs(mod.temperature.grad.item() - tau.grad.item()) <= GRAD_RTOL * abs(tau.grad.item())
    # outside the clamp range the temperature gets no gradient (torch.clamp semantics)
    mod2 = pg.TemperatureScaledSimilarity(temperature=0.05, learnable=True).to(dev)
    mod2(v.detach(), t.detach()).sum().backward()
    assert mod2.temperature.grad.item() == 0.0


def test_opcheck(pg, cuda_device):
    from preference_guided_image_captioning_alignment_b200 import ops
    dev = cuda_device
    a = torch.randn(32, 64, device=dev, requires_grad=True)
    b = torch.randn(32, 64, device=dev, requires_grad=True)
    torch.library.opcheck(ops.ntxent, (a, b, 2.0, True), test_utils=("test_schema", "test_faketensor"))
    h = torch.randn(2, 5, 64, device=dev, requires_grad=True)
    W = torch.randn(30, 64, device=dev, requires_grad=True)
    y = torch.randint(0, 30, (2, 5), device=dev)
    torch.library.opcheck(ops.lmhead_seq_logprob, (h, W, y, None, False), test_utils=("test_schema", "test_faketensor"))


# ================================================================================================ backward core
def _sgg_reference(x, y, scale, row, col):
    """fp32 torch restatement of the softmax-gradient GEMM (include/pgica.h), G rounded to bf16 like the kernel."""
    xf, yf = x.float(), y.float()
    z = (xf @ yf.t()) * scale
    g = torch.zeros_like(z)
    rows = torch.arange(x.shape[0], device=x.device)
    if row is not None:
        lse, coef, tgt = row
        oh = torch.zeros_like(z)
        v = tgt >= 0
        oh[rows[v], tgt[v].long()] = 1
        g += coef[:, None] * (torch.exp(z - lse[:, None]) - oh)
    if col is not None:
        lse, coef, tgt = col
        oh = (tgt[None, :].long() == rows[:, None]).float()
        g += coef[None, :] * (torch.exp(z - lse[None, :]) - oh)
    return g @ yf, g.to(torch.bfloat16).float() @ yf


@pytest.mark.parametrize("mx,my,k,mode", [
    (128, 128, 256, "row"),        # single-CTA kernel (k = 256)
    (77, 90, 64, "col"),           # k < 256, ragged rows and columns
    (300, 1000, 512, "both"),      # 2-CTA clusters, last 256-wide tile half empty
    (200, 333, 1024, "row"),       # 4-CTA clusters, fewer 256-wide tiles than CTAs
    (128 * 37 + 5, 700, 1024, "col"),   # more row blocks than resident clusters: persistent loop over items
    (520, 128 * 9 + 1, 1024, "both"),   # odd number of G tiles, one-column tail
    (4500, 260, 512, "row"),       # 2-CTA clusters, one round per item, many items
])
def test_softmax_grad_gemm_shapes(pg, cuda_device, mx, my, k, mode):
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    torch.manual_seed(mx * 7 + my)
    x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
    y = (torch.randn(my, k, device=dev) * 0.2).to(torch.bfloat16)
    row = col = None
    if mode in ("row", "both"):
        lse_r, _ = F.gemm_lse(x, y, 1.0)
        coef_r = torch.randn(mx, device=dev)
        coef_r[::5] = 0
        tgt_r = torch.randint(0, my, (mx,), device=dev, dtype=torch.int32)
        tgt_r[::3] = -1
        row = (lse_r, coef_r, tgt_r)
    if mode in ("col", "both"):
        lse_c, _ = F.gemm_lse(y, x, 1.0)
        coef_c = torch.randn(my, device=dev)
        coef_c[::5] = 0
        tgt_c = torch.randint(0, mx, (my,), device=dev, dtype=torch.int32)
        tgt_c[::3] = -1
        col = (lse_c, coef_c, tgt_c)
    out = F.softmax_grad_gemm(x, y, 1.0, row=row, col=col)
    out2 = F.softmax_grad_gemm(x, y, 1.0, row=row, col=col)
    assert torch.equal(out, out2)  # deterministic: no atomics anywhere
    exact, emul = _sgg_reference(x, y, 1.0, row, col)
    scale = exact.abs().max().item()
    assert (out - emul).abs().max().item() < 2e-3 * scale + 1e-5    # same arithmetic as the kernel (bf16 G)
    assert rel(out, exact) < GRAD_RTOL                               # and within the gradient tolerance of fp32
    ob = F.softmax_grad_gemm(x, y, 1.0, row=row, col=col, out_dtype=torch.bfloat16)
    assert rel(ob.float(), exact) < GRAD_RTOL


def test_softmax_grad_gemm_single_cta_variant_matches(pg, cuda_device):
    """The single-CTA kernel (option sgg_cluster = 1; also the path for k not a multiple of 512) and the default
    cluster kernel with its L2 exchange ring agree."""
    from preference_guided_image_captioning_alignment_b200 import _lib
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    torch.manual_seed(9)
    x = (torch.randn(640, 1024, device=dev) * 0.5).to(torch.bfloat16)
    y = (torch.randn(1500, 1024, device=dev) * 0.2).to(torch.bfloat16)
    lse, _ = F.gemm_lse(x, y, 1.0)
    row = (lse, torch.randn(640, device=dev), torch.randint(0, 1500, (640,), device=dev, dtype=torch.int32))
    a = F.softmax_grad_gemm(x, y, 1.0, row=row)
    _lib.set_option("sgg_cluster", 1)
    b = F.softmax_grad_gemm(x, y, 1.0, row=row)
    _lib.set_option("sgg_cluster", 0)
    assert rel(a, b) < 1e-5


def test_cfg4_slice_properties(pg, cuda_device):
    """One rank's slice of BASELINE config 4 (T = 512, 8 pairs -> 8176 scored rows, 64 row blocks > resident
    clusters): size-independent properties through the persistent backward."""
    dev = cuda_device
    B, T = 4, 512
    W, h, y, m = _cfg2_inputs(dev, B, T=T)
    head = pg.FusedDPOHead(beta=0.1)
    loss, met = head(h[:B], h[B:], W, y[:B], y[B:], m[:B], m[B:], h[:B], h[B:], W)          # KA1
    assert loss.item() == pytest.approx(math.log(2), rel=1e-6)
    Wg, hg = W.clone().requires_grad_(True), h.clone().requires_grad_(True)
    seq = pg.lmhead_sequence_logprobs(hg, Wg, y, m)
    g1 = torch.autograd.grad(seq.sum(), (hg, Wg), retain_graph=True)
    g2 = torch.autograd.grad(seq.sum(), (hg, Wg))
    assert torch.equal(g1[0], g2[0]) and torch.equal(g1[1], g2[1])
    assert torch.count_nonzero(g1[0][(m[:, 1:] == 0).nonzero(as_tuple=True)]).item() == 0
    # gradient of sum_b seq_logp[b] w.r.t. hidden on one sequence against the oracle
    o = cf.lmhead_sequence_logprobs(h[:1].double().cpu().numpy(), W.double().cpu().numpy(), y[:1].cpu().numpy(),
                                    m[:1].cpu().numpy(), False, grad_seq=np.ones(1))
    np.testing.assert_allclose(seq.detach()[:1].cpu().numpy(), o["seq_logp"], rtol=LOSS_RTOL)
    assert rel(g1[0][:1].float(), o["dhidden"]) < GRAD_RTOL


# ================================================================================================ dual backward
def _dual_inputs(dev, mx, my, k, mode):
    from preference_guided_image_captioning_alignment_b200 import functional as F
    torch.manual_seed(mx * 11 + my)
    x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
    y = (torch.randn(my, k, device=dev) * 0.2).to(torch.bfloat16)
    row = col = None
    if mode in ("row", "both"):
        lse_r, _ = F.gemm_lse(x, y, 1.0)
        coef_r = torch.randn(mx, device=dev)
        coef_r[::5] = 0
        tgt_r = torch.randint(0, my, (mx,), device=dev, dtype=torch.int32)
        tgt_r[::3] = -1
        row = (lse_r, coef_r, tgt_r)
    if mode in ("col", "both"):
        lse_c, _ = F.gemm_lse(y, x, 1.0)
        coef_c = torch.randn(my, device=dev)
        coef_c[::5] = 0
        tgt_c = torch.randint(0, mx, (my,), device=dev, dtype=torch.int32)
        tgt_c[::3] = -1
        col = (lse_c, coef_c, tgt_c)
    return x, y, row, col


@pytest.mark.parametrize("mx,my,k,mode,plan", [
    (128, 128, 512, "row", None),          # one tile, one split
    (300, 1000, 512, "both", None),        # ragged rows and columns, odd tile count
    (200, 333, 1024, "row", None),         # two 512-column splits
    (128 * 5 + 7, 128 * 9 + 1, 1024, "row", "2,4"),   # 3 chunks of 2 row blocks (OutY accumulated), 3 passes, 1-col tail
    (128 * 7, 128 * 6, 512, "both", "3,2"),           # chunks of 3/3/1, passes of 2: Rc < column pairs never, odd rows
    (128 * 2, 128 * 11, 1024, "col", "4,6"),          # fewer row blocks than the chunk; column term only
    (2048, 5003, 1024, "row", None),       # the planner's own split on a mid-size head
])
def test_softmax_grad_gemm_dual(pg, cuda_device, monkeypatch, mx, my, k, mode, plan):
    """Both backward products from one recomputation (sgg_f.cu) against fp32 torch, the bf16-G emulation and the
    two single-product launches; deterministic."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    x, y, row, col = _dual_inputs(cuda_device, mx, my, k, mode)
    set_plan(plan)
    ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
    ox2, oy2 = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
    assert torch.equal(ox, ox2) and torch.equal(oy, oy2)
    exact_x, emul_x = _sgg_reference(x, y, 1.0, row, col)
    exact_y, emul_y = _sgg_reference(y, x, 1.0, col, row)
    for out, exact, emul in ((ox, exact_x, emul_x), (oy, exact_y, emul_y)):
        scale = exact.abs().max().item()
        assert (out - emul).abs().max().item() < 2e-3 * scale + 1e-5
        assert rel(out, exact) < GRAD_RTOL
    sx = F.softmax_grad_gemm(x, y, 1.0, row=row, col=col)
    sy = F.softmax_grad_gemm(y, x, 1.0, row=col, col=row)
    assert rel(ox, sx) < 1e-3 and rel(oy, sy) < 1e-3
    bx, _ = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col, out_x_dtype=torch.bfloat16)
    assert rel(bx.float(), exact_x) < GRAD_RTOL
    if plan is None:  # one chunk: OutY is written once and may be bf16 as well
        bx, by = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col, out_x_dtype=torch.bfloat16,
                                          out_y_dtype=torch.bfloat16)
        assert rel(bx.float(), exact_x) < GRAD_RTOL and rel(by.float(), exact_y) < GRAD_RTOL


def test_lmhead_backward_dual_matches_split(pg, cuda_device, monkeypatch):
    """pgica_lmhead_logprob_bwd: the one-launch dual backward and the two-launch backward give the same gradients."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    B, T, d, V = 6, 64, 1024, 5003
    g = torch.Generator().manual_seed(5)
    W = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    H = torch.randn(B, T, d, generator=g).to(torch.bfloat16).to(dev)
    y = torch.randint(0, V, (B, T), generator=g).to(dev)
    m = torch.ones(B, T, dtype=torch.long, device=dev)
    m[:, T - 9:] = 0
    seq, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(H, W, y, m, False)
    gseq = torch.randn(B, device=dev)
    from preference_guided_image_captioning_alignment_b200 import _lib
    _lib.set_option("sgg_fused", 1)
    dh1, dw1 = F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
    _lib.set_option("sgg_fused", 0)
    dh0, dw0 = F.lmhead_logprob_bwd(H, W, rl, rw, lse, gseq, False)
    _lib.set_option("sgg_fused", 1)
    assert rel(dh1.float(), dh0.float()) < 2e-3 and rel(dw1, dw0) < 2e-3
    assert torch.count_nonzero(dh1[:, T - 10:]).item() == 0  # rows that score nothing get exactly zero


def test_graphed_dpo_step_matches_eager(pg, cuda_device):
    """GraphedDPOStep (CUDA graph of FusedDPOHead.forward_stacked + backward) replays to the eager result, also after
    the static inputs are overwritten."""
    dev = cuda_device
    B, T, d, V = 3, 32, 512, 3001
    g = torch.Generator().manual_seed(11)
    W = (torch.randn(V, d, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    Wr = (torch.randn(V, d, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    hs = [torch.randn(2 * B, T, d, generator=g).to(torch.bfloat16).to(dev) for _ in range(4)]
    ys = [torch.randint(0, V, (2 * B, T), generator=g).to(dev) for _ in range(2)]
    m = torch.ones(2 * B, T, dtype=torch.long, device=dev)
    head = pg.FusedDPOHead(beta=0.1)
    step = pg.GraphedDPOStep(head, W.clone().requires_grad_(True), Wr, hs[0], ys[0], m, hs[1])
    for h, hr, y in ((hs[0], hs[1], ys[0]), (hs[2], hs[3], ys[1])):
        step.copy_inputs(h, y, m, hr)
        step.launch()
        value = step.loss_value()
        loss = step.loss
        Wg, hg = W.clone().requires_grad_(True), h.clone().requires_grad_(True)
        ref_loss, _ = head.forward_stacked(hg, Wg, y, m, hr, Wr)
        ref_loss.backward()
        assert loss.item() == pytest.approx(ref_loss.item(), rel=1e-6) and value == loss.item()
        assert torch.equal(step.dweight, Wg.grad) and torch.equal(step.dhidden, hg.grad)


def test_causal_lm_loss_from_hidden_states(pg, cuda_device):
    """the HF causal-LM cross-entropy that GPT2LMHeadModel computes from `labels`
    (transformers loss_utils.py:45-67, reached from pkg/models/model.py:604-610) out of the fused LM-head kernel:
    mean over the shifted positions whose label is not -100, and its gradients, against torch cross_entropy on the
    materialised logits (fp64, same bf16-rounded inputs)."""
    from preference_guided_image_captioning_alignment_b200.install import lazy_causal_lm_loss
    dev = cuda_device
    B, T, d, V = 3, 40, 256, 5003
    g = torch.Generator().manual_seed(21)
    W = (torch.randn(V, d, generator=g) * 0.05).to(torch.bfloat16)
    H = torch.randn(B, T, d, generator=g).to(torch.bfloat16)
    y = torch.randint(0, V, (B, T), generator=g)
    y[0, 25:] = -100   # padded tail, HF convention
    y[2, :5] = -100    # ignored prefix
    Wg, Hg = W.to(dev).requires_grad_(True), H.to(dev).requires_grad_(True)
    loss = lazy_causal_lm_loss(pg.LazyLogits(Hg, Wg), y.to(dev))
    loss.backward()
    Wd, Hd = W.double().requires_grad_(True), H.double().requires_grad_(True)
    logits = Hd @ Wd.t()
    ref = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, V), y[:, 1:].reshape(-1), ignore_index=-100)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= LOSS_RTOL * abs(ref.item())
    assert rel(Hg.grad.float(), Hd.grad) < GRAD_RTOL and rel(Wg.grad.float(), Wd.grad) < GRAD_RTOL
    assert torch.count_nonzero(Hg.grad[0, 25:]).item() == 0  # positions that score nothing get exactly zero


# ================================================================================ SURVEY 8(f) row 2: grad norm + clip
@pytest.mark.parametrize("tag", ["clip", "noclip", "big", "nan"])
def test_grad_norm_clip_golden(pg, cuda_device, golden_dir, tag):
    """pg.NaNSafeGradientNorm (one multi-tensor pass) against the outputs of the real reference module."""
    g = np.load(os.path.join(golden_dir, "grad_clip.npz"))
    n = 2 if tag == "nan" else 5
    params = []
    for i in range(n):
        p = torch.nn.Parameter(torch.zeros(g[f"{tag}_g{i}"].shape, device=cuda_device))
        p.grad = torch.from_numpy(g[f"{tag}_g{i}"]).to(cuda_device)
        params.append(p)
    max_norm = 1.0 if tag == "nan" else float(g[f"{tag}_max_norm"])
    total, finite = pg.NaNSafeGradientNorm(max_norm=max_norm)(params)
    assert finite == bool(g[f"{tag}_finite"])
    if tag != "nan":
        assert abs(total.item() - float(g[f"{tag}_total"])) <= 2e-6 * float(g[f"{tag}_total"])
    for i, p in enumerate(params):
        np.testing.assert_allclose(p.grad.cpu().numpy(), g[f"{tag}_c{i}"], rtol=3e-6, atol=0, equal_nan=True)


def test_grad_norm_clip_large_mixed(pg, cuda_device):
    """Many chunks, bf16 and fp32 gradients, unaligned views, against the float64 oracle; deterministic; Inf detected."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    gen = torch.Generator().manual_seed(9)
    shapes = [(50257, 64), (3 * 32768 + 5,), (1023, 513), (7,), (32768,)]
    grads = [torch.randn(*s, generator=gen).to(dev) * 0.01 for s in shapes]
    grads[2] = grads[2].to(torch.bfloat16)
    base = torch.randn(100003, generator=gen).to(dev)
    grads.append(base[3:])  # 4-byte aligned only: the scalar path of the kernel
    ref = cf.grad_norm_clip([x.double().cpu().numpy() for x in grads], 0.25)
    work = [x.clone() for x in grads[:-1]] + [base.clone()[3:]]
    s1 = F.grad_norm_clip(work, 0.25)
    assert abs(s1[0].item() - ref["total_norm"]) <= 2e-6 * ref["total_norm"]
    assert abs(s1[1].item() - ref["clip_coef"]) <= 2e-6 * ref["clip_coef"] and s1[2].item() == 1.0
    for w, r in zip(work, ref["clipped"]):
        tol = 1e-2 if w.dtype == torch.bfloat16 else 1e-5
        assert rel(w.float(), r) < tol
    s2 = F.grad_norm_clip([x.clone() for x in grads[:-1]] + [base.clone()[3:]], 0.25)
    assert torch.equal(s1, s2)
    small = [x.clone() * 1e-4 for x in grads[:2]]
    keep = [x.clone() for x in small]
    s3 = F.grad_norm_clip(small, 1.0)  # below max_norm: untouched, bit for bit
    assert s3[1].item() == 1.0 and all(torch.equal(a, b) for a, b in zip(small, keep))
    bad = [x.clone() for x in grads[:2]]
    bad[1][12345] = float("inf")
    keep = [x.clone() for x in bad]
    s4 = F.grad_norm_clip(bad, 0.25)
    assert s4[2].item() == 0.0 and all(torch.equal(a, b) for a, b in zip(bad, keep))


def test_softmax_grad_gemm_dual_random_shapes(pg, cuda_device, monkeypatch):
    """Randomised shapes / terms / role splits (k = 512 ... 2048) of the dual kernel against the single-product
    launches; tools/dual_stress.py runs the longer version (profiles/r1_dual_stress_60cases.log)."""
    import random

    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    rng = random.Random(7)
    for i in range(14):
        k = rng.choice([512, 1024, 1536, 2048])
        S = k // 512
        mx = rng.choice([rng.randint(1, 300), rng.randint(300, 2000), 128 * rng.randint(1, 12) + 1])
        my = rng.choice([rng.randint(1, 300), rng.randint(300, 4000), 256 * rng.randint(1, 12) + 129])
        mode = rng.choice(["row", "col", "both"])
        set_plan(f"{rng.randint(1, 10 // S + 2)},{rng.randint(1, 10 // S + 2)}" if rng.random() < 0.6 else None)
        torch.manual_seed(i)
        x = (torch.randn(mx, k, device=dev) * 0.3).to(torch.bfloat16)
        y = (torch.randn(my, k, device=dev) * 0.3).to(torch.bfloat16)
        row = col = None
        if mode in ("row", "both"):
            row = (F.gemm_lse(x, y, 1.0)[0], torch.randn(mx, device=dev),
                   torch.randint(-1, my, (mx,), device=dev, dtype=torch.int32))
        if mode in ("col", "both"):
            col = (F.gemm_lse(y, x, 1.0)[0], torch.randn(my, device=dev),
                   torch.randint(-1, mx, (my,), device=dev, dtype=torch.int32))
        ox, oy = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
        sx = F.softmax_grad_gemm(x, y, 1.0, row=row, col=col)
        sy = F.softmax_grad_gemm(y, x, 1.0, row=col, col=row)
        assert rel(ox, sx) < 1e-3 and rel(oy, sy) < 1e-3, (i, mx, my, k, mode)


@pytest.mark.parametrize("B,D", [(8, 512), (64, 512), (128, 256), (33, 128), (1, 384), (100, 512)])
@pytest.mark.parametrize("tau,mean", [(0.07, True), (0.5, False)])
def test_ntxent_small_single_launch(pg, cuda_device, B, D, tau, mean):
    """The one-launch small-batch NT-Xent (trainer batch 8, cfg1 batch 64) against the float64 oracle: loss, both LSE
    vectors and the gradients; and through losses.ContrastiveLoss with a non-unit upstream gradient."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    g = torch.Generator().manual_seed(B * 7 + D)
    a = bf16r(torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1))
    b = bf16r(torch.nn.functional.normalize(a + 0.3 * torch.randn(B, D, generator=g), dim=-1))
    o = cf.ntxent(a.double().numpy(), b.double().numpy(), tau, normalize=False, clamp_tau=False,
                  reduction="mean" if mean else "sum")
    assert F.ntxent_small_supported(B, D)
    loss, lr, lc, da, db = F.ntxent_small(a.to(dev).to(torch.bfloat16), b.to(dev).to(torch.bfloat16), 1.0 / tau, mean)
    assert loss_close(loss.item(), o["loss"], 1.0 / tau)
    assert rel(da, o["dx"]) < GRAD_RTOL and rel(db, o["dy"]) < GRAD_RTOL
    z = (a.double() @ b.double().T) / tau
    assert rel(lr, torch.logsumexp(z, 1)) < 1e-5 and rel(lc, torch.logsumexp(z, 0)) < 1e-5
    if mean:
        ag, bg = a.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
        (3.0 * pg.ContrastiveLoss(temperature=tau)(ag, bg)).backward()
        assert rel(ag.grad, 3.0 * o["dx"]) < GRAD_RTOL and rel(bg.grad, 3.0 * o["dy"]) < GRAD_RTOL


def test_graphed_contrastive_step(pg, cuda_device):
    """GraphedContrastiveStep replays to the eager result of both NT-Xent flavours, also after new inputs are copied in."""
    from preference_guided_image_captioning_alignment_b200 import components
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    for mod, B in ((pg.ContrastiveLoss(temperature=0.07), 8), (components.ContrastiveLoss(temperature=0.5), 64)):
        xs = [torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).to(dev) for _ in range(4)]
        step = pg.GraphedContrastiveStep(mod, xs[0], xs[1])
        for a, b in ((xs[0], xs[1]), (xs[2], xs[3])):
            step.copy_inputs(a, b)
            loss = step.replay()
            ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
            ref = mod(ag, bg)
            ref.backward()
            assert loss.item() == pytest.approx(ref.item(), rel=1e-6)
            assert torch.equal(step.da, ag.grad) and torch.equal(step.db, bg.grad)


# ================================================================================================ compacted rows
def test_compact_rows_bit_exact(pg, cuda_device):
    """Index / mask plumbing of the compacted Stage-2 head, bit-exact: the scored-row list equals
    nonzero(mask[:, 1:]) of the reference's shift (components.py:339-344), gathered rows are the bf16 roundings of the
    source rows, labels follow their rows, and scatter puts every row back where it came from, zeros elsewhere."""
    import ctypes

    from preference_guided_image_captioning_alignment_b200 import _lib
    from preference_guided_image_captioning_alignment_b200 import functional as F
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    for nseq, T, d, kind in ((5, 37, 64, "i64"), (16, 128, 1024, "i64"), (3, 1500, 128, "f32"), (2, 9, 8, "none")):
        V = 1000
        labels = torch.randint(0, V + 3, (nseq, T), generator=g).to(cuda_device)        # some labels >= V
        lens = torch.randint(0, T + 1, (nseq,), generator=g)
        m01 = (torch.arange(T)[None] < lens[:, None])
        mask = {"i64": m01.long(), "f32": m01.float() * torch.rand(nseq, T, generator=g), "none": None}[kind]
        mask = None if mask is None else mask.to(cuda_device)
        rl, rw = F.prep_rows(labels, mask, V)
        idx = torch.full((nseq * T,), -7, dtype=torch.int32, device=cuda_device)
        cnt = torch.zeros(1, dtype=torch.int32, device=cuda_device)
        _lib.check(lib.pgica_compact_rows(ctypes.c_void_p(rw.data_ptr()), nseq * T, ctypes.c_void_p(idx.data_ptr()),
                                          ctypes.c_void_p(cnt.data_ptr()), None))
        # expected: shifted mask, last position of every sequence never scored
        w = torch.ones(nseq, T, device=cuda_device) if mask is None else mask.float()
        wshift = torch.cat([w[:, 1:], torch.zeros(nseq, 1, device=cuda_device)], 1).reshape(-1)
        yshift = torch.cat([labels[:, 1:], torch.zeros(nseq, 1, dtype=torch.long, device=cuda_device)], 1).reshape(-1)
        scored = (wshift != 0)
        expect = torch.nonzero(scored).flatten().int()
        n = int(cnt.item())
        assert n == expect.numel()
        assert torch.equal(idx[:n], expect) and bool((idx[n:] == -7).all())
        oov = scored & (yshift >= V)
        assert bool(torch.isnan(rw[oov]).all()) and torch.equal(rw[scored & ~oov], wshift[scored & ~oov])
        if n == 0:
            continue
        src = torch.randn(nseq * T, d, generator=g).to(cuda_device)
        dst = torch.empty(n, d, dtype=torch.bfloat16, device=cuda_device)
        lab_c = torch.empty(n, dtype=torch.int32, device=cuda_device)
        _lib.check(lib.pgica_gather_rows_bf16(ctypes.c_void_p(src.data_ptr()), 0, ctypes.c_void_p(idx.data_ptr()), n, d,
                                              ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(rl.data_ptr()),
                                              ctypes.c_void_p(lab_c.data_ptr()), None))
        assert torch.equal(dst, src[expect.long()].to(torch.bfloat16)) and torch.equal(lab_c, rl[expect.long()])
        for out_dtype in (torch.float32, torch.bfloat16):
            back = torch.full((nseq * T, d), 5.0, dtype=out_dtype, device=cuda_device)
            _lib.check(lib.pgica_scatter_rows(ctypes.c_void_p(dst.data_ptr()), 1, ctypes.c_void_p(idx.data_ptr()), n, d,
                                              ctypes.c_void_p(back.data_ptr()), 1 if out_dtype == torch.bfloat16 else 0,
                                              nseq * T, None))
            ref = torch.zeros(nseq * T, d, dtype=out_dtype, device=cuda_device)
            ref[expect.long()] = dst.to(out_dtype)
            assert torch.equal(back, ref)


@pytest.mark.parametrize("length_normalize", [False, True])
def test_compact_head_equals_full_head(pg, cuda_device, length_normalize):
    """The compacted, pair-stacked LM head (what PreferenceLoss runs on LazyLogits) against the full-row op on the same
    inputs: per-row statistics do not depend on which other rows share the GEMM, so sequence log-probs and dhidden agree
    to fp32 rounding; dweight is summed over rows in another order (1e-3).  Also: all-masked sequence -> NaN (mean) / 0
    (sum), zero gradient at masked rows, and both against the float64 oracle."""
    from preference_guided_image_captioning_alignment_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, T, d, V = 6, 40, 512, 3001
    W = (torch.randn(V, d, generator=g) * 0.05).to(cuda_device).requires_grad_(True)          # fp32, not bf16-exact
    hs = [torch.randn(B, T, d, generator=g).to(cuda_device).requires_grad_(True) for _ in range(2)]
    ys = [torch.randint(0, V, (B, T), generator=g).to(cuda_device) for _ in range(2)]
    lens = torch.randint(2, T + 1, (2, B), generator=g)
    ms = [(torch.arange(T)[None] < lens[i][:, None]).long().to(cuda_device) for i in range(2)]
    up = [torch.randn(B, generator=g).to(cuda_device) for _ in range(2)]

    def grads(seqs):
        W.grad = None
        for h in hs:
            h.grad = None
        (seqs[0] * up[0]).sum().add((seqs[1] * up[1]).sum()).backward()
        return W.grad.clone(), hs[0].grad.clone(), hs[1].grad.clone()

    full = [ops.lmhead_seq_logprob(hs[i], W, ys[i], ms[i], length_normalize)[0] for i in range(2)]
    gw_f, g0_f, g1_f = grads(full)
    comp = ops.lmhead_seq_logprob_compact(hs, W, ys, ms, length_normalize)
    gw_c, g0_c, g1_c = grads(comp)
    for a, b in zip(comp, full):
        assert rel(a, b) < 1e-6
    assert rel(g0_c, g0_f) < 2e-3 and rel(g1_c, g1_f) < 2e-3       # full path hands dhidden back through bf16
    assert rel(gw_c, gw_f) < 2e-3
    for i in range(2):
        masked = (torch.cat([ms[i][:, 1:], torch.zeros(B, 1, dtype=torch.long, device=cuda_device)], 1) == 0)
        assert bool((hs[i].grad[masked] == 0).all())
    # float64 oracle on the bf16-rounded operands
    n = lambda x: bf16r(x.detach()).double().cpu().numpy()
    o = cf.dpo_head(n(hs[0]), n(hs[1]), n(W), ys[0].cpu().numpy(), ys[1].cpu().numpy(), ms[0].cpu().numpy(),
                    ms[1].cpu().numpy(), beta=0.1)
    key = "pc_mean" if length_normalize else "pc"
    if key in o:
        assert rel(comp[0], o[key]) < 1e-5
    # an all-masked sequence: NaN in mean mode, exactly 0 in sum mode, like the reference
    ms0 = ms[0].clone()
    ms0[1] = 0
    out = ops.lmhead_seq_logprob_compact([hs[0]], W, [ys[0]], [ms0], length_normalize)[0]
    assert (math.isnan(out[1].item()) if length_normalize else out[1].item() == 0.0)
    assert rel(out[[0, 2, 3, 4, 5]], full[0][[0, 2, 3, 4, 5]]) < 1e-6
    # everything masked: no GEMM at all, zero gradients
    out = ops.lmhead_seq_logprob_compact([hs[0]], W, [ys[0]], [torch.zeros_like(ms0)], False)[0]
    W.grad = None
    out.sum().backward()
    assert bool((out == 0).all()) and bool((W.grad == 0).all())


# ================================================================================================ fp32 inputs
@pytest.mark.parametrize("seed", [5, 6])
@pytest.mark.parametrize("tau", [0.5, 0.07])
def test_ntxent_fp32_inputs_golden(pg, cuda_device, golden_dir, seed, tau):
    """Inputs that are NOT bf16-representable (fp32 unit vectors, exactly what the model hands the trainer's loss):
    the fp32 path (two-term bf16 split, similarity of depth 3*D) keeps the 1e-4 loss bar against the reference run in
    float64 on the same fp32 values — trainer flavour through the one-launch kernel, components flavour through the
    general kernels."""
    from preference_guided_image_captioning_alignment_b200 import components
    g = np.load(os.path.join(golden_dir, "fp32_inputs.npz"))
    vn = torch.from_numpy(g[f"s{seed}_vn"]).to(cuda_device).requires_grad_(True)
    tn = torch.from_numpy(g[f"s{seed}_tn"]).to(cuda_device).requires_grad_(True)
    loss = pg.ContrastiveLoss(temperature=tau)(vn, tn)
    k = f"s{seed}_trainer_tau{tau}"
    assert loss_close(loss.item(), float(g[k + "_loss"]), 1.0 / tau), (loss.item(), float(g[k + "_loss"]))
    loss.backward()
    assert rel(vn.grad, g[k + "_dv"]) < GRAD_RTOL and rel(tn.grad, g[k + "_dt"]) < GRAD_RTOL
    v = torch.from_numpy(g[f"s{seed}_v"]).to(cuda_device).requires_grad_(True)
    t = torch.from_numpy(g[f"s{seed}_t"]).to(cuda_device).requires_grad_(True)
    loss = components.ContrastiveLoss(temperature=tau)(v, t)
    k = f"s{seed}_comp_tau{tau}"
    assert loss_close(loss.item(), float(g[k + "_loss"]), 10.0), (loss.item(), float(g[k + "_loss"]))
    loss.backward()
    assert rel(v.grad, g[k + "_dv"]) < GRAD_RTOL and rel(t.grad, g[k + "_dt"]) < GRAD_RTOL


@pytest.mark.parametrize("B", [300, 1024])
def test_ntxent_fp32_inputs_general_path(pg, cuda_device, B):
    """B > 128 with fp32 unit vectors: the general kernels on split operands against the float64 oracle (tau = 0.07,
    where a bf16 rounding of the inputs would cost ~1e-3 of the loss)."""
    g = torch.Generator().manual_seed(B)
    a = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1)
    b = torch.nn.functional.normalize(a + 0.3 * torch.randn(B, 512, generator=g), dim=-1)
    ag, bg = a.to(cuda_device).requires_grad_(True), b.to(cuda_device).requires_grad_(True)
    loss = pg.ContrastiveLoss(temperature=0.07)(ag, bg)
    loss.backward()
    o = cf.ntxent(a.double().numpy(), b.double().numpy(), 0.07, normalize=False, clamp_tau=False)
    assert loss_close(loss.item(), o["loss"], 1.0 / 0.07), (loss.item(), o["loss"])
    assert rel(ag.grad, o["dx"]) < GRAD_RTOL and rel(bg.grad, o["dy"]) < GRAD_RTOL


def test_preference_loss_fp32_hidden_golden(pg, cuda_device, golden_dir):
    """LM head on fp32 hidden states and an fp32 weight (what install() sees in the trainer): PreferenceLoss on
    LazyLogits against the real reference run in float64 on the same fp32 values."""
    from preference_guided_image_captioning_alignment_b200.losses import LazyLogits
    g = np.load(os.path.join(golden_dir, "fp32_inputs.npz"))
    t = lambda k: torch.from_numpy(g[k]).to(cuda_device)
    W, hw, hl = (t(k).requires_grad_(True) for k in ("lm_W", "lm_hw", "lm_hl"))
    loss = pg.PreferenceLoss(0.1)(LazyLogits(hw, W), LazyLogits(hl, W), t("lm_yw"), t("lm_yl"), t("lm_mw"), t("lm_ml"))
    assert abs(loss.item() - float(g["lm_loss"])) <= LOSS_RTOL * abs(float(g["lm_loss"]))
    loss.backward()
    assert rel(W.grad, g["lm_dW"]) < GRAD_RTOL
    assert rel(hw.grad, g["lm_dhw"]) < GRAD_RTOL and rel(hl.grad, g["lm_dhl"]) < GRAD_RTOL
    lp = pg.PreferenceLoss(0.1)._compute_log_probs(LazyLogits(hw, W), t("lm_yw"), t("lm_mw"))
    assert rel(lp, g["lm_lpw"]) < 1e-4


# ================================================================================================ a1 / scoring
def test_dense_similarity_backward_own_gemm(pg, cuda_device):
    """TemperatureScaledSimilarity (components.py:61-83) as a differentiable dense matrix: forward and the backward of
    a dense upstream gradient (both GEMMs on the library's tcgen05 kernel) against float64 autograd of the reference
    arithmetic; learnable tau inside and outside the clamp range; B not a multiple of 8."""
    from oracle import torch_port as tp
    from preference_guided_image_captioning_alignment_b200 import _lib, components
    for B, D, tau in ((37, 128, 0.5), (64, 512, 0.07), (130, 256, 1.3)):
        g = torch.Generator().manual_seed(B)
        v, t = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
        up = torch.randn(B, B, generator=g)
        vd, td = v.double().requires_grad_(True), t.double().requires_grad_(True)
        taud = torch.tensor(tau, dtype=torch.float64, requires_grad=True)
        vh, th = torch.nn.functional.normalize(vd, dim=-1), torch.nn.functional.normalize(td, dim=-1)
        Sd = vh @ th.T / torch.clamp(taud, 0.1, 2.0)
        (Sd * up.double()).sum().backward()
        mod = components.TemperatureScaledSimilarity(tau, learnable=True).to(cuda_device)
        vg, tg = v.to(cuda_device).requires_grad_(True), t.to(cuda_device).requires_grad_(True)
        n0 = _lib.load().pgica_kernel_launches()
        S = mod(vg, tg)
        (S * up.to(cuda_device)).sum().backward()
        assert _lib.load().pgica_kernel_launches() - n0 >= 7   # 2 norms + S, 2 GEMMs + 2 norm backwards: all ours
        assert rel(S, Sd.detach()) < 5e-3
        assert rel(vg.grad, vd.grad) < GRAD_RTOL and rel(tg.grad, td.grad) < GRAD_RTOL
        assert abs(mod.temperature.grad.item() - taud.grad.item()) <= 2e-2 * abs(taud.grad.item()) + 1e-6
    # the temperature is read once, not per call: no device->host copy on the second forward
    mod = components.TemperatureScaledSimilarity(0.5).to(cuda_device)
    mod(vg.detach(), tg.detach())
    key = mod.__dict__["_tau_cache"][0]
    mod(vg.detach(), tg.detach())
    assert mod.__dict__["_tau_cache"][0] == key
    with torch.no_grad():
        mod.temperature.fill_(0.25)
    assert mod.effective_temperature() == 0.25


def test_scoring_ops(pg, cuda_device):
    """compute_similarity (model.py:925-954), per-pair CLIP-style scores (metrics.py:380-439) and
    retrieval ranks against fp32/fp64 torch."""
    g = torch.Generator().manual_seed(4)
    for B, Bt, D in ((50, 50, 512), (130, 77, 256)):
        img = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1)
        txt = torch.nn.functional.normalize(torch.randn(Bt, D, generator=g), dim=-1)
        ref = img.double() @ txt.double().T / 0.07
        S = pg.compute_similarity(img.to(cuda_device), txt.to(cuda_device), 0.07)
        assert S.shape == (B, Bt) and rel(S, ref) < 1e-5
        Sb = pg.compute_similarity(img.to(cuda_device).bfloat16(), txt.to(cuda_device).bfloat16(), 0.07)
        assert rel(Sb, img.bfloat16().double() @ txt.bfloat16().double().T / 0.07) < 1e-5
    raw_i, raw_t = torch.randn(64, 512, generator=g), torch.randn(64, 512, generator=g)
    sc = pg.paired_scores(raw_i.to(cuda_device), raw_t.to(cuda_device), 100.0)
    ref = 100.0 * torch.nn.functional.cosine_similarity(raw_i.double(), raw_t.double(), dim=-1)
    assert rel(sc, ref) < 1e-5
    a = torch.nn.functional.normalize(torch.randn(40, 128, generator=g), dim=-1)
    b = torch.nn.functional.normalize(a + 0.8 * torch.randn(40, 128, generator=g), dim=-1)
    ranks = pg.retrieval_ranks(a.to(cuda_device), b.to(cuda_device))
    sim = a.double() @ b.double().T
    assert torch.equal(ranks.cpu(), (sim > sim.diagonal()[:, None]).sum(1))


def test_options_api_selects_kernels(pg, cuda_device):
    """pgica_set_option replaces the per-call getenv of round 1: the plan and the kernel choice are explicit
    process-wide options, and both backward variants agree."""
    from preference_guided_image_captioning_alignment_b200 import _lib
    from preference_guided_image_captioning_alignment_b200 import functional as F
    g = torch.Generator().manual_seed(2)
    B, T, d, V = 4, 32, 512, 2000
    h = torch.randn(B, T, d, generator=g).bfloat16().to(cuda_device)
    W = (torch.randn(V, d, generator=g) * 0.05).bfloat16().to(cuda_device)
    y = torch.randint(0, V, (B, T), generator=g).to(cuda_device)
    seq, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(h, W, y, None, False)
    gs = torch.randn(B, generator=g).to(cuda_device)
    assert _lib.get_option("sgg_fused") == 1 and _lib.get_option("sggf_coop") in (0, 1)
    outs = {}
    try:
        for fused in (1, 0):
            _lib.set_option("sgg_fused", fused)
            outs[fused] = F.lmhead_logprob_bwd(h, W, rl, rw, lse, gs, False, dhidden_dtype=torch.float32)
        _lib.set_option("sgg_fused", 1)
        _lib.set_option("sggf_plan_r2", 1)
        _lib.set_option("sggf_plan_c2", 3)
        outs[2] = F.lmhead_logprob_bwd(h, W, rl, rw, lse, gs, False, dhidden_dtype=torch.float32)
    finally:
        _lib.set_option("sgg_fused", 1)
        _lib.set_option("sggf_plan_r2", 0)
        _lib.set_option("sggf_plan_c2", 0)
    for k in (0, 2):
        assert rel(outs[k][0], outs[1][0]) < 1e-3 and rel(outs[k][1], outs[1][1]) < 1e-3
    with pytest.raises(_lib.PgicaError):
        _lib.set_option("no_such_option", 1)


# ================================================================================================ progress / peer all-reduce
def test_dual_backward_progress_and_peer_allreduce_single_rank(pg, cuda_device):
    """The pieces of the overlapped dW all-reduce on ONE GPU: the dual kernel with progress counters gives the same
    gradients as the plain launch and every segment counter ends exactly at its target; the peer all-reduce kernel,
    running beside it on a second stream (world = 1: it sums one buffer, but waits on the counters and walks every
    segment / flag barrier), leaves dW intact and terminates; three epochs on the same counters and flags."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    g = torch.Generator().manual_seed(12)
    for (B, T, d, V, rows_per_seg) in ((4, 64, 1024, 5003, 1024), (16, 128, 1024, 50257, 6400), (3, 50, 512, 1000, 256)):
        h = torch.randn(B, T, d, generator=g).bfloat16().to(cuda_device)
        W = (torch.randn(V, d, generator=g) * 0.03).bfloat16().to(cuda_device)
        y = torch.randint(0, V, (B, T), generator=g).to(cuda_device)
        _, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(h, W, y, None, False)
        gs = torch.randn(B, generator=g).to(cuda_device)
        dh_ref, dw_ref = F.lmhead_logprob_bwd(h, W, rl, rw, lse, gs, False)
        pairs = (V + 255) // 256
        rows = pairs * 256
        nseg = (rows + rows_per_seg - 1) // rows_per_seg
        buf = torch.zeros(rows * d + 64 * (nseg + 1), dtype=torch.float32, device=cuda_device)
        dw = buf[: V * d].view(V, d)
        progress = torch.zeros(nseg, dtype=torch.int32, device=cuda_device)
        local_sync = torch.zeros(2, dtype=torch.int32, device=cuda_device)
        seg_begin = [min(s * rows_per_seg, rows) * d for s in range(nseg + 1)]
        seg_pairs = [(min((s + 1) * rows_per_seg, rows) - s * rows_per_seg) // 256 for s in range(nseg)]
        side = torch.cuda.Stream(device=cuda_device)
        for epoch in (1, 2, 3):
            dw.zero_()
            torch.cuda.synchronize()
            dh, inc = F.lmhead_logprob_bwd_progress(h, W, rl, rw, lse, gs, dw, progress, rows_per_seg, False)
            assert inc == 8 * (d // 512)
            targets = [epoch * inc * n for n in seg_pairs]
            F.peer_allreduce_progress([buf.data_ptr()], [buf.data_ptr() + 4 * rows * d], 0, progress, targets, seg_begin,
                                      epoch, local_sync, 0, stream=side)
            side.synchronize()
            torch.cuda.synchronize()
            assert progress.tolist() == targets, (progress.tolist(), targets)
            assert rel(dw, dw_ref) < 1e-3 and rel(dh.float(), dh_ref.float()) < 1e-3
            assert local_sync[0].item() == epoch * nseg


# ================================================================================================ full-size parity
def _torch_lmhead_reference(h, W, y, m, gseq, chunk_rows=8192):
    """fp32 torch (TF32 off) reference of the Stage-2 head at full size with the logits MATERIALISED row-chunk by
    row-chunk on the GPU: seq_logp[b] = sum_t m[b,t+1] log softmax(W h[b,t])[y[b,t+1]], and the gradients of
    sum_b gseq[b] seq_logp[b] w.r.t. h and W from the closed form dZ = coef (onehot - softmax)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    nseq, T, d = h.shape
    V = W.shape[0]
    Wf = W.float()
    hf = h.float().reshape(nseq * T, d)
    tgt = torch.cat([y[:, 1:], torch.zeros(nseq, 1, dtype=y.dtype, device=y.device)], 1).reshape(-1)
    wt = torch.cat([m[:, 1:].float(), torch.zeros(nseq, 1, device=h.device)], 1).reshape(-1)
    coef = wt * gseq.float().repeat_interleave(T)
    logp = torch.empty(nseq * T, device=h.device)
    dh = torch.zeros(nseq * T, d, device=h.device)
    dW = torch.zeros(V, d, device=h.device)
    for r0 in range(0, nseq * T, chunk_rows):
        sl = slice(r0, min(r0 + chunk_rows, nseq * T))
        z = hf[sl] @ Wf.T
        lp = torch.log_softmax(z, dim=-1)
        logp[sl] = lp.gather(1, tgt[sl, None]).squeeze(1)
        dz = -torch.exp(lp) * coef[sl, None]
        dz.scatter_add_(1, tgt[sl, None], coef[sl, None])
        dh[sl] = dz @ Wf
        dW += dz.T @ hf[sl]
    seq = (logp * wt).reshape(nseq, T).sum(1)
    return seq, dh.reshape(nseq, T, d), dW


@pytest.mark.parametrize("plan", [None, "8,11", "16,9"])
def test_cfg2_full_size_gradients_vs_torch(pg, cuda_device, plan):
    """The EXACT launch bench.py times (cfg2: 16 pairs, seq 128, d = 1024, V = 50257: 4096 x 50257 logits; planner's own
    split, the two-chunk split of round 1 and a pinned one-chunk split) against fp32 torch with the logits
    materialised: every sequence log-prob, all of dH and all of dW."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    W, h, y, m = _cfg2_inputs(dev, 16)
    gseq = torch.randn(32, generator=torch.Generator().manual_seed(3)).to(dev)
    seq_ref, dh_ref, dW_ref = _torch_lmhead_reference(h, W, y, m, gseq)
    set_plan(plan)
    seq, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(h, W, y, m, False)
    dh, dw = F.lmhead_logprob_bwd(h, W, rl, rw, lse, gseq, False, dhidden_dtype=torch.float32)
    set_plan(None)
    assert rel(seq, seq_ref) < 1e-5
    assert rel(dh, dh_ref) < GRAD_RTOL and rel(dw, dW_ref) < GRAD_RTOL
    assert torch.count_nonzero(dh[:, -1]).item() == 0


def test_cfg4_rank_slice_gradients_vs_torch(pg, cuda_device):
    """One rank's share of BASELINE config 4 at FULL size (32 pairs, seq 512: 32768 rows x 50257 logits, 8 chunks of the
    dual kernel with dW accumulated in place) against fp32 torch with materialised logits."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    dev = cuda_device
    W, h, y, m = _cfg2_inputs(dev, 32, T=512, seed=4321)
    gseq = torch.randn(64, generator=torch.Generator().manual_seed(4)).to(dev)
    seq_ref, dh_ref, dW_ref = _torch_lmhead_reference(h, W, y, m, gseq, chunk_rows=4096)
    seq, lse, _, rl, rw, _ = F.lmhead_logprob_fwd(h, W, y, m, False)
    dh, dw = F.lmhead_logprob_bwd(h, W, rl, rw, lse, gseq, False, dhidden_dtype=torch.float32)
    assert rel(seq, seq_ref) < 1e-5
    assert rel(dh, dh_ref) < GRAD_RTOL and rel(dw, dW_ref) < GRAD_RTOL


def test_cfg3_rank_slice_vs_torch(pg, cuda_device):
    """One rank's share of BASELINE config 3 at FULL size: 4096 local rows against 32768 gathered rows (D = 512,
    tau = 0.5, positives on diagonal offset 3 * 4096): row LSE, diagonal, column-LSE partial, dA and the dB partial
    against fp32 torch with the 4096 x 32768 slice materialised; loss terms within 1e-4."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = cuda_device
    nb, Bg, D, tau, off = 4096, 32768, 512, 0.5, 3 * 4096
    g = torch.Generator().manual_seed(33)
    b_all = torch.nn.functional.normalize(torch.randn(Bg, D, generator=g), dim=-1).bfloat16().to(dev)
    a = torch.nn.functional.normalize(b_all[off:off + nb].float().cpu() + 0.5 * torch.randn(nb, D, generator=g),
                                      dim=-1).bfloat16().to(dev)
    S = (a.float() @ b_all.float().T) / tau
    lse_row_ref = torch.logsumexp(S, 1)
    lse_col_ref = torch.logsumexp(S, 0)
    idx = torch.arange(nb, device=dev)
    diag_ref = S[idx, idx + off]
    lse_row, diag, lse_col = F.ntxent_fwd(a, b_all, 1.0 / tau, off)
    assert rel(lse_row, lse_row_ref) < 1e-5 and rel(lse_col, lse_col_ref) < 1e-5 and rel(diag, diag_ref) < 1e-5
    loss_terms = (lse_row - diag).sum().item()
    assert abs(loss_terms - (lse_row_ref - diag_ref).sum().item()) <= LOSS_RTOL * abs(loss_terms)
    # backward with the rank-local column LSE standing in for the merged one (same closed form, one rank)
    one = torch.ones((), device=dev)
    mult = 1.0 / (2.0 * Bg)
    da, db = F.ntxent_bwd(a, b_all, 1.0 / tau, off, lse_row, lse_col, one, mult)
    dS = torch.exp(S - lse_row_ref[:, None]) + torch.exp(S - lse_col_ref[None, :])
    dS[idx, idx + off] -= 2.0
    dS *= mult / tau
    assert rel(da, dS @ b_all.float()) < GRAD_RTOL and rel(db, dS.T @ a.float()) < GRAD_RTOL


@pytest.mark.parametrize("mx,my,k", [(100, 5003, 1024), (470, 50257, 1024), (300, 3000, 512), (513, 9000, 1536)])
def test_dual_backward_column_groups(pg, cuda_device, mx, my, k):
    """Few row blocks (the compacted Stage-2 batches): several X-holder pairs share a row pair's sweep over the vocabulary
    ("column groups", partial OutX add-reduced into zeros).  Planner's choice, pinned group counts and the un-grouped
    launch give the same gradients; OutX rows past mx stay untouched; fp32 torch agrees."""
    from preference_guided_image_captioning_alignment_b200 import _lib
    from preference_guided_image_captioning_alignment_b200 import functional as F
    x, y, row, col = _dual_inputs(cuda_device, mx, my, k, "row")
    outs = {}
    try:
        for groups in (0, 1, 2, 5):
            _lib.set_option("sggf_col_groups", groups)
            outs[groups] = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col)
    finally:
        _lib.set_option("sggf_col_groups", 1)
    for groups in (1, 2, 5):
        # same tiles, another order of fp32 partial sums over up to 197 column pairs
        assert rel(outs[groups][0], outs[0][0]) < 2e-4 and rel(outs[groups][1], outs[0][1]) < 1e-5, groups
    exact_x, _ = _sgg_reference(x, y, 1.0, row, col)
    assert rel(outs[1][0], exact_x) < GRAD_RTOL
    # bf16 OutX cannot be add-reduced: the launch falls back to one group and still works
    bx, _ = F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col, out_x_dtype=torch.bfloat16)
    assert rel(bx.float(), exact_x) < GRAD_RTOL


# ================================================================================================ SURVEY 8(f) rows 3, 4
@pytest.mark.parametrize("B,T,E,H", [(8, 128, 1024, 8), (3, 17, 256, 4)])
def test_collapsed_cross_attention_matches_multihead_attention(pg, cuda_device, B, T, E, H):
    """Row 4: the decoder's one-key cross-attention + residual LayerNorm (pkg/models/model.py:528-535, 594-601) as one
    fused launch against the real nn.MultiheadAttention + nn.LayerNorm — forward and every gradient the reference block
    produces (text, image, in_proj, out_proj, LayerNorm), evaluation mode and training mode with dropout 0; with
    dropout 0.1 the same keep-mask through a dense restatement."""
    from preference_guided_image_captioning_alignment_b200 import prologue
    torch.manual_seed(B * T)
    mha = torch.nn.MultiheadAttention(E, H, dropout=0.0, batch_first=True).to(cuda_device)
    norm = torch.nn.LayerNorm(E).to(cuda_device)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_(0, 0.1)
    text = torch.randn(B, T, E, device=cuda_device, requires_grad=True)
    vis = torch.randn(B, 1, E, device=cuda_device, requires_grad=True)
    up = torch.randn(B, T, E, device=cuda_device)
    params = list(mha.parameters()) + list(norm.parameters())

    def grads(out):
        for t in [text, vis] + params:
            t.grad = None
        (out * up).sum().backward()
        return [t.grad.clone() if t.grad is not None else None for t in [text, vis] + params]

    for mode in ("eval", "train"):
        mha.train(mode == "train")
        ref = norm(text + mha(query=text, key=vis, value=vis)[0])
        g_ref = grads(ref)
        out = prologue.collapsed_cross_attention_ln(text, vis, mha, norm)
        g = grads(out)
        assert rel(out, ref) < 1e-5
        for a, b in zip(g, g_ref):
            assert (a is None) == (b is None)
            if a is not None and b.abs().max() > 0:
                assert rel(a, b) < 1e-4
            elif a is not None:
                assert a.abs().max().item() == 0.0            # query / key projections: exact zeros, like the reference
    # the instance patches: the reference's own two lines (model.py:594-601) run the fused kernel and give the same
    dec = types.SimpleNamespace(cross_attention=mha, attention_norm=norm)
    mha.eval()
    ref = norm(text + mha(query=text, key=vis, value=vis)[0]).detach()
    prologue.fuse_cross_attention(dec)
    attended, _ = dec.cross_attention(query=text, key=vis, value=vis)
    assert isinstance(attended, prologue.CollapsedAttention)
    out = dec.attention_norm(text + attended)
    assert rel(out, ref) < 1e-5
    assert rel(dec.attention_norm(text.detach() * 2), torch.nn.functional.layer_norm(text.detach() * 2, (E,), norm.weight, norm.bias)) < 1e-6
    assert rel(attended.materialize(), type(mha).forward(mha, text, vis, vis)[0]) < 1e-5
    two_keys = torch.randn(B, 2, E, device=cuda_device)
    assert torch.is_tensor(dec.cross_attention(query=text, key=two_keys, value=two_keys)[0])   # not collapsible: stock path
    prologue.unfuse_cross_attention(dec)
    # attention dropout: keep-mask per (batch, position, head), against a dense restatement with the same mask
    mha.train()
    mha.dropout = 0.1
    keep = (torch.rand(B, T, H, device=cuda_device) >= 0.1).float() / 0.9
    hd = E // H
    v = vis.reshape(B, E) @ mha.in_proj_weight[2 * E:].t() + mha.in_proj_bias[2 * E:]
    heads = v.view(B, 1, H, hd) * keep.unsqueeze(-1)                          # (B, T, H, hd): dropped attention weights
    ref = norm(text + heads.reshape(B, T, E) @ mha.out_proj.weight.t() + mha.out_proj.bias)
    g_ref = grads(ref)
    out = prologue.collapsed_cross_attention_ln(text, vis, mha, norm, keep_weights=keep)
    g = grads(out)
    assert rel(out, ref) < 1e-5
    for a, b in zip(g, g_ref):
        if a is not None and b is not None and b.abs().max() > 0:
            assert rel(a, b) < 1e-4
    # and statistically, with its own mask: the mean over many draws approaches the dropout-free output
    mha.dropout = 0.1
    acc = sum(prologue.collapsed_cross_attention_ln(text, vis, mha, norm).detach() for _ in range(8)) / 8
    assert acc.shape == (B, T, E) and bool(torch.isfinite(acc).all())


@pytest.mark.parametrize("rows,D", [(8, 512), (64, 512), (100, 256), (3, 1024)])
def test_ln_l2norm_matches_layernorm_then_normalize(pg, cuda_device, rows, D):
    """Row 3: LayerNorm -> F.normalize (pkg/models/model.py:141/343 and 828-829) from one launch, both outputs and the
    backward of both, against torch; and the carrier that lets the reference's own F.normalize call pick up the twin."""
    from preference_guided_image_captioning_alignment_b200 import prologue
    torch.manual_seed(rows)
    norm = torch.nn.LayerNorm(D).to(cuda_device)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_(0, 0.2)
    z = torch.randn(rows, D, device=cuda_device, requires_grad=True)
    ue, un = torch.randn(rows, D, device=cuda_device), torch.randn(rows, D, device=cuda_device)

    def grads(e, n):
        z.grad = norm.weight.grad = norm.bias.grad = None
        ((e * ue).sum() + (n * un).sum()).backward()
        return z.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone()

    e_ref = norm(z)
    n_ref = torch.nn.functional.normalize(e_ref, p=2, dim=-1)
    g_ref = grads(e_ref, n_ref)
    e, n = prologue.ln_l2norm(z, norm)
    g = grads(e, n)
    assert rel(e, e_ref) < 1e-5 and rel(n, n_ref) < 1e-5
    for a, b in zip(g, g_ref):
        assert rel(a, b) < 1e-4
    # only one of the outputs used downstream
    z.grad = None
    (prologue.ln_l2norm(z, norm)[1] * un).sum().backward()
    z2 = z.grad.clone()
    z.grad = None
    (torch.nn.functional.normalize(norm(z), dim=-1) * un).sum().backward()
    assert rel(z2, z.grad) < 1e-4
    # instance patch on a projection head shaped like the reference's (Linear-ReLU-Dropout-Linear-LayerNorm)
    enc = types.SimpleNamespace(projection=torch.nn.Sequential(torch.nn.Linear(D, D), torch.nn.ReLU(), torch.nn.Dropout(0.0),
                                                               torch.nn.Linear(D, D), norm).to(cuda_device))
    x = torch.randn(rows, D, device=cuda_device)
    ref_e = enc.projection(x)
    ref_n = torch.nn.functional.normalize(ref_e, p=2, dim=-1)
    prologue.fuse_projection_tail(enc)
    out = enc.projection(x)
    assert isinstance(out, prologue.NormalizedCarrier)
    got_n = torch.nn.functional.normalize(out, p=2, dim=-1)
    assert got_n is out._pgica_normalized and rel(got_n, ref_n) < 1e-5 and rel(out, ref_e) < 1e-5
    assert type(out.float()) is torch.Tensor and type(out + 1) is torch.Tensor          # everything else: plain tensors
    assert rel(torch.nn.functional.normalize(out, p=2, dim=0), torch.nn.functional.normalize(ref_e, p=2, dim=0)) < 1e-6
    prologue.unfuse_projection_tail(enc)
    assert type(enc.projection(x)) is torch.Tensor


def test_grad_norm_clip_odd_layouts(pg, cuda_device):
    """NaNSafeGradientNorm on gradients the multi-tensor kernel cannot take in place (a transposed view, fp16, fp64):
    same norm and the same clipped values as torch.nn.utils.clip_grad_norm_ (do not raise)."""
    g = torch.Generator().manual_seed(21)
    shapes = [(33, 65), (128,), (7, 9, 5), (40, 24)]
    ps = [torch.nn.Parameter(torch.zeros(*s, device=cuda_device)) for s in shapes]
    ps[3] = torch.nn.Parameter(torch.zeros(40, 24, device=cuda_device, dtype=torch.float16))
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    for p, r in zip(ps, ref):
        gr = torch.randn(*p.shape, generator=g).to(cuda_device) * 3
        p.grad = gr.to(p.dtype)
        r.grad = gr.to(p.dtype).clone()
    ps[0].grad = ps[0].grad.t().contiguous().t()          # same values, non-contiguous layout
    assert not ps[0].grad.is_contiguous()
    norm_ref = torch.nn.utils.clip_grad_norm_(ref, 1.0)
    total, finite = pg.NaNSafeGradientNorm(max_norm=1.0)(ps)
    assert finite and abs(total.item() - norm_ref.item()) <= 1e-4 * norm_ref.item()
    for p, r in zip(ps, ref):
        assert rel(p.grad.float(), r.grad.float()) < 2e-3


@pytest.mark.parametrize("ra,rb,D,tau,off", [(64, 64, 512, 0.5, 0), (300, 1000, 256, 0.07, 200), (4096, 4096, 512, 0.5, 0),
                                             (1000, 3001, 128, 0.1, 1500), (4096, 32768, 512, 0.5, 8192)])
def test_ntxent_forward_one_pass_matches_two_pass(pg, cuda_device, ra, rb, D, tau, off):
    """Row and column log-sum-exp from ONE pass over the similarity tiles (unit-norm rows: bounded logits) against the
    exact two-pass forward and fp64 torch; ragged shapes, a diagonal offset (one rank's slice), tau down to 0.07; with
    tau too small for the bounded scheme the call falls back to the two-pass form by itself."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    g = torch.Generator().manual_seed(ra + rb)
    b = torch.nn.functional.normalize(torch.randn(rb, D, generator=g), dim=-1).bfloat16().to(cuda_device)
    a = torch.nn.functional.normalize(b[off:off + ra].float().cpu() + 0.7 * torch.randn(ra, D, generator=g), dim=-1) \
        .bfloat16().to(cuda_device)
    lr2, dg2, lc2 = F.ntxent_fwd(a, b, 1.0 / tau, off)
    lr1, dg1, lc1 = F.ntxent_fwd(a, b, 1.0 / tau, off, bounded=True)
    # the row side is the same arithmetic up to the grouping of the column partials (two epilogue groups per tile)
    assert rel(lr1, lr2) < 1e-6 and float((lr1 - lr2).abs().max()) < 1e-5 and torch.equal(dg1, dg2)
    assert rel(lc1, lc2) < 2e-6 and float((lc1 - lc2).abs().max()) < 2e-5 * max(1.0, 1.0 / tau)
    if ra * rb <= 5_000_000:
        S = a.double() @ b.double().T / tau
        assert rel(lc1, torch.logsumexp(S, 0)) < 1e-6 and rel(lr1, torch.logsumexp(S, 1)) < 1e-6
    lr3, _, lc3 = F.ntxent_fwd(a, b, 1.0 / 0.01, off, bounded=True)   # 2 * 144 binades: exact path taken
    lr4, _, lc4 = F.ntxent_fwd(a, b, 1.0 / 0.01, off)
    assert torch.equal(lc3, lc4) and torch.equal(lr3, lr4)


@pytest.mark.parametrize("ra,rb,D,tau,off", [(256, 256, 512, 0.5, 0), (300, 1000, 512, 0.07, 200), (4096, 4096, 512, 0.5, 0),
                                             (1000, 3001, 512, 0.1, 1500), (4096, 32768, 512, 0.5, 8192)])
def test_ntxent_backward_one_exponential_matches_two(pg, cuda_device, ra, rb, D, tau, off):
    """Unit-norm rows: the backward with one shared exponential per element (bounded) against the two-exponential form
    and, where the similarity matrix fits, fp64 torch autograd; ragged shapes, diagonal offsets, tau down to 0.07; a
    temperature outside the bound takes the exact path."""
    from preference_guided_image_captioning_alignment_b200 import functional as F
    g = torch.Generator().manual_seed(ra * 3 + rb)
    b = torch.nn.functional.normalize(torch.randn(rb, D, generator=g), dim=-1).bfloat16().to(cuda_device)
    a = torch.nn.functional.normalize(b[off:off + ra].float().cpu() + 0.7 * torch.randn(ra, D, generator=g), dim=-1) \
        .bfloat16().to(cuda_device)
    lr, dg, lc = F.ntxent_fwd(a, b, 1.0 / tau, off, bounded=True)
    gl = torch.full((1,), 1.7, device=cuda_device)
    mult = 0.5 / rb
    da2, db2 = F.ntxent_bwd(a, b, 1.0 / tau, off, lr, lc, gl, mult)
    da1, db1 = F.ntxent_bwd(a, b, 1.0 / tau, off, lr, lc, gl, mult, bounded=True)
    assert rel(da1, da2) < 2e-3 and rel(db1, db2) < 2e-3
    if ra * rb <= 5_000_000:
        A, B = a.double().requires_grad_(), b.double().requires_grad_()
        S = A @ B.T / tau
        idx = torch.arange(ra, device=cuda_device)
        # the full column log-sum-exp is what lc holds here (one rank: rows_a may be fewer than rows_b, then the column
        # term of the missing rows is simply absent, exactly as in the two-exponential kernel)
        loss = (torch.logsumexp(S, 1) - S[idx, idx + off]).sum() + (torch.logsumexp(S, 0)[off:off + ra] - S[idx, idx + off]).sum()
        if ra == rb:
            (loss * 1.7 * mult).backward()
            assert rel(da1, A.grad) < 5e-3 and rel(db1, B.grad) < 5e-3
    lr9, _, lc9 = F.ntxent_fwd(a, b, 100.0, off)
    da3, db3 = F.ntxent_bwd(a, b, 100.0, off, lr9, lc9, gl, mult, bounded=True)  # 2 * 144 binades: two-exponential path
    da4, db4 = F.ntxent_bwd(a, b, 100.0, off, lr9, lc9, gl, mult)
    # (dA of a few-row problem is add-reduced over column groups in arrival order: equal to rounding, not bit for bit)
    assert rel(da3, da4) < 1e-6 and torch.equal(db3, db4) and bool(torch.isfinite(da3).all())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_contrastive_loss_assume_normalized(pg, cuda_device, dtype):
    """model-flavour ContrastiveLoss beyond the single-CTA batch: the unit-norm promise (one-pass forward, shared
    exponential backward) gives the loss and gradients of the default path and of fp64 torch."""
    B, D, tau = 1024, 512, 0.07
    g = torch.Generator().manual_seed(9)
    t = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1)
    i = torch.nn.functional.normalize(t + 0.5 * torch.randn(B, D, generator=g), dim=-1)
    res = {}
    for flag in (False, True):
        a = i.to(dtype).to(cuda_device).requires_grad_()
        b = t.to(dtype).to(cuda_device).requires_grad_()
        head = pg.ContrastiveLoss(temperature=tau)
        head.assume_normalized = flag
        loss = head(a, b)
        loss.backward()
        res[flag] = (loss.detach().double(), a.grad.double(), b.grad.double())
    A = i.to(dtype).double().to(cuda_device).requires_grad_()
    Bm = t.to(dtype).double().to(cuda_device).requires_grad_()
    S = A @ Bm.T / tau
    idx = torch.arange(B, device=cuda_device)
    ref = 0.5 * (torch.nn.functional.cross_entropy(S, idx) + torch.nn.functional.cross_entropy(S.T, idx))
    ref.backward()
    tol_l, tol_g = (1e-6, 8e-3) if dtype == torch.bfloat16 else (1e-6, 3e-3)
    for flag in (False, True):
        assert abs(res[flag][0] - ref.detach()) / abs(ref.detach()) < tol_l
        assert rel(res[flag][1], A.grad) < tol_g and rel(res[flag][2], Bm.grad) < tol_g
    assert abs(res[True][0] - res[False][0]) / abs(res[False][0]) < 1e-6
