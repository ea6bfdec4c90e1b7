"""Bring-up checks for the backward kernel and the head-level entry points (B200, via gpurun).
Each case runs in its own subprocess.  python tools/gpu_check2.py [CASE ARGS...]"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case_sgg(mx, my, k, mode, scale, time_it):
    import torch
    from preference_guided_image_captioning_alignment_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(2)
    dev = "cuda"
    x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
    y = (torch.randn(my, k, device=dev) * (0.05 if my > 4096 or mx > 4096 else 0.2)).to(torch.bfloat16)
    row = col = None
    if mode in ("row", "both"):
        lse_r, _ = F.gemm_lse(x, y, scale)
        coef_r = torch.randn(mx, device=dev)
        coef_r[::5] = 0
        tgt_r = torch.randint(0, my, (mx,), device=dev, dtype=torch.int32)
        tgt_r[::3] = -1
        row = (lse_r, coef_r, tgt_r)
    if mode in ("col", "both"):
        lse_c, _ = F.gemm_lse(y, x, scale)
        coef_c = torch.randn(my, device=dev)
        coef_c[::5] = 0
        tgt_c = torch.randint(0, mx, (my,), device=dev, dtype=torch.int32)
        tgt_c[::3] = -1
        col = (lse_c, coef_c, tgt_c)
    out = F.softmax_grad_gemm(x, y, scale, row=row, col=col)
    torch.cuda.synchronize()
    # reference in chunks of rows: exact G and bf16-rounded G
    err_exact = err_emul = ref_max = 0.0
    yf = y.float()
    chunk = 256
    for r0 in range(0, mx, chunk):
        z = (x[r0:r0 + chunk].float() @ yf.t()) * scale
        g = torch.zeros_like(z)
        rows = torch.arange(r0, min(r0 + chunk, mx), device=dev)
        if row is not None:
            p = torch.exp(z - row[0][rows][:, None])
            oh = torch.zeros_like(z)
            t = row[2][rows].long()
            v = t >= 0
            oh[v.nonzero()[:, 0], t[v]] = 1
            g += row[1][rows][:, None] * (p - oh)
        if col is not None:
            p = torch.exp(z - col[0][None, :])
            oh = (col[2][None, :].long() == rows[:, None]).float()
            g += col[1][None, :] * (p - oh)
        ref = g @ yf
        emu = g.to(torch.bfloat16).float() @ yf
        o = out[r0:r0 + chunk]
        err_exact = max(err_exact, (o - ref).abs().max().item())
        err_emul = max(err_emul, (o - emu).abs().max().item())
        ref_max = max(ref_max, ref.abs().max().item())
    res = {"case": "sgg", "mx": mx, "my": my, "k": k, "mode": mode, "err_vs_exact": err_exact,
           "err_vs_bf16G": err_emul, "ref_absmax": ref_max,
           "ok": bool(err_emul < 2e-3 * max(ref_max, 1e-6) + 1e-5 and err_exact < 2e-2 * max(ref_max, 1e-6))}
    if time_it:
        ms = timeit(lambda: F.softmax_grad_gemm(x, y, scale, row=row, col=col))
        res["ms"] = ms
        res["alg_tflops"] = 4.0 * mx * my * k / ms / 1e9  # recompute + product, once each
    print(json.dumps(res))


def case_lmhead(nseq, T, d, V, ln, time_it):
    import torch
    from preference_guided_image_captioning_alignment_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    dev = "cuda"
    h = torch.randn(nseq, T, d, device=dev).to(torch.bfloat16)
    w = (torch.randn(V, d, device=dev) * 0.02).to(torch.bfloat16)
    labels = torch.randint(0, V, (nseq, T), device=dev)
    lens = torch.randint(T // 2, T + 1, (nseq,), device=dev)
    mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).long()
    seq, lse, ztgt, rl, rw, nll = F.lmhead_logprob_fwd(h, w, labels, mask, bool(ln), want_nll=True)
    gseq = torch.randn(nseq, device=dev)
    dh, dw = F.lmhead_logprob_bwd(h, w, rl, rw, lse, gseq, bool(ln), dhidden_dtype=torch.float32)
    torch.cuda.synchronize()
    res = {"case": "lmhead", "nseq": nseq, "T": T, "d": d, "V": V, "ln": ln}
    if nseq * T * V <= 2 ** 28:
        hf = h.float().requires_grad_(True)
        wf = w.float().requires_grad_(True)
        logits = hf @ wf.t()
        lp = torch.log_softmax(logits[:, :-1], -1).gather(-1, labels[:, 1:, None])[..., 0]
        m = mask[:, 1:].float()
        ref = (lp * m).sum(1)
        if ln:
            ref = ref / m.sum(1)
        (ref * gseq).sum().backward()
        rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        res.update(seq_rel=rel(seq, ref.detach()), dh_rel=rel(dh, hf.grad), dw_rel=rel(dw, wf.grad),
                   nll_rel=abs(nll.item() - (-lp.detach().sum().item())) / abs(lp.detach().sum().item()),
                   seq_maxabs=(seq - ref.detach()).abs().max().item())
        res["ok"] = bool(res["seq_rel"] < 1e-4 and res["dh_rel"] < 1e-2 and res["dw_rel"] < 1e-2)
    if time_it:
        res["fwd_ms"] = timeit(lambda: F.lmhead_logprob_fwd(h, w, labels, mask, bool(ln)))
        res["bwd_ms"] = timeit(lambda: F.lmhead_logprob_bwd(h, w, rl, rw, lse, gseq, bool(ln)))
        res["bwd_dh_ms"] = timeit(lambda: F.lmhead_logprob_bwd(h, w, rl, rw, lse, gseq, bool(ln), need_dweight=False))
        res["bwd_dw_ms"] = timeit(lambda: F.lmhead_logprob_bwd(h, w, rl, rw, lse, gseq, bool(ln), need_dhidden=False))
    print(json.dumps(res))


def case_ntxent(B, D, tau, time_it):
    import torch
    from preference_guided_image_captioning_alignment_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(4)
    dev = "cuda"
    a = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1).to(torch.bfloat16)
    b = torch.nn.functional.normalize(a.float() + 0.5 * torch.randn(B, D, device=dev), dim=-1).to(torch.bfloat16)
    lse_r, diag, lse_c = F.ntxent_fwd(a, b, 1.0 / tau)
    loss = F.ntxent_loss(lse_r, diag, lse_c, 1.0 / B)
    g = torch.ones((), device=dev)
    da, db = F.ntxent_bwd(a, b, 1.0 / tau, 0, lse_r, lse_c, g, 1.0 / (2 * B))
    torch.cuda.synchronize()
    res = {"case": "ntxent", "B": B, "D": D, "tau": tau, "loss": loss.item()}
    if B <= 8192:
        af = a.double().requires_grad_(True)
        bf = b.double().requires_grad_(True)
        S = af @ bf.t() / tau
        lab = torch.arange(B, device=dev)
        ref = (torch.nn.functional.cross_entropy(S, lab) + torch.nn.functional.cross_entropy(S.t(), lab)) / 2
        ref.backward()
        rel = lambda x, y: ((x.double() - y).norm() / y.norm()).item()
        res.update(loss_rel=abs(loss.item() - ref.item()) / abs(ref.item()), da_rel=rel(da, af.grad),
                   db_rel=rel(db, bf.grad))
        res["ok"] = bool(res["loss_rel"] < 1e-4 and res["da_rel"] < 1e-2 and res["db_rel"] < 1e-2)
    if time_it:
        res["fwd_ms"] = timeit(lambda: F.ntxent_fwd(a, b, 1.0 / tau))
        res["bwd_ms"] = timeit(lambda: F.ntxent_bwd(a, b, 1.0 / tau, 0, lse_r, lse_c, g, 1.0 / (2 * B)))
    print(json.dumps(res))


def case_small():
    import torch
    from preference_guided_image_captioning_alignment_b200 import functional as F
    torch.manual_seed(5)
    dev = "cuda"
    res = {"case": "small"}
    pc, pr, rc, rr = [torch.randn(16, device=dev) * 20 - 500 for _ in range(4)]
    for ls in (0.0, 0.1):
        pcg = pc.clone().requires_grad_(True)
        x = 0.1 * ((pcg - pr) - (rc - rr))
        ref = (torch.nn.functional.binary_cross_entropy_with_logits(x, (1 - ls) * torch.ones_like(x)) if ls > 0
               else -torch.nn.functional.logsigmoid(x).mean())
        ref.backward()
        loss, met, dpc = F.dpo_loss_fwd(pc, pr, rc, rr, 0.1, ls)
        res[f"dpo_ls{ls}_loss_err"] = abs(loss.item() - ref.item())
        res[f"dpo_ls{ls}_grad_err"] = (dpc - pcg.grad).abs().max().item()
    met_ref = [((pc - pr) - (rc - rr)).mean().item(), ((pc - pr) > (rc - rr)).float().mean().item(), pc.mean().item(),
               pr.mean().item()]
    res["metrics_err"] = max(abs(a - b) for a, b in zip(met[1:].tolist(), met_ref))
    x = torch.randn(100, 512, device=dev)
    y, inv, _ = F.rownorm_fwd(x)
    res["norm_err"] = (y.float() - torch.nn.functional.normalize(x, dim=-1)).abs().max().item()
    g = torch.randn(100, 512, device=dev)
    xg = x.clone().requires_grad_(True)
    (torch.nn.functional.normalize(xg, dim=-1) * g).sum().backward()
    res["norm_bwd_err"] = (F.rownorm_bwd(x, inv, g) - xg.grad).abs().max().item()
    # logits path
    nseq, T, V = 3, 9, 1000
    logits = torch.randn(nseq, T, V, device=dev)
    labels = torch.randint(0, V, (nseq, T), device=dev)
    mask = torch.ones(nseq, T, device=dev)
    mask[0, 6:] = 0
    rl, rw = F.prep_rows(labels, mask, V)
    lse, zt = F.logits_lse(logits, rl)
    seq = F.seq_reduce(lse, zt, rw, nseq, T, True)
    lg = logits.clone().requires_grad_(True)
    lp = torch.log_softmax(lg[:, :-1], -1).gather(-1, labels[:, 1:, None])[..., 0]
    ref = (lp * mask[:, 1:]).sum(1) / mask[:, 1:].sum(1)
    gs = torch.randn(nseq, device=dev)
    (ref * gs).sum().backward()
    coef = F.row_coef(gs, rw, nseq, T, True)
    dl = F.logits_grad(logits, rl, lse, coef)
    res["logits_seq_err"] = (seq - ref.detach()).abs().max().item()
    res["logits_grad_err"] = (dl - lg.grad).abs().max().item()
    res["ok"] = bool(max(v for k, v in res.items() if k.endswith("err")) < 2e-3)
    print(json.dumps(res))


def driver():
    only = os.environ.get("CHECK2_ONLY", "")
    cases = [["small"],
             ["sgg", 128, 128, 256, "row", 1.0, 0], ["sgg", 128, 128, 256, "col", 1.0, 0],
             ["sgg", 128, 128, 256, "both", 1.0, 0], ["sgg", 300, 1000, 512, "both", 2.0, 0],
             ["sgg", 200, 333, 1024, "row", 1.0, 0], ["sgg", 77, 90, 64, "col", 1.0, 0],
             ["lmhead", 2, 16, 256, 1000, 0, 0], ["lmhead", 3, 33, 1024, 5000, 1, 0],
             ["ntxent", 64, 512, 0.5, 1], ["ntxent", 256, 512, 0.1, 0], ["ntxent", 4096, 512, 0.5, 1],
             ["sgg", 4064, 50257, 1024, "row", 1.0, 1], ["sgg", 50257, 4064, 1024, "col", 1.0, 1],
             ["lmhead", 32, 128, 1024, 50257, 0, 1]]
    if only:
        cases = [c for c in cases if c[0] in only.split(",")]
    t0 = time.time()
    print("PGICA_SGG_CLUSTER =", os.environ.get("PGICA_SGG_CLUSTER", "(default)"))
    for c in cases:
        cmd = [sys.executable, os.path.abspath(__file__)] + [str(x) for x in c]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            if r.returncode != 0:
                print(json.dumps({"case": c, "rc": r.returncode, "stderr": r.stderr[-800:], "stdout": r.stdout[-400:]}))
            else:
                print(tail)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": c, "timeout": True}))
        sys.stdout.flush()
    print("elapsed", time.time() - t0)


if __name__ == "__main__":
    a = sys.argv
    if len(a) == 1:
        driver()
    elif a[1] == "sgg":
        case_sgg(int(a[2]), int(a[3]), int(a[4]), a[5], float(a[6]), int(a[7]))
    elif a[1] == "lmhead":
        case_lmhead(int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6]), int(a[7]))
    elif a[1] == "ntxent":
        case_ntxent(int(a[2]), int(a[3]), float(a[4]), int(a[5]))
    elif a[1] == "small":
        case_small()
