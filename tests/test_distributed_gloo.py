"""World-size-2 tests of the multi-GPU plumbing on CPU (gloo).  The collectives, slicing, offsets and gradient
conventions in distributed.py run for real; the kernels are replaced by an oracle-backed stand-in (test double,
tests only) because the product has no CPU compute path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form as cf


class OracleCompute:
    """Same interface as distributed.CudaCompute, arithmetic from oracle/closed_form.py (float64)."""

    name = "oracle"

    def to_operand(self, x):
        return x.detach().double()

    def ntxent_fwd(self, a, b_all, inv_tau, diag_offset):
        S = a.numpy() @ b_all.numpy().T * inv_tau
        lse_row = torch.from_numpy(cf._logsumexp(S, 1))
        lse_col = torch.from_numpy(cf._logsumexp(S, 0))
        idx = np.arange(a.shape[0])
        return lse_row, torch.from_numpy(S[idx, idx + diag_offset].copy()), lse_col

    def lse_combine(self, parts):
        return torch.logsumexp(parts, dim=0)

    def ntxent_loss(self, lse_row, diag, lse_col_owned, inv_denom):
        return 0.5 * inv_denom * ((lse_row - diag).sum() + (lse_col_owned - diag).sum())

    def ntxent_bwd(self, a, b_all, inv_tau, diag_offset, lse_row, lse_col, grad_loss, mult):
        S = a @ b_all.T * inv_tau
        onehot = torch.zeros_like(S)
        idx = torch.arange(a.shape[0])
        onehot[idx, idx + diag_offset] = 1.0
        dS = grad_loss * mult * (torch.exp(S - lse_row[:, None]) + torch.exp(S - lse_col[None, :]) - 2 * onehot)
        return dS @ b_all * inv_tau, dS.T @ a * inv_tau


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from preference_guided_image_captioning_alignment_b200 import distributed as D
        nb, Dm, tau = 5, 16, 0.5
        gen = torch.Generator().manual_seed(7)
        A = torch.nn.functional.normalize(torch.randn(world * nb, Dm, generator=gen, dtype=torch.float64), dim=-1)
        Bm = torch.nn.functional.normalize(torch.randn(world * nb, Dm, generator=gen, dtype=torch.float64), dim=-1)
        a = A[rank * nb:(rank + 1) * nb].clone().requires_grad_(True)
        b = Bm[rank * nb:(rank + 1) * nb].clone().requires_grad_(True)
        loss = D.global_ntxent(a, b, tau, True, None, OracleCompute())
        loss.backward()
        ref = cf.ntxent(A.numpy(), Bm.numpy(), tau)
        out = dict(rank=rank, loss=loss.item(), ref_loss=ref["loss"],
                   da_err=float(np.abs(a.grad.numpy() - ref["dx"][rank * nb:(rank + 1) * nb]).max()),
                   db_err=float(np.abs(b.grad.numpy() - ref["dy"][rank * nb:(rank + 1) * nb]).max()))
        # module form + sum reduction
        l2 = D.global_ntxent(a.detach(), b.detach(), tau, False, None, OracleCompute())
        out["sum_loss"] = l2.item()
        out["ref_sum_loss"] = cf.ntxent(A.numpy(), Bm.numpy(), tau, reduction="sum")["loss"]
        # DPO: shard pairs, global mean, scalar all-reduce
        n_global = 7
        gen = torch.Generator().manual_seed(11)
        pc, pr, rc, rr = [torch.randn(n_global, generator=gen, dtype=torch.float64) * 5 for _ in range(4)]
        s, e = D.shard_pairs(n_global, rank, world)
        loc = cf.dpo_loss(pc[s:e].numpy(), pr[s:e].numpy(), rc[s:e].numpy(), rr[s:e].numpy(), beta=0.1)
        n_loc = e - s
        local_loss = torch.tensor(loc["loss"] * n_loc / n_global, dtype=torch.float64)
        local_metrics = torch.tensor([loc["metrics"][k] * n_loc / n_global for k in
                                      ("dpo_loss", "reward_margin", "reward_accuracy", "policy_chosen_logprob",
                                       "policy_rejected_logprob")], dtype=torch.float64)
        gl, gm = D.allreduce_scalars(local_loss, local_metrics)
        full = cf.dpo_loss(pc.numpy(), pr.numpy(), rc.numpy(), rr.numpy(), beta=0.1)
        out["dpo_loss"], out["dpo_ref"] = gl.item(), full["loss"]
        out["dpo_margin"], out["dpo_margin_ref"] = gm[1].item(), full["metrics"]["reward_margin"]
        # dW all-reduce
        dw = torch.full((4, 3), float(rank + 1), dtype=torch.float64)
        D.allreduce_dweight(dw)
        out["dw"] = dw[0, 0].item()
        # reduce_scatter_rows / all_gather_rows round trip
        x = torch.arange(world * 2 * 3, dtype=torch.float64).reshape(world * 2, 3) * (rank + 1)
        out["rs"] = D.reduce_scatter_rows(x).tolist()
        out["ag"] = D.all_gather_rows(torch.full((1, 2), float(rank))).tolist()
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert r["loss"] == pytest.approx(r["ref_loss"], rel=1e-12)
        assert r["sum_loss"] == pytest.approx(r["ref_sum_loss"], rel=1e-12)
        assert r["da_err"] < 1e-13 and r["db_err"] < 1e-13
        assert r["dpo_loss"] == pytest.approx(r["dpo_ref"], rel=1e-12)
        assert r["dpo_margin"] == pytest.approx(r["dpo_margin_ref"], rel=1e-12)
        assert r["dw"] == 3.0
        assert r["ag"] == [[0.0, 0.0], [1.0, 1.0]]
        base = np.arange(12, dtype=np.float64).reshape(4, 3) * 3  # (1 + 2) * x
        assert r["rs"] == base[r["rank"] * 2:(r["rank"] + 1) * 2].tolist()
