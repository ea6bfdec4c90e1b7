"""Multi-GPU parity check (run under torchrun on N >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

* NT-Xent with global negatives (distributed.global_ntxent: all-gather, kernels, column-LSE merge, reduce-scatter)
  against the float64 oracle applied to the concatenated global batch.
* DPO head with the pair batch sharded over ranks (global-mean loss, all-reduced dW) against the float64 oracle on
  the full batch.
Prints one JSON line per check on rank 0; exit code 1 if any check fails."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import closed_form as cf


KEYS = ("dpo_loss", "reward_margin", "reward_accuracy", "policy_chosen_logprob", "policy_rejected_logprob")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import preference_guided_image_captioning_alignment_b200 as pg
    from preference_guided_image_captioning_alignment_b200 import distributed as D
    ok = True

    # ---------------------------------------------------------------- NT-Xent, global negatives
    for nb, Dm, tau in ((96, 512, 0.5), (300, 256, 0.2)):
        B = nb * world
        gen = torch.Generator().manual_seed(11)
        A = torch.nn.functional.normalize(torch.randn(B, Dm, generator=gen), dim=-1).to(torch.bfloat16)
        Bm = torch.nn.functional.normalize(A.float() + 0.4 * torch.randn(B, Dm, generator=gen), dim=-1).to(torch.bfloat16)
        a = A[rank * nb:(rank + 1) * nb].to(dev).requires_grad_(True)
        b = Bm[rank * nb:(rank + 1) * nb].to(dev).requires_grad_(True)
        loss = D.global_ntxent(a, b, tau)
        loss.backward()
        ref = cf.ntxent(A.double().numpy(), Bm.double().numpy(), tau)
        da = [torch.empty_like(a.grad) for _ in range(world)]
        db = [torch.empty_like(b.grad) for _ in range(world)]
        dist.all_gather(da, a.grad)
        dist.all_gather(db, b.grad)
        res = {"check": "global_ntxent", "world": world, "rows_per_rank": nb, "dim": Dm, "tau": tau,
               "loss": loss.item(), "loss_rel": abs(loss.item() - ref["loss"]) / abs(ref["loss"]),
               "da_rel": rel(torch.cat(da).float().cpu().numpy(), ref["dx"]),
               "db_rel": rel(torch.cat(db).float().cpu().numpy(), ref["dy"])}
        res["ok"] = bool(res["loss_rel"] < 1e-4 and res["da_rel"] < 1e-2 and res["db_rel"] < 1e-2)
        ok &= res["ok"]
        if rank == 0:
            print(json.dumps(res), flush=True)

    # ---------------------------------------------------------------- DPO head, pairs sharded
    n_global, T, d, V, beta = 2 * world, 24, 256, 3001, 0.1
    gen = torch.Generator().manual_seed(5)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=gen) * sc).to(torch.bfloat16)
    W, Wr = r(V, d, sc=0.05), r(V, d, sc=0.05)
    hc, hr, rhc, rhr = r(n_global, T, d), r(n_global, T, d), r(n_global, T, d), r(n_global, T, d)
    yc = torch.randint(0, V, (n_global, T), generator=gen)
    yr = torch.randint(0, V, (n_global, T), generator=gen)
    lens = torch.randint(T // 2, T + 1, (n_global,), generator=gen)
    m = (torch.arange(T)[None, :] < lens[:, None]).long()
    lo, hi = D.shard_pairs(n_global, rank, world)
    sl = slice(lo, hi)
    Wg = W.to(dev).requires_grad_(True)
    hcg, hrg = hc[sl].to(dev).requires_grad_(True), hr[sl].to(dev).requires_grad_(True)
    head = pg.FusedDPOHead(beta=beta)
    hstack = torch.cat([hcg, hrg])
    hcg.retain_grad()
    loss, metrics = head.forward_stacked(hstack, Wg, torch.cat([yc[sl], yr[sl]]).to(dev), torch.cat([m[sl], m[sl]]).to(dev),
                                         torch.cat([rhc[sl], rhr[sl]]).to(dev), Wr.to(dev), n_global)
    loss.backward()
    D.allreduce_dweight(Wg.grad)
    gl, gm = D.allreduce_scalars(loss, metrics)
    o = cf.dpo_head(hc.double().numpy(), hr.double().numpy(), W.double().numpy(), yc.numpy(), yr.numpy(), m.numpy(),
                    m.numpy(), dict(hc=rhc.double().numpy(), hr=rhr.double().numpy(), W=Wr.double().numpy()), beta)
    dh = [torch.empty_like(hcg.grad) for _ in range(world)]
    dist.all_gather(dh, hcg.grad)
    res = {"check": "dpo_sharded", "world": world, "pairs": n_global, "loss": gl.item(),
           "loss_rel": abs(gl.item() - o["loss"]) / abs(o["loss"]),
           "dw_rel": rel(Wg.grad.float().cpu().numpy(), o["dW"]),
           "dhc_rel": rel(torch.cat(dh).float().cpu().numpy(), o["dhc"]),
           "metrics_maxabs": float(np.max(np.abs(gm.cpu().numpy() - np.asarray([o["metrics"][k] for k in KEYS]))))}
    res["ok"] = bool(res["loss_rel"] < 1e-4 and res["dw_rel"] < 1e-2 and res["dhc_rel"] < 1e-2 and res["metrics_maxabs"] < 1e-3)
    ok &= res["ok"]
    if rank == 0:
        print(json.dumps(res), flush=True)
    # ---------------------------------------------------------------- overlapped peer all-reduce of dW == NCCL all-reduce
    from preference_guided_image_captioning_alignment_b200 import distributed as D2
    from preference_guided_image_captioning_alignment_b200 import functional as Fn
    for (B2, T2, d2, V2, nseg) in ((4, 48, 512, 3001, 3), (6, 64, 1024, 5003, 8), (16, 128, 1024, 50257, 8)):
        gen2 = torch.Generator().manual_seed(200 + rank)
        gw = torch.Generator().manual_seed(7)
        W2 = (torch.randn(V2, d2, generator=gw) * 0.05).to(torch.bfloat16).to(dev)
        H2 = torch.randn(B2, T2, d2, generator=gen2).to(torch.bfloat16).to(dev)
        y2 = torch.randint(0, V2, (B2, T2), generator=gen2).to(dev)
        m2 = torch.ones(B2, T2, dtype=torch.long, device=dev)
        gs = torch.randn(B2, generator=gen2).to(dev)
        _, lse2, _, rl2, rw2, _ = Fn.lmhead_logprob_fwd(H2, W2, y2, m2, False)
        dh_ref, dw_ref = Fn.lmhead_logprob_bwd(H2, W2, rl2, rw2, lse2, gs, False)
        dw_ref = dw_ref.clone()
        dist.all_reduce(dw_ref)
        for multicast in (False, None):  # unicast peer loads, then the NVSwitch multicast path where the fabric has one
            red = D2.OverlappedDWAllReduce(V2, d2, dev, segments=nseg, multicast=multicast)
            if multicast is None and not red.multicast_ptr:
                if rank == 0:
                    print(json.dumps({"check": "overlapped_dw_allreduce", "multicast": "not available"}), flush=True)
                del red
                continue
            for trial in range(3):
                scal = torch.arange(6, device=dev, dtype=torch.float32) + rank + trial
                dh2, dw2, done, scal_sum = red.backward(H2, W2, rl2, rw2, lse2, gs, False, scalars=scal)
                torch.cuda.current_stream().wait_event(done)
                torch.cuda.synchronize()
                e_w = rel(dw2.cpu().numpy(), dw_ref.cpu().numpy())
                e_h = rel(dh2.float().cpu().numpy(), dh_ref.float().cpu().numpy())
                want_scal = sum(torch.arange(6, dtype=torch.float32) + q + trial for q in range(world))
                e_s = float((scal_sum.cpu() - want_scal).abs().max())
                # every rank must hold bit-identical sums (fixed rank order inside the kernel)
                chk = dw2.double().sum().reshape(1)
                allchk = [torch.empty_like(chk) for _ in range(world)]
                dist.all_gather(allchk, chk)
                same = all(bool(c.item() == allchk[0].item()) for c in allchk)
                res = {"check": "overlapped_dw_allreduce", "multicast": bool(red.multicast_ptr), "world": world,
                       "shape": [B2, T2, d2, V2], "segments": red.nseg, "trial": trial, "dw_rel": e_w, "dh_rel": e_h, "scalars_maxabs": e_s, "identical_on_all_ranks": same,
                       "ok": bool(e_w < 1e-5 and e_h < 1e-5 and e_s < 1e-4 and same)}
                ok &= res["ok"]
                if rank == 0:
                    print(json.dumps(res), flush=True)
            del red
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
