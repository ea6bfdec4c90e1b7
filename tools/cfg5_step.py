#!/usr/bin/env python
"""BASELINE config 5: end-to-end Stage-2 micro-step of the reference's own model (random-init CLIP ViT-B/32 + GPT-2
Medium cross-attention decoder, 867 M parameters) on one B200, unpatched vs with `install()` applied.

    python tools/cfg5_step.py [--batch 8] [--steps 5] [--warmup 2] [--json out.json]

A micro-step is what pkg/training/trainer.py:575-633 does per batch: two generation-mode forwards (preferred, rejected),
PreferenceLoss, loss.item(), backward, clip_grad_norm_, AdamW step, zero_grad.  Same model object, same batch, same
seeds for both arms; the patched arm differs only by `pg.install()` + `fuse_decoder` (loss names rebound, lm_head lazy).
Needs the reference package (baseline/_ref or /root/reference/src); test infrastructure, not product code.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(batch=8, steps=5, warmup=2, seq_len=128, log=print):
    import importlib

    import torch

    import preference_guided_image_captioning_alignment_b200 as pg
    from oracle import ref_model
    inst = importlib.import_module("preference_guided_image_captioning_alignment_b200.install")
    if not ref_model.available():
        return {"unavailable": "reference package not importable (no baseline/_ref, no /root/reference)"}
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    mm = ref_model.load_package()
    t0 = time.time()
    model = ref_model.build_model().to(dev)
    model.train()
    log(f"model built in {time.time() - t0:.1f} s: {sum(p.numel() for p in model.parameters()) / 1e6:.1f} M parameters")
    data = ref_model.stage2_batch(batch, seq_len=seq_len, vocab=50257, seed=11, device=dev)
    params = [p for p in model.parameters() if p.requires_grad]
    valid = int(data["preferred_mask"][:, 1:].sum() + data["rejected_mask"][:, 1:].sum())

    def arm(make_loss, label):
        opt = torch.optim.AdamW(params, lr=1e-6, weight_decay=0.01)
        pl = make_loss()
        losses, times = [], []
        torch.cuda.reset_peak_memory_stats()
        for it in range(warmup + steps):
            torch.manual_seed(1000 + it)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss, _, _ = ref_model.stage2_micro_step(model, pl, data)
            val = loss.item()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad()
            e1.record()
            torch.cuda.synchronize()
            if it >= warmup:
                times.append(e0.elapsed_time(e1))
            losses.append(val)
        times.sort()
        out = {"ms_per_step_median": times[len(times) // 2], "ms_per_step_min": times[0], "first_loss": losses[0],
               "last_loss": losses[-1], "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        log(label, json.dumps(out))
        del opt
        return out

    state = {k: v.clone() for k, v in model.state_dict().items()}
    ref = arm(lambda: mm.PreferenceLoss(0.1), "reference step:")
    model.load_state_dict(state)
    def eval_loss(make_loss):  # dropout off: the two arms must agree (the fused cross-attention draws its own masks)
        model.eval()
        with torch.no_grad():
            val = ref_model.stage2_micro_step(model, make_loss(), data)[0].item()
        model.train()
        return val

    ref_eval = eval_loss(lambda: mm.PreferenceLoss(0.1))
    try:
        pg.install()
        pg.fuse_model(model)   # lazy LM head + one-key cross-attention + LayerNorm/normalise tails
        ours_eval = eval_loss(lambda: mm.PreferenceLoss(0.1))
        ours = arm(lambda: mm.PreferenceLoss(0.1), "patched step:  ")
    finally:
        pg.unfuse_model(model)
        pg.uninstall()
    return {"workload": f"cfg5: Stage-2 micro-step (2 forwards, PreferenceLoss, backward, clip, AdamW), batch {batch}, "
                        f"seq {seq_len}, right-padded captions U[10,20] ({valid} scored rows of {2 * batch * seq_len}), "
                        "fp32, TF32 off, random-init 867 M reference model",
            "reference": ref, "patched": ours,
            "speedup": ref["ms_per_step_median"] / ours["ms_per_step_median"],
            "patched_with": "install() + fuse_model(): lazy LM head + compacted pair-stacked PreferenceLoss, one-key "
                            "cross-attention + LayerNorm fused, LayerNorm + L2-normalise tails fused",
            "eval_loss_reference": ref_eval, "eval_loss_patched": ours_eval,
            "first_loss_rel_diff": abs(ours_eval - ref_eval) / abs(ref_eval)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = run(a.batch, a.steps, a.warmup, log=lambda *x: print(*x, file=sys.stderr, flush=True))
    print(json.dumps(res))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)
