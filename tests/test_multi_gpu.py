"""Multi-GPU parity (`-m gpu`; self-skips below 2 visible B200s): tools/dist_check.py under torchrun over NCCL —
global-negatives NT-Xent and the pair-sharded DPO head against the float64 oracle on the concatenated batch, and the
progress-gated peer-memory all-reduce of dW (csrc/peer_ar.cu beside the dual backward kernel) against NCCL."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dist_check_under_torchrun(cuda_device):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (single-GPU box)")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    checks = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert checks and all(c["ok"] for c in checks)
    assert {c["check"] for c in checks} >= {"global_ntxent", "dpo_sharded", "overlapped_dw_allreduce"}
