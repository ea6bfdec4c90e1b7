"""Cycle accounting of the dual softmax-gradient GEMM (sgg_f.cu), diagnostics build (-DPGICA_TRACE).

    python tools/trace_sgg.py build       # here: compile libpgica_trace.so
    python tools/trace_sggf.py [R,Cw]     # on the B200: cfg2 shape, per-role wait breakdown (kilo-cycles per CTA)
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "preference_guided_image_captioning_alignment_b200")
os.environ["PGICA_LIB_PATH"] = os.path.join(PKG, "libpgica_trace.so")
if len(sys.argv) > 1:
    os.environ["PGICA_SGGF_PLAN"] = sys.argv[1]
import torch

from preference_guided_image_captioning_alignment_b200 import _lib
from preference_guided_image_captioning_alignment_b200 import functional as F

PRODUCER = ["mma_issue", "mma_wait_zempty", "mma_wait_full", "-", "tma_issue", "tma_wait_empty", "xw_other",
            "xw_wait_stfull", "xw_wait_done_flag", "xw_store_fence", "epi_other", "epi_wait_zfull", "epi_compute",
            "epi_wait_stfree", "epi_stage_write"]
CONSUMER = ["mma_issue", "mma_wait_outfree", "mma_wait_gfull", "mma_wait_full", "tma_issue", "tma_wait_gempty",
            "tma_wait_ready_flag", "tma_wait_empty", "-", "-", "drain_other", "drain_wait_outfull", "drain_copy"]

lib = _lib.load()
lib.pgica_debug_set_sggf_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
torch.manual_seed(0)
mx, my, k = int(os.environ.get("MX", 4096)), int(os.environ.get("MY", 50257)), int(os.environ.get("K", 1024))
x = (torch.randn(mx, k, device=dev) * 0.5).to(torch.bfloat16)
y = (torch.randn(my, k, device=dev) * 0.02).to(torch.bfloat16)
row = (torch.full((mx,), 11.0, device=dev), torch.randn(mx, device=dev),
       torch.randint(0, my, (mx,), device=dev, dtype=torch.int32))
col = None
if os.environ.get("MODE") == "both":  # NT-Xent form: row and column term
    col = (torch.full((my,), 11.0, device=dev), torch.randn(my, device=dev),
           torch.randint(0, mx, (my,), device=dev, dtype=torch.int32))
OX = torch.float32 if os.environ.get("OX") == "f32" else torch.bfloat16
buf = torch.zeros(1024, 24, dtype=torch.int64, device=dev)
OY = torch.bfloat16 if os.environ.get('OY') == 'bf16' else torch.float32
for _ in range(2):
    F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col, out_x_dtype=OX, out_y_dtype=OY)
torch.cuda.synchronize()
assert lib.pgica_debug_set_sggf_trace(ctypes.c_void_p(buf.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
F.softmax_grad_gemm_dual(x, y, 1.0, row=row, col=col, out_x_dtype=OX, out_y_dtype=OY)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
t = buf.cpu().double()
S = k // 512
plan = os.environ.get("PGICA_SGGF_PLAN")
print(f"shape {mx}x{my}x{k} plan {plan}: {ms:.3f} ms")
if plan:
    R2, C2 = (int(v) for v in plan.split(","))
    nH, nW = 2 * R2 * S, 2 * C2 * S  # CTAs (the plan counts pairs)
    used = int((t.abs().sum(1) > 0).nonzero().max().item()) + 1
    groups = (("X-holders", t[:nH], CONSUMER), ("Y-holders", t[nH:nH + nW], CONSUMER),
              ("producers", t[nH + nW:used], PRODUCER))
    out = {"ms": ms, "plan": plan, "ctas": used}
    for name, rows, names in groups:
        lead = rows[0::2]  # leaders issue the MMAs; TMA / epilogue / drain columns are averaged over both CTAs
        mean = rows.mean(0) / 1e3
        mean[:4] = lead.mean(0)[:4] / 1e3
        d = {n: round(mean[i].item(), 1) for i, n in enumerate(names) if n != "-"}
        out[name] = d
        print(name, json.dumps(d))
    print(json.dumps(out))
