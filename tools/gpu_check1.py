"""Bring-up checks for the tcgen05 building blocks (run on a B200 through gpurun).

Each case runs in its own subprocess so a faulting descriptor cannot poison later cases.
Usage: python tools/gpu_check1.py            (driver: runs all cases)
       python tools/gpu_check1.py CASE ARGS  (one case)
"""
import ctypes
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def case_probe(n, k, b_mn, a_manual, lbo, sbo):
    import torch
    from preference_guided_image_captioning_alignment_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    a = torch.randn(128, k, device="cuda").to(torch.bfloat16)
    if b_mn:
        b = torch.randn(k, n, device="cuda").to(torch.bfloat16)
        ref = a.float() @ b.float()
    else:
        b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        ref = a.float() @ b.float().t()
    d = torch.full((128, n), float("nan"), device="cuda")
    rc = lib.pgica_probe_umma(_ptr(a), _ptr(b), n, k, b_mn, a_manual, lbo, sbo, _ptr(d), None)
    _lib.check(rc)
    torch.cuda.synchronize()
    err = (d - ref).abs().max().item()
    print(json.dumps({"case": "probe", "n": n, "k": k, "b_mn": b_mn, "a_manual": a_manual, "lbo": lbo, "sbo": sbo,
                      "max_abs_err": err, "ref_absmax": ref.abs().max().item(), "ok": bool(err < 1e-2)}))


def gemm_lse(lib, a, b, scale, labels=None, diag_offset=0):
    import torch
    rows, k = a.shape
    cols = b.shape[0]
    need = ctypes.c_size_t(0)
    from preference_guided_image_captioning_alignment_b200 import _lib
    _lib.check(lib.pgica_gemm_lse_workspace_bytes(rows, cols, k, ctypes.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    lse = torch.empty(rows, device="cuda")
    tgt = torch.empty(rows, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.pgica_gemm_lse(_ptr(a), _ptr(b), rows, cols, k, scale, _ptr(labels) if labels is not None else None,
                            diag_offset, _ptr(lse), _ptr(tgt), _ptr(ws), need.value, ctypes.c_void_p(st))
    _lib.check(rc)
    return lse, tgt, ws


def case_lse(rows, cols, k, scale, time_it):
    import torch
    from preference_guided_image_captioning_alignment_b200 import _lib
    lib = _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    a = torch.randn(rows, k, device="cuda").to(torch.bfloat16)
    b = (torch.randn(cols, k, device="cuda") * (0.02 if cols > 4096 else 0.2)).to(torch.bfloat16)
    labels = torch.randint(0, cols, (rows,), device="cuda", dtype=torch.int32)
    labels[::7] = -1
    lse, tgt, _ = gemm_lse(lib, a, b, scale, labels)
    torch.cuda.synchronize()
    errs = []
    chunk = 512
    for r0 in range(0, rows, chunk):
        z = (a[r0:r0 + chunk].float() @ b.float().t()) * scale
        ref_lse = torch.logsumexp(z.double(), dim=-1)
        lab = labels[r0:r0 + chunk].long()
        ref_t = torch.where(lab >= 0, z.gather(1, lab.clamp(min=0)[:, None])[:, 0], torch.zeros_like(z[:, 0]))
        errs.append(((lse[r0:r0 + chunk].double() - ref_lse).abs().max().item(),
                     (tgt[r0:r0 + chunk] - ref_t).abs().max().item()))
    e_lse = max(e[0] for e in errs)
    e_tgt = max(e[1] for e in errs)
    out = {"case": "lse", "rows": rows, "cols": cols, "k": k, "scale": scale, "lse_err": e_lse, "tgt_err": e_tgt,
           "ok": bool(e_lse < 2e-4 and e_tgt < 2e-4)}
    if time_it:
        for _ in range(3):
            gemm_lse(lib, a, b, scale, labels)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        iters = 20
        ev[0].record()
        for _ in range(iters):
            gemm_lse(lib, a, b, scale, labels)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / iters
        out["ms"] = ms
        out["tflops"] = 2.0 * rows * cols * k / ms / 1e9
    print(json.dumps(out))


def driver():
    cases = []
    for n, k in [(64, 64), (128, 64), (256, 64), (256, 256), (192, 128)]:
        cases.append(["probe", n, k, 0, 0, 0, 0])
    cases.append(["probe", 256, 128, 0, 1, 0, 0])
    for n, k in [(64, 64), (256, 64), (256, 128), (128, 256)]:
        cases.append(["probe", n, k, 1, 0, k * 128, 1024])
        cases.append(["probe", n, k, 1, 0, 1024, k * 128])
    cases += [["lse", 128, 256, 64, 1.0, 0], ["lse", 200, 1000, 512, 2.0, 0], ["lse", 64, 64, 512, 2.0, 0],
              ["lse", 1000, 50257, 1024, 1.0, 0], ["lse", 4064, 50257, 1024, 1.0, 1],
              ["lse", 4096, 32768, 512, 2.0, 1], ["lse", 32704, 50260, 1024, 1.0, 1]]
    t0 = time.time()
    for c in cases:
        cmd = [sys.executable, os.path.abspath(__file__)] + [str(x) for x in c]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            if r.returncode != 0:
                print(json.dumps({"case": c, "rc": r.returncode, "stderr": r.stderr[-600:], "stdout": r.stdout[-300:]}))
            else:
                print(tail)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": c, "timeout": True}))
        sys.stdout.flush()
    print("elapsed", time.time() - t0)


if __name__ == "__main__":
    if len(sys.argv) == 1:
        driver()
    elif sys.argv[1] == "probe":
        case_probe(*[int(x) for x in sys.argv[2:8]])
    elif sys.argv[1] == "lse":
        case_lse(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6]))
