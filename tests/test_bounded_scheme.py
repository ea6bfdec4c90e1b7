"""CPU restatement (numpy, fp32 arithmetic) of the two shortcuts the kernels take for unit-norm rows, checked against
float64: (1) the column log-sum-exp built from the row pass's exponentials with a per-warp shift (gemm_lse.cu,
gemm_lse_kernel<true> + col_merge_kernel), (2) the backward's single shared exponential (sgg_f.cu, kShared).  These
tests pin the ARITHMETIC (no underflow, fp32 accuracy) on the host; the kernels themselves are compared with the
two-pass / two-exponential forms and with float64 torch in tests/test_gpu_parity.py."""
import numpy as np
import pytest

LOG2E = np.float32(1.4426950408889634)


def _unit_rows(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def _one_pass_column_lse(S, inv_tau, tile=256, group_cols=128):
    """S fp32 [rows][cols] of raw similarities.  Per 32-row group and 32-column chunk: exponentials relative to the
    row's running maximum (running over the chunks of ONE column group of ONE tile, as the epilogue warp does), rescaled
    to the group's largest running maximum, summed down the rows; groups merged with an online (max, sum)."""
    rows, cols = S.shape
    c = np.float32(inv_tau) * LOG2E
    n_groups = (rows + 31) // 32
    col_sum = np.zeros((n_groups, cols), np.float32)
    col_shift = np.full((n_groups, (cols + 31) // 32), -np.inf, np.float32)
    for g in range(n_groups):
        r0, r1 = g * 32, min(rows, g * 32 + 32)
        for t0 in range(0, cols, group_cols):        # one epilogue warp: a 128-column group of a tile
            run_m = np.full(r1 - r0, -np.inf, np.float32)
            for c0 in range(t0, min(cols, t0 + group_cols), 32):
                v = S[r0:r1, c0:c0 + 32]
                m_new = np.maximum(run_m, v.max(axis=1) * c)
                e = np.exp2(v * c - m_new[:, None]).astype(np.float32)
                sh = m_new.max()
                f = np.exp2(m_new - sh).astype(np.float32)
                col_sum[g, c0:c0 + v.shape[1]] = (e * f[:, None]).sum(axis=0, dtype=np.float32)
                col_shift[g, c0 // 32] = sh
                run_m = m_new
    M = col_shift.max(axis=0)                                             # per 32-column chunk
    w = np.exp2(col_shift - M[None]).astype(np.float32)                   # [groups][chunks]
    total = (col_sum * np.repeat(w, 32, axis=1)[:, :cols]).sum(axis=0, dtype=np.float32)
    return ((np.repeat(M, 32)[:cols] + np.log2(total)) / LOG2E).astype(np.float32), col_sum


@pytest.mark.parametrize("tau", [0.5, 0.07, 0.03])
def test_one_pass_column_lse_matches_float64(tau):
    rng = np.random.default_rng(3)
    rows, cols, d = 200, 700, 64
    b = _unit_rows(rng, cols, d)
    a = _unit_rows(rng, rows, d)
    a[:64] = b[100:164] * 0.9 + 0.1 * a[:64]           # near-positives: logits close to +1/tau next to ones near 0
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    S = (a @ b.T).astype(np.float32)
    assert 2.0 / tau * float(LOG2E) < 100.0            # the bound pgica_ntxent_fwd_bounded checks
    got, col_sum = _one_pass_column_lse(S, 1.0 / tau)
    want = np.log(np.exp(S.astype(np.float64) / tau).sum(axis=0))
    assert np.all(col_sum > 0)                         # nothing underflowed: every partial carries its column
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("tau", [0.5, 0.07])
def test_shared_exponential_gradient_matches_float64(tau):
    rng = np.random.default_rng(4)
    n, d = 160, 64
    b = _unit_rows(rng, n, d)
    a = _unit_rows(rng, n, d) * 0.3 + b * 0.7
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    S64 = (a.astype(np.float64) @ b.astype(np.float64).T) / tau
    lse_row = np.log(np.exp(S64).sum(axis=1))
    lse_col = np.log(np.exp(S64).sum(axis=0))
    want = np.exp(S64 - lse_row[:, None]) + np.exp(S64 - lse_col[None]) - 2.0 * np.eye(n)   # dS up to the scalar
    # the kernel's fp32 arithmetic: e = 2^(z c - c), row factor 2^(c - lse_row log2e), column factor likewise
    c = np.float32(1.0 / tau) * LOG2E
    z = (a @ b.T).astype(np.float32)
    e = np.exp2(z * c - c).astype(np.float32)
    rs = np.exp2(c - lse_row.astype(np.float32) * LOG2E).astype(np.float32)
    cs = np.exp2(c - lse_col.astype(np.float32) * LOG2E).astype(np.float32)
    assert np.isfinite(rs).all() and np.isfinite(cs).all() and (e > 0).all()
    got = e * (rs[:, None] + cs[None]) - 2.0 * np.eye(n, dtype=np.float32)
    # the two-exponential form the general kernel uses, same fp32 arithmetic
    two = (np.exp2(z * c - lse_row.astype(np.float32)[:, None] * LOG2E)
           + np.exp2(z * c - lse_col.astype(np.float32)[None] * LOG2E)).astype(np.float32) - 2.0 * np.eye(n, dtype=np.float32)
    err_shared, err_two = np.abs(got - want).max(), np.abs(two - want).max()
    # absolute error of a probability: fp32 rounding of the logit times 1/tau — the same class for both forms, and far
    # below the 8 bits the G tiles are stored with
    assert err_shared < 2e-5 * max(1.0, 1.0 / tau)
    assert err_shared < 3.0 * err_two + 1e-7
