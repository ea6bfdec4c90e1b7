"""torch.library registration of the fused loss-head ops (namespace `pgica`).

Each op is a thin shim over functional.py (-> C ABI -> CUDA kernels); autograd formulas call the backward
kernels.  Fake (meta) implementations make the ops traceable; there is deliberately no CPU implementation:
calling an op on CPU tensors raises.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import functional as F

_lib_def = torch.library.Library("pgica", "DEF")  # noqa: F841  (keeps the namespace alive)


def _f32(n, like):
    return torch.empty(n, dtype=torch.float32, device=like.device)


def _grad_dtype(t):
    return torch.bfloat16 if t.dtype == torch.bfloat16 else torch.float32


# ============================================================================================ NT-Xent
@torch.library.custom_op("pgica::ntxent", mutates_args=())
def ntxent(a: Tensor, b: Tensor, inv_tau: float, reduce_mean: bool, bounded: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """Symmetric NT-Xent on (B, D) x (B, D) embeddings as given (no normalisation here).
    bounded: the caller promises unit-norm rows — one pass over the similarity tiles gives both log-sum-exps.
    Returns (loss[], lse_row[B], lse_col[B])."""
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError(f"ntxent expects two (B, D) tensors of equal shape, got {tuple(a.shape)} {tuple(b.shape)}")
    n = a.shape[0]
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        lse_row, diag, lse_col = F.ntxent_fwd(a.contiguous(), b.contiguous(), inv_tau, 0, bounded=bounded)
    else:  # fp32 embeddings: two-term bf16 split, one GEMM of depth 3*D per direction (as ntxent_cosine does)
        al, ar = F.split3(a)
        bl, br = F.split3(b)
        if bounded:  # a_left3 . b_right3 and b_left3 . a_right3 are the same three products: one pass serves both
            lse_row, diag, lse_col = F.ntxent_fwd(al, br, inv_tau, 0, bounded=True)
        else:
            lse_row, diag = F.gemm_lse(al, br, inv_tau, None, 0)
            lse_col, _ = F.gemm_lse(bl, ar, inv_tau, None, 0, want_tgt=False)
    loss = F.ntxent_loss(lse_row, diag, lse_col, 1.0 / n if reduce_mean else 1.0)
    return loss, lse_row, lse_col


@ntxent.register_fake
def _(a, b, inv_tau, reduce_mean, bounded=False):
    return a.new_empty((), dtype=torch.float32), _f32(a.shape[0], a), _f32(b.shape[0], a)


@torch.library.custom_op("pgica::ntxent_small", mutates_args=())
def ntxent_small(a: Tensor, b: Tensor, inv_tau: float, reduce_mean: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Small-batch form of `ntxent` (B <= 128, D in {128, 256, 384, 512}): loss, both LSE vectors and the gradients
    of the loss w.r.t. a and b from ONE single-CTA launch (csrc/ntxent_small.cu)."""
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError(f"ntxent expects two (B, D) tensors of equal shape, got {tuple(a.shape)} {tuple(b.shape)}")
    return _ntxent_small_impl(a, b, inv_tau, reduce_mean)


def _ntxent_small_impl(a, b, inv_tau, reduce_mean):
    """bf16 inputs are the tensor-core operands as they are; anything wider (the trainer hands over fp32 unit vectors,
    pkg/models/model.py:826-829) goes in as its two-term bf16 split so the loss keeps the fp32 accuracy the reference
    computes it with (similarity of depth 3*D; the gradient products use the hi parts)."""
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        return F.ntxent_small(a.contiguous(), b.contiguous(), inv_tau, reduce_mean)
    al3, _ = F.split3(a)
    _, br3 = F.split3(b)
    return F.ntxent_small_split(al3, br3, a.shape[1], inv_tau, reduce_mean)


@ntxent_small.register_fake
def _(a, b, inv_tau, reduce_mean):
    return (a.new_empty((), dtype=torch.float32), _f32(a.shape[0], a), _f32(b.shape[0], a),
            torch.empty_like(a, dtype=torch.float32), torch.empty_like(b, dtype=torch.float32))


def _ntxent_small_setup(ctx, inputs, output):
    ctx.save_for_backward(output[3], output[4])
    ctx.dtypes = (inputs[0].dtype, inputs[1].dtype)


def _ntxent_small_backward(ctx, g_loss, g_lr, g_lc, g_da, g_db):
    da, db = ctx.saved_tensors  # gradients of the loss itself: scale by the upstream scalar
    return (da * g_loss).to(ctx.dtypes[0]), (db * g_loss).to(ctx.dtypes[1]), None, None


ntxent_small.register_autograd(_ntxent_small_backward, setup_context=_ntxent_small_setup)


class _NTXentSmallEager(torch.autograd.Function):
    """Eager-mode twin of the `pgica::ntxent_small` op: the same kernel behind a plain autograd.Function.  At B = 8 the
    step is ~10 us of device time; the dispatcher and schema machinery of a torch.library op would more than double
    the host time around it (the registered op remains what torch.compile / opcheck see)."""

    @staticmethod
    def forward(ctx, a, b, inv_tau, reduce_mean):
        loss, lse_row, lse_col, da, db = _ntxent_small_impl(a, b, inv_tau, reduce_mean)
        ctx.save_for_backward(da, db)
        ctx.dtypes = (a.dtype, b.dtype)
        ctx.mark_non_differentiable(lse_row, lse_col)
        return loss, lse_row, lse_col

    @staticmethod
    def backward(ctx, g_loss, g_lr, g_lc):
        da, db = ctx.saved_tensors
        return (da * g_loss).to(ctx.dtypes[0]), (db * g_loss).to(ctx.dtypes[1]), None, None


def ntxent_auto(a: Tensor, b: Tensor, inv_tau: float, reduce_mean: bool, bounded: bool = False):
    """`ntxent_small` when the batch fits one CTA, else the general kernels (bounded: unit-norm rows promised, see
    `ntxent`).  -> (loss, lse_row, lse_col)"""
    if a.dim() == 2 and a.shape == b.shape and a.is_cuda and F.ntxent_small_supported(a.shape[0], a.shape[1]):
        if torch.compiler.is_compiling():
            return ntxent_small(a, b, inv_tau, reduce_mean)[:3]
        return _NTXentSmallEager.apply(a, b, inv_tau, reduce_mean)
    return ntxent(a, b, inv_tau, reduce_mean, bounded)


@torch.library.custom_op("pgica::ntxent_bwd", mutates_args=())
def ntxent_bwd(a: Tensor, b: Tensor, lse_row: Tensor, lse_col: Tensor, grad_loss: Tensor, inv_tau: float,
               reduce_mean: bool, bounded: bool = False) -> Tuple[Tensor, Tensor]:
    n, d = a.shape
    mult = 1.0 / (2.0 * n) if reduce_mean else 0.5
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        return F.ntxent_bwd(a.contiguous(), b.contiguous(), inv_tau, 0, lse_row, lse_col, grad_loss, mult,
                            da_dtype=torch.bfloat16, db_dtype=torch.bfloat16, bounded=bounded)
    # the backward must recompute exactly the logits the forward saw: same split operands, depth 3*D; columns
    # [0, D) and [2D, 3D) of each product are G.hi and G.lo of the other side
    al, ar = F.split3(a)
    bl, br = F.split3(b)
    g3a, _ = F.ntxent_bwd(al, br, inv_tau, 0, lse_row, lse_col, grad_loss, mult, need_db=False)
    _, g3b = F.ntxent_bwd(ar, bl, inv_tau, 0, lse_row, lse_col, grad_loss, mult, need_da=False)
    return g3a[:, :d] + g3a[:, 2 * d:], g3b[:, :d] + g3b[:, 2 * d:]


@ntxent_bwd.register_fake
def _(a, b, lse_row, lse_col, grad_loss, inv_tau, reduce_mean, bounded=False):
    dt = torch.bfloat16 if (a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16) else torch.float32
    return torch.empty_like(a, dtype=dt), torch.empty_like(b, dtype=dt)


def _ntxent_setup(ctx, inputs, output):
    a, b, inv_tau, reduce_mean = inputs[:4]
    _, lse_row, lse_col = output
    ctx.save_for_backward(a, b, lse_row, lse_col)
    ctx.inv_tau, ctx.reduce_mean = inv_tau, reduce_mean
    ctx.bounded = bool(inputs[4]) if len(inputs) > 4 else False


def _ntxent_backward(ctx, g_loss, g_lr, g_lc):
    a, b, lse_row, lse_col = ctx.saved_tensors
    da, db = ntxent_bwd(a, b, lse_row, lse_col, g_loss.contiguous(), ctx.inv_tau, ctx.reduce_mean, ctx.bounded)
    return da.to(a.dtype), db.to(b.dtype), None, None, None


ntxent.register_autograd(_ntxent_backward, setup_context=_ntxent_setup)


# ----------------------------------------------------------------- NT-Xent on cosine similarity (normalise inside)
@torch.library.custom_op("pgica::ntxent_cosine", mutates_args=())
def ntxent_cosine(x: Tensor, y: Tensor, inv_tau: float, reduce_mean: bool,
                  eps: float) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """components.ContrastiveLoss arithmetic: L2-normalise both inputs, then symmetric NT-Xent.  The unit vectors are
    not bf16-representable, so the similarity uses their two-term bf16 split (one GEMM of depth 3*D): the loss keeps
    fp32-level accuracy, and the backward recomputes exactly the same logits from the same split operands.
    -> (loss, lse_row, lse_col, x_left3, x_right3, y_left3, y_right3, inv_x, inv_y)."""
    if x.shape != y.shape or x.dim() != 2:
        raise ValueError(f"ntxent_cosine expects two (B, D) tensors of equal shape, got {tuple(x.shape)} {tuple(y.shape)}")
    _, inv_x, _, xl, xr = F.rownorm_fwd(x, eps, split=True)
    _, inv_y, _, yl, yr = F.rownorm_fwd(y, eps, split=True)
    # the rows were normalised right here: logits are bounded by inv_tau, so one pass gives both log-sum-exps
    lse_row, diag, lse_col = F.ntxent_fwd(xl, yr, inv_tau, 0, bounded=True)
    n = x.shape[0]
    loss = F.ntxent_loss(lse_row, diag, lse_col, 1.0 / n if reduce_mean else 1.0)
    return loss, lse_row, lse_col, xl, xr, yl, yr, inv_x, inv_y


@ntxent_cosine.register_fake
def _(x, y, inv_tau, reduce_mean, eps):
    n, d = x.shape
    s3 = lambda: torch.empty((n, 3 * d), dtype=torch.bfloat16, device=x.device)
    return (x.new_empty((), dtype=torch.float32), _f32(n, x), _f32(n, x), s3(), s3(), s3(), s3(), _f32(n, x),
            _f32(n, x))


@torch.library.custom_op("pgica::ntxent_cosine_bwd", mutates_args=())
def ntxent_cosine_bwd(x: Tensor, y: Tensor, xl: Tensor, xr: Tensor, yl: Tensor, yr: Tensor, inv_x: Tensor,
                      inv_y: Tensor, lse_row: Tensor, lse_col: Tensor, grad_loss: Tensor, inv_tau: float,
                      reduce_mean: bool) -> Tuple[Tensor, Tensor]:
    """Backward of ntxent_cosine: two softmax-gradient GEMMs over the split operands (depth 3*D; columns [0, D) and
    [2D, 3D) of each result are G.hi and G.lo of the other side), then the normalisation backward."""
    n, d = x.shape
    mult = 1.0 / (2.0 * n) if reduce_mean else 0.5
    g3x, _ = F.ntxent_bwd(xl, yr, inv_tau, 0, lse_row, lse_col, grad_loss, mult, need_db=False)
    _, g3y = F.ntxent_bwd(xr, yl, inv_tau, 0, lse_row, lse_col, grad_loss, mult, need_da=False)
    xx = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
    yy = y if y.dtype in (torch.float32, torch.bfloat16) else y.float()
    return F.rownorm_bwd(xx.contiguous(), inv_x, g3x, second=2 * d), F.rownorm_bwd(yy.contiguous(), inv_y, g3y,
                                                                                     second=2 * d)


@ntxent_cosine_bwd.register_fake
def _(x, y, xl, xr, yl, yr, inv_x, inv_y, lse_row, lse_col, grad_loss, inv_tau, reduce_mean):
    return torch.empty_like(x, dtype=torch.float32), torch.empty_like(y, dtype=torch.float32)


def _ntxc_setup(ctx, inputs, output):
    x, y, inv_tau, reduce_mean, eps = inputs
    _, lse_row, lse_col, xl, xr, yl, yr, inv_x, inv_y = output
    ctx.save_for_backward(x, y, xl, xr, yl, yr, inv_x, inv_y, lse_row, lse_col)
    ctx.inv_tau, ctx.reduce_mean = inv_tau, reduce_mean


def _ntxc_backward(ctx, g_loss, *unused):
    x, y, xl, xr, yl, yr, inv_x, inv_y, lse_row, lse_col = ctx.saved_tensors
    dx, dy = ntxent_cosine_bwd(x, y, xl, xr, yl, yr, inv_x, inv_y, lse_row, lse_col, g_loss.contiguous(), ctx.inv_tau,
                               ctx.reduce_mean)
    return dx.to(x.dtype), dy.to(y.dtype), None, None, None


ntxent_cosine.register_autograd(_ntxc_backward, setup_context=_ntxc_setup)


# ------------------------------------------------------------------------------------- L2 normalisation
@torch.library.custom_op("pgica::l2_normalize", mutates_args=())
def l2_normalize(x: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    """F.normalize(x, dim=-1) -> (bf16 unit rows, 1/max(||x||, eps))."""
    y, inv, _ = F.rownorm_fwd(x, eps)
    return y, inv


@l2_normalize.register_fake
def _(x, eps):
    return torch.empty_like(x, dtype=torch.bfloat16), _f32(x.shape[0], x)


@torch.library.custom_op("pgica::l2_normalize_bwd", mutates_args=())
def l2_normalize_bwd(x: Tensor, inv_norm: Tensor, g: Tensor) -> Tensor:
    xx = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
    gg = g if g.dtype in (torch.float32, torch.bfloat16) else g.float()
    return F.rownorm_bwd(xx.contiguous(), inv_norm, gg)


@l2_normalize_bwd.register_fake
def _(x, inv_norm, g):
    return torch.empty_like(x, dtype=torch.float32)


def _l2_setup(ctx, inputs, output):
    x, _ = inputs
    ctx.save_for_backward(x, output[1])


def _l2_backward(ctx, g_y, g_inv):
    x, inv = ctx.saved_tensors
    return l2_normalize_bwd(x, inv, g_y).to(x.dtype), None


l2_normalize.register_autograd(_l2_backward, setup_context=_l2_setup)


# ============================================================================================ LM head
@torch.library.custom_op("pgica::lmhead_seq_logprob", mutates_args=())
def lmhead_seq_logprob(hidden: Tensor, weight: Tensor, labels: Tensor, mask: Optional[Tensor],
                       length_normalize: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """hidden (nseq, T, d), weight (V, d), labels (nseq, T) -> (seq_logp[nseq], lse[nseq*T], ztgt[nseq*T],
    row_label[nseq*T] int32, row_weight[nseq*T]).  The (nseq, T, V) logits are never formed."""
    if hidden.dim() != 3 or weight.dim() != 2 or hidden.shape[-1] != weight.shape[-1]:
        raise ValueError("lmhead_seq_logprob: hidden (nseq,T,d) and weight (V,d) expected")
    if tuple(labels.shape) != tuple(hidden.shape[:2]):
        raise ValueError("lmhead_seq_logprob: labels must be (nseq, T)")
    hb, wb = F.as_bf16(hidden), cached_bf16(weight)
    seq, lse, ztgt, rl, rw, _ = F.lmhead_logprob_fwd(hb, wb, labels, mask, length_normalize)
    return seq, lse, ztgt, rl, rw


@lmhead_seq_logprob.register_fake
def _(hidden, weight, labels, mask, length_normalize):
    n = hidden.shape[0] * hidden.shape[1]
    return (_f32(hidden.shape[0], hidden), _f32(n, hidden), _f32(n, hidden),
            torch.empty(n, dtype=torch.int32, device=hidden.device), _f32(n, hidden))


@torch.library.custom_op("pgica::lmhead_seq_logprob_bwd", mutates_args=())
def lmhead_seq_logprob_bwd(hidden: Tensor, weight: Tensor, row_label: Tensor, row_weight: Tensor, lse: Tensor,
                           grad_seq: Tensor, length_normalize: bool, need_dhidden: bool,
                           need_dweight: bool) -> Tuple[Tensor, Tensor]:
    hb, wb = F.as_bf16(hidden), cached_bf16(weight)
    dh, dw = F.lmhead_logprob_bwd(hb, wb, row_label, row_weight, lse, grad_seq, length_normalize,
                                  need_dhidden=need_dhidden, need_dweight=need_dweight,
                                  dhidden_dtype=_grad_dtype(hidden), dweight_dtype=_grad_dtype(weight))
    if dh is None:
        dh = hidden.new_empty(0)
    if dw is None:
        dw = weight.new_empty(0)
    return dh, dw


@lmhead_seq_logprob_bwd.register_fake
def _(hidden, weight, row_label, row_weight, lse, grad_seq, length_normalize, need_dhidden, need_dweight):
    dh = torch.empty_like(hidden, dtype=_grad_dtype(hidden)) if need_dhidden else hidden.new_empty(0)
    dw = torch.empty_like(weight, dtype=_grad_dtype(weight)) if need_dweight else weight.new_empty(0)
    return dh, dw


def _lm_setup(ctx, inputs, output):
    hidden, weight, labels, mask, length_normalize = inputs
    _, lse, _, row_label, row_weight = output
    ctx.save_for_backward(hidden, weight, row_label, row_weight, lse)
    ctx.length_normalize = length_normalize


def _lm_backward(ctx, g_seq, g_lse, g_zt, g_rl, g_rw):
    hidden, weight, row_label, row_weight, lse = ctx.saved_tensors
    need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    dh, dw = lmhead_seq_logprob_bwd(hidden, weight, row_label, row_weight, lse, g_seq.contiguous(),
                                    ctx.length_normalize, need_h, need_w)
    return (dh.to(hidden.dtype) if need_h else None, dw.to(weight.dtype) if need_w else None, None, None, None)


lmhead_seq_logprob.register_autograd(_lm_backward, setup_context=_lm_setup)


# ==================================================================================== materialised logits
@torch.library.custom_op("pgica::logits_seq_logprob", mutates_args=())
def logits_seq_logprob(logits: Tensor, labels: Tensor, mask: Optional[Tensor],
                       length_normalize: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(nseq, T, V) logits -> (seq_logp[nseq], lse[nseq*T], row_label, row_weight); streaming, HBM-bound."""
    if logits.dim() != 3:
        raise ValueError("logits_seq_logprob: logits must be (nseq, T, V)")
    lg = logits if logits.dtype in (torch.float32, torch.bfloat16) else logits.float()
    lg = lg.contiguous()
    nseq, T, V = lg.shape
    rl, rw = F.prep_rows(labels, mask, V)
    lse, zt = F.logits_lse(lg, rl)
    seq = F.seq_reduce(lse, zt, rw, nseq, T, length_normalize)
    return seq, lse, rl, rw


@logits_seq_logprob.register_fake
def _(logits, labels, mask, length_normalize):
    n = logits.shape[0] * logits.shape[1]
    return (_f32(logits.shape[0], logits), _f32(n, logits),
            torch.empty(n, dtype=torch.int32, device=logits.device), _f32(n, logits))


@torch.library.custom_op("pgica::logits_seq_logprob_bwd", mutates_args=())
def logits_seq_logprob_bwd(logits: Tensor, row_label: Tensor, row_weight: Tensor, lse: Tensor, grad_seq: Tensor,
                           length_normalize: bool) -> Tensor:
    lg = logits if logits.dtype in (torch.float32, torch.bfloat16) else logits.float()
    lg = lg.contiguous()
    nseq, T, _ = lg.shape
    coef = F.row_coef(grad_seq, row_weight, nseq, T, length_normalize, 1.0)
    return F.logits_grad(lg, row_label, lse, coef)


@logits_seq_logprob_bwd.register_fake
def _(logits, row_label, row_weight, lse, grad_seq, length_normalize):
    return torch.empty_like(logits, dtype=logits.dtype if logits.dtype == torch.bfloat16 else torch.float32)


def _lg_setup(ctx, inputs, output):
    logits, labels, mask, length_normalize = inputs
    _, lse, row_label, row_weight = output
    ctx.save_for_backward(logits, row_label, row_weight, lse)
    ctx.length_normalize = length_normalize


def _lg_backward(ctx, g_seq, g_lse, g_rl, g_rw):
    logits, row_label, row_weight, lse = ctx.saved_tensors
    d = logits_seq_logprob_bwd(logits, row_label, row_weight, lse, g_seq.contiguous(), ctx.length_normalize)
    return d.to(logits.dtype), None, None, None


logits_seq_logprob.register_autograd(_lg_backward, setup_context=_lg_setup)


# ============================================================================================ DPO scalar
@torch.library.custom_op("pgica::dpo_loss", mutates_args=())
def dpo_loss(pc: Tensor, pr: Tensor, rc: Optional[Tensor], rr: Optional[Tensor], beta: float,
             label_smoothing: float, n_global: int) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (loss[], metrics[5], dloss/dpc[n]).  metrics = (dpo_loss, reward_margin, reward_accuracy,
    mean policy chosen log-prob, mean policy rejected log-prob)."""
    f = lambda t: None if t is None else t.contiguous().float()
    return F.dpo_loss_fwd(f(pc), f(pr), f(rc), f(rr), beta, label_smoothing, n_global)


@dpo_loss.register_fake
def _(pc, pr, rc, rr, beta, label_smoothing, n_global):
    return pc.new_empty((), dtype=torch.float32), _f32(5, pc), _f32(pc.numel(), pc)


@torch.library.custom_op("pgica::scale_by_scalar", mutates_args=())
def scale_by_scalar(a: Tensor, scalar: Tensor, mult: float) -> Tensor:
    return F.scale_by_scalar(a.contiguous().float(), scalar.reshape(1).contiguous().float(), mult)


@scale_by_scalar.register_fake
def _(a, scalar, mult):
    return torch.empty_like(a, dtype=torch.float32)


def _dpo_setup(ctx, inputs, output):
    pc, pr, rc, rr = inputs[:4]
    ctx.save_for_backward(output[2])
    ctx.has_ref = rc is not None
    ctx.dtypes = (pc.dtype, pr.dtype, rc.dtype if rc is not None else None, rr.dtype if rr is not None else None)


def _dpo_backward(ctx, g_loss, g_metrics, g_dpc):
    (dpc,) = ctx.saved_tensors
    need = ctx.needs_input_grad
    pos = scale_by_scalar(dpc, g_loss, 1.0)
    neg = scale_by_scalar(dpc, g_loss, -1.0)
    dt = ctx.dtypes
    return (pos.to(dt[0]) if need[0] else None, neg.to(dt[1]) if need[1] else None,
            neg.to(dt[2]) if ctx.has_ref and need[2] else None, pos.to(dt[3]) if ctx.has_ref and need[3] else None,
            None, None, None)


dpo_loss.register_autograd(_dpo_backward, setup_context=_dpo_setup)


# ============================================================================== LM head on compacted rows
_BF16_CACHE = {}  # id(weight) -> (weakref, data_ptr, _version, bf16 copy)


def cached_bf16(weight: Tensor) -> Tensor:
    """bf16 operand copy of an fp32 LM-head weight, recast only when the Parameter has been written to (its
    `_version` moves with every in-place update, i.e. once per optimiser step): the tied wte / lm_head matrix of GPT-2
    Medium is 206 MB in fp32, and a Stage-2 micro-step would otherwise cast it once per forward and once per backward."""
    import weakref
    if weight.dtype == torch.bfloat16:
        return weight.detach().contiguous()
    key = id(weight)
    hit = _BF16_CACHE.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight.data_ptr() and hit[2] == weight._version:
        return hit[3]
    copy = F.as_bf16(weight.detach())
    try:
        ref = weakref.ref(weight, lambda _r, k=key: _BF16_CACHE.pop(k, None))
    except TypeError:
        return copy
    _BF16_CACHE[key] = (ref, weight.data_ptr(), weight._version, copy)
    return copy


class _LMHeadCompact(torch.autograd.Function):
    """Sequence log-probs of several sequence sets (preferred, rejected, ...) against one LM-head weight on the scored
    rows only (functional.lmhead_compact_fwd / _bwd): one forward GEMM and ONE dual backward launch for all sets, so the
    weight gradient is accumulated once.  A plain autograd.Function: the number of scored rows is read on the host to
    size the launches, which a shape-static custom op cannot express."""

    @staticmethod
    def forward(ctx, weight, length_normalize, nset, *rest):
        hiddens, labels, masks = rest[:nset], rest[nset:2 * nset], rest[2 * nset:3 * nset]
        wb = cached_bf16(weight)
        seqs, saved = F.lmhead_compact_fwd(list(hiddens), wb, list(labels), list(masks), length_normalize)
        ctx.saved, ctx.nset = saved, nset
        ctx.weight = weight
        ctx.h_dtypes = [h.dtype for h in hiddens]
        return tuple(seqs)

    @staticmethod
    def backward(ctx, *grad_seqs):
        nset = ctx.nset
        need_w = ctx.needs_input_grad[0]
        need_h = any(ctx.needs_input_grad[3:3 + nset])
        wb = cached_bf16(ctx.weight)
        gs = [g.contiguous().float() for g in grad_seqs]
        dts = [dt if dt in (torch.float32, torch.bfloat16) else torch.float32 for dt in ctx.h_dtypes]
        dhs, dw = F.lmhead_compact_bwd(ctx.saved, wb, gs, need_dhidden=need_h, need_dweight=need_w,
                                       dhidden_dtypes=dts, dweight_dtype=_grad_dtype(ctx.weight))
        ctx.saved = None
        out_h = [None] * nset
        if need_h:
            out_h = [dh.to(dt) if ctx.needs_input_grad[3 + s] else None
                     for s, (dh, dt) in enumerate(zip(dhs, ctx.h_dtypes))]
        return (dw.to(ctx.weight.dtype) if need_w else None, None, None, *out_h, *([None] * (2 * nset)))


def lmhead_seq_logprob_compact(hiddens, weight: Tensor, labels, masks, length_normalize: bool):
    """[(B_s, T_s, d)] hidden sets, (V, d) weight, [(B_s, T_s)] labels, [(B_s, T_s) or None] masks -> [seq_logp_s]."""
    nset = len(hiddens)
    return list(_LMHeadCompact.apply(weight, bool(length_normalize), nset, *hiddens, *labels, *masks))
