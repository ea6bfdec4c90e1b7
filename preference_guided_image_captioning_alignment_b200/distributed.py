"""Multi-GPU forms of the two heads — only where they shard naturally.

NT-Xent with global negatives: rank r owns rows [r*b, (r+1)*b) of both embedding matrices.
    forward : all-gather text rows -> local (b x B) slice on the tensor cores -> row LSE is complete, column
              LSE is a partial over b rows -> all-gather the W partial vectors (B floats each) and merge ->
              local loss terms -> all-reduce of one scalar.
    backward: dA_r is final locally; dB is a (B x D) partial -> reduce-scatter to the owners.
    The result equals the reference module applied to the concatenated global batch on one device (the
    reference itself, under DDP, only ever sees local negatives).  Gradients are those of ONE copy of the
    global loss (no extra 1/world factor): averaging across replicas is the caller's / DDP's business.

DPO: preference pairs are independent -> shard pairs, replicate W; the loss is a mean over the GLOBAL batch
    (1/B_global inside the kernel), scalars are all-reduced, and dW is all-reduced by `allreduce_dweight`
    (stand-alone use) or by DDP's bucket reducer when the head sits inside a DDP-wrapped model.

The collective plumbing is written against a small `compute` interface so that the CPU test-suite can drive
it under gloo with the oracle standing in for the kernels; the product always uses CudaCompute.
"""
from typing import Optional

import torch
import torch.distributed as dist

from . import functional as F


class CudaCompute:
    """The kernels (functional.py -> C ABI).  No CPU implementation exists."""

    name = "cuda"

    def to_operand(self, x):
        return F.as_bf16(x)

    def __init__(self, assume_normalized: bool = False):
        # unit-norm rows (what the reference model hands its loss, pkg/models/model.py:826-829): row and column
        # log-sum-exp from one pass over the similarity tiles instead of two, one exponential per element in the backward
        self.assume_normalized = bool(assume_normalized)

    def ntxent_fwd(self, a, b_all, inv_tau, diag_offset):
        return F.ntxent_fwd(a, b_all, inv_tau, diag_offset, bounded=self.assume_normalized)

    def lse_combine(self, parts):
        return F.lse_combine(parts)

    def ntxent_loss(self, lse_row, diag, lse_col_owned, inv_denom):
        return F.ntxent_loss(lse_row, diag, lse_col_owned, inv_denom)

    def ntxent_bwd(self, a, b_all, inv_tau, diag_offset, lse_row, lse_col, grad_loss, mult):
        return F.ntxent_bwd(a, b_all, inv_tau, diag_offset, lse_row, lse_col, grad_loss, mult,
                            da_dtype=torch.float32, db_dtype=torch.float32, bounded=self.assume_normalized)


def _world(group):
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """(b, D) per rank -> (W*b, D), rank-major."""
    world, _ = _world(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def reduce_scatter_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """(W*b, D) partial sums per rank -> (b, D) owned rows, summed over ranks."""
    world, rank = _world(group)
    b = x.shape[0] // world
    out = torch.empty((b,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if dist.get_backend(group) == "gloo":  # gloo has no reduce_scatter: all-reduce and keep the owned slice
        full = x.contiguous().clone()
        dist.all_reduce(full, group=group)
        out.copy_(full[rank * b:(rank + 1) * b])
    else:
        dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
    return out


class _GlobalNTXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, inv_tau, reduce_mean, group, compute):
        world, rank = _world(group)
        nb = a.shape[0]
        B = nb * world
        a_op = compute.to_operand(a)
        b_all = all_gather_rows(compute.to_operand(b), group)
        off = rank * nb
        lse_row, diag, lse_col_part = compute.ntxent_fwd(a_op, b_all, inv_tau, off)
        parts = all_gather_rows(lse_col_part.reshape(1, B), group)  # (W, B)
        lse_col = compute.lse_combine(parts)
        loss = compute.ntxent_loss(lse_row, diag, lse_col[off:off + nb].contiguous(), (1.0 / B) if reduce_mean else 1.0)
        loss = loss.clone()
        dist.all_reduce(loss, group=group)
        ctx.save_for_backward(a_op, b_all, lse_row, lse_col)
        ctx.meta = (inv_tau, off, (1.0 / (2.0 * B)) if reduce_mean else 0.5, group, compute, a.dtype, b.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        a_op, b_all, lse_row, lse_col = ctx.saved_tensors
        inv_tau, off, mult, group, compute, a_dtype, b_dtype = ctx.meta
        da, db_part = compute.ntxent_bwd(a_op, b_all, inv_tau, off, lse_row, lse_col, grad_loss.contiguous(), mult)
        db = reduce_scatter_rows(db_part, group)
        return da.to(a_dtype), db.to(b_dtype), None, None, None, None


def global_ntxent(a: torch.Tensor, b: torch.Tensor, temperature: float, reduce_mean: bool = True, group=None,
                  compute=None, assume_normalized: bool = False) -> torch.Tensor:
    """NT-Xent over the GLOBAL batch; a, b are this rank's (b, D) rows, used as given (normalise first if needed).
    assume_normalized=True promises unit-norm rows and lets the forward make one pass instead of two.
    Returns the global loss (identical on every rank)."""
    return _GlobalNTXent.apply(a, b, 1.0 / float(temperature), reduce_mean, group,
                               compute or CudaCompute(assume_normalized))


class GlobalContrastiveLoss(torch.nn.Module):
    """ContrastiveLoss (pkg/models/model.py:957-1000 semantics) with negatives from every rank."""

    def __init__(self, temperature: float = 0.07, group=None, compute=None, assume_normalized: bool = False):
        super().__init__()
        self.temperature = temperature
        self.group = group
        self.compute = compute
        self.assume_normalized = assume_normalized

    def forward(self, image_embeddings, text_embeddings):
        return global_ntxent(image_embeddings, text_embeddings, self.temperature, True, self.group, self.compute,
                             self.assume_normalized)


# ------------------------------------------------------------------------------------------------ DPO
def shard_pairs(n_global: int, rank: int, world: int):
    """Contiguous, balanced slice of the global pair batch owned by `rank`."""
    base, rem = divmod(n_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_scalars(loss: torch.Tensor, metrics: Optional[torch.Tensor], group=None):
    """Local sums / B_global -> global mean (the kernel already divides by n_global)."""
    packed = torch.cat([loss.detach().reshape(1), metrics.detach()]) if metrics is not None else loss.detach().reshape(1)
    dist.all_reduce(packed, group=group)
    return packed[0], (packed[1:] if metrics is not None else None)


def allreduce_dweight(dweight: torch.Tensor, group=None, async_op: bool = False):
    """Sum the LM-head weight gradient over the data-parallel ranks (206 MB fp32 for GPT-2 Medium).
    Skip inside a DDP-wrapped model: DDP's reducer already owns the tied wte / lm_head Parameter."""
    return dist.all_reduce(dweight, group=group, async_op=async_op)


class OverlappedDWAllReduce:
    """Data-parallel sum of the LM-head weight gradient that runs WHILE the backward kernel is still producing it.

    The dual backward kernel (csrc/sgg_f.cu) finalises dW in vocabulary order and bumps a per-segment progress counter
    (release at GPU scope) whenever the stores of a finished tile have completed.  Next to it, on a second stream, runs
    `pgica_peer_allreduce_progress` (csrc/peer_ar.cu): one small CTA per SM that fits beside the persistent kernel.  Per
    segment it waits for the local counter, meets the other ranks at a flag barrier in peer memory, and every rank sums
    its 1/W slice of the segment straight out of all W symmetric buffers over NVLink and stores the result into all of
    them (reduce-scatter + all-gather of a two-shot all-reduce, one pass, deterministic).  Only the last segment's
    share of the transfer is exposed after the backward kernel ends; the NCCL all-reduce it replaces cost 0.61 ms of a
    2.2 ms step on 8 GPUs (profiles/r1_scaling_notes.md).

        red = OverlappedDWAllReduce(vocab, d, device)
        dh, dw, done = red.backward(hidden, weight, row_label, row_weight, lse, grad_seq)   # per step
        torch.cuda.current_stream().wait_event(done)      # dw = the summed (V, d) gradient, identical on every rank

    dW lives in torch symmetric memory; the buffer must not be rewritten (next backward) before `done` has fired on
    every rank — `backward` itself orders that."""

    def __init__(self, vocab: int, d: int, device, group=None, segments: int = 8, max_ctas: int = -1,
                 multicast: Optional[bool] = None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = _world(group)
        self.vocab, self.d = int(vocab), int(d)
        pairs = (self.vocab + 255) // 256                       # 256-row column pairs of the kernel
        self.rows = pairs * 256
        segments = max(1, min(int(segments), pairs, 32))
        self.pairs_per_seg = (pairs + segments - 1) // segments
        self.nseg = (pairs + self.pairs_per_seg - 1) // self.pairs_per_seg
        self.rows_per_seg = self.pairs_per_seg * 256
        n = self.rows * self.d
        nflag = (self.nseg + 1) * self.world
        # one symmetric allocation: [dW: rows x d fp32][flags: (nseg + 1) x world uint32, padded]
        self.flag_off = n
        self.buf = symm.empty(n + ((nflag + 63) // 64) * 64, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        total = self.buf.numel()
        flat = [self.hdl.get_buffer(p, (total,), torch.float32) for p in range(self.world)]
        self.buf_ptrs = [f.data_ptr() for f in flat]
        self.flag_ptrs = [f.data_ptr() + 4 * self.flag_off for f in flat]
        self._flat = flat
        # NVSwitch multicast mapping of the same allocation (NVLS), when the fabric offers one: the sum then happens
        # inside the switch.  multicast=None: use it if available; False: never; True: require it.
        mc_base = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if multicast is True and not mc_base:
            raise RuntimeError("OverlappedDWAllReduce(multicast=True): this symmetric allocation has no multicast mapping")
        self.multicast_ptr = 0
        if multicast is None and self.world <= 2:
            multicast = False  # two ranks: nothing to save (2 loads -> 1) and measured slower, 1.65-1.84 vs 1.60 ms per step
        if mc_base and multicast is not False and self.world > 1:
            self.multicast_ptr = mc_base + (flat[self.rank].data_ptr() - int(self.hdl.buffer_ptrs[self.rank]))
        self.local = flat[self.rank][:n].view(self.rows, self.d)
        self.progress = torch.zeros(self.nseg, dtype=torch.int32, device=device)
        self.local_sync = torch.zeros(2, dtype=torch.int32, device=device)
        unit = 4 * self.world
        assert (self.rows_per_seg * self.d) % unit == 0
        self.seg_begin = [min(s * self.rows_per_seg, self.rows) * self.d for s in range(self.nseg + 1)]
        self.seg_pairs = [(min((s + 1) * self.rows_per_seg, self.rows) - s * self.rows_per_seg) // 256
                          for s in range(self.nseg)]
        self.stream = torch.cuda.Stream(device=device)
        # Grid of the all-reduce kernel.  Measured on 8 B200s (profiles/r2_n8_overlap_tuning.jsonl): one CTA per SM slows
        # the co-resident backward kernel by 22 % (1.32 vs 1.08 ms), one per two SMs by 9 % with the same exposed tail
        # (0.125 ms), one per four SMs no longer keeps up with NVLink (0.34 ms exposed).  With 2 ranks a rank moves half
        # as much and the full grid is the faster one (1.75 vs 1.80 ms).  Default: every SM up to 2 ranks, half beyond.
        # Through the multicast mapping a CTA issues a `world`-th of the memory instructions; 8 B200s, same box:
        # 74 CTAs 1.79 ms per step, 37: 1.71, 24: 1.69 (unicast with 74: 1.75) — profiles/r2_n8_multicast_tuning.jsonl.
        sms = F._lib.load().pgica_sm_count()
        self.max_ctas = int(max_ctas) if int(max_ctas) >= 0 else (sms if self.world <= 2 else max(1, sms // 2))
        if int(max_ctas) < 0 and self.multicast_ptr:
            self.max_ctas = max(1, sms // 6)
        self.epoch = 0
        self.done = torch.cuda.Event()
        self.done.record()
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)  # every rank's flags are zero before anybody's first epoch

    @property
    def view(self):
        """The reduced gradient, (V, d) fp32 (rows past the vocabulary belong to the padding of the last tile pair)."""
        return self.local[: self.vocab]

    def backward(self, hidden, weight, row_label, row_weight, lse, grad_seq, length_normalize=False,
                 dhidden_dtype=torch.bfloat16, scalars: Optional[torch.Tensor] = None):
        """-> (dhidden, dweight view, event[, summed scalars]): dweight is complete on the event (recorded on the
        reducer's stream).  `scalars` (a few fp32 values: the local loss / metric sums) ride along in the padding rows
        of the last tile pair — rows past the vocabulary that the kernel never writes — and come back summed over the
        ranks, so the step needs no second collective."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self.done)  # the previous all-reduce has left the buffer (on every rank: its final barrier)
        self.epoch += 1
        pad = None
        if scalars is not None:
            if self.rows == self.vocab or scalars.numel() > self.d:
                raise ValueError("no padding row to carry the scalars (vocabulary is a multiple of 256)")
            pad = self.local[self.rows - 1, : scalars.numel()]
            pad.copy_(scalars.detach().float().reshape(-1))
        dh, inc = F.lmhead_logprob_bwd_progress(hidden, weight, row_label, row_weight, lse, grad_seq, self.view,
                                                self.progress, self.rows_per_seg, length_normalize, dhidden_dtype)
        targets = [self.epoch * inc * np_ for np_ in self.seg_pairs]
        F.peer_allreduce_progress(self.buf_ptrs, self.flag_ptrs, self.rank, self.progress, targets, self.seg_begin,
                                  self.epoch, self.local_sync, self.max_ctas, stream=self.stream,
                                  multicast_ptr=self.multicast_ptr)
        self.done = torch.cuda.Event()
        self.done.record(self.stream)
        if scalars is not None:
            return dh, self.view, self.done, pad
        return dh, self.view, self.done
