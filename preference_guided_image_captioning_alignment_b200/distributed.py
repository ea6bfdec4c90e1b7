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

    def ntxent_fwd(self, a, b_all, inv_tau, diag_offset):
        return F.ntxent_fwd(a, b_all, inv_tau, diag_offset)

    def lse_combine(self, parts):
        return F.lse_combine(parts)

    def ntxent_loss(self, lse_row, diag, lse_col_owned, inv_denom):
        return F.ntxent_loss(lse_row, diag, lse_col_owned, inv_denom)

    def ntxent_bwd(self, a, b_all, inv_tau, diag_offset, lse_row, lse_col, grad_loss, mult):
        return F.ntxent_bwd(a, b_all, inv_tau, diag_offset, lse_row, lse_col, grad_loss, mult,
                            da_dtype=torch.float32, db_dtype=torch.float32)


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
                  compute=None) -> torch.Tensor:
    """NT-Xent over the GLOBAL batch; a, b are this rank's (b, D) rows, used as given (normalise first if needed).
    Returns the global loss (identical on every rank)."""
    return _GlobalNTXent.apply(a, b, 1.0 / float(temperature), reduce_mean, group, compute or CudaCompute())


class GlobalContrastiveLoss(torch.nn.Module):
    """ContrastiveLoss (pkg/models/model.py:957-1000 semantics) with negatives from every rank."""

    def __init__(self, temperature: float = 0.07, group=None, compute=None):
        super().__init__()
        self.temperature = temperature
        self.group = group
        self.compute = compute

    def forward(self, image_embeddings, text_embeddings):
        return global_ntxent(image_embeddings, text_embeddings, self.temperature, True, self.group, self.compute)


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


class PeerAllReduce:
    """Sum-all-reduce of ONE fixed fp32 buffer across the GPUs of an NVLink node, moved by the COPY ENGINES.

    Why not NCCL here: the gradient kernels of the Stage-2 head are persistent 4-CTA-cluster kernels that keep 128-132
    of the 148 SMs resident; every SM an NCCL kernel takes displaces a whole cluster, so an "overlapped" NCCL
    all-reduce of the 206 MB LM-head weight gradient made the dH kernel wait for it (measured at N=2: 1.27 ms
    overlapped vs 0.82 + 0.34 ms back to back).  Peer copies through symmetric memory use no SM at all:

        barrier -> pull my 1/W chunk of every peer's buffer (W-1 peer copies into scratch)
                -> pgica_sum_into_f32 (a few CTAs: fits beside the resident clusters)
                -> push the reduced chunk into every peer's buffer -> barrier

    Traffic per GPU and direction: 2 (W-1)/W of the buffer, the same as a ring all-reduce; measured peer-copy rate on
    this pool ~700 GB/s.  The buffer lives in symmetric memory (torch.distributed._symmetric_memory); the kernels
    write the gradient straight into it (`functional.lmhead_logprob_bwd(..., dweight_out=reducer.view)`)."""

    def __init__(self, shape, device, group=None, sum_ctas: int = 32):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = _world(group)
        self.shape = tuple(shape)
        numel = 1
        for s in self.shape:
            numel *= int(s)
        unit = 4 * self.world  # chunks are float4-aligned
        self.numel = numel
        self.padded = (numel + unit - 1) // unit * unit
        self.chunk = self.padded // self.world
        self.buf = symm.empty(self.padded, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.peers = [self.hdl.get_buffer(p, (self.padded,), torch.float32) for p in range(self.world)]
        self.scratch = torch.empty((max(self.world - 1, 1), self.chunk), dtype=torch.float32, device=device)
        self.stream = torch.cuda.Stream(device=device)
        self.sum_ctas = sum_ctas
        self.trace = False       # set True to keep timing events of the last all_reduce in .last_trace
        self.last_trace = None
        self.buf.zero_()

    @property
    def view(self):
        """The local buffer with the caller's shape (what the gradient kernel writes into)."""
        return self.buf[: self.numel].view(self.shape)

    def all_reduce(self, after: Optional[torch.cuda.Event] = None) -> torch.cuda.Event:
        """Enqueue the all-reduce on the reducer's own stream, ordered after `after` (default: everything enqueued
        on the current stream so far).  Returns the event that marks the reduced buffer complete on this rank; the
        buffer must not be rewritten before every rank's event has fired (wait on it before the next producer)."""
        if after is None:
            after = torch.cuda.Event()
            after.record()
        s = self.stream
        s.wait_event(after)
        lo, hi = self.rank * self.chunk, (self.rank + 1) * self.chunk
        others = [p for p in range(self.world) if p != self.rank]
        trace = [] if self.trace else None

        def mark(name):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                trace.append((name, e))

        with torch.cuda.stream(s):
            mark("start")
            self.hdl.barrier(channel=0)  # every rank's buffer is complete
            mark("barrier0")
            for i, p in enumerate(others):
                self.scratch[i].copy_(self.peers[p][lo:hi])
            mark("pull")
            if others:
                F.sum_into(self.buf[lo:hi], [self.scratch[i] for i in range(len(others))], self.sum_ctas)
            mark("sum")
            for p in others:
                self.peers[p][lo:hi].copy_(self.buf[lo:hi])
            mark("push")
            self.hdl.barrier(channel=1)  # every rank's pushes have landed
            mark("barrier1")
            done = torch.cuda.Event()
            done.record()
        if trace is not None:
            self.last_trace = trace
        return done


class FusedDWReduce:
    """Data-parallel sum of the LM-head weight gradient with the SCATTER done by the backward kernel itself.

    The dual backward kernel (csrc/sgg_f.cu) keeps 128-row tiles of dW resident in TMEM; in scatter mode it drains
    every finished tile with a TMA store straight into this rank's slot in the memory of the rank that OWNS those rows
    — peer memory over NVLink / NVSwitch, mapped through torch symmetric memory — while the tensor cores are already
    on the next tiles.  When the kernels of all ranks have ended, rank r holds W partial copies of its rows
    [r * own, (r+1) * own), sums them (pgica_sum_into_f32, HBM-bound, 1/W of dW), and the all-gather that completes
    the all-reduce is W-1 peer copies of 1/W of the buffer on the copy engines:

        barrier -> backward kernel (dH local, dW tiles -> owners' slots) -> barrier -> sum my slots
                -> push my rows to every peer -> barrier

    Per GPU the kernel sends (W-1)/W of dW spread over its whole run time (~0.2 TB/s at W=8, a fraction of NVLink)
    and only the sum and the all-gather are exposed.  (Remote TMA add-reductions instead of per-source slots were
    measured an order of magnitude slower: 2.5 ms instead of 1.0 ms for the kernel at W=8.)  Needs a problem whose
    rows fit one chunk of the kernel (<= 32 row blocks at d = 1024); larger ones accumulate dW locally and all-reduce.

    `backward(...)` returns (dhidden, dweight_view); dweight_view is this rank's copy of the reduced (V, d) fp32
    gradient inside the symmetric buffer, valid on the current stream when the call returns."""

    def __init__(self, vocab: int, d: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = _world(group)
        self.vocab, self.d = int(vocab), int(d)
        blocks = (self.vocab + 127) // 128
        self.own_rows = 128 * ((blocks + self.world - 1) // self.world)
        self.rows = self.own_rows * self.world
        n_own = self.own_rows * self.d
        # one symmetric allocation: [reduced gradient: rows x d][slots: world x own_rows x d]
        self.buf = symm.empty(self.rows * self.d + self.world * n_own, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        total = self.buf.numel()
        flat = [self.hdl.get_buffer(p, (total,), torch.float32) for p in range(self.world)]
        self.out_peers = [f[: self.rows * self.d].view(self.rows, self.d) for f in flat]
        self.local = self.out_peers[self.rank]
        self.slots = flat[self.rank][self.rows * self.d:].view(self.world, self.own_rows, self.d)
        # where rank p keeps the slot for MY tiles
        base = self.rows * self.d + self.rank * n_own
        self.peer_ptrs = [f[base: base + n_own].data_ptr() for f in flat]
        raw = torch.empty(128 * self.world + 128, dtype=torch.uint8, device=device)
        off = (-raw.data_ptr()) % 128
        self.tmaps = raw[off: off + 128 * self.world]
        self.buf.zero_()

    @property
    def view(self):
        return self.local[: self.vocab]

    def backward(self, hidden, weight, row_label, row_weight, lse, grad_seq, length_normalize=False,
                 dhidden_dtype=torch.bfloat16):
        lo, hi = self.rank * self.own_rows, (self.rank + 1) * self.own_rows
        self.hdl.barrier(channel=0)  # every owner has summed last step's slots: they may be overwritten
        dh = F.lmhead_logprob_bwd_scatter(hidden, weight, row_label, row_weight, lse, grad_seq, self.peer_ptrs,
                                          self.own_rows, self.tmaps, length_normalize, dhidden_dtype)
        self.hdl.barrier(channel=1)  # every rank's tiles have landed in my slots
        mine = self.local[lo:hi]
        mine.zero_()
        F.sum_into(mine, [self.slots[s] for s in range(self.world)])
        for p in range(self.world):
            if p != self.rank:
                self.out_peers[p][lo:hi].copy_(mine)
        self.hdl.barrier(channel=2)  # everybody's rows have arrived here
        return dh, self.view
