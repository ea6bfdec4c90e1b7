"""CUDA-graph capture of a whole DPO-head training step (forward + backward through the module API).

The head's step is ~20 short launches around three long ones; issued one by one from Python the host needs longer
to enqueue them than the B200 needs to run them (bench.py: 2.2 ms of host time against 1.7 ms of device time on
BASELINE config 2).  `GraphedDPOStep` records `FusedDPOHead.forward_stacked(...)` and its backward once into two CUDA
graphs over static input buffers and replays them: two `cudaGraphLaunch` calls per step, same kernels, same results.
The loss is copied to pinned host memory at the end of the FORWARD graph, so the host can read it (and go on to
enqueue the next step) while the backward graph is still running — the device never waits for the host.

    step = GraphedDPOStep(head, weight, ref_weight, hidden_like, labels_like, mask_like, ref_hidden_like, n_global)
    step.copy_inputs(hidden, labels, mask, ref_hidden)      # async copies into the static buffers
    step.launch()                                           # forward graph, event, backward graph
    value = step.loss_value()                               # waits for the forward graph only
    step.dweight, step.dhidden                              # gradients (valid once the stream has caught up)
"""
from typing import Optional

import torch


class GraphedDPOStep:
    def __init__(self, head, weight: torch.Tensor, ref_weight: Optional[torch.Tensor], hidden: torch.Tensor,
                 labels: torch.Tensor, mask: Optional[torch.Tensor], ref_hidden: Optional[torch.Tensor] = None,
                 n_global: Optional[int] = None, warmup: int = 2):
        if not hidden.is_cuda:
            raise RuntimeError("GraphedDPOStep needs CUDA tensors (there is no CPU path)")
        self.head = head
        self.weight = weight if weight.requires_grad else weight.detach().requires_grad_(True)
        self.ref_weight = ref_weight
        self.n_global = n_global
        # static inputs: the graphs read these addresses on every replay
        self.hidden = hidden.detach().clone().requires_grad_(True)
        self.labels = labels.clone()
        self.mask = None if mask is None else mask.clone()
        self.ref_hidden = None if ref_hidden is None else ref_hidden.detach().clone()
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self.forward_done = torch.cuda.Event()
        side = torch.cuda.Stream(device=hidden.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # first-use work (attribute setting, occupancy queries, allocator growth)
                loss, _ = self._forward()
                torch.autograd.grad(loss, (self.hidden, self.weight))
        torch.cuda.current_stream().wait_stream(side)
        self.fwd_graph = torch.cuda.CUDAGraph()
        self.bwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd_graph):
            self.loss, self.metrics = self._forward()
            self.loss_host.copy_(self.loss.detach(), non_blocking=True)
        with torch.cuda.graph(self.bwd_graph, pool=self.fwd_graph.pool()):
            self.dhidden, self.dweight = torch.autograd.grad(self.loss, (self.hidden, self.weight))

    def _forward(self):
        return self.head.forward_stacked(self.hidden, self.weight, self.labels, self.mask, self.ref_hidden,
                                         self.ref_weight, self.n_global)

    def copy_inputs(self, hidden, labels=None, mask=None, ref_hidden=None):
        """Asynchronous copies (e.g. from pinned host memory) into the graphs' static input buffers."""
        with torch.no_grad():
            self.hidden.copy_(hidden, non_blocking=True)
            if labels is not None:
                self.labels.copy_(labels, non_blocking=True)
            if mask is not None and self.mask is not None:
                self.mask.copy_(mask, non_blocking=True)
            if ref_hidden is not None and self.ref_hidden is not None:
                self.ref_hidden.copy_(ref_hidden, non_blocking=True)

    def launch(self):
        """Enqueue the step on the current stream: forward (+ loss to pinned memory), then backward."""
        self.fwd_graph.replay()
        self.forward_done.record()
        self.bwd_graph.replay()

    def loss_value(self) -> float:
        """The step's loss as a Python float; waits for the forward graph only."""
        self.forward_done.synchronize()
        return self.loss_host.item()

    def replay(self) -> torch.Tensor:
        """launch() and return the loss as a device tensor (no host synchronisation)."""
        self.launch()
        return self.loss


class GraphedContrastiveStep:
    """CUDA-graph capture of a contrastive-loss step (forward + backward of `loss_module(a, b)`), for the small batches
    the Stage-1 trainer uses: through the eager module API the step costs ~230 us of host time (custom-op dispatch,
    casts, autograd) around ~10 us of device time; replayed it is one `cudaGraphLaunch`.

        step = GraphedContrastiveStep(pg.ContrastiveLoss(temperature=0.07), image_embeddings_like, text_embeddings_like)
        step.copy_inputs(image_embeddings, text_embeddings)
        loss = step.replay()                 # device tensor; step.da / step.db hold the gradients
    """

    def __init__(self, loss_module, a: torch.Tensor, b: torch.Tensor, warmup: int = 2):
        if not a.is_cuda:
            raise RuntimeError("GraphedContrastiveStep needs CUDA tensors (there is no CPU path)")
        self.loss_module = loss_module
        self.a = a.detach().clone().requires_grad_(True)
        self.b = b.detach().clone().requires_grad_(True)
        side = torch.cuda.Stream(device=a.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                torch.autograd.grad(self.loss_module(self.a, self.b), (self.a, self.b))
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self.loss_module(self.a, self.b)
            self.da, self.db = torch.autograd.grad(self.loss, (self.a, self.b))

    def copy_inputs(self, a, b):
        with torch.no_grad():
            self.a.copy_(a, non_blocking=True)
            self.b.copy_(b, non_blocking=True)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
