"""Whole-step execution: forward + loss + backward (+ gradient all-reduce) + optimizer update as ONE CUDA graph.

The reference trains with a Python loop (experiments/train_helper.py:90-145): per step it builds a graph, moves
it to the device, runs ``model(graph)``, ``loss = sqrt(MSE_sum(pred, graph.y))``, ``loss.backward()``,
``optimizer.step()``.  The drop-in modules run that loop unchanged; this class is the fast path for the same
step.  On the reference's fixed grids the topology (``edge_index``, ``batch``) is identical for every step
(common/utils.py:365-377 depends only on the grid and the batch size), so the whole step is captured once
and replayed: per step the host only copies the new ``x, y, pos`` (+ PDE parameters) into static device
buffers and launches one graph -- no per-kernel launch or Python overhead on the critical path.

Data parallelism: whole graphs are sharded across ranks; because the loss is the
square root of the *batch-global* summed squared error (train_helper.py:126,138) the local sum is all-reduced
before the backward seed is formed, and the flat gradient bucket is SUM-all-reduced (NCCL) inside the same
captured graph, followed by the identical AdamW update on every rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _GlobalSqrtSSE(torch.autograd.Function):
    """loss = sqrt(sum over ALL ranks of the squared error); backward scales by 1 / (2 loss)."""

    @staticmethod
    def forward(ctx, sse_local, group):
        total = sse_local.detach().clone()
        if group is not None:
            dist.all_reduce(total, group=group)
        loss = torch.sqrt(total)
        ctx.save_for_backward(loss)
        return loss

    @staticmethod
    def backward(ctx, g):
        (loss,) = ctx.saved_tensors
        return g * 0.5 / loss, None


def global_rmse_loss(pred, y, group=None):
    """sqrt(MSELoss(reduction='sum')) over the global (all-rank) batch -- train_helper.py:126,138."""
    sse = ((pred - y) ** 2).sum()
    if group is None and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return torch.sqrt(sse)
    return _GlobalSqrtSSE.apply(sse, group if group is not None else dist.group.WORLD)


class FlatGradBucket:
    """One contiguous buffer (the parameters' dtype; fp32 for the msmp modules) holding every parameter gradient (``p.grad`` are views into it), so the
    data-parallel reduction is a single NCCL all-reduce on a fixed address."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=self.params[0].dtype, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        dist.all_reduce(self.flat, group=group)          # SUM: the loss already is the global one


class GraphedTrainStep:
    """Captures ``model(graph) -> loss -> backward -> (all-reduce) -> optimizer.step()`` into CUDA graphs.

    >>> step = GraphedTrainStep(model, optimizer, example_graph_on_device)
    >>> loss = step(host_or_device_graph)        # copies the float fields into the static graph and replays

    Single GPU: ONE graph for the whole step.  Data parallel (world > 1): three graphs -- forward (+ local
    squared error), backward (seeded with 1 / (2 sqrt(global SSE))), optimizer -- with the two NCCL all-reduces
    (one scalar, one flat gradient bucket) launched eagerly between them on the same stream.

    ``optimizer`` must be capturable (``torch.optim.AdamW(..., capturable=True)``).  The topology of later
    graphs must equal the example's (only floating-point fields are refreshed)."""

    def __init__(self, model, optimizer, example, group=None, warmup: int = 3, use_graph: bool = True):
        self.model, self.opt, self.group = model, optimizer, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.static = example.clone()
        # floating-point fields are refreshed every step; integer fields (edge_index, batch) are the static topology
        self.fields = [k for k in self.static.keys()
                       if torch.is_tensor(getattr(self.static, k)) and getattr(self.static, k).is_floating_point()]
        self.bucket = FlatGradBucket(model.parameters())
        self.gplan = None          # gradsink.GradPlan, built after the first eager step (the packs exist by then)
        self.use_graph = use_graph
        self.graph = None
        self.graphs = None
        self.loss = None
        dev = self.bucket.flat.device
        self._sse_total = torch.zeros((), dtype=torch.float64, device=dev)
        self._seed = torch.zeros((), dtype=torch.float64, device=dev)
        side = torch.cuda.Stream(priority=-1)      # high priority: weight-gradient side streams run below it
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(max(warmup, 1)):
                self._eager_step()
                if i == 0:
                    self._make_grad_plan()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if not use_graph:
            return
        if self.world == 1:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._eager_step()
        else:
            g_fwd, g_bwd, g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fwd, stream=side):
                self._fwd()
            self._reduce_loss()
            with torch.cuda.graph(g_bwd, pool=g_fwd.pool(), stream=side):
                self._bwd()
            self.bucket.all_reduce(self.group)
            with torch.cuda.graph(g_opt, pool=g_fwd.pool(), stream=side):
                self.opt.step()
            self.graphs = (g_fwd, g_bwd, g_opt)
        torch.cuda.synchronize()

    # ---- pieces of one step -------------------------------------------------------------------------
    def _make_grad_plan(self):
        """Route the weight gradients of the msmp modules through one unpack launch (gradsink.py); models without
        tensor-core packs (or foreign modules only) keep plain autograd accumulation."""
        from . import ops
        if ops.GEMM_MODE != "tc" or self.model.__dict__.get("_msmp_pack_plan") is None and not any(
                m.__dict__.get("_msmp_pack_plan") is not None for m in self.model.modules()):
            return
        from .gradsink import GradPlan
        self.gplan = GradPlan(self.model)

    def _fwd(self):
        if self.gplan is None:
            self.bucket.zero_()          # with a GradPlan the covered gradients are overwritten, the others zeroed in begin()
        self._fwd_stream = torch.cuda.current_stream()
        pred = self.model(self.static)
        self._sse_local = ((pred - self.static.y) ** 2).sum()          # train_helper.py:126 (reduction='sum')
        self._sse_val = self._sse_local.detach()                       # same storage, no autograd graph

    def _reduce_loss(self):
        """global SSE (all ranks) -> loss and the backward seed d loss / d sse_local = 1 / (2 loss)."""
        self._sse_total.copy_(self._sse_val)
        if self.world > 1:
            dist.all_reduce(self._sse_total, group=self.group)
        self.loss = torch.sqrt(self._sse_total)
        self._seed.copy_(0.5 / self.loss)

    def _bwd(self):
        # The AccumulateGrad nodes of this step were created on the forward's stream and the engine syncs the
        # caller's stream with it after the backward pass even when a node only saw an undefined gradient (the
        # sink returns None for the parameters).  When the backward runs on another stream (separate CUDA graphs
        # in the data-parallel step) that stream is forked here so the sync stays inside the capture.
        cur = torch.cuda.current_stream()
        if self._fwd_stream != cur:
            self._fwd_stream.wait_stream(cur)
        if self.gplan is not None:
            self.gplan.begin()
        try:
            self._sse_local.backward(self._seed.to(self._sse_local.dtype))
        finally:
            if self.gplan is not None:
                self.gplan.finish()
            # drop the autograd graph now: AccumulateGrad nodes kept alive across steps stay bound to the stream of
            # the step that created them (a warm-up stream outside any later capture)
            self._sse_local = None

    def _eager_step(self):
        self._fwd()
        self._reduce_loss()
        self._bwd()
        if self.world > 1:
            self.bucket.all_reduce(self.group)
        self.opt.step()

    def load(self, graph):
        """Copy the floating-point fields of ``graph`` (host or device) into the static device buffers."""
        for k in self.fields:
            src = getattr(graph, k)
            dst = getattr(self.static, k)
            if src is dst:
                continue
            if src.shape != dst.shape:
                raise ValueError(f"graph.{k} has shape {tuple(src.shape)}, the captured step expects {tuple(dst.shape)}")
            dst.copy_(src, non_blocking=True)

    def __call__(self, graph=None):
        if graph is not None and graph is not self.static:
            self.load(graph)
        if self.graph is not None:
            self.graph.replay()
        elif self.graphs is not None:
            self.graphs[0].replay()
            self._reduce_loss()
            self.graphs[1].replay()
            self.bucket.all_reduce(self.group)
            self.graphs[2].replay()
        else:
            self._eager_step()
        return self.loss
