"""Whole-step execution: forward + loss + backward (+ gradient all-reduce) + optimizer update as CUDA graph replays.

The reference trains with a Python loop (experiments/train_helper.py:90-145): per step it builds a graph, moves
it to the device, runs ``model(graph)``, ``loss = sqrt(MSE_sum(pred, graph.y))``, ``loss.backward()``,
``optimizer.step()``.  The drop-in modules run that loop unchanged; this class is the fast path for the same
step.  On the reference's fixed grids the topology (``edge_index``, ``batch``) is identical for every step
(common/utils.py:365-377 depends only on the grid and the batch size), so the whole step is captured once
and replayed: per step the host only copies the new ``x, y, pos`` (+ PDE parameters) into static device
buffers, refreshes eight floats of optimizer hyper-parameters and launches one graph.

How the loss enters (train_helper.py:126,138): ``loss = sqrt(SSE)`` with ``SSE`` the squared error summed over the
whole (all-rank) batch, so ``d loss / d theta = (1 / (2 sqrt(SSE))) * d SSE / d theta``.  The backward pass is seeded with
1 (it produces ``d SSE_local / d theta``), the local ``SSE`` rides along as the last two elements (an exact float pair) of
the flat gradient bucket, and the factor ``1 / (2 sqrt(SSE))`` is applied by the optimizer kernel, which also writes the
scaled gradient back -- after the step ``param.grad`` is the gradient of the loss, as in the reference.

Data parallelism: whole graphs are sharded across ranks.  Because the squared error travels
inside the bucket, a step needs exactly ONE collective: a SUM all-reduce (NCCL) of the bucket, launched between the two
captured graphs (forward + backward | optimizer); every rank then applies the identical update.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import optim as moptim
from ._lib import check, lib


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
    """One contiguous buffer (the parameters' dtype; fp32 for the msmp modules) holding every parameter gradient
    (``p.grad`` are views into it) plus ``extra`` trailing scalars, so the data-parallel reduction is a single NCCL
    all-reduce on a fixed address."""

    def __init__(self, params, extra: int = 0):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.n = n
        self.flat = torch.zeros(n + extra, dtype=self.params[0].dtype, device=dev)
        self.tail = self.flat[n:]
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        dist.all_reduce(self.flat, group=group)          # SUM: gradients of the summed squared error + the error itself


class GraphedTrainStep:
    """Captures ``model(graph) -> loss -> backward -> (all-reduce) -> optimizer update`` into CUDA graphs.

    >>> step = GraphedTrainStep(model, optimizer, example_graph_on_device)
    >>> loss = step(host_or_device_graph)        # copies the float fields into the static graph and replays

    Single GPU: ONE graph for the whole step.  Data parallel (world > 1): two graphs -- forward + backward | optimizer --
    with the one NCCL all-reduce of the gradient bucket (which carries the local squared error) launched between them.

    ``optimizer``: any ``torch.optim.AdamW`` without amsgrad / maximize, e.g. exactly the one experiments/train.py:410
    builds, is driven by the package's own update kernel (optim.FusedAdamW: learning-rate schedulers keep working);
    other optimizers must be ``capturable`` and are captured as they are.  ``group=False`` forces the single-process
    step inside an initialised process group.  The topology of later graphs must equal the example's (only floating-point
    fields are refreshed; a different ``edge_index`` / ``batch`` raises)."""

    def __init__(self, model, optimizer, example, group=None, warmup: int = 3, use_graph: bool = True,
                 preserve_state: bool = False):
        self.model, self.opt = model, optimizer
        if group is False:
            self.group, self.world = None, 1
        else:
            self.group = group
            self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.static = example.clone()
        self._topo_src = (example.edge_index, example.batch)          # strong references: identity = same topology
        self._topo_ok = {(example.edge_index.data_ptr(), example.edge_index._version)}
        # floating-point fields are refreshed every step; integer fields (edge_index, batch) are the static topology
        self.fields = [k for k in self.static.keys()
                       if torch.is_tensor(getattr(self.static, k)) and getattr(self.static, k).is_floating_point()]
        self.bucket = FlatGradBucket(model.parameters(), extra=2)
        self._staging, self._have_staged = None, False          # prefetch()
        self.gplan = None          # gradsink.GradPlan, built after the first eager step (the packs exist by then)
        self.use_graph = use_graph
        self.graph = None
        self.graphs = None
        dev = self.bucket.flat.device
        self.loss = torch.zeros((), dtype=torch.float64, device=dev)
        self.gscale = torch.zeros((), dtype=torch.float32, device=dev)
        self.fused = moptim.FusedAdamW(optimizer) if moptim.supported(optimizer) else None
        if self.fused is None and use_graph and not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise ValueError("GraphedTrainStep captures torch.optim.AdamW (any flavour) with its own update kernel; "
                             "other optimizers must be built with capturable=True")
        if self.fused is None and use_graph:
            # a Python-float lr would be baked into the captured kernels: schedulers fill_ a tensor lr in place instead
            for g in optimizer.param_groups:
                if not torch.is_tensor(g["lr"]):
                    g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device=dev)
        # preserve_state: the warm-up steps below really train; a caller that wants its first call to be the first
        # update (train_helper.training_loop) gets parameters and optimizer state put back afterwards
        snap = self._snapshot() if preserve_state else None
        side = torch.cuda.Stream(priority=-1)      # high priority: weight-gradient side streams run below it
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(max(warmup, 1)):
                self._host_prologue()
                self._eager_step()
                if i == 0:
                    self._make_grad_plan()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            self._restore(snap)
        if not use_graph:
            return
        if self.world == 1:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._fwd_bwd()
                self._update()
        else:
            g_a, g_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_a, stream=side):
                self._fwd_bwd()
            self.bucket.all_reduce(self.group)
            with torch.cuda.graph(g_b, pool=g_a.pool(), stream=side):
                self._update()
            self.graphs = (g_a, g_b)
        torch.cuda.synchronize()

    def _snapshot(self):
        params = [p.detach().clone() for p in self.model.parameters()]
        bufs = [b.detach().clone() for b in self.model.buffers()]
        state = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                 for p, st in self.opt.state.items()}
        return params, bufs, state

    def _restore(self, snap):
        params, bufs, state = snap
        with torch.no_grad():
            for p, q in zip(self.model.parameters(), params):
                p.copy_(q)
            for b, q in zip(self.model.buffers(), bufs):
                b.copy_(q)
            for p, st in self.opt.state.items():
                old = state.get(id(p))
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()          # state created by the warm-up: back to "never stepped"
                    elif old is not None and k in old:
                        st[k] = old[k]
        if self.fused is not None:
            self.fused.resync()

    # ---- pieces of one step -------------------------------------------------------------------------
    def _make_grad_plan(self):
        """Route the weight gradients of the msmp modules through one unpack launch (gradsink.py); models without
        tensor-core packs (or foreign modules only) keep plain autograd accumulation."""
        from . import ops
        if ops.GEMM_MODE != "tc" or self.model.__dict__.get("_msmp_pack_plan") is None and not any(
                m.__dict__.get("_msmp_pack_plan") is not None for m in self.model.modules()):
            return
        from .gradsink import GradPlan
        # row counts of the step's weight-gradient GEMMs: the sink then holds split-M partials that the unpack launch sums
        tw = getattr(self.model, "time_window", None)
        self.gplan = GradPlan(self.model, n_nodes=int(self.static.x.shape[0]), n_edges=int(self.static.edge_index.shape[1]),
                              n_steps=int(tw) if tw else None)

    def _host_prologue(self):
        if self.fused is not None:
            self.fused.host_update()

    def _fwd_bwd(self):
        """forward, local squared error, backward with unit seed; leaves d SSE_local / d theta in the bucket and the
        squared error (fp64 split exactly into two floats) in its two trailing elements."""
        if self.gplan is None:
            self.bucket.zero_()          # with a GradPlan the covered gradients are overwritten, the others zeroed in begin()
        from . import ops
        y = self.static.y
        ops.RAW_OUTPUT = True          # the solvers hand over their float32 output; foreign modules ignore the switch
        try:
            pred = self.model(self.static)
        finally:
            ops.RAW_OUTPUT = False
        fused = (pred.dtype == torch.float32 and y.dtype == torch.float64 and pred.shape == y.shape and pred.is_contiguous()
                 and y.is_contiguous())
        if fused:          # train_helper.py:126 (reduction='sum') on float64 labels: msmp_sse_fwd / msmp_sse_bwd
            sse = ops.sse_loss(pred, y, self.bucket.tail)
        else:
            sse = ((pred.to(self.static.x.dtype) - y) ** 2).sum()
        if self.gplan is not None:
            self.gplan.begin()
        try:
            sse.backward()
        finally:
            if self.gplan is not None:
                self.gplan.finish()
        if not fused:
            s64 = sse.detach().double()
            hi = s64.float()
            self.bucket.tail[0] = hi
            self.bucket.tail[1] = (s64 - hi.double()).float()

    def _update(self):
        """loss = sqrt(SSE), gradient scale 1 / (2 loss), optimizer update (and p.grad <- gradient of the loss)."""
        check(lib.msmp_loss_scalars(self.bucket.tail.data_ptr(), self.loss.data_ptr(), self.gscale.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "msmp_loss_scalars")
        if self.fused is not None:
            self.fused.launch(self.gscale)
        else:
            self.bucket.flat[:self.bucket.n].mul_(self.gscale)
            self.opt.step()

    def eager(self, graph=None):
        """The same step launched kernel by kernel (no graph replay); used for per-kernel timing."""
        if graph is not None and graph is not self.static:
            self.load(graph)
        self._host_prologue()
        self._eager_step()
        return self.loss

    def _eager_step(self):
        self._fwd_bwd()
        if self.world > 1:
            self.bucket.all_reduce(self.group)
        self._update()

    def _check_topology(self, graph):
        ei = graph.edge_index
        if ei is self._topo_src[0] and graph.batch is self._topo_src[1]:
            return
        key = (ei.data_ptr(), ei._version)
        if key in self._topo_ok and ei.shape == self.static.edge_index.shape:
            return
        same = (ei.shape == self.static.edge_index.shape and graph.batch.shape == self.static.batch.shape
                and bool(torch.equal(ei.to(self.static.edge_index.device), self.static.edge_index))
                and bool(torch.equal(graph.batch.to(self.static.batch.device), self.static.batch)))
        if not same:
            raise ValueError("GraphedTrainStep: this graph's edge_index / batch differ from the captured topology; build "
                             "a new GraphedTrainStep for it")
        if len(self._topo_ok) > 256:
            self._topo_ok.clear()
        self._topo_ok.add(key)

    def load(self, graph):
        """Copy the floating-point fields of ``graph`` (host or device) into the static device buffers."""
        self._check_topology(graph)
        for k in self.fields:
            src = getattr(graph, k)
            dst = getattr(self.static, k)
            if src is dst:
                continue
            if src.shape != dst.shape:
                raise ValueError(f"graph.{k} has shape {tuple(src.shape)}, the captured step expects {tuple(dst.shape)}")
            dst.copy_(src, non_blocking=True)

    def prefetch(self, graph):
        """Start copying the NEXT step's inputs (e.g. pinned host tensors) into device-side staging buffers on a copy
        stream while the current step still runs; the next ``step()`` (called without a graph) moves them into the static
        buffers with device-to-device copies before it replays.  What a prefetching data loader does for the reference's
        loop (``graph.to(device)``, experiments/train_helper.py:99)."""
        self._check_topology(graph)
        if self._staging is None:
            self._staging = {k: torch.empty_like(getattr(self.static, k)) for k in self.fields}
            self._copy_stream = torch.cuda.Stream()
            self._staged, self._consumed = torch.cuda.Event(), None
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)          # the previous contents have been moved into the static buffers
        with torch.cuda.stream(cs):
            for k in self.fields:
                src = getattr(graph, k)
                if src.shape != self._staging[k].shape:
                    raise ValueError(f"graph.{k} has shape {tuple(src.shape)}, the captured step expects "
                                     f"{tuple(self._staging[k].shape)}")
                self._staging[k].copy_(src, non_blocking=True)
            self._staged.record(cs)
        self._have_staged = True

    def __call__(self, graph=None):
        if graph is not None and graph is not self.static:
            self.load(graph)
        elif graph is None and self._have_staged:
            torch.cuda.current_stream().wait_event(self._staged)
            for k in self.fields:
                getattr(self.static, k).copy_(self._staging[k], non_blocking=True)
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
            self._have_staged = False
        if self.fused is not None and not self.fused.valid():
            raise RuntimeError("GraphedTrainStep: parameter / gradient / optimizer-state storage was replaced after the "
                               "capture (e.g. optimizer.load_state_dict or zero_grad(set_to_none=True)); build a new step")
        self._host_prologue()
        if self.graph is not None:
            self.graph.replay()
        elif self.graphs is not None:
            self.graphs[0].replay()
            self.bucket.all_reduce(self.group)
            self.graphs[1].replay()
        else:
            self._eager_step()
        return self.loss


class GraphedForward:
    """The forward pass ``model(graph)`` of the evaluation rollouts (experiments/train_helper.py:150-292,362-471) as one
    CUDA-graph replay: on a fixed grid every window of an autoregressive rollout has the same topology, so the host only
    refreshes the floating-point fields (x, pos, the PDE parameters) of a static graph and replays.  The returned tensor is
    the captured output buffer: it is overwritten by the next call (the rollouts consume it before, through
    ``create_next_graph`` and the loss)."""

    def __init__(self, model, example, warmup: int = 2):
        self.model = model
        self.static = example.clone()
        self._topo_src = (example.edge_index, example.batch)
        self.fields = [k for k in self.static.keys()
                       if torch.is_tensor(getattr(self.static, k)) and getattr(self.static, k).is_floating_point()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                self.out = model(self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=side):
            self.out = model(self.static)
        torch.cuda.synchronize()

    def matches(self, graph) -> bool:
        ei, st = graph.edge_index, self.static
        if ei is self._topo_src[0] and graph.batch is self._topo_src[1]:
            return tuple(graph.x.shape) == tuple(st.x.shape)
        return (ei.shape == st.edge_index.shape and graph.batch.shape == st.batch.shape
                and tuple(graph.x.shape) == tuple(st.x.shape) and bool(torch.equal(ei.to(st.edge_index.device), st.edge_index))
                and bool(torch.equal(graph.batch.to(st.batch.device), st.batch)))

    def __call__(self, graph):
        for k in self.fields:
            src = getattr(graph, k)
            dst = getattr(self.static, k)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"GraphedForward: field {k} has shape {tuple(src.shape)}, captured {tuple(dst.shape)}")
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out
