"""Message-passing layers on the msmp_b200 kernels: weight packing, autograd glue, drop-in modules.

``GNN_Layer`` / ``GNN_LayerLin`` keep the reference's constructor, forward signature and state_dict keys
(experiments/models_gnn.py:23-149).  The forward/backward math is Appendix A of SURVEY.md, executed by
``msmp_linear_*`` (per-node GEMMs), ``msmp_edge_*`` (gather + message MLP + deterministic segmented
mean), ``msmp_segment_reduce`` (by-source gradient scatter) and ``msmp_instnorm_*``.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .graph import Topology, get_topology

H = 128
SIDE_LD = 8          # [pos_x, v0..v_{V-1}, 0...] per node


_SIDE_STREAMS: dict = {}


def _side_stream(cur, device, tag):
    """A helper stream per (device, current stream, role); created once.  Weight-gradient streams get the lowest
    priority and everything else (gate layers) a high one: the dgrad chain is the critical path of the backward pass,
    so when both have thread blocks pending the chain's are scheduled first (GraphedTrainStep also runs its main
    stream at high priority; stream priorities are recorded in captured kernel nodes)."""
    if ops.SERIALIZE:          # per-kernel timing passes: everything on the caller's stream, one kernel at a time
        return cur
    key = (device.index, cur.cuda_stream, tag)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=0 if tag == "wgrad" else -1)
        _SIDE_STREAMS[key] = st
    return st


class Swish(nn.Module):
    """models_gnn.py:12-21 (kept for state_dict / Sequential index compatibility and CPU-side glue)."""

    def __init__(self, beta=1):
        super().__init__()
        self.beta = beta

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


def pad32(n: int) -> int:
    return (n + 31) // 32 * 32


class NodeFeatures:
    """Per-forward constant node inputs, shared by every layer of the stack:
    ``upad`` [N, pad32(F_u)] (zero padded u) and ``side`` [N, 8] = [pos_x, variables..., 0]."""

    def __init__(self, u: torch.Tensor, pos_x: torch.Tensor, variables: torch.Tensor):
        N, F_u = u.shape
        V = variables.shape[1]
        if 1 + V > SIDE_LD:
            raise ValueError("at most 7 equation variables supported")
        dev = u.device
        self.F_u, self.V, self.N = F_u, V, N
        if u.is_cuda and V >= 1 and all(t.dtype == torch.float32 for t in (u, pos_x, variables)):
            # one launch (msmp_node_features) instead of two fills and three slice copies
            self.upad, self.side = ops.node_features(u.contiguous(), pos_x.reshape(N, 1).contiguous(), variables.contiguous(),
                                                     pad32(F_u))
            return
        self.upad = torch.zeros(N, pad32(F_u), dtype=torch.float32, device=dev)
        self.upad[:, :F_u] = u
        self.side = torch.zeros(N, SIDE_LD, dtype=torch.float32, device=dev)
        self.side[:, 0:1] = pos_x
        self.side[:, 1:1 + V] = variables


class LayerPack:
    """Kernel-side layouts of one layer's weights (private caches; refreshed when a parameter changes)."""

    def __init__(self, W1, b1, W3, W2, W4, F_u, V):
        dev = W1.device
        Kp = H + pad32(F_u)
        W1xi, W1xj = W1[:, :H], W1[:, H:2 * H]
        W1u = W1[:, 2 * H:2 * H + F_u]
        W1p = W1[:, 2 * H + F_u:2 * H + F_u + 1]
        W1v = W1[:, 2 * H + F_u + 1:]
        Wpq_t = torch.zeros(Kp, 2 * H, dtype=torch.float32, device=dev)
        Wpq_t[:H, :H] = W1xi.t()
        Wpq_t[:H, H:] = W1xj.t()
        Wpq_t[H:H + F_u, :H] = W1u.t()
        Wpq_t[H:H + F_u, H:] = -W1u.t()
        side = torch.zeros(SIDE_LD, 2 * H, dtype=torch.float32, device=dev)
        side[0, :H] = W1p[:, 0]
        side[0, H:] = -W1p[:, 0]
        side[1:1 + V, :H] = W1v.t()
        self.Wpq_t, self.Wpq_side = Wpq_t, side
        self.bias_pq = torch.cat([b1, torch.zeros_like(b1)])
        self.W1hq = torch.cat([W1xi, W1xj], 0).contiguous()            # [256][128]: dgrad operand
        self.W2t = W2.t().contiguous()
        self.W3t = W3[:, :2 * H].t().contiguous()                      # [256][128]
        self.W3hx = W3[:, :2 * H].contiguous()                         # [128][256]: dgrad operand
        w3s = torch.zeros(SIDE_LD, H, dtype=torch.float32, device=dev)
        w3s[:V] = W3[:, 2 * H:].t()
        self.W3side = w3s
        self.W4t = W4.t().contiguous()
        self.W4d = W4.clone()                                          # [n][k]: dgrad operand
        self.W2d = W2.clone()                                          # [n][k]: edge-backward operand


def _pack_generation(pk):
    """Refresh count of the PackPlan a layer pack is a view into (None for packs built per call)."""
    plan = getattr(pk, "plan", None)
    return None if plan is None else plan.generation


class _Aux:
    """Non-tensor arguments of the layer function."""
    __slots__ = ("topo", "feat", "pack", "final", "gsink")

    def __init__(self, topo: Topology, feat: NodeFeatures, pack: LayerPack, final: bool, gsink=None):
        self.topo, self.feat, self.pack, self.final, self.gsink = topo, feat, pack, final, gsink


class _LayerCoreFn(torch.autograd.Function):
    """propagate() of one layer without the norm: h -> y (pre-InstanceNorm)."""

    @staticmethod
    def forward(ctx, h, W1, b1, W2, b2, W3, b3, W4, b4, aux: _Aux):
        pk, topo, ft = aux.pack, aux.topo, aux.feat
        V = ft.V
        h = h.contiguous()
        PQ = ops.linear_fwd([h, ft.upad], pk.Wpq_t, bias=pk.bias_pq, side=ft.side, r=1 + V, Wside=pk.Wpq_side)
        agg, z2 = ops.edge_fwd(PQ[:, :H], PQ[:, H:], topo, pk.W2t, b2, W2raw=W2)
        z3 = ops.linear_fwd([h, agg], pk.W3t, bias=b3, side=ft.side[:, 1:], r=V, Wside=pk.W3side)
        if aux.final:       # GNN_Layer: y = h + swish(z4)
            z4 = torch.empty_like(h)
            y = ops.linear_fwd([z3], pk.W4t, bias=b4, Ypre=z4, act=True, R=h, aswish=[1])
        else:               # GNN_LayerLin: y = z4
            z4 = None
            y = ops.linear_fwd([z3], pk.W4t, bias=b4, aswish=[1])
        ctx.aux = aux
        ctx.pack_gen = _pack_generation(pk)
        ctx.save_for_backward(h, PQ, z2, agg, z3, z4, W2, W4)
        return y

    @staticmethod
    def backward(ctx, dy):
        h, PQ, z2, agg, z3, z4, W2, W4 = ctx.saved_tensors
        aux = ctx.aux
        pk, topo, ft = aux.pack, aux.topo, aux.feat
        if ctx.pack_gen != _pack_generation(pk):
            raise RuntimeError("the model's packed weights were refreshed by a later forward pass before this backward pass ran: "
                               "run forward and backward of a step back to back (or keep one model per concurrent graph)")
        V, F_u, N = ft.V, ft.F_u, ft.N
        dev = h.device
        dy = dy.contiguous()
        dz4 = ops.mul_dswish(dy, z4) if aux.final else dy
        # The weight-gradient GEMMs do not feed the dgrad chain: they run on a side stream and overlap it (each of these
        # launches fills only part of the GPU at the reference's graph sizes).
        cur = torch.cuda.current_stream()
        wst = _side_stream(cur, h.device, "wgrad")
        tc = ops.GEMM_MODE == "tc"
        # GraphedTrainStep: raw weight gradients stay in the model-wide sink (gradsink.GradPlan), nothing is joined here
        plan = ops.GRAD_SINK
        if plan is not None and not (tc and topo.E > 0):
            raise RuntimeError("GraphedTrainStep's gradient sink needs the tensor-core path and a graph with edges")
        gs = aux.gsink if plan is not None else None
        if gs is not None:
            plan.streams.add(wst)

        def on_side(fn, *deps):
            wst.wait_stream(cur)
            for t in deps:
                t.record_stream(wst)
            with torch.cuda.stream(wst):
                return fn()

        def raw(name):
            return getattr(gs, name) if gs is not None else None

        # update_net_2
        dW4t, dW4s = on_side(lambda: ops.linear_wgrad(z3, dz4, xswish=True, has_bias=True, dWt=raw("dW4t"),
                                                      dWside=raw("dW4s")), z3, dz4)
        dz3 = ops.linear_fwd([dz4], pk.W4d, Zmul=z3)
        # update_net_1
        dW3t = raw("dW3t") if gs is not None else torch.empty(2 * H, H, dtype=torch.float32, device=dev)
        _, dW3s = on_side(lambda: ops.linear_wgrad(h, dz3, X1=agg, side=ft.side[:, 1:], r=V, has_bias=True, dWt=dW3t,
                                                   dWside=raw("dW3s")), h, agg, dz3, dW3t)
        dcat = ops.linear_fwd([dz3], pk.W3hx)                          # [N,256] = [dh (via x) | dagg]
        # message path
        dPQ = torch.empty(N, 2 * H, dtype=torch.float32, device=dev)
        if tc and topo.E > 0:
            dz1, a1, dz2 = ops.edge_bwd(PQ[:, :H], PQ[:, H:], topo, pk.W2d, z2, dcat[:, H:], dPQ[:, :H], defer_wgrad=True,
                                        W2raw=W2)
            dW2t_, db2s = on_side(lambda: ops.linear_wgrad(a1, dz2, has_bias=True, dWt=raw("dW2t"), dWside=raw("db2s")),
                                  a1, dz2)
            dW2, db2 = (None, None) if gs is not None else (dW2t_.t(), db2s[0])
        else:
            dz1, dW2, db2 = ops.edge_bwd(PQ[:, :H], PQ[:, H:], topo, pk.W2d, z2, dcat[:, H:], dPQ[:, :H], W2raw=W2)
        ops.segment_reduce(dz1, topo.colptr, perm=topo.csc_perm, out=dPQ[:, H:], N=N)
        Kp = H + ft.upad.shape[1]
        dWpq_t = raw("dWpq_t") if gs is not None else torch.empty(Kp, 2 * H, dtype=torch.float32, device=dev)
        _, dWs = on_side(lambda: ops.linear_wgrad(h, dPQ, X1=ft.upad, side=ft.side, r=1 + V, has_bias=True, dWt=dWpq_t,
                                                  dWside=raw("dWs")), dPQ, dWpq_t)
        dh = ops.linear_fwd([dPQ], pk.W1hq, R=dcat[:, :H])
        if aux.final:
            dh = dh + dy
        if gs is not None:
            return dh, None, None, None, None, None, None, None, None, None
        cur.wait_stream(wst)
        for t in (dW4t, dW4s, dW3t, dW3s, dWpq_t, dWs, dW2, db2):
            t.record_stream(cur)
        # gradients in parameter layout
        dWu = dWpq_t[H:H + F_u]
        dW1 = torch.cat([dWpq_t[:H, :H].t(), dWpq_t[:H, H:].t(), (dWu[:, :H] - dWu[:, H:]).t(),
                         (dWs[0:1, :H] - dWs[0:1, H:]).t(), dWs[1:1 + V, :H].t()], 1)
        db1 = dWs[1 + V, :H]
        dW3 = torch.cat([dW3t.t(), dW3s[:V].t()], 1)
        db3 = dW3s[V]
        # GNN_LayerLin: b4 sits directly in front of a non-affine InstanceNorm, its gradient is identically
        # zero (SURVEY.md appendix A "structural zero"); return exact zeros instead of rounding noise.
        db4 = dW4s[0] if aux.final else torch.zeros_like(dW4s[0])
        return dh, dW1, db1, dW2, db2, dW3, db3, dW4t.t(), db4, None


class _InstNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, topo: Topology):
        y = y.contiguous()
        out, stat = ops.instnorm_fwd(y, topo)
        ctx.topo = topo
        ctx.save_for_backward(y, stat)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, stat = ctx.saved_tensors
        return ops.instnorm_bwd(dout, y, ctx.topo, stat), None


class _GateBlendFn(torch.autograd.Function):
    """h_new = (1 - s) h + s * swish(IN(y_main)),  s = sigmoid(IN(y_gate))  (models_gnn.py:1365-1368)."""

    @staticmethod
    def forward(ctx, y_gate, y_main, h, topo: Topology):
        y_gate, y_main, h = y_gate.contiguous(), y_main.contiguous(), h.contiguous()
        out, stat = ops.instnorm_fwd(y_gate, topo, y1=y_main, h=h)
        ctx.topo = topo
        ctx.save_for_backward(y_gate, y_main, h, stat)
        return out

    @staticmethod
    def backward(ctx, dout):
        y_gate, y_main, h, stat = ctx.saved_tensors
        dyg, dym, dh = ops.instnorm_bwd(dout, y_gate, ctx.topo, stat, y1=y_main, h=h)
        return dyg, dym, dh, None


def instance_norm(y, topo):
    return _InstNormFn.apply(y, topo)


def gate_blend(y_gate, y_main, h, topo):
    return _GateBlendFn.apply(y_gate, y_main, h, topo)


class _LayerBase(nn.Module):
    """Shared implementation of GNN_Layer / GNN_LayerLin (parameters exactly as in the reference)."""
    final_swish = True

    def __init__(self, in_features: int, out_features: int, hidden_features: int, time_window: int, n_variables: int):
        super().__init__()
        if not (in_features == out_features == hidden_features == H):
            raise ValueError("the msmp_b200 kernels are specialised for in = out = hidden = 128 features "
                             "(every model in the reference uses 128, models_gnn.py:158)")
        self.in_features, self.out_features, self.hidden_features = in_features, out_features, hidden_features
        self.time_window, self.n_variables = time_window, n_variables
        f32 = dict(dtype=torch.float32)      # the reference's default dtype is float64; ours is explicit
        self.message_net_1 = nn.Sequential(nn.Linear(2 * in_features + time_window + 1 + n_variables, hidden_features, **f32), Swish())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, hidden_features, **f32), Swish())
        self.update_net_1 = nn.Sequential(nn.Linear(in_features + hidden_features + n_variables, hidden_features, **f32), Swish())
        if self.final_swish:
            self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features, **f32), Swish())
        else:
            self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features, **f32))
        self._plan_pack = None

    def _params(self):
        return (self.message_net_1[0].weight, self.message_net_1[0].bias, self.message_net_2[0].weight,
                self.message_net_2[0].bias, self.update_net_1[0].weight, self.update_net_1[0].bias,
                self.update_net_2[0].weight, self.update_net_2[0].bias)

    def pack(self):
        """Kernel-side weight layouts.  Tensor-core mode inside a solver: views into the model-wide PackPlan (re-packed
        by one launch per forward).  Otherwise built here with framework ops on every call -- parameter ``_version``
        counters cannot be used to cache them (fused optimizers and CUDA-graph replays do not move them)."""
        if self._plan_pack is not None and ops.GEMM_MODE == "tc":
            return self._plan_pack
        with torch.no_grad():
            W1, b1, W2, b2, W3, b3, W4, b4 = [p.detach() for p in self._params()]
            return LayerPack(W1, b1, W3, W2, W4, self.time_window, self.n_variables)

    def core(self, h, feat: NodeFeatures, topo: Topology):
        """h -> propagate(h) (before the norm)."""
        if feat.F_u != self.time_window or feat.V != self.n_variables:
            raise ValueError("node feature widths do not match the layer")
        aux = _Aux(topo, feat, self.pack(), self.final_swish, self.__dict__.get("_msmp_gsink"))
        return _LayerCoreFn.apply(h, *self._params(), aux)

    def forward_prepared(self, h, feat, topo):
        return instance_norm(self.core(h, feat, topo), topo)

    def forward(self, x, u, pos, variables, edge_index, batch):
        """Reference signature (models_gnn.py:61): returns norm(propagate(x)) in x's dtype."""
        if not x.is_cuda:
            raise RuntimeError("msmp_pde_b200 layers run on CUDA only (no CPU fallback)")
        topo = get_topology(edge_index, batch, x.shape[0])
        feat = NodeFeatures(u.float(), pos.float(), variables.float())
        out = self.forward_prepared(x.float(), feat, topo)
        return out.to(x.dtype)


class GNN_Layer(_LayerBase):
    """models_gnn.py:23-86"""
    final_swish = True


class GNN_LayerLin(_LayerBase):
    """models_gnn.py:88-149 (no final Swish, no residual)"""
    final_swish = False


def prepare_packs(model, layers, lem=None, linears=()) -> None:
    """Tensor-core mode: (re)build the model-wide PackPlan when parameters moved and re-pack every weight of the model
    with ONE launch (called at the start of each forward pass)."""
    if ops.GEMM_MODE != "tc":
        return
    from .packing import PackPlan
    plan = model.__dict__.get("_msmp_pack_plan")
    if plan is None or not plan.valid():
        plan = PackPlan(layers[0].message_net_1[0].weight.device)
        for l in layers:
            l._plan_pack = plan.add_layer(l)
        if lem is not None:
            lem.__dict__["_plan_pack"] = plan.add_lem(lem)
        for lin in linears:
            lin.__dict__["_msmp_tcw"] = plan.add_linear(lin)
        plan.finalize()
        model.__dict__["_msmp_pack_plan"] = plan
    plan.refresh()




def gated_pair(gate_layer, main_layer, h, feat, topo):
    """One gated layer pair (models_gnn.py:1365-1368): tau = sigmoid(norm(gate(h))), h' = (1-tau) h + tau sw(norm(main(h))).
    The two message-passing layers are independent until the blend and each fills only 50-100 of the 148 SMs at the
    reference's graph sizes, so the gate layer is issued on a side stream and overlaps the main layer (autograd replays
    the same stream assignment in the backward pass; CUDA-graph capture records the fork/join as parallel branches)."""
    cur = torch.cuda.current_stream()
    side = _side_stream(cur, h.device, "gate")
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        yg = gate_layer.core(h, feat, topo)
    ym = main_layer.core(h, feat, topo)
    cur.wait_stream(side)
    yg.record_stream(cur)
    h.record_stream(side)
    return gate_blend(yg, ym, h, topo)
