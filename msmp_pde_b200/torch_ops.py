"""``torch.ops.msmp.*``: the C-ABI entry points of the hot path as torch custom ops (north_star: "a thin C-ABI torch custom-op
layer").

``ops.py`` binds ``libmsmp_b200.so`` with ctypes and is what the modules call (no dispatcher overhead inside the captured
step).  This module registers the same calls with ``torch.library`` -- schemas, CUDA implementations that forward to
``ops.py`` and fake (meta) implementations -- so that they are visible to the dispatcher, ``torch.export`` / ``make_fx``
tracing and ``torch.library.opcheck``; a reference maintainer can call ``torch.ops.msmp.scatter_mean(...)`` the way the
reference calls ``torch_scatter.scatter`` (experiments/models_gnn2D.py:600-601) or ``lem_cuda.forward``
(experiments/models_gnn.py:290).  Tensors in, tensors out, current CUDA stream; CUDA only (no CPU kernels are registered,
so a CPU tensor fails in the dispatcher)."""
from __future__ import annotations

import torch
from torch.library import custom_op

from . import ops
from .compat import lem_cuda as _lem_cuda
from .graph import get_topology

H = 128


@custom_op("msmp::scatter_mean", mutates_args=(), device_types="cuda")
def scatter_mean(src: torch.Tensor, rowptr: torch.Tensor, perm: torch.Tensor | None, mean: bool) -> torch.Tensor:
    """out[n] = mean (or sum) of src[perm[k] or k] over k in [rowptr[n], rowptr[n+1])  -- msmp_segment_reduce: the
    deterministic replacement of torch_scatter.scatter(reduce='mean'); rows of 128 floats, int32 offsets."""
    n = rowptr.numel() - 1
    scale = None
    if mean:
        cnt = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(torch.float32)
        scale = 1.0 / cnt
    return ops.segment_reduce(src, rowptr, perm=perm, scale=scale, N=n)


@scatter_mean.register_fake
def _(src, rowptr, perm, mean):
    return src.new_empty(rowptr.numel() - 1, src.shape[1])


@custom_op("msmp::edge_mlp_scatter", mutates_args=(), device_types="cuda")
def edge_mlp_scatter(P: torch.Tensor, Q: torch.Tensor, edge_index: torch.Tensor, batch: torch.Tensor, W2: torch.Tensor,
                     b2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """agg[i] = mean_{e -> i} swish(swish(P[dst e] + Q[src e]) W2^T + b2), z2 = the second pre-activation per edge (CSR
    order) -- msmp_edge_ws_fwd: gather + second message layer on the tensor pipe + segmented mean in one kernel."""
    topo = get_topology(edge_index, batch, P.shape[0])
    agg, z2 = ops.edge_fwd(P, Q, topo, W2.t().contiguous(), b2, W2raw=W2.contiguous())
    return agg, z2


@edge_mlp_scatter.register_fake
def _(P, Q, edge_index, batch, W2, b2):
    return P.new_empty(P.shape[0], H), P.new_empty(edge_index.shape[1], H)


@custom_op("msmp::linear", mutates_args=(), device_types="cuda")
def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, swish_out: bool) -> torch.Tensor:
    """act(x weight^T + bias) with nn.Linear's weight layout [out, in] (in % 32 == 0) -- msmp_linear_tc_fwd."""
    return ops.linear_fwd([x.contiguous()], weight.t().contiguous(), bias=bias, act=swish_out)


@linear.register_fake
def _(x, weight, bias, swish_out):
    return x.new_empty(x.shape[0], weight.shape[0])


@custom_op("msmp::linear_wgrad", mutates_args=(), device_types="cuda")
def linear_wgrad(x: torch.Tensor, dy: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """(dW [out, in], db [out]) of y = x W^T + b -- msmp_wgrad_ws / msmp_linear_wgrad_tc2 (deterministic split-M)."""
    dWt, dbs = ops.linear_wgrad(x.contiguous(), dy.contiguous(), has_bias=True)
    return dWt.t().contiguous(), dbs[0].contiguous()


@linear_wgrad.register_fake
def _(x, dy):
    return x.new_empty(dy.shape[1], x.shape[1]), x.new_empty(dy.shape[1])


@custom_op("msmp::instance_norm", mutates_args=(), device_types="cuda")
def instance_norm(x: torch.Tensor, edge_index: torch.Tensor, batch: torch.Tensor) -> torch.Tensor:
    """PyG InstanceNorm(128) semantics (per graph and channel, biased variance, eps 1e-5, no affine) -- msmp_instnorm_fwd."""
    topo = get_topology(edge_index, batch, x.shape[0])
    out, _ = ops.instnorm_fwd(x.contiguous(), topo)
    return out


@instance_norm.register_fake
def _(x, edge_index, batch):
    return torch.empty_like(x)


@custom_op("msmp::lem_forward", mutates_args=(), device_types="cuda")
def lem_forward(inputs: torch.Tensor, weights: torch.Tensor, weights_lin_z: torch.Tensor, bias: torch.Tensor,
                bias_lin_z: torch.Tensor, y0: torch.Tensor, z0: torch.Tensor, dt: float) -> list[torch.Tensor]:
    """lem_cuda.forward (experiments/models_gnn.py:290): [all_y, all_z, 4 saved tensors] -- msmp_lem_tc_fwd."""
    return [t.clone() for t in _lem_cuda.forward(inputs, weights, weights_lin_z, bias, bias_lin_z, y0, z0, torch.tensor(dt))]


@lem_forward.register_fake
def _(inputs, weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt):
    T, N, ninp = inputs.shape
    Npad = (N + ops.LEM_TILE - 1) // ops.LEM_TILE * ops.LEM_TILE
    f = lambda *s: inputs.new_empty(*s, dtype=torch.float32)
    return [inputs.new_empty(T, N, H), inputs.new_empty(T, N, H), f(T + 1, N, H), f(T + 1, N, H), f(T, Npad, 4 * H),
            f(T, N, (ninp + 31) // 32 * 32)]


@custom_op("msmp::sse", mutates_args=(), device_types="cuda")
def sse(pred: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """sum((pred.double() - labels) ** 2) as a float64 scalar: MSELoss(reduction='sum') of the training loop
    (experiments/train_helper.py:126, train.py:413) on float32 predictions and float64 labels -- msmp_sse_fwd
    (deterministic block sums)."""
    return ops.sse_fwd(pred.contiguous(), labels.contiguous())


@sse.register_fake
def _(pred, labels):
    return pred.new_empty((), dtype=torch.float64)


@custom_op("msmp::sse_grad", mutates_args=(), device_types="cuda")
def sse_grad(pred: torch.Tensor, labels: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """d sse / d pred = 2 g (pred - labels), evaluated in float64 and rounded once to float32 -- msmp_sse_bwd."""
    return ops.sse_bwd(pred.contiguous(), labels.contiguous(), g)


@sse_grad.register_fake
def _(pred, labels, g):
    return torch.empty_like(pred)


def _sse_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _sse_backward(ctx, g):
    pred, labels = ctx.saved_tensors
    return torch.ops.msmp.sse_grad(pred, labels, g), None


sse.register_autograd(_sse_backward, setup_context=_sse_setup)


def names():
    return ["scatter_mean", "edge_mlp_scatter", "linear", "linear_wgrad", "instance_norm", "lem_forward", "sse", "sse_grad"]
