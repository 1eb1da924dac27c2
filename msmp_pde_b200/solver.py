"""Building blocks shared by the drop-in solver classes (encoder MLPs, decoder, input prep)."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .layers import H, NodeFeatures, Swish, _side_stream, pad32


def ops_raw_output() -> bool:
    """True while GraphedTrainStep wants the solver's float32 output uncast (ops.RAW_OUTPUT)."""
    return ops.RAW_OUTPUT

_EQ_ORDER_1F = ("alpha", "beta", "gamma", "bc_left", "bc_right", "c", "D", "r")


def variables_1field(data, pos_t, eq_variables):
    """models_gnn.py:250-266: [t/tmax, then parameters in this fixed order]; bc_left / bc_right are
    not divided by their maximum."""
    v = pos_t
    for k in _EQ_ORDER_1F:
        if k in eq_variables:
            col = getattr(data, k)
            if k not in ("bc_left", "bc_right"):
                col = col / eq_variables[k]
            v = torch.cat((v, col), -1)
    return v


class _LinearActFn(torch.autograd.Function):
    """y = act(x W^T + b) on msmp_linear_fwd / msmp_linear_wgrad.  x is [M, Kp] with Kp % 32 == 0 (zero padded)."""

    @staticmethod
    def forward(ctx, x, W, b, act: bool, packs, gsink=None):
        Nout, K = W.shape
        x = x.contiguous()
        Wt, Wd = packs
        z = torch.empty(x.shape[0], Nout, dtype=torch.float32, device=x.device) if act else None
        y = ops.linear_fwd([x], Wt, bias=b, Ypre=z, act=act)
        ctx.act, ctx.K, ctx.Wd, ctx.gsink = act, K, Wd, gsink
        ctx.save_for_backward(x, z)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, z = ctx.saved_tensors
        dy = dy.contiguous()
        dz = ops.mul_dswish(dy, z) if ctx.act else dy
        gs = ctx.gsink if (ops.GRAD_SINK is not None and ops.GEMM_MODE == "tc") else None
        if gs is not None:       # GraphedTrainStep: raw gradient stays in the sink, side stream joined once per step
            cur = torch.cuda.current_stream()
            wst = _side_stream(cur, x.device, "wgrad")
            ops.GRAD_SINK.streams.add(wst)
            wst.wait_stream(cur)
            x.record_stream(wst)
            dz.record_stream(wst)
            with torch.cuda.stream(wst):
                ops.linear_wgrad(x, dz, has_bias=True, dWt=gs.dWt, dWside=gs.dbs)
        else:
            dWt, dbs = ops.linear_wgrad(x, dz, has_bias=True)
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.K != x.shape[1]:
                raise RuntimeError("input gradient of a zero-padded linear layer is not needed on this path")
            dx = ops.linear_fwd([dz], ctx.Wd)       # W [Nout][K] is the reduction-major dgrad operand
        if gs is not None:
            return dx, None, None, None, None, None
        return dx, dWt[:ctx.K].t(), dbs[0], None, None, None


def _linear_packs(linear: nn.Linear, Kp: int):
    """(forward operand, dgrad operand): PackPlan images in tensor-core mode inside a solver, else k-major tensors
    rebuilt on every call (parameter versions cannot be trusted as a cache key, see packing.py)."""
    tcw = linear.__dict__.get("_msmp_tcw")
    if tcw is not None and ops.GEMM_MODE == "tc":
        return tcw.fwd, tcw.dgrad
    with torch.no_grad():
        Wd = linear.weight.detach()
        Wt = Wd.new_zeros(Kp, Wd.shape[0])
        Wt[:Wd.shape[1]] = Wd.t()
        return Wt, Wd.clone()


def linear_act(x, linear: nn.Linear, act: bool = True):
    return _LinearActFn.apply(x, linear.weight, linear.bias, act, _linear_packs(linear, x.shape[1]),
                              linear.__dict__.get("_msmp_gsink"))


def pad_cols(x: torch.Tensor) -> torch.Tensor:
    """Zero-pad the feature dimension to a multiple of 32 (fp32)."""
    n, k = x.shape
    kp = pad32(k)
    if kp == k and x.dtype == torch.float32:
        return x.contiguous()
    out = torch.zeros(n, kp, dtype=torch.float32, device=x.device)
    out[:, :k] = x
    return out


def mlp2(x, seq: nn.Sequential):
    """Sequential(Linear, Swish, Linear, Swish) -- embedding_mlp / lemoutput_mlp (models_gnn.py:201-206,1288-1292)."""
    return linear_act(linear_act(x, seq[0]), seq[2])


def make_decoder(time_window: int, channels: int):
    """Conv1d decoder geometry of models_gnn.py:208-224 (1 field) / models_gnn2D.py:382-391 (2 fields)."""
    f32 = dict(dtype=torch.float32)
    if time_window == 20 and channels == 1:
        return nn.Sequential(nn.Conv1d(channels, 8, 15, stride=4, **f32), Swish(), nn.Conv1d(8, channels, 10, stride=1, **f32))
    if time_window == 25:
        return nn.Sequential(nn.Conv1d(channels, 8, 16, stride=3, **f32), Swish(), nn.Conv1d(8, channels, 14, stride=1, **f32))
    if time_window == 50:
        return nn.Sequential(nn.Conv1d(channels, 8, 12, stride=2, **f32), Swish(), nn.Conv1d(8, channels, 10, stride=1, **f32))
    raise AssertionError("unsupported time_window")


class _DecoderFn(torch.autograd.Function):
    """out = base(u) + dt * Conv1d(Swish(Conv1d(h)))  on msmp_decoder_fwd / msmp_decoder_bwd."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2, b2, u, dt, geom, gsink=None):
        h = h.contiguous()
        out, za = ops.decoder_fwd(h, w1, b1, w2, b2, u, dt, geom)
        ctx.geom, ctx.gsink = geom, gsink
        ctx.save_for_backward(h, za, w1, w2, dt)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, za, w1, w2, dt = ctx.saved_tensors
        C, K1, S1, L1, K2, TW = ctx.geom
        gs = ctx.gsink if ops.GRAD_SINK is not None else None
        dh, dW = ops.decoder_bwd(dout.contiguous(), h, za, w1, w2, dt, ctx.geom, dW=gs.dW if gs is not None else None)
        if gs is not None:
            return dh, None, None, None, None, None, None, None, None
        n1 = 8 * C * K1
        n2 = C * 8 * K2
        return (dh, dW[:n1].view(8, C, K1), dW[n1:n1 + 8], dW[n1 + 8:n1 + 8 + n2].view(C, 8, K2),
                dW[n1 + 8 + n2:], None, None, None, None)


def decode(h, output_mlp: nn.Sequential, u, dt, channels: int, time_window: int):
    """h [N, channels*128] -> out [N, channels*time_window] (formula 10 of the paper + time stepping)."""
    c1, c2 = output_mlp[0], output_mlp[2]
    K1, S1, K2 = c1.kernel_size[0], c1.stride[0], c2.kernel_size[0]
    L1 = (H - K1) // S1 + 1
    geom = (channels, K1, S1, L1, K2, time_window)
    return _DecoderFn.apply(h, c1.weight, c1.bias, c2.weight, c2.bias, u, dt.reshape(-1).contiguous(), geom,
                            output_mlp.__dict__.get("_msmp_gsink"))


def cumulative_dt(pde, time_window, device):
    # cumulated in float64 like the reference (models_gnn.py:275-276), then rounded once to fp32
    return torch.cumsum(torch.ones(1, time_window, dtype=torch.float64, device=device) * float(pde.dt), dim=1).float()


def require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("msmp_pde_b200 models run on CUDA only: move the model and the Data object to a GPU "
                           "(there is no CPU fallback)")
