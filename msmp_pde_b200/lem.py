"""LEM (Long Expressive Memory) recurrent encoder -- replaces the absent native extension ``lem_cuda``.

Reference boundary: ``LEMFunction`` / ``LEMcuda`` / ``LEM`` / ``LEMS`` (experiments/models_gnn.py:285-361)
call ``lem_cuda.forward(inputs, weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt)`` and
``lem_cuda.backward`` (upstream tk-rusch/LEM ``src/lem_cuda``; source not in the reference tree).

Equations, per step t with X = [y_{t-1} | I_t]:
    G = X W^T + b -> (G0, G1, G2);  dt_bar = dt*sigmoid(G0);  dt_z = dt*sigmoid(G1)
    z_t = (1 - dt_z) z_{t-1} + dt_z tanh(G2)
    y_t = (1 - dt_bar) y_{t-1} + dt_bar tanh([z_t | I_t] Wz^T + bz)
Column order of ``weights`` / ``weights_lin_z`` is [state(H) | input(ninp)]; this (and the chunk-to-gate
assignment) cannot be verified without the lem_cuda source and is documented in DESIGN.md.
No gradient flows to the inputs (LEMFunction.backward returns None for them, models_gnn.py:302).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import ops
from .layers import H, _side_stream, pad32


def make_packs(W, Wz, ninp):
    """(Wt, Wzt, Wh, Wzh, Wt_h, Wzt_h, Wt_in, Wzt_in) from the raw fp32 parameters (framework ops, rebuilt per call):
    k-major packs W^T / Wz^T (rows [state | input]), dgrad operands W[:, :H] / Wz[:, :H], state rows and input rows of
    the k-major packs."""
    with torch.no_grad():
        W, Wz = W.detach(), Wz.detach()
        ip = pad32(ninp)
        Wt = W.new_zeros(H + ip, 3 * H)
        Wt[:H] = W[:, :H].t()
        Wt[H:H + ninp] = W[:, H:].t()
        Wzt = W.new_zeros(H + ip, H)
        Wzt[:H] = Wz[:, :H].t()
        Wzt[H:H + ninp] = Wz[:, H:].t()
        return (Wt, Wzt, W[:, :H].contiguous(), Wz[:, :H].contiguous(), Wt[:H].contiguous(), Wzt[:H].contiguous(),
                Wt[H:], Wzt[H:])


def use_persistent(ninp: int) -> bool:
    return ops.GEMM_MODE == "tc" and ops.LEM_PERSISTENT and pad32(ninp) == 32 and ninp <= 8


def lem_forward(inputs, bias, bias_lin_z, y0, z0, dt, packs, ninp=None):
    """All T steps.  Returns (inp [T,N,32k] zero-padded inputs, Y, Z [T+1,N,128] with slab 0 = initial state, gates,
    persistent).  tensor-core mode: one persistent kernel (msmp_lem_tc_fwd); otherwise one msmp_linear_fwd + one fused
    gate kernel per GEMM and step.  ``ninp`` < inputs.shape[2] = pad32(ninp): the slab arrives zero padded already
    (ops.lem_inputs) and is used as it is."""
    T, N, width = inputs.shape
    ninp = width if ninp is None else ninp
    dev = inputs.device
    ip = pad32(ninp)
    if width == ip and inputs.dtype == torch.float32 and inputs.is_contiguous():
        inp = inputs
    else:
        inp = torch.zeros(T, N, ip, dtype=torch.float32, device=dev)
        inp[:, :, :ninp] = inputs
    Wt, Wzt, Wh, Wzh, Wt_h, Wzt_h, Wt_in, Wzt_in = packs
    Y = torch.empty(T + 1, N, H, dtype=torch.float32, device=dev)
    Z = torch.empty(T + 1, N, H, dtype=torch.float32, device=dev)
    Y[0], Z[0] = y0, z0
    persistent = use_persistent(ninp)
    if persistent:
        gates = ops.lem_tc_fwd(inp, ninp, Wt_in, Wzt_in, Wt_h, Wzt_h, bias, bias_lin_z, Y, Z, dt)
    else:
        gates = torch.empty(T, 4, N, H, dtype=torch.float32, device=dev)   # dt_bar, dt_z, tanh(G2), tanh(L)
        G = torch.empty(N, 3 * H, dtype=torch.float32, device=dev)
        L = torch.empty(N, H, dtype=torch.float32, device=dev)
        for t in range(T):
            ops.linear_fwd([Y[t], inp[t]], Wt, bias=bias, out=G)
            ops.lem_gate_z(G, Z[t], dt, gates[t], Z[t + 1])
            ops.linear_fwd([Z[t + 1], inp[t]], Wzt, bias=bias_lin_z, out=L)
            ops.lem_gate_y(L, Y[t], gates[t], Y[t + 1])
    return inp, Y, Z, gates, persistent


def lem_backward(inp, Y, Z, gates, gY, gZ, dt, packs, persistent, last_only, gs=None):
    """Reverse recurrence + the four weight gradients.  Returns (dWt [160,384], dWzt [160,128], dbias [1,384],
    dbz [1,128], dy0, dz0, joined): k-major raw gradients (rows [state | input, zero padded]); ``joined`` is False
    when they were left on the weight-gradient side stream for the gradient sink ``gs`` to collect."""
    Wh, Wzh = packs[2], packs[3]
    T, N, ip = inp.shape
    dev = inp.device
    gY, gZ = gY.contiguous(), gZ.contiguous()
    dG = torch.empty(T, N, 3 * H, dtype=torch.float32, device=dev)
    dL = torch.empty(T, N, H, dtype=torch.float32, device=dev)
    Kp = H + ip
    if gs is not None:       # GraphedTrainStep: raw gradients stay in the sink, the side stream is joined once per step
        dWt, dWzt, dbias, dbz = gs.dWt, gs.dWzt, gs.dbias, gs.dbz
    else:
        dWt = torch.empty(Kp, 3 * H, dtype=torch.float32, device=dev)
        dWzt = torch.empty(Kp, H, dtype=torch.float32, device=dev)
        dbias = torch.empty(1, 3 * H, dtype=torch.float32, device=dev)
        dbz = torch.empty(1, H, dtype=torch.float32, device=dev)

    def wgrads(t0, t1, accumulate):
        """weight gradients of steps [t0, t1) (M = (t1 - t0) * N rows per GEMM)"""
        rows = (t1 - t0) * N
        inpf = inp[t0:t1].view(rows, ip)
        ops.linear_wgrad(Y[t0:t1].view(rows, H), dG[t0:t1].view(rows, 3 * H), X1=inpf, has_bias=True, dWt=dWt,
                         dWside=dbias, accumulate=accumulate)
        ops.linear_wgrad(Z[1 + t0:1 + t1].view(rows, H), dL[t0:t1].view(rows, H), X1=inpf, has_bias=True, dWt=dWzt,
                         dWside=dbz, accumulate=accumulate)

    if persistent:
        # The recurrence can be cut into LEM_BWD_SEGMENTS launches so that the weight gradient GEMMs of the steps
        # already walked run on a side stream next to the remaining steps.
        state = ops.lem_tc_bwd_state(gates)
        cur = torch.cuda.current_stream()
        wst = _side_stream(cur, dev, "wgrad")
        # GraphedTrainStep's gradient sink holds split-M partials sized for the whole recurrence: one segment there
        nseg = 1 if gs is not None else max(1, min(ops.LEM_BWD_SEGMENTS, T))
        bounds = [T * i // nseg for i in range(nseg + 1)]
        for i in range(nseg - 1, -1, -1):
            t0, t1 = bounds[i], bounds[i + 1]
            ops.lem_tc_bwd(Wzh, Wh, Y, Z, gates, gY, gZ, last_only, dG, dL, dt, N, state, t0, t1)
            wst.wait_stream(cur)
            with torch.cuda.stream(wst):
                wgrads(t0, t1, accumulate=i != nseg - 1)
        for t_ in (inp, Y, Z, dG, dL, dWt, dWzt, dbias, dbz):
            t_.record_stream(wst)
        dy, dz = state[0][:N], state[1][:N]
        if gs is not None:
            ops.GRAD_SINK.streams.add(wst)
            return dWt, dWzt, dbias, dbz, dy, dz, False
        cur.wait_stream(wst)
        return dWt, dWzt, dbias, dbz, dy, dz, True
    if last_only:
        gy_full = torch.zeros(T, N, H, dtype=torch.float32, device=dev)
        gz_full = torch.zeros(T, N, H, dtype=torch.float32, device=dev)
        gy_full[T - 1], gz_full[T - 1] = gY, gZ
        gY, gZ = gy_full, gz_full
    dy = torch.zeros(N, H, dtype=torch.float32, device=dev)      # carried d/dy_t
    dz = torch.zeros(N, H, dtype=torch.float32, device=dev)      # carried d/dz_t
    dz_tot = torch.empty(N, H, dtype=torch.float32, device=dev)
    for t in range(T - 1, -1, -1):
        # through y_t = (1-a) y_{t-1} + a tanh(L):  dL, dG0, dy <- dy*(1-a)
        ops.lem_bwd_y(dy, gY[t], Y[t], gates[t], dt, dL[t], dG[t])
        # dz_t total = carried + dL Wz[:, :H]  (+ external gZ[t], added inside lem_bwd_z)
        ops.linear_fwd([dL[t]], Wzh, R=dz, out=dz_tot)
        # through z_t = (1-b) z_{t-1} + b tanh(G2): dG1, dG2, dz <- d*(1-b)
        ops.lem_bwd_z(dz_tot, gZ[t], Z[t], gates[t], dt, dG[t], dz)
        # dy_{t-1} += dG W[:, :H]
        ops.linear_fwd([dG[t]], Wh, R=dy, out=dy)
    wgrads(0, T, accumulate=False)
    return dWt, dWzt, dbias, dbz, dy, dz, True


class _LEMFn(torch.autograd.Function):
    """All T steps (lem_forward / lem_backward).  ``last_only`` returns (y_T, z_T) instead of the whole history (what
    LEM / LEMS consume, models_gnn.py:340-342,354-357)."""

    @staticmethod
    def forward(ctx, inputs, weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt, packs, last_only, gsink=None, ninp=None):
        T, N, width = inputs.shape
        ninp = width if ninp is None else ninp
        inp, Y, Z, gates, persistent = lem_forward(inputs, bias, bias_lin_z, y0, z0, dt, packs, ninp)
        ctx.save_for_backward(inp, Y, Z, gates)
        ctx.dt, ctx.ninp, ctx.packs, ctx.persistent, ctx.last_only = dt, ninp, packs, persistent, last_only
        ctx.gsink = gsink
        if last_only:
            return Y[T], Z[T]
        return Y[1:], Z[1:]

    @staticmethod
    def backward(ctx, gY, gZ):
        inp, Y, Z, gates = ctx.saved_tensors
        ninp = ctx.ninp
        gs = ctx.gsink if (ops.GRAD_SINK is not None and ctx.persistent) else None
        dWt, dWzt, dbias, dbz, dy, dz, joined = lem_backward(inp, Y, Z, gates, gY, gZ, ctx.dt, ctx.packs, ctx.persistent,
                                                             ctx.last_only, gs)
        if not joined:
            return None, None, None, None, None, dy, dz, None, None, None, None, None
        dW = torch.cat([dWt[:H].t(), dWt[H:H + ninp].t()], 1)
        dWz = torch.cat([dWzt[:H].t(), dWzt[H:H + ninp].t()], 1)
        return None, dW, dWz, dbias[0], dbz[0], dy, dz, None, None, None, None, None


class LEMcuda(nn.Module):
    """models_gnn.py:305-330 -- same parameter names/shapes/init; fp32 parameters; device follows .to()."""

    def __init__(self, ninp, nhid, dt):
        super().__init__()
        if nhid != H:
            raise ValueError("msmp_b200 LEM kernels are specialised for nhid = 128")
        self.ninp, self.nhid = ninp, nhid
        f32 = dict(dtype=torch.float32)
        self.weights = nn.Parameter(torch.empty(3 * nhid, ninp + nhid, **f32))
        self.weights_lin_z = nn.Parameter(torch.empty(nhid, ninp + nhid, **f32))
        self.bias = nn.Parameter(torch.empty(3 * nhid, **f32))
        self.bias_lin_z = nn.Parameter(torch.empty(nhid, **f32))
        self.dt = float(dt)
        self.reset_parameters()

    def packs(self):
        """(Wt, Wzt, Wh, Wzh, Wt_h, Wzt_h, Wt_in, Wzt_in): k-major packs W^T / Wz^T (rows [state | input]), dgrad
        operands W[:, :H] / Wz[:, :H], state rows and input rows of the k-major packs.  Tensor-core mode inside a
        solver: images from the model-wide PackPlan; otherwise rebuilt here on every call."""
        pk = self.__dict__.get("_plan_pack")
        if pk is not None and use_persistent(self.ninp):
            return (None, None, pk.Wh, pk.Wzh, pk.Wt_h, pk.Wzt_h, pk.Wt_in, pk.Wzt_in)
        return make_packs(self.weights, self.weights_lin_z, self.ninp)

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.nhid)
        for w in self.parameters():
            w.data.uniform_(-stdv, +stdv)

    def forward(self, input, states=None, last_only=False):
        if not input.is_cuda:
            raise RuntimeError("msmp_pde_b200 LEM runs on CUDA only (no CPU fallback)")
        x = input.detach().float().contiguous()
        # a slab assembled by ops.lem_inputs arrives zero padded to 32 columns and says how many of them are inputs
        padded = getattr(input, "_msmp_lem_ninp", None)
        if padded is not None and (padded != self.ninp or x.shape[2] != pad32(self.ninp)):
            raise ValueError("padded LEM input slab does not match this module's ninp")
        if padded is None and x.shape[2] != self.ninp:
            raise ValueError(f"LEM input has {x.shape[2]} features, the module expects {self.ninp}")
        if states is None:
            y = x.new_zeros(x.size(1), self.nhid)
            z = x.new_zeros(x.size(1), self.nhid)
        else:
            y, z = states[0].float().contiguous(), states[1].float().contiguous()
        return _LEMFn.apply(x, self.weights, self.weights_lin_z, self.bias, self.bias_lin_z, y, z, self.dt,
                            self.packs(), last_only, self.__dict__.get("_msmp_gsink"), self.ninp)


class LEM(nn.Module):
    """models_gnn.py:333-342"""

    def __init__(self, ninp, nhid, dt=1.):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = LEMcuda(ninp, nhid, dt)

    def forward(self, input):
        y_last, _ = self.rnn(input, last_only=True)
        return y_last


class LEMS(nn.Module):
    """models_gnn.py:345-361 -- carries (y_T, z_T) into the next call until reset_states()."""

    def __init__(self, ninp, nhid, dt=1.):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = LEMcuda(ninp, nhid, dt)
        self.states = None

    def forward(self, input):
        y_last, z_last = self.rnn(input, self.states, last_only=True)
        self.states = (y_last, z_last)
        return y_last

    def reset_states(self):
        self.states = None
