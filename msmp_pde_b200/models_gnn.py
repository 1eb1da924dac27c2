"""Drop-in replacements for the classes of ``experiments/models_gnn.py`` (1-field models).

Same class names, constructor arguments, ``forward(data)`` contract, ``__repr__`` and state_dict layout as
the reference, computed by the msmp_b200 CUDA kernels.  Parameters are fp32 (the
reference's are float64 because of ``temporal/solvers.py:10``); inputs of any float dtype are cast once at
entry and the result is returned in ``data.x.dtype`` so ``criterion(pred, graph.y)`` (train_helper.py:126)
and ``torch.cat((graph.x, pred), 1)`` (common/utils.py:448) behave as before.
"""
from __future__ import annotations

import torch
from torch import nn

from .graph import get_topology
from .layers import GNN_Layer, GNN_LayerLin, H, NodeFeatures, Swish, gate_blend, gated_pair, prepare_packs  # noqa: F401 (re-exported)
from . import ops
from .lem import LEM, LEMS, LEMcuda, use_persistent  # noqa: F401 (re-exported)
from .solver import (cumulative_dt, decode, linear_act, make_decoder, ops_raw_output, mlp2, pad_cols, require_cuda, variables_1field)


class _LSTMNoTF32Fn(torch.autograd.Function):
    """cuDNN LSTM with TF32 off in BOTH directions and the process-wide flag left as the caller set it: the forward and
    the backward each run inside ``torch.backends.cudnn.flags(allow_tf32=False)`` (a plain context manager around the
    module call would have closed before autograd runs the backward kernels)."""

    @staticmethod
    def forward(ctx, x, rnn, *params):
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False), torch.enable_grad():
            out, _ = rnn(x.detach())
        ctx.out, ctx.params = out, params
        return out.detach()

    @staticmethod
    def backward(ctx, gout):
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            grads = torch.autograd.grad(ctx.out, ctx.params, gout.contiguous(), allow_unused=True)
        ctx.out = None
        return (None, None) + tuple(grads)


class LSTM(nn.Module):
    """models_gnn.py:758-767 (cuDNN LSTM encoder of the LSTM variants; not on the BASELINE path).  fp32 parity needs
    cuDNN's TF32 RNN kernels off; that is scoped to this module's own forward / backward (_LSTMNoTF32Fn).  The inputs
    carry no gradient in the reference either (they are built from data)."""

    def __init__(self, ninp, nhid):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = nn.LSTM(ninp, nhid, dtype=torch.float32)

    def forward(self, input):
        output = _LSTMNoTF32Fn.apply(input.float(), self.rnn, *self.rnn.parameters())
        return output[-1]


class _Solver1F(nn.Module):
    """Shared skeleton of the 1-field solvers: encoder -> (gated) message passing stack -> Conv1d decoder."""
    layer_cls = GNN_Layer
    gated = False
    encoder = "mlp"            # 'mlp' | 'lem' | 'lems' (state kept across calls) | 'lstm'
    lem_mlp = False            # lemoutput_mlp / lstmoutput_mlp after the recurrent encoder
    diff_only = False          # MSSMP sub-network: return the decoder output without the time stepping

    def __init__(self, pde, time_window: int = 25, hidden_features: int = 128, hidden_layer: int = 6,
                 eq_variables: dict = {}):
        super().__init__()
        assert time_window in (20, 25, 50)           # models_gnn.py:176
        if hidden_features != H:
            raise ValueError("msmp_pde_b200 supports hidden_features = 128 (the reference's only value)")
        self.pde = pde
        self.out_features = time_window
        self.hidden_features, self.hidden_layer = hidden_features, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        nv = len(eq_variables) + 1
        mk = lambda: self.layer_cls(hidden_features, hidden_features, hidden_features, time_window, nv)
        self.gnn_layers = nn.ModuleList(mk() for _ in range(hidden_layer))
        if self.gated:
            self.gnn_layers_gate = nn.ModuleList(mk() for _ in range(hidden_layer))
        f32 = dict(dtype=torch.float32)
        if self.encoder == "mlp":
            self.embedding_mlp = nn.Sequential(nn.Linear(time_window + 2 + len(eq_variables), hidden_features, **f32),
                                               Swish(), nn.Linear(hidden_features, hidden_features, **f32), Swish())
        else:
            ninp = 2 + len(eq_variables) + 1
            mlp = lambda: nn.Sequential(nn.Linear(hidden_features, hidden_features, **f32), Swish(),
                                        nn.Linear(hidden_features, hidden_features, **f32), Swish())
            if self.encoder == "lstm":
                self.embedding_lstm = LSTM(ninp, hidden_features)
                if self.lem_mlp:
                    self.lstmoutput_mlp = mlp()
            else:
                self.embedding_lem = (LEMS if self.encoder == "lems" else LEM)(ninp, hidden_features)
                if self.lem_mlp:
                    self.lemoutput_mlp = mlp()
        if self.gated:
            self.swish = Swish()
        self.output_mlp = make_decoder(time_window, 1)

    def __repr__(self):
        return 'GNN'

    def _prepare_packs(self):
        layers = list(self.gnn_layers) + (list(self.gnn_layers_gate) if self.gated else [])
        if self.encoder == "mlp":
            lem, linears = None, [self.embedding_mlp[0], self.embedding_mlp[2]]
        elif self.encoder == "lstm":
            lem = None
            linears = [self.lstmoutput_mlp[0], self.lstmoutput_mlp[2]] if self.lem_mlp else []
        else:
            lem = self.embedding_lem.rnn
            linears = [self.lemoutput_mlp[0], self.lemoutput_mlp[2]] if self.lem_mlp else []
        prepare_packs(self, layers, lem, linears)

    def forward(self, data) -> torch.Tensor:
        u_in = data.x
        require_cuda(u_in)
        self._prepare_packs()
        pos = data.pos
        pos_x = pos[:, 1][:, None] / self.pde.L
        pos_t = pos[:, 0][:, None] / self.pde.tmax
        variables = variables_1field(data, pos_t, self.eq_variables)
        u = u_in.float().contiguous()
        pos_xf, variables_f = pos_x.float(), variables.float()
        feat = NodeFeatures(u, pos_xf, variables_f)
        topo = get_topology(data.edge_index, data.batch, u.shape[0])

        if self.encoder == "mlp":
            node_input = pad_cols(torch.cat((u_in, pos_x, variables), -1))
            h = mlp2(node_input, self.embedding_mlp)
        elif self.encoder != "lstm" and use_persistent(2 + variables.shape[1]):
            # I_t = [pos_x, u[:, t], variables]  (models_gnn.py:1357-1360), written as the recurrence's zero-padded
            # [T, N, 32] slab by one launch
            T, V = u.shape[1], variables.shape[1]
            cols = [("static", pos_xf, 0), ("time", u, 0)] + [("static", variables_f, k) for k in range(V)]
            lem_in = ops.lem_inputs(T, u.shape[0], cols)
            lem_in._msmp_lem_ninp = 2 + V
            h = self.embedding_lem(lem_in)
            if self.lem_mlp:
                h = mlp2(h, self.lemoutput_mlp)
        else:
            T = u.shape[1]
            static = torch.cat((pos_x, variables), -1).float()
            lem_in = torch.empty(T, u.shape[0], 2 + variables.shape[1], dtype=torch.float32, device=u.device)
            lem_in[:, :, 0] = static[:, 0]
            lem_in[:, :, 1] = u.t()
            lem_in[:, :, 2:] = static[:, 1:]
            if self.encoder == "lstm":
                h = self.embedding_lstm(lem_in).contiguous()
                if self.lem_mlp:
                    h = mlp2(h, self.lstmoutput_mlp)
            else:
                h = self.embedding_lem(lem_in)
                if self.lem_mlp:
                    h = mlp2(h, self.lemoutput_mlp)

        for i in range(self.hidden_layer):
            if self.gated:
                h = gated_pair(self.gnn_layers_gate[i], self.gnn_layers[i], h, feat, topo)
            else:
                h = self.gnn_layers[i].forward_prepared(h, feat, topo)

        if self.diff_only:             # MSSMP_PDE_Solver_sub returns diff (models_gnn.py:1676-1680)
            out = decode(h, self.output_mlp, torch.zeros_like(u), torch.ones(self.time_window, dtype=torch.float32, device=h.device), 1,
                         self.time_window)
            return out if ops_raw_output() else out.to(u_in.dtype)
        dt = cumulative_dt(self.pde, self.time_window, h.device)
        out = decode(h, self.output_mlp, u, dt, 1, self.time_window)     # models_gnn.py:278-279
        return out if ops_raw_output() else out.to(u_in.dtype)


class MP_PDE_Solver(_Solver1F):
    """models_gnn.py:151-281 (`--model MP-PDE`)."""
    layer_cls, gated, encoder = GNN_Layer, False, "mlp"


class MP_PDE_SolverLEM(_Solver1F):
    """models_gnn.py:365-497 (LEM encoder, plain GNN_Layer stack)."""
    layer_cls, gated, encoder, lem_mlp = GNN_Layer, False, "lem", False


class MP_PDE_SolverLEMLinGated(_Solver1F):
    """models_gnn.py:1220-1377 (`--model MSMP-PDE`)."""
    layer_cls, gated, encoder, lem_mlp = GNN_LayerLin, True, "lem", True


class MP_PDE_SolverLEMLin(_Solver1F):
    """models_gnn.py:619-756 (LEM + lemoutput_mlp, plain GNN_Layer stack)."""
    layer_cls, gated, encoder, lem_mlp = GNN_Layer, False, "lem", True


class MP_PDE_SolverLSTMLin(_Solver1F):
    """models_gnn.py:770-907 (cuDNN LSTM encoder + lstmoutput_mlp, plain stack)."""
    layer_cls, gated, encoder, lem_mlp = GNN_Layer, False, "lstm", True


class MP_PDE_SolverLSTMLinGated(_Solver1F):
    """models_gnn.py:909-1065"""
    layer_cls, gated, encoder, lem_mlp = GNN_LayerLin, True, "lstm", True


class MP_PDE_SolverGated(_Solver1F):
    """models_gnn.py:1067-1218 (MLP encoder, gated GNN_LayerLin stack)."""
    layer_cls, gated, encoder = GNN_LayerLin, True, "mlp"


class MP_PDE_SolverLEMLinGatedSave(_Solver1F):
    """models_gnn.py:1747-1904 (LEMS keeps (y, z) across calls; train_helper.py:144-145 resets it)."""
    layer_cls, gated, encoder, lem_mlp = GNN_LayerLin, True, "lems", True


class MSSMP_PDE_Solver_sub(_Solver1F):
    """models_gnn.py:1525-1682 (returns the decoder output only)."""
    layer_cls, gated, encoder, lem_mlp, diff_only = GNN_LayerLin, True, "lem", True, True


class MSSMP_PDE_Solver(nn.Module):
    """models_gnn.py:1684-1745: out = (1 - scale) * u_last + cumsum(dt) * scale * diff with two sub-networks."""

    def __init__(self, pde, time_window: int = 25, hidden_features: int = 128, hidden_layer: int = 6,
                 eq_variables: dict = {}):
        super().__init__()
        assert time_window in (20, 25, 50)
        self.pde, self.out_features = pde, time_window
        self.hidden_features, self.hidden_layer = hidden_features, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        self.diff = MSSMP_PDE_Solver_sub(pde, time_window, hidden_features, hidden_layer, eq_variables)
        self.scale = MSSMP_PDE_Solver_sub(pde, time_window, hidden_features, hidden_layer, eq_variables)

    def __repr__(self):
        return 'GNN'

    def forward(self, data) -> torch.Tensor:
        from . import ops
        raw, ops.RAW_OUTPUT = ops.RAW_OUTPUT, False          # the blend below is the reference's arithmetic in the input dtype
        try:
            scale = self.scale(data)
            diff = self.diff(data)
        finally:
            ops.RAW_OUTPUT = raw
        u = data.x
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=u.dtype, device=u.device) * self.pde.dt, dim=1)
        return (1 - scale) * u[:, -1:].expand(-1, self.time_window) + dt * (scale * diff)


# hidden_features = 164: not a multiple of the kernels' 128-channel block -> torch-operator implementation (glu.py)
from .glu import MP_PDE_SolverLEMLinGatedGLU  # noqa: E402,F401
