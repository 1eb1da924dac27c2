"""Drop-in replacements for the classes of ``experiments/models_gnn2D.py`` (two-field models).

"2D" is a two-field 1-D system: node features are [N, 2*time_window] with the two fields in
consecutive column blocks (common/utils.py:350-354).
"""
from __future__ import annotations

import torch
from torch import nn

from .graph import get_topology
from .layers import GNN_Layer, GNN_LayerLin, H, NodeFeatures, Swish, gate_blend, gated_pair, prepare_packs  # noqa: F401
from . import ops
from .lem import LEM, LEMS, use_persistent  # noqa: F401
from .models_gnn import LSTM  # noqa: F401
from .solver import decode, linear_act, make_decoder, ops_raw_output, mlp2, pad_cols, require_cuda


def unflatten_u(u: torch.Tensor, time_window: int):
    """models_gnn2D.py:9-14: [*, n*time_window] -> [*, n, time_window]"""
    return u.unflatten(1, (u.size(1) // time_window, time_window))


class _Solver2F(nn.Module):
    layer_cls = GNN_LayerLin
    gated = True
    encoder = "lem"            # 'lem' (LEM, or LEMS when save_state) | 'lstm' | 'mlp'
    g2 = False                 # G^2-style gate (models_gnn2D.py:598-603)

    def __init__(self, pde, time_window: int = 25, hidden_features: int = 128, hidden_layer: int = 6,
                 eq_variables: dict = {}, save_state=None):
        super().__init__()
        assert time_window in (25, 50)               # models_gnn2D.py:322
        if hidden_features != H:
            raise ValueError("msmp_pde_b200 supports hidden_features = 128 (the reference's only value)")
        self.pde = pde
        self.out_features = time_window
        self.hidden_features, self.hidden_layer = hidden_features, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        self.save_state = save_state
        nv = len(eq_variables) + 1
        mk = lambda: self.layer_cls(hidden_features, hidden_features, hidden_features, 2 * time_window, nv)
        self.gnn_layers = nn.ModuleList(mk() for _ in range(hidden_layer))
        if self.gated:
            self.gnn_layers_gate = nn.ModuleList(mk() for _ in range(hidden_layer))
        f32 = dict(dtype=torch.float32)
        if self.encoder == "lem":
            lem_cls = LEM if save_state is None else LEMS          # models_gnn2D.py:358-363
            self.embedding_lem = lem_cls(2 + len(eq_variables) + 2, hidden_features)
            self.lemoutput_mlp = nn.Sequential(nn.Linear(hidden_features, hidden_features, **f32), Swish(),
                                               nn.Linear(hidden_features, hidden_features, **f32), Swish())
        elif self.encoder == "lstm":
            self.embedding_lstm = LSTM(2 + len(eq_variables) + 2, hidden_features)
            self.lstmoutput_mlp = nn.Sequential(nn.Linear(hidden_features, hidden_features, **f32), Swish(),
                                                nn.Linear(hidden_features, hidden_features, **f32), Swish())
        else:
            self.embedding_mlp = nn.Sequential(nn.Linear(2 * time_window + 2 + len(eq_variables), hidden_features, **f32),
                                               Swish(), nn.Linear(hidden_features, hidden_features, **f32), Swish())
        if self.gated:
            self.swish = Swish()
        self.double_mlp = nn.Sequential(nn.Linear(hidden_features, 2 * hidden_features, **f32), Swish(),
                                        nn.Unflatten(1, (2, hidden_features)))
        self.output_mlp = make_decoder(time_window, 2)

    def __repr__(self):
        return 'GNN'

    def _prepare_packs(self):
        layers = list(self.gnn_layers) + (list(self.gnn_layers_gate) if self.gated else [])
        if self.encoder == "lem":
            lem, linears = self.embedding_lem.rnn, [self.lemoutput_mlp[0], self.lemoutput_mlp[2]]
        elif self.encoder == "lstm":
            lem, linears = None, [self.lstmoutput_mlp[0], self.lstmoutput_mlp[2]]
        else:
            lem, linears = None, [self.embedding_mlp[0], self.embedding_mlp[2]]
        prepare_packs(self, layers, lem, linears + [self.double_mlp[0]])

    def _clock(self, device):
        """(cumsum(dt) in float64 [1, tw], the same rounded to float32): constants of the model, built once per device."""
        c = self.__dict__.get("_msmp_clock")
        key = (device, float(self.pde.dt), self.time_window)
        if c is None or c[0] != key:
            dt64 = torch.cumsum(torch.ones(1, self.time_window, dtype=torch.float64, device=device) * float(self.pde.dt), dim=1)
            c = (key, dt64, dt64.float())
            self.__dict__["_msmp_clock"] = c
        return c[1], c[2]

    def forward(self, data) -> torch.Tensor:
        tw = self.time_window
        u_in = data.x
        require_cuda(u_in)
        self._prepare_packs()
        pos = data.pos
        pos_x = pos[:, 1][:, None] / self.pde.L
        pos_t = pos[:, 0][:, None] / self.pde.tmax
        variables = pos_t
        if "a" in self.eq_variables:
            variables = torch.cat((variables, data.a / self.eq_variables["a"]), -1)
        if "b" in self.eq_variables:          # sic: data.a also feeds the 'b' column (models_gnn2D.py:419)
            variables = torch.cat((variables, data.a / self.eq_variables["b"]), -1)
        u = u_in.float().contiguous()
        N = u.shape[0]
        pos_xf, variables_f = pos_x.float(), variables.float()
        feat = NodeFeatures(u, pos_xf, variables_f)
        topo = get_topology(data.edge_index, data.batch, N)
        dt64, dt = self._clock(u.device)

        nvar = variables.shape[1] - 1
        if self.encoder == "lem" and use_persistent(4 + nvar):
            # I_t = [pos_x, u1[:, t], u2[:, t], cumsum(dt)_t + pos_t, variables[:, 1:]]  (models_gnn2D.py:421-433), written
            # as the recurrence's zero-padded [T, N, 32] slab by one launch
            cols = [("static", pos_xf, 0), ("time", u, 0), ("time", u, tw), ("clock",)]
            cols += [("static", variables_f, 1 + k) for k in range(nvar)]
            lem_in = ops.lem_inputs(tw, N, cols, clock=dt64.view(-1), node_t=pos_t.double().reshape(-1).contiguous())
            lem_in._msmp_lem_ninp = 4 + nvar
            h = mlp2(self.embedding_lem(lem_in), self.lemoutput_mlp)
        elif self.encoder in ("lem", "lstm"):
            lem_in = torch.empty(tw, N, 4 + nvar, dtype=torch.float32, device=u.device)
            lem_in[:, :, 0] = pos_x.float()[:, 0]
            lem_in[:, :, 1] = u[:, :tw].t()
            lem_in[:, :, 2] = u[:, tw:].t()
            lem_in[:, :, 3] = (dt64 + pos_t.double()).float().t()
            if nvar:
                lem_in[:, :, 4:] = variables[:, 1:].float()
            if self.encoder == "lem":
                h = mlp2(self.embedding_lem(lem_in), self.lemoutput_mlp)
            else:
                h = mlp2(self.embedding_lstm(lem_in).contiguous(), self.lstmoutput_mlp)
        else:
            h = mlp2(pad_cols(torch.cat((u_in, pos_x, variables), -1)), self.embedding_mlp)

        for i in range(self.hidden_layer):
            if self.g2:
                h = _g2_pair(self.gnn_layers_gate[i], self.gnn_layers[i], h, feat, topo)
            elif self.gated:
                h = gated_pair(self.gnn_layers_gate[i], self.gnn_layers[i], h, feat, topo)
            else:
                h = self.gnn_layers[i].forward_prepared(h, feat, topo)

        h2 = linear_act(h, self.double_mlp[0])                            # [N, 2*128] (models_gnn2D.py:444)
        out = decode(h2, self.output_mlp, u, dt, 2, tw)                   # models_gnn2D.py:448-458
        return out if ops_raw_output() else out.to(u_in.dtype)


class MP_PDE_Solver2DLEMLinGated(_Solver2F):
    """models_gnn2D.py:290-458 (`--model MSMP-PDE2D`; BASELINE configs 2-4)."""
    layer_cls, gated, encoder = GNN_LayerLin, True, "lem"


class _SourceMeanFn(torch.autograd.Function):
    """torch_scatter.scatter(src, edge_index[0], reduce='mean') on the deterministic segmented kernel
    (msmp_segment_reduce over the CSC permutation); backward is the matching gather."""

    @staticmethod
    def forward(ctx, src, topo):
        from . import ops
        outdeg = (topo.colptr[1:] - topo.colptr[:-1]).clamp(min=1).float()
        inv = 1.0 / outdeg
        ctx.topo, ctx.inv = topo, inv
        return ops.segment_reduce(src.contiguous(), topo.colptr, perm=topo.csc_perm, scale=inv, N=topo.N)

    @staticmethod
    def backward(ctx, g):
        topo = ctx.topo
        return (g * ctx.inv[:, None])[topo.src.long()], None


class _G2MeanFn(torch.autograd.Function):
    """mean over a node's out-edges of (t[src] - t[dst])^2 in one launch per direction (msmp_g2_fwd / msmp_g2_bwd): no
    [E,128] gathers, no atomics in the backward pass."""

    @staticmethod
    def forward(ctx, t, topo):
        outdeg = (topo.colptr[1:] - topo.colptr[:-1]).clamp(min=1).float()
        inv = (1.0 / outdeg).contiguous()
        t = t.contiguous()
        ctx.topo = topo
        ctx.save_for_backward(t, inv)
        return ops.g2_fwd(t, topo, inv)

    @staticmethod
    def backward(ctx, g):
        t, inv = ctx.saved_tensors
        return ops.g2_bwd(t, g, ctx.topo, inv), None


def _g2_pair(gate_layer, main_layer, h, feat, topo):
    """models_gnn2D.py:598-603: tau = tanh(mean over out-edges of |t_src - t_dst|^2), t = swish(gate layer output)."""
    from .layers import instance_norm
    t = gate_layer.forward_prepared(h, feat, topo)
    t = t * torch.sigmoid(t)
    tau = torch.tanh(_G2MeanFn.apply(t, topo))
    m = main_layer.forward_prepared(h, feat, topo)
    return (1 - tau) * h + tau * (m * torch.sigmoid(m))


class MP_PDE_Solver2D(_Solver2F):
    """models_gnn2D.py:17-141 (MLP encoder, plain GNN_Layer stack)."""
    layer_cls, gated, encoder = GNN_Layer, False, "mlp"


class MP_PDE_Solver2DGated(_Solver2F):
    """models_gnn2D.py:143-288"""
    layer_cls, gated, encoder = GNN_LayerLin, True, "mlp"


class MP_PDE_Solver2DLEMLinG2(_Solver2F):
    """models_gnn2D.py:460-620 (`--model MSG2-PDE2D`)."""
    layer_cls, gated, encoder, g2 = GNN_LayerLin, True, "lem", True


class MP_PDE_Solver2DLSTMLinGated(_Solver2F):
    """models_gnn2D.py:622-780"""
    layer_cls, gated, encoder = GNN_LayerLin, True, "lstm"


class MP_PDE_Solver2DLSTMLin(_Solver2F):
    """models_gnn2D.py:782-918"""
    layer_cls, gated, encoder = GNN_Layer, False, "lstm"


class MP_PDE_Solver2DLEMLin(_Solver2F):
    """models_gnn2D.py:920-1056"""
    layer_cls, gated, encoder = GNN_Layer, False, "lem"


# hidden_features = 164: not a multiple of the kernels' 128-channel block -> torch-operator implementation (glu.py)
from .glu import MP_PDE_Solver2DLEMLinGatedGLU  # noqa: E402,F401


class G_PDE_Solver2DLEMLinGated(nn.Module):
    """models_gnn2D.py:1058-1196 cannot be constructed in the reference either (RGATConv without num_relations)."""

    def __init__(self, *a, **k):
        raise NotImplementedError("G_PDE_Solver2DLEMLinGated is dead code in the reference (SURVEY.md section 2)")
