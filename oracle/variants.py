"""TEST INFRASTRUCTURE ONLY -- CPU float64 restatement of the reference's variant solver classes.

Every variant recombines the blocks of ``oracle/models.py`` (layers, LEM, decoder) with a different
encoder ('mlp' | 'lem' | 'lems' | 'lstm'), an optional output MLP after the recurrent encoder, and a
plain / sigmoid-gated / G^2-gated processor.  Each class cites the reference definition it follows;
all of them are pinned against fixtures written from the reference's own classes
(tests/golden/var_*.npz, tests/test_oracle_golden.py).  Never imported by the product.
"""
from __future__ import annotations

import torch
from torch import nn

from .models import (GNN_Layer, GNN_LayerLin, LEM, LEMS, Swish, _decoder_1field, unflatten_u, variables_1field)
from .pyg_semantics import scatter_mean


class LSTM(nn.Module):
    """models_gnn.py:758-767 (last output of a single-layer nn.LSTM over [T, N, ninp])."""

    def __init__(self, ninp, nhid):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = nn.LSTM(ninp, nhid)

    def forward(self, input):
        return self.rnn(input)[0][-1]


def _mlp(h):
    return nn.Sequential(nn.Linear(h, h), Swish(), nn.Linear(h, h), Swish())


class _Processor(nn.Module):
    layer_cls = GNN_LayerLin
    gate = "sigmoid"           # None | 'sigmoid' | 'g2'

    def _make_layers(self, H, L, F_u, nv):
        mk = lambda: self.layer_cls(H, H, H, F_u, nv)
        self.gnn_layers = nn.ModuleList(mk() for _ in range(L))
        if self.gate:
            self.gnn_layers_gate = nn.ModuleList(mk() for _ in range(L))
            self.swish = Swish()

    def _process(self, h, u, pos_x, variables, edge_index, batch):
        for i in range(self.hidden_layer):
            if not self.gate:
                h = self.gnn_layers[i](h, u, pos_x, variables, edge_index, batch)
                continue
            t = self.gnn_layers_gate[i](h, u, pos_x, variables, edge_index, batch)
            if self.gate == "sigmoid":           # models_gnn.py:1365-1368
                tau = torch.sigmoid(t)
            else:                                # models_gnn2D.py:598-603
                t = self.swish(t)
                tau = torch.tanh(scatter_mean((t[edge_index[0]] - t[edge_index[1]]).abs() ** 2, edge_index[0],
                                              t.shape[0]))
            h = (1 - tau) * h + tau * self.swish(self.gnn_layers[i](h, u, pos_x, variables, edge_index, batch))
        return h


class _Solver1F(_Processor):
    encoder = "lem"
    out_mlp = True
    diff_only = False

    def __init__(self, pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables={}):
        super().__init__()
        assert time_window in (20, 25, 50)
        H = hidden_features
        self.pde, self.out_features, self.hidden_features, self.hidden_layer = pde, time_window, H, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        self._make_layers(H, hidden_layer, time_window, len(eq_variables) + 1)
        ninp = 2 + len(eq_variables) + 1
        if self.encoder == "mlp":
            self.embedding_mlp = nn.Sequential(nn.Linear(time_window + 2 + len(eq_variables), H), Swish(),
                                               nn.Linear(H, H), Swish())
        elif self.encoder == "lstm":
            self.embedding_lstm = LSTM(ninp, H)
            if self.out_mlp:
                self.lstmoutput_mlp = _mlp(H)
        else:
            self.embedding_lem = (LEMS if self.encoder == "lems" else LEM)(ninp, H)
            if self.out_mlp:
                self.lemoutput_mlp = _mlp(H)
        self.output_mlp = _decoder_1field(time_window)

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        u = data.x
        pos_x = data.pos[:, 1][:, None] / self.pde.L
        pos_t = data.pos[:, 0][:, None] / self.pde.tmax
        variables = variables_1field(data, pos_t, self.eq_variables)
        if self.encoder == "mlp":
            h = self.embedding_mlp(torch.cat((u, pos_x, variables), -1))
        else:
            seq = torch.stack([torch.cat((pos_x, u[:, t:t + 1], variables), -1) for t in range(u.shape[1])])
            if self.encoder == "lstm":
                h = self.embedding_lstm(seq)
                h = self.lstmoutput_mlp(h) if self.out_mlp else h
            else:
                h = self.embedding_lem(seq)
                h = self.lemoutput_mlp(h) if self.out_mlp else h
        h = self._process(h, u, pos_x, variables, data.edge_index, data.batch)
        diff = self.output_mlp(h[:, None]).squeeze(1)
        if self.diff_only:
            return diff
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=h.dtype, device=h.device) * self.pde.dt, dim=1)
        return u[:, -1].repeat(self.time_window, 1).transpose(0, 1) + dt * diff


class MP_PDE_SolverLEM(_Solver1F):
    """models_gnn.py:469-617"""
    layer_cls, gate, encoder, out_mlp = GNN_Layer, None, "lem", False


class MP_PDE_SolverLEMLin(_Solver1F):
    """models_gnn.py:619-756"""
    layer_cls, gate, encoder = GNN_Layer, None, "lem"


class MP_PDE_SolverLSTMLin(_Solver1F):
    """models_gnn.py:770-907"""
    layer_cls, gate, encoder = GNN_Layer, None, "lstm"


class MP_PDE_SolverLSTMLinGated(_Solver1F):
    """models_gnn.py:909-1065"""
    encoder = "lstm"


class MP_PDE_SolverGated(_Solver1F):
    """models_gnn.py:1067-1218"""
    encoder = "mlp"


class MP_PDE_SolverLEMLinGatedSave(_Solver1F):
    """models_gnn.py:1747-1904"""
    encoder = "lems"


class MSSMP_PDE_Solver_sub(_Solver1F):
    """models_gnn.py:1525-1682"""
    diff_only = True


class MSSMP_PDE_Solver(nn.Module):
    """models_gnn.py:1684-1745"""

    def __init__(self, pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables={}):
        super().__init__()
        self.pde, self.time_window = pde, time_window
        self.diff = MSSMP_PDE_Solver_sub(pde, time_window, hidden_features, hidden_layer, eq_variables)
        self.scale = MSSMP_PDE_Solver_sub(pde, time_window, hidden_features, hidden_layer, eq_variables)

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        scale, diff = self.scale(data), self.diff(data)
        u = data.x
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=u.dtype, device=u.device) * self.pde.dt, dim=1)
        return (1 - scale) * u[:, -1].repeat(self.time_window, 1).transpose(0, 1) + dt * (scale * diff)


class _Solver2F(_Processor):
    encoder = "lem"

    def __init__(self, pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables={}, save_state=None):
        super().__init__()
        assert time_window in (25, 50)
        H = hidden_features
        self.pde, self.out_features, self.hidden_features, self.hidden_layer = pde, time_window, H, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        self._make_layers(H, hidden_layer, 2 * time_window, len(eq_variables) + 1)
        ninp = 2 + len(eq_variables) + 2
        if self.encoder == "mlp":
            self.embedding_mlp = nn.Sequential(nn.Linear(2 * time_window + 2 + len(eq_variables), H), Swish(),
                                               nn.Linear(H, H), Swish())
        elif self.encoder == "lstm":
            self.embedding_lstm = LSTM(ninp, H)
            self.lstmoutput_mlp = _mlp(H)
        else:
            self.embedding_lem = LEM(ninp, H)
            self.lemoutput_mlp = _mlp(H)
        self.double_mlp = nn.Sequential(nn.Linear(H, 2 * H), Swish(), nn.Unflatten(1, (2, H)))
        if time_window == 25:
            self.output_mlp = nn.Sequential(nn.Conv1d(2, 8, 16, stride=3), Swish(), nn.Conv1d(8, 2, 14, stride=1))
        else:
            self.output_mlp = nn.Sequential(nn.Conv1d(2, 8, 12, stride=2), Swish(), nn.Conv1d(8, 2, 10, stride=1))

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        tw = self.time_window
        u = data.x
        pos_x = data.pos[:, 1][:, None] / self.pde.L
        pos_t = data.pos[:, 0][:, None] / self.pde.tmax
        variables = pos_t
        if "a" in self.eq_variables:
            variables = torch.cat((variables, data.a / self.eq_variables["a"]), -1)
        if "b" in self.eq_variables:      # sic: data.a (models_gnn2D.py:116, :419)
            variables = torch.cat((variables, data.a / self.eq_variables["b"]), -1)
        dt = torch.cumsum(torch.ones(1, 1, tw, dtype=u.dtype, device=u.device) * self.pde.dt, dim=2)
        if self.encoder == "mlp":
            h = self.embedding_mlp(torch.cat((u, pos_x, variables), -1))
        else:
            ts = (dt + pos_t).squeeze(0)
            seq = torch.stack([
                torch.cat((pos_x, u[:, t:t + 1], u[:, t + tw:t + tw + 1], ts[:, t:t + 1], variables[:, 1:]), -1)
                for t in range(tw)])
            if self.encoder == "lstm":
                h = self.lstmoutput_mlp(self.embedding_lstm(seq))
            else:
                h = self.lemoutput_mlp(self.embedding_lem(seq))
        h = self._process(h, u, pos_x, variables, data.edge_index, data.batch)
        diff = self.output_mlp(self.double_mlp(h))
        return torch.flatten(unflatten_u(u, tw) + dt * diff, 1, 2)


class MP_PDE_Solver2D(_Solver2F):
    """models_gnn2D.py:17-141"""
    layer_cls, gate, encoder = GNN_Layer, None, "mlp"


class MP_PDE_Solver2DGated(_Solver2F):
    """models_gnn2D.py:143-288"""
    encoder = "mlp"


class MP_PDE_Solver2DLEMLinG2(_Solver2F):
    """models_gnn2D.py:460-620"""
    gate = "g2"


class MP_PDE_Solver2DLSTMLinGated(_Solver2F):
    """models_gnn2D.py:622-780"""
    encoder = "lstm"


class MP_PDE_Solver2DLSTMLin(_Solver2F):
    """models_gnn2D.py:782-918"""
    layer_cls, gate, encoder = GNN_Layer, None, "lstm"


class MP_PDE_Solver2DLEMLin(_Solver2F):
    """models_gnn2D.py:920-1056"""
    layer_cls, gate, encoder = GNN_Layer, None, "lem"


def _glu_decoder(c):
    """models_gnn.py:1455-1456, models_gnn2D.py:1284-1291: Conv1d(c, 8, 6, stride 2) -> Swish -> Conv1d(8, c, 15) on half a row
    (82 -> 39 -> 25 for hidden_features = 164)."""
    return nn.Sequential(nn.Conv1d(c, 8, 6, stride=2), Swish(), nn.Conv1d(8, c, 15, stride=1))


class MP_PDE_SolverLEMLinGatedGLU(_Solver1F):
    """models_gnn.py:1379-1523: hidden_features = 164; two decoders on the two halves of h (gate / diff),
    out = (1 - scale) u[:, -1] + dt (scale * diff)."""

    def __init__(self, pde, time_window=25, hidden_features=164, hidden_layer=6, eq_variables={}):
        super().__init__(pde, time_window, hidden_features, hidden_layer, eq_variables)
        del self.output_mlp
        self.output_mlp_gate = _glu_decoder(1)
        self.output_mlp_diff = _glu_decoder(1)

    def forward(self, data):
        u = data.x
        pos_x = data.pos[:, 1][:, None] / self.pde.L
        pos_t = data.pos[:, 0][:, None] / self.pde.tmax
        variables = variables_1field(data, pos_t, self.eq_variables)
        seq = torch.stack([torch.cat((pos_x, u[:, t:t + 1], variables), -1) for t in range(u.shape[1])])
        h = self.lemoutput_mlp(self.embedding_lem(seq))
        h = self._process(h, u, pos_x, variables, data.edge_index, data.batch)
        half = h.size(1) // 2
        scale = self.output_mlp_gate(h[:, :half][:, None]).squeeze(1)
        diff = self.output_mlp_diff(h[:, half:][:, None]).squeeze(1)
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=h.dtype, device=h.device) * self.pde.dt, dim=1)
        return (1 - scale) * u[:, -1].repeat(self.time_window, 1).transpose(0, 1) + dt * (scale * diff)


class MP_PDE_Solver2DLEMLinGatedGLU(_Solver2F):
    """models_gnn2D.py:1198-1366: hidden_features = 164; double_mlp -> [N, 2, 164]; the gate decoder reads the first half of
    every field's row, the diff decoder the second half; out = (1 - scale) u + dt scale diff."""

    def __init__(self, pde, time_window=25, hidden_features=164, hidden_layer=6, eq_variables={}, save_state=None):
        super().__init__(pde, time_window, hidden_features, hidden_layer, eq_variables, save_state)
        del self.output_mlp
        if time_window == 25:          # (the reference defines no decoder for 50: models_gnn2D.py:1283-1292)
            self.output_mlp_diff = _glu_decoder(2)
            self.output_mlp_gate = _glu_decoder(2)

    def forward(self, data):
        tw = self.time_window
        u = data.x
        pos_x = data.pos[:, 1][:, None] / self.pde.L
        pos_t = data.pos[:, 0][:, None] / self.pde.tmax
        variables = pos_t
        if "a" in self.eq_variables:
            variables = torch.cat((variables, data.a / self.eq_variables["a"]), -1)
        if "b" in self.eq_variables:      # sic: data.a (models_gnn2D.py:1326)
            variables = torch.cat((variables, data.a / self.eq_variables["b"]), -1)
        dt = torch.cumsum(torch.ones(1, 1, tw, dtype=u.dtype, device=u.device) * self.pde.dt, dim=2)
        ts = (dt + pos_t).squeeze(0)
        seq = torch.stack([
            torch.cat((pos_x, u[:, t:t + 1], u[:, t + tw:t + tw + 1], ts[:, t:t + 1], variables[:, 1:]), -1)
            for t in range(tw)])
        h = self.lemoutput_mlp(self.embedding_lem(seq))
        h = self._process(h, u, pos_x, variables, data.edge_index, data.batch)
        h = self.double_mlp(h)
        half = h.size(2) // 2
        diff = self.output_mlp_diff(h[:, :, half:])
        scale = self.output_mlp_gate(h[:, :, :half])
        return torch.flatten((1 - scale) * unflatten_u(u, tw) + dt * scale * diff, 1, 2)
