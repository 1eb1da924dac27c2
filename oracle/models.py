"""Pure-PyTorch CPU restatement of the reference model classes (oracle; test infrastructure only).

Every class keeps the reference's constructor signature and state_dict layout so the same weights
load into the reference classes, this oracle and the CUDA product.  Default dtype is whatever
torch's default is when the class is constructed (tests construct under float64, matching
temporal/solvers.py:10).

Follows:  experiments/models_gnn.py:12-149 (Swish, GNN_Layer, GNN_LayerLin), :151-281
(MP_PDE_Solver), :285-361 (LEM wrappers), :1220-1377 (MP_PDE_SolverLEMLinGated);
experiments/models_gnn2D.py:9-14 (unflatten_u), :290-458 (MP_PDE_Solver2DLEMLinGated).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .pyg_semantics import instance_norm, propagate_mean


class Swish(nn.Module):
    """models_gnn.py:12-21"""

    def __init__(self, beta=1):
        super().__init__()
        self.beta = beta

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


class _LayerBase(nn.Module):
    final_swish = True
    residual = True

    def __init__(self, in_features, out_features, hidden_features, time_window, n_variables):
        super().__init__()
        self.in_features, self.out_features, self.hidden_features = in_features, out_features, hidden_features
        self.message_net_1 = nn.Sequential(
            nn.Linear(2 * in_features + time_window + 1 + n_variables, hidden_features), Swish())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, hidden_features), Swish())
        self.update_net_1 = nn.Sequential(
            nn.Linear(in_features + hidden_features + n_variables, hidden_features), Swish())
        if self.final_swish:
            self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), Swish())
        else:
            self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features))

    def forward(self, x, u, pos, variables, edge_index, batch):
        x = propagate_mean(self, edge_index, x, u, pos, variables)
        return instance_norm(x, batch)

    def message(self, x_i, x_j, u_i, u_j, pos_i, pos_j, variables_i):
        m = self.message_net_1(torch.cat((x_i, x_j, u_i - u_j, pos_i - pos_j, variables_i), dim=-1))
        return self.message_net_2(m)

    def update(self, message, x, variables):
        upd = self.update_net_1(torch.cat((x, message, variables), dim=-1))
        upd = self.update_net_2(upd)
        if self.residual and self.in_features == self.out_features:
            return x + upd
        return upd


class GNN_Layer(_LayerBase):
    """models_gnn.py:23-86 (Swish after update_net_2, residual when in == out)."""
    final_swish, residual = True, True


class GNN_LayerLin(_LayerBase):
    """models_gnn.py:88-149 (no final Swish, never a residual)."""
    final_swish, residual = False, False


# ---------------------------------------------------------------------------------------------
# LEM  (restates the absent native extension `lem_cuda`, upstream tk-rusch/LEM src/lem_cuda)
# ---------------------------------------------------------------------------------------------

def lem_forward(inputs, weights, weights_lin_z, bias, bias_lin_z, y, z, dt):
    """All-steps LEM recurrence with plain differentiable torch ops.

    inputs [T, N, ninp]; weights [3H, H+ninp] with columns ordered [state | input];
    chunk 0 of the 3H rows gates y (dt_bar), chunk 1 gates z (dt), chunk 2 is the z candidate.
    Returns (all_y [T,N,H], all_z [T,N,H])."""
    H = weights_lin_z.shape[0]
    ys, zs = [], []
    for t in range(inputs.shape[0]):
        X = torch.cat((y, inputs[t]), 1)
        G = torch.addmm(bias, X, weights.t())
        ms_dt_bar = dt * torch.sigmoid(G[:, 0:H])
        ms_dt = dt * torch.sigmoid(G[:, H:2 * H])
        z = (1.0 - ms_dt) * z + ms_dt * torch.tanh(G[:, 2 * H:3 * H])
        X2 = torch.cat((z, inputs[t]), 1)
        lin = torch.addmm(bias_lin_z, X2, weights_lin_z.t())
        y = (1.0 - ms_dt_bar) * y + ms_dt_bar * torch.tanh(lin)
        ys.append(y)
        zs.append(z)
    return torch.stack(ys), torch.stack(zs)


class LEMcuda(nn.Module):
    """models_gnn.py:305-330 (parameter holder; no .cuda() so it runs on the CPU oracle)."""

    def __init__(self, ninp, nhid, dt):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.weights = nn.Parameter(torch.empty(3 * nhid, ninp + nhid))
        self.weights_lin_z = nn.Parameter(torch.empty(nhid, ninp + nhid))
        self.bias = nn.Parameter(torch.empty(3 * nhid))
        self.bias_lin_z = nn.Parameter(torch.empty(nhid))
        self.dt = float(dt)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.nhid)
        for w in self.parameters():
            w.data.uniform_(-stdv, +stdv)

    def forward(self, input, states=None):
        if states is None:
            y = input.new_zeros(input.size(1), self.nhid)
            z = input.new_zeros(input.size(1), self.nhid)
            states = (y, z)
        # LEMFunction.backward returns None for the inputs (models_gnn.py:302): inputs get no grad.
        return lem_forward(input.detach(), self.weights, self.weights_lin_z, self.bias, self.bias_lin_z,
                           states[0], states[1], self.dt)


class LEM(nn.Module):
    """models_gnn.py:333-342"""

    def __init__(self, ninp, nhid, dt=1.):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = LEMcuda(ninp, nhid, dt)

    def forward(self, input):
        all_y, _ = self.rnn(input)
        return all_y[-1]


class LEMS(nn.Module):
    """models_gnn.py:345-361 (keeps (y, z) for the next call until reset_states())."""

    def __init__(self, ninp, nhid, dt=1.):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = LEMcuda(ninp, nhid, dt)
        self.states = None

    def forward(self, input):
        all_y, all_z = self.rnn(input, self.states)
        self.states = (all_y[-1], all_z[-1])
        return all_y[-1]

    def reset_states(self):
        self.states = None


# ---------------------------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------------------------

def _decoder_1field(tw, c=1):
    if tw == 20:
        return nn.Sequential(nn.Conv1d(c, 8, 15, stride=4), Swish(), nn.Conv1d(8, c, 10, stride=1))
    if tw == 25:
        return nn.Sequential(nn.Conv1d(c, 8, 16, stride=3), Swish(), nn.Conv1d(8, c, 14, stride=1))
    if tw == 50:
        return nn.Sequential(nn.Conv1d(c, 8, 12, stride=2), Swish(), nn.Conv1d(8, c, 10, stride=1))
    raise AssertionError


_EQ_ORDER_1F = ("alpha", "beta", "gamma", "bc_left", "bc_right", "c", "D", "r")


def variables_1field(data, pos_t, eq_variables):
    """models_gnn.py:250-266: time first, then parameters in this fixed order; bc_left/bc_right are
    NOT divided by their maximum."""
    v = pos_t
    for k in _EQ_ORDER_1F:
        if k in eq_variables:
            col = getattr(data, k)
            if k not in ("bc_left", "bc_right"):
                col = col / eq_variables[k]
            v = torch.cat((v, col), -1)
    return v


class MP_PDE_Solver(nn.Module):
    """models_gnn.py:151-281"""

    def __init__(self, pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables={}):
        super().__init__()
        assert time_window in (20, 25, 50)
        self.pde, self.out_features = pde, time_window
        self.hidden_features, self.hidden_layer = hidden_features, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        nv = len(eq_variables) + 1
        self.gnn_layers = nn.ModuleList(
            GNN_Layer(hidden_features, hidden_features, hidden_features, time_window, nv)
            for _ in range(hidden_layer))
        self.embedding_mlp = nn.Sequential(
            nn.Linear(time_window + 2 + len(eq_variables), hidden_features), Swish(),
            nn.Linear(hidden_features, hidden_features), Swish())
        self.output_mlp = _decoder_1field(time_window)

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        u = data.x
        pos = data.pos
        pos_x = pos[:, 1][:, None] / self.pde.L
        pos_t = pos[:, 0][:, None] / self.pde.tmax
        variables = variables_1field(data, pos_t, self.eq_variables)
        h = self.embedding_mlp(torch.cat((u, pos_x, variables), -1))
        for layer in self.gnn_layers:
            h = layer(h, u, pos_x, variables, data.edge_index, data.batch)
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=h.dtype, device=h.device) * self.pde.dt, dim=1)
        diff = self.output_mlp(h[:, None]).squeeze(1)
        return u[:, -1].repeat(self.time_window, 1).transpose(0, 1) + dt * diff


class MP_PDE_SolverLEMLinGated(nn.Module):
    """models_gnn.py:1220-1377 (`--model MSMP-PDE`)."""

    def __init__(self, pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables={}):
        super().__init__()
        assert time_window in (20, 25, 50)
        self.pde, self.out_features = pde, time_window
        self.hidden_features, self.hidden_layer = hidden_features, hidden_layer
        self.time_window, self.eq_variables = time_window, eq_variables
        nv = len(eq_variables) + 1
        mk = lambda: GNN_LayerLin(hidden_features, hidden_features, hidden_features, time_window, nv)
        self.gnn_layers = nn.ModuleList(mk() for _ in range(hidden_layer))
        self.gnn_layers_gate = nn.ModuleList(mk() for _ in range(hidden_layer))
        self.embedding_lem = LEM(2 + len(eq_variables) + 1, hidden_features)
        self.lemoutput_mlp = nn.Sequential(nn.Linear(hidden_features, hidden_features), Swish(),
                                           nn.Linear(hidden_features, hidden_features), Swish())
        self.swish = Swish()
        self.output_mlp = _decoder_1field(time_window)

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        u = data.x
        pos = data.pos
        pos_x = pos[:, 1][:, None] / self.pde.L
        pos_t = pos[:, 0][:, None] / self.pde.tmax
        variables = variables_1field(data, pos_t, self.eq_variables)
        # models_gnn.py:1357-1360: I_t = [pos_x, u[:, t], variables]
        lem_in = torch.stack([torch.cat((pos_x, u[:, t:t + 1], variables), -1) for t in range(u.shape[1])])
        h = self.lemoutput_mlp(self.embedding_lem(lem_in))
        for i in range(self.hidden_layer):
            tau = torch.sigmoid(self.gnn_layers_gate[i](h, u, pos_x, variables, data.edge_index, data.batch))
            h = (1 - tau) * h + tau * self.swish(self.gnn_layers[i](h, u, pos_x, variables, data.edge_index, data.batch))
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=h.dtype, device=h.device) * self.pde.dt, dim=1)
        diff = self.output_mlp(h[:, None]).squeeze(1)
        return u[:, -1].repeat(self.time_window, 1).transpose(0, 1) + dt * diff


def unflatten_u(u, time_window):
    """models_gnn2D.py:9-14"""
    return u.unflatten(1, (u.size(1) // time_window, time_window))


class MP_PDE_Solver2DLEMLinGated(nn.Module):
    """models_gnn2D.py:290-458 (`--model MSMP-PDE2D`)."""

    def __init__(self, pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables={}, save_state=None):
        super().__init__()
        assert time_window in (25, 50)
        self.pde, self.out_features = pde, time_window
        self.hidden_features, self.hidden_layer = hidden_features, hidden_layer
        self.time_window, self.eq_variables, self.save_state = time_window, eq_variables, save_state
        nv = len(eq_variables) + 1
        mk = lambda: GNN_LayerLin(hidden_features, hidden_features, hidden_features, 2 * time_window, nv)
        self.gnn_layers = nn.ModuleList(mk() for _ in range(hidden_layer))
        self.gnn_layers_gate = nn.ModuleList(mk() for _ in range(hidden_layer))
        lem_cls = LEM if save_state is None else LEMS      # models_gnn2D.py:358-363
        self.embedding_lem = lem_cls(2 + len(eq_variables) + 2, hidden_features)
        self.lemoutput_mlp = nn.Sequential(nn.Linear(hidden_features, hidden_features), Swish(),
                                           nn.Linear(hidden_features, hidden_features), Swish())
        self.swish = Swish()
        self.double_mlp = nn.Sequential(nn.Linear(hidden_features, 2 * hidden_features), Swish(),
                                        nn.Unflatten(1, (2, hidden_features)))
        if time_window == 25:
            self.output_mlp = nn.Sequential(nn.Conv1d(2, 8, 16, stride=3), Swish(), nn.Conv1d(8, 2, 14, stride=1))
        else:
            self.output_mlp = nn.Sequential(nn.Conv1d(2, 8, 12, stride=2), Swish(), nn.Conv1d(8, 2, 10, stride=1))

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        tw = self.time_window
        u = data.x
        pos = data.pos
        pos_x = pos[:, 1][:, None] / self.pde.L
        pos_t = pos[:, 0][:, None] / self.pde.tmax
        variables = pos_t
        if "a" in self.eq_variables:
            variables = torch.cat((variables, data.a / self.eq_variables["a"]), -1)
        if "b" in self.eq_variables:      # sic: data.a is used for 'b' (models_gnn2D.py:419)
            variables = torch.cat((variables, data.a / self.eq_variables["b"]), -1)
        dt = torch.cumsum(torch.ones(1, 1, tw, dtype=u.dtype, device=u.device) * self.pde.dt, dim=2)
        ts = (dt + pos_t).squeeze(0)                                   # [N, tw] (un-normalised dt + normalised t)
        lem_in = torch.stack([
            torch.cat((pos_x, u[:, t:t + 1], u[:, t + tw:t + tw + 1], ts[:, t:t + 1], variables[:, 1:]), -1)
            for t in range(tw)])
        h = self.lemoutput_mlp(self.embedding_lem(lem_in))
        for i in range(self.hidden_layer):
            tau = torch.sigmoid(self.gnn_layers_gate[i](h, u, pos_x, variables, data.edge_index, data.batch))
            h = (1 - tau) * h + tau * self.swish(self.gnn_layers[i](h, u, pos_x, variables, data.edge_index, data.batch))
        h = self.double_mlp(h)
        diff = self.output_mlp(h)
        out = unflatten_u(u, tw) + dt * diff
        return torch.flatten(out, 1, 2)
