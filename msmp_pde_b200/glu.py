"""The two GLU solver variants of the reference (hidden_features = 164, two decoders on the two halves of the hidden row):
``MP_PDE_SolverLEMLinGatedGLU`` (experiments/models_gnn.py:1379-1523) and ``MP_PDE_Solver2DLEMLinGatedGLU``
(experiments/models_gnn2D.py:1198-1366).

NOT on the hand-written path: every kernel of ``csrc/`` is specialised for 128 hidden channels (one TMEM lane block, one
512-byte row), and 164 does not decompose into 128-wide blocks through the message MLP (its second layer sits behind a
non-linearity).  These classes exist so that ``train.py --model ...GLU`` constructs, trains and loads checkpoints after
``install()``: same constructor, ``forward(data)`` contract, ``__repr__`` and state_dict layout as the reference, computed with
plain torch operators on the tensors' own device and dtype (cuBLAS matmuls, index_add for the scatter; the convolutions are
unfold + matmul so that no cuDNN TF32 path is involved).  Any width works.  Parity: tests/golden/var_*GLU.npz, written from
the reference's own classes.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class Swish(nn.Module):
    """x * sigmoid(beta x) (models_gnn.py:12-21)."""

    def __init__(self, beta=1):
        super().__init__()
        self.beta = beta

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


def _segment_mean(src, index, n):
    """mean of the rows of `src` per value of `index` (rows without entries: 0) -- PyG aggr='mean' / scatter(reduce='mean')."""
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


def _instance_norm(x, batch, eps=1e-5):
    """PyG InstanceNorm(affine=False): per graph and channel, biased variance of the centred values."""
    nb = int(batch.max()) + 1 if batch.numel() else 1
    cnt = torch.bincount(batch, minlength=nb).clamp(min=1).to(x.dtype).view(-1, 1)
    mean = torch.zeros(nb, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x) / cnt
    xc = x - mean[batch]
    var = torch.zeros(nb, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, xc * xc) / cnt
    return xc / (var + eps).sqrt()[batch]


class GNN_LayerLin(nn.Module):
    """Message passing layer without final activation (models_gnn.py:88-149), any width."""

    def __init__(self, in_features, out_features, hidden_features, time_window, n_variables):
        super().__init__()
        self.in_features, self.out_features, self.hidden_features = in_features, out_features, hidden_features
        self.message_net_1 = nn.Sequential(nn.Linear(2 * in_features + time_window + 1 + n_variables, hidden_features), Swish())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, hidden_features), Swish())
        self.update_net_1 = nn.Sequential(nn.Linear(in_features + hidden_features + n_variables, hidden_features), Swish())
        self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features))

    def forward(self, x, u, pos, variables, edge_index, batch):
        j, i = edge_index[0], edge_index[1]                   # source, target (flow = source_to_target)
        m = self.message_net_1(torch.cat((x[i], x[j], u[i] - u[j], pos[i] - pos[j], variables[i]), dim=-1))
        m = _segment_mean(self.message_net_2(m), i, x.shape[0])
        upd = self.update_net_2(self.update_net_1(torch.cat((x, m, variables), dim=-1)))
        return _instance_norm(upd, batch)


class LEMcuda(nn.Module):
    """Parameters and recurrence of the reference's LEMcuda / lem_cuda (models_gnn.py:304-330), any width, torch operators."""

    def __init__(self, ninp, nhid, dt):
        super().__init__()
        self.ninp, self.nhid, self.dt = ninp, nhid, float(dt)
        self.weights = nn.Parameter(torch.empty(3 * nhid, ninp + nhid))
        self.weights_lin_z = nn.Parameter(torch.empty(nhid, ninp + nhid))
        self.bias = nn.Parameter(torch.empty(3 * nhid))
        self.bias_lin_z = nn.Parameter(torch.empty(nhid))
        stdv = 1.0 / math.sqrt(nhid)
        for w in self.parameters():
            w.data.uniform_(-stdv, +stdv)

    def forward(self, inputs, states=None):
        T, N, _ = inputs.shape
        H = self.nhid
        if states is None:
            y = inputs.new_zeros(N, H)
            z = inputs.new_zeros(N, H)
        else:
            y, z = states
        ys, zs = [], []
        for t in range(T):
            g = torch.addmm(self.bias, torch.cat((y, inputs[t]), 1), self.weights.t())
            a = self.dt * torch.sigmoid(g[:, :H])              # multi-scale for y
            b = self.dt * torch.sigmoid(g[:, H:2 * H])         # multi-scale for z
            z = (1 - b) * z + b * torch.tanh(g[:, 2 * H:])
            lin = torch.addmm(self.bias_lin_z, torch.cat((z, inputs[t]), 1), self.weights_lin_z.t())
            y = (1 - a) * y + a * torch.tanh(lin)
            ys.append(y)
            zs.append(z)
        return torch.stack(ys), torch.stack(zs)


class LEM(nn.Module):
    def __init__(self, ninp, nhid, dt=1.):
        super().__init__()
        self.ninp, self.nhid = ninp, nhid
        self.rnn = LEMcuda(ninp, nhid, dt)

    def forward(self, input):
        return self.rnn(input)[0][-1]


class LEMS(nn.Module):
    """Keeps the last (y, z) for the next call (models_gnn.py:345-361)."""

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


def _conv1d(x, conv):
    """nn.Conv1d forward as unfold + matmul (stride from the module, no padding / dilation / groups)."""
    K, S = conv.kernel_size[0], conv.stride[0]
    cols = x.unfold(2, K, S)                                       # [N, C_in, L_out, K]
    return torch.einsum("nclk,ock->nol", cols, conv.weight) + conv.bias.view(1, -1, 1)


class _GluDecoder(nn.Sequential):
    """Conv1d(c, 8, 6, stride 2) -> Swish -> Conv1d(8, c, 15): 82 -> 39 -> 25."""

    def __init__(self, c):
        super().__init__(nn.Conv1d(c, 8, 6, stride=2), Swish(), nn.Conv1d(8, c, 15, stride=1))

    def forward(self, x):
        return _conv1d(self[1](_conv1d(x, self[0])), self[2])


def _mlp(h):
    return nn.Sequential(nn.Linear(h, h), Swish(), nn.Linear(h, h), Swish())


def _gated_stack(model, h, u, pos_x, variables, edge_index, batch):
    for i in range(model.hidden_layer):                       # models_gnn.py:1510-1513
        tau = torch.sigmoid(model.gnn_layers_gate[i](h, u, pos_x, variables, edge_index, batch))
        h = (1 - tau) * h + tau * model.swish(model.gnn_layers[i](h, u, pos_x, variables, edge_index, batch))
    return h


class MP_PDE_SolverLEMLinGatedGLU(nn.Module):
    """experiments/models_gnn.py:1379-1523."""

    def __init__(self, pde, time_window: int = 25, hidden_features: int = 164, hidden_layer: int = 6, eq_variables: dict = {}):
        super().__init__()
        assert time_window in (20, 25, 50)
        self.pde, self.out_features, self.hidden_features = pde, time_window, hidden_features
        self.hidden_layer, self.time_window, self.eq_variables = hidden_layer, time_window, eq_variables
        H, nv = hidden_features, len(eq_variables) + 1
        self.gnn_layers = nn.ModuleList(GNN_LayerLin(H, H, H, time_window, nv) for _ in range(hidden_layer))
        self.gnn_layers_gate = nn.ModuleList(GNN_LayerLin(H, H, H, time_window, nv) for _ in range(hidden_layer))
        self.embedding_lem = LEM(2 + len(eq_variables) + 1, H)
        self.lemoutput_mlp = _mlp(H)
        self.swish = Swish()
        self.output_mlp_gate = _GluDecoder(1)
        self.output_mlp_diff = _GluDecoder(1)

    def __repr__(self):
        return 'GNN'

    def forward(self, data):
        u = data.x
        pos_x = data.pos[:, 1][:, None] / self.pde.L
        pos_t = data.pos[:, 0][:, None] / self.pde.tmax
        variables = pos_t
        for key in ("alpha", "beta", "gamma"):
            if key in self.eq_variables:
                variables = torch.cat((variables, getattr(data, key) / self.eq_variables[key]), -1)
        for key in ("bc_left", "bc_right"):
            if key in self.eq_variables:
                variables = torch.cat((variables, getattr(data, key)), -1)
        for key in ("c", "D", "r"):
            if key in self.eq_variables:
                variables = torch.cat((variables, getattr(data, key) / self.eq_variables[key]), -1)
        seq = torch.stack([torch.cat((pos_x, u[:, t:t + 1], variables), -1) for t in range(u.shape[1])])
        h = self.lemoutput_mlp(self.embedding_lem(seq))
        h = _gated_stack(self, h, u, pos_x, variables, data.edge_index, data.batch)
        half = h.size(1) // 2
        scale = self.output_mlp_gate(h[:, :half][:, None]).squeeze(1)
        diff = self.output_mlp_diff(h[:, half:][:, None]).squeeze(1)
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=h.dtype, device=h.device) * self.pde.dt, dim=1)
        return (1 - scale) * u[:, -1].repeat(self.time_window, 1).transpose(0, 1) + dt * (scale * diff)


class MP_PDE_Solver2DLEMLinGatedGLU(nn.Module):
    """experiments/models_gnn2D.py:1198-1366."""

    def __init__(self, pde, time_window: int = 25, hidden_features: int = 164, hidden_layer: int = 6, eq_variables: dict = {},
                 save_state=None):
        super().__init__()
        assert time_window in (25, 50)
        self.pde, self.out_features, self.hidden_features = pde, time_window, hidden_features
        self.hidden_layer, self.time_window, self.eq_variables = hidden_layer, time_window, eq_variables
        self.save_state = save_state
        H, nv = hidden_features, len(eq_variables) + 1
        self.gnn_layers = nn.ModuleList(GNN_LayerLin(H, H, H, 2 * time_window, nv) for _ in range(hidden_layer))
        self.gnn_layers_gate = nn.ModuleList(GNN_LayerLin(H, H, H, 2 * time_window, nv) for _ in range(hidden_layer))
        self.embedding_lem = (LEMS if save_state else LEM)(2 + len(eq_variables) + 2, H)
        self.lemoutput_mlp = _mlp(H)
        self.swish = Swish()
        self.double_mlp = nn.Sequential(nn.Linear(H, 2 * H), Swish(), nn.Unflatten(1, (2, H)))
        if time_window == 25:          # (the reference defines no decoders for 50: models_gnn2D.py:1283-1292)
            self.output_mlp_diff = _GluDecoder(2)
            self.output_mlp_gate = _GluDecoder(2)

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
        if "b" in self.eq_variables:          # sic: data.a (models_gnn2D.py:1326)
            variables = torch.cat((variables, data.a / self.eq_variables["b"]), -1)
        dt = torch.cumsum(torch.ones(1, 1, tw, dtype=u.dtype, device=u.device) * self.pde.dt, dim=2)
        ts = (dt + pos_t).squeeze(0)
        seq = torch.stack([
            torch.cat((pos_x, u[:, t:t + 1], u[:, t + tw:t + tw + 1], ts[:, t:t + 1], variables[:, 1:]), -1)
            for t in range(tw)])
        h = self.lemoutput_mlp(self.embedding_lem(seq))
        h = _gated_stack(self, h, u, pos_x, variables, data.edge_index, data.batch)
        h = self.double_mlp(h)
        half = h.size(2) // 2
        diff = self.output_mlp_diff(h[:, :, half:])
        scale = self.output_mlp_gate(h[:, :, :half])
        u_unf = torch.stack((u[:, :tw], u[:, tw:]), dim=1)           # unflatten_u (models_gnn2D.py:9-14)
        return torch.flatten((1 - scale) * u_unf + dt * scale * diff, 1, 2)
