#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REFERENCE's own model code.

Runs only in the build container (needs /root/reference; the GPU box never runs this).  It imports
``experiments/models_gnn.py`` and ``experiments/models_gnn2D.py`` unchanged from /root/reference and
executes their classes (GNN_Layer, GNN_LayerLin, MP_PDE_Solver, MP_PDE_SolverLEMLinGated,
MP_PDE_Solver2DLEMLinGated) in float64 -- the reference's native dtype -- on seeded inputs.  The
third-party packages those files import are not installed, so they are stubbed with
the restated semantics of ``oracle/pyg_semantics.py`` / ``oracle.models.lem_forward``: the fixtures
pin everything the reference's own files compute (feature order, quirks, decoder, gating, time
stepping) while the third-party boundary stays "parity unpinned".

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import pyg_semantics as pg            # noqa: E402
from oracle.models import lem_forward             # noqa: E402
from tests.util import formula_weights_, grads_digest, random_directed_graph  # noqa: E402
from msmp_pde_b200 import synth                   # noqa: E402  (input generators only; no kernels)
from msmp_pde_b200.compat.torch_geometric.data import Data  # noqa: E402


# ------------------------------------------------------------------ third-party stubs
def _install_reference_stubs():
    tg = types.ModuleType("torch_geometric")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data = Data
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_rand = types.ModuleType("torch_geometric.utils.random")
    tg_rand.erdos_renyi_graph = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())

    class MessagePassing(torch.nn.Module):
        def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
            super().__init__()
            assert aggr == "mean" and flow == "source_to_target" and node_dim == -2
            self._msg_args = list(inspect.signature(self.message).parameters)
            self._upd_args = list(inspect.signature(self.update).parameters)[1:]

        def propagate(self, edge_index, **kwargs):
            j, i = edge_index[0], edge_index[1]
            margs = []
            for a in self._msg_args:
                base, side = a[:-2], a[-2:]
                margs.append(kwargs[base][i if side == "_i" else j])
            msg = self.message(*margs)
            n = kwargs["x"].shape[0]
            aggr = pg.scatter_mean(msg, i, n)
            return self.update(aggr, *[kwargs[a] for a in self._upd_args])

    class InstanceNorm(torch.nn.Module):
        def __init__(self, in_channels, eps=1e-5, affine=False, track_running_stats=False):
            super().__init__()
            assert not affine and not track_running_stats
            self.eps = eps

        def forward(self, x, batch=None):
            return pg.instance_norm(x, batch, self.eps)

    for n in ("BatchNorm", "GCNConv", "GATConv", "SAGEConv", "TransformerConv", "RGATConv"):
        setattr(tg_nn, n, type(n, (), {}))
    tg_nn.MessagePassing, tg_nn.InstanceNorm = MessagePassing, InstanceNorm
    tg_nn.global_mean_pool = tg_nn.avg_pool_x = None
    tg.data, tg.nn, tg.utils = tg_data, tg_nn, tg_utils
    tg_utils.random = tg_rand

    tc = types.ModuleType("torch_cluster")
    tc.radius_graph, tc.knn_graph = pg.radius_graph, pg.knn_graph
    ts = types.ModuleType("torch_scatter")
    ts.scatter = lambda src, index, dim=0, dim_size=None, reduce="sum": (
        pg.scatter_mean(src, index, dim_size) if reduce == "mean" else pg.scatter_sum(src, index, dim_size))

    lem = types.ModuleType("lem_cuda")

    def lem_fwd(inputs, weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt):
        with torch.no_grad():
            ys, zs = lem_forward(inputs, weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt)
        e = torch.empty(0)
        return ys, zs, inputs, e, e, e          # all_X slot carries the inputs for the stub backward

    def lem_bwd(gy, gz, all_X, all_X2, all_ms, all_lin, weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt):
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in (weights, weights_lin_z, bias, bias_lin_z, y0, z0)]
            ys, zs = lem_forward(all_X, *leaves[:4], leaves[4], leaves[5], dt)
            gs = torch.autograd.grad([ys, zs], leaves, [gy, gz], allow_unused=True)
        return (None,) + tuple(gs)

    lem.forward, lem.backward = lem_fwd, lem_bwd

    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    h5 = types.ModuleType("h5py")
    for name, mod in {"torch_geometric": tg, "torch_geometric.data": tg_data, "torch_geometric.nn": tg_nn,
                      "torch_geometric.utils": tg_utils, "torch_geometric.utils.random": tg_rand,
                      "torch_cluster": tc, "torch_scatter": ts, "lem_cuda": lem, "matplotlib": mpl,
                      "matplotlib.pyplot": plt, "h5py": h5}.items():
        sys.modules[name] = mod


def _load_reference():
    _install_reference_stubs()
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # LEMcuda.__init__ calls .cuda() (models_gnn.py:314)
    import experiments.models_gnn as mg
    import experiments.models_gnn2D as mg2
    assert torch.get_default_dtype() == torch.float64        # temporal/solvers.py:10 side effect
    return mg, mg2


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def _data_arrays(data):
    return {"in_" + k: getattr(data, k) for k in data.keys()}


def layer_case(mg, cls_name, seed, F_u, V, fname):
    torch.manual_seed(seed)
    ei, batch = random_directed_graph([23, 40, 9], avg_deg=4.0, seed=seed)
    n = batch.numel()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, 128, generator=g, dtype=torch.float64, requires_grad=True)
    u = torch.randn(n, F_u, generator=g, dtype=torch.float64)
    pos = torch.rand(n, 1, generator=g, dtype=torch.float64)
    var = torch.rand(n, V, generator=g, dtype=torch.float64)
    wout = torch.randn(n, 128, generator=g, dtype=torch.float64)
    layer = getattr(mg, cls_name)(128, 128, 128, F_u, V)
    formula_weights_(layer)
    out = layer(x, u, pos, var, ei, batch)
    (out * wout).sum().backward()
    arrs = dict(in_x=x, in_u=u, in_pos=pos, in_variables=var, in_edge_index=ei, in_batch=batch, in_wout=wout,
                out=out, grad_x=x.grad)
    arrs.update({"gdig_" + k: v for k, v in grads_digest(layer).items()})
    np.savez_compressed(os.path.join(os.path.dirname(__file__), fname), **_np(arrs))
    print(fname, "out", tuple(out.shape), "E", ei.shape[1])


def model_case(cls, cfg_fn, cfg_kwargs, fname, extra_params=None, model_kwargs=None, second_call=False):
    pde, data, meta = cfg_fn(**cfg_kwargs)
    if extra_params:
        g = torch.Generator().manual_seed(7)
        B = meta["B"]
        for k in extra_params:
            vals = 0.2 + torch.rand(B, generator=g, dtype=torch.float64)
            setattr(data, k, vals.repeat_interleave(meta["nx"])[:, None])
    eq = meta["eq_variables"] if extra_params is None else extra_params
    mk = dict(hidden_features=128)
    mk.update(model_kwargs or {})
    model = cls(pde, time_window=meta["tw"], hidden_layer=6, eq_variables=eq, **mk)
    formula_weights_(model)
    out = model(data)
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, data.y, reduction="sum"))   # train_helper.py:126,138
    loss.backward()
    arrs = _data_arrays(data)
    arrs.update(out=out, loss=loss, pde_L=pde.L, pde_tmax=pde.tmax, pde_dt=pde.dt)
    if second_call:            # stateful encoders (LEMS): a second forward continues from the stored (y, z)
        with torch.no_grad():
            arrs["out2"] = model(data)
    arrs.update({"gdig_" + k: v for k, v in grads_digest(model).items()})
    np.savez_compressed(os.path.join(os.path.dirname(__file__), fname), **_np(arrs))
    print(fname, "N", data.x.shape[0], "E", data.edge_index.shape[1], "loss", float(loss),
          "params", sum(p.numel() for p in model.parameters()))


VARIANTS_1F = ["MP_PDE_SolverLEM", "MP_PDE_SolverLEMLin", "MP_PDE_SolverLSTMLin", "MP_PDE_SolverLSTMLinGated",
               "MP_PDE_SolverGated", "MP_PDE_SolverLEMLinGatedSave", "MSSMP_PDE_Solver"]
VARIANTS_2F = ["MP_PDE_Solver2D", "MP_PDE_Solver2DGated", "MP_PDE_Solver2DLEMLinG2", "MP_PDE_Solver2DLSTMLinGated",
               "MP_PDE_Solver2DLSTMLin", "MP_PDE_Solver2DLEMLin"]


def graph_cases():
    """Fixtures of the reference's own GraphCreator (common/utils.py:267-471): every field of create_data,
    create_graph and create_next_graph on seeded trajectories, for the four graph-construction branches."""
    import common.utils as cu
    from msmp_pde_b200.synth import SyntheticPDE, pseudo_random_grid
    out = {}
    for tag, name, unstructured, fields, n, dt_in in (
            ("ce", "CE", False, 1, 2, torch.float64), ("ad", "AD", False, 2, 3, torch.float32),
            ("adu", "AD", True, 2, 3, torch.float64), ("we", "WE", False, 1, 2, torch.float64)):
        nt, nx, tw, B = 90, 28, 10, 3
        pde = SyntheticPDE(name, L=16.0, tmax=4.0, grid_size=(nt, nx), untructured_grid=unstructured)
        g = torch.Generator().manual_seed(100 + len(out))
        shape = (B, nt, nx) if fields == 1 else (B, nt, fields, nx)
        traj = torch.randn(*shape, generator=g, dtype=torch.float64).to(dt_in)
        grid = pseudo_random_grid(0.0, 16.0, nx) if unstructured else np.linspace(0.0, 16.0, nx)
        x = torch.tensor(grid, dtype=torch.float64).repeat(B, 1)
        keys = {"CE": ("alpha", "beta", "gamma"), "AD": ("a", "b"), "WE": ("bc_left", "bc_right", "c")}[name]
        variables = {k: torch.rand(B, generator=g, dtype=torch.float64) for k in keys}
        steps, steps2 = [10, 37, 52], [20, 47, 62]
        gc = cu.GraphCreator(pde=pde, neighbors=n, time_window=tw, t_resolution=nt, x_resolution=nx)
        data, labels = gc.create_data(traj, steps)
        graph = gc.create_graph(data, labels, x, variables, steps)
        arrs = {"traj": traj, "x": x, "data": data, "labels": labels, "steps": np.array(steps), "steps2": np.array(steps2)}
        arrs.update({"var_" + k: v for k, v in variables.items()})
        arrs.update({"g_" + k: graph[k].clone() for k in graph.keys()})
        pred = torch.randn(graph.x.shape[0], fields * tw, generator=g, dtype=torch.float64)
        _, labels2 = gc.create_data(traj, steps2)
        nxt = gc.create_next_graph(graph, pred, labels2, steps2)
        arrs.update({"pred": pred, "labels2": labels2})
        arrs.update({"n_" + k: nxt[k].clone() for k in ("x", "y", "pos")})
        np.savez_compressed(os.path.join(os.path.dirname(__file__), f"graph_{tag}.npz"), **_np(arrs))
        out[tag] = {k: str(v.dtype) for k, v in arrs.items() if torch.is_tensor(v)}
        print(f"graph_{tag}.npz", "E", graph.edge_index.shape[1], out[tag]["g_x"], out[tag]["g_pos"])
    return out


def training_loop_case(mg2):
    """Three batches of the reference's own training_loop (experiments/train_helper.py:66-148) with its own
    GraphCreator and MP_PDE_Solver2DLEMLinGated on seeded synthetic AD trajectories (float64, CPU): the per-batch
    losses it returns, with and without the pushforward unrolling."""
    import random
    import common.utils as cu
    import experiments.train_helper as th
    from msmp_pde_b200.synth import SyntheticPDE
    nt, nx, tw, B, nb = 120, 40, 25, 4, 3
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(nt, nx))
    g = torch.Generator().manual_seed(77)
    loader = []
    for _ in range(nb):
        traj = torch.randn(B, nt, 2, nx, generator=g, dtype=torch.float64)
        x = torch.linspace(0.0, 16.0, nx, dtype=torch.float64).repeat(B, 1)
        variables = {"a": 0.1 + 0.9 * torch.rand(B, generator=g, dtype=torch.float64),
                     "b": 1.0 + 9.0 * torch.rand(B, generator=g, dtype=torch.float64)}
        loader.append((traj, traj, x, variables))
    arrs = {}
    for i, (u_b, u_s, x, v) in enumerate(loader):
        arrs[f"traj{i}"], arrs[f"x{i}"], arrs[f"a{i}"], arrs[f"b{i}"] = u_s, x, v["a"], v["b"]
    for tag, unrolling in (("u0", [0]), ("u01", [0, 1])):
        model = mg2.MP_PDE_Solver2DLEMLinGated(pde, time_window=tw, hidden_features=128, hidden_layer=6,
                                                eq_variables={"a": 1.0, "b": 1.0})
        formula_weights_(model)
        gc = cu.GraphCreator(pde=pde, neighbors=3, time_window=tw, t_resolution=nt, x_resolution=nx)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        random.seed(5)
        losses = th.training_loop(model, unrolling, B, opt, loader, gc, torch.nn.MSELoss(reduction="sum"), "cpu")
        arrs["losses_" + tag] = losses
        print("training_loop", tag, [float(l) for l in losses])
    # full autoregressive rollout of the same trajectories (experiments/train_helper.py:205-292), closed-form weights
    model = mg2.MP_PDE_Solver2DLEMLinGated(pde, time_window=tw, hidden_features=128, hidden_layer=6,
                                            eq_variables={"a": 1.0, "b": 1.0})
    formula_weights_(model)
    arrs["unrolled"] = th.test_unrolled_losses(model, [], B, 1, nx, loader, gc, torch.nn.MSELoss(reduction="sum"), "cpu")
    print("test_unrolled_losses", [float(l) for l in arrs["unrolled"]])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "training_loop_ad.npz"), **_np(arrs))


def l2_norms_case(mg2):
    """The reference's evaluation metric (experiments/train_helper.py:299-471): compute_L2_norms of the full rollout of
    the seeded AD trajectories of training_loop_case (same model, closed-form weights), and the two tensor helpers
    compute_spacetime_L2_norms / compute_space_L2_norms on seeded [B, n_t, d, n_x] inputs."""
    import common.utils as cu
    import experiments.train_helper as th
    from msmp_pde_b200.synth import SyntheticPDE
    nt, nx, tw, B, nb = 120, 40, 25, 4, 3
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(nt, nx))
    g = torch.Generator().manual_seed(77)
    loader = []
    for _ in range(nb):          # the same draws as training_loop_case
        traj = torch.randn(B, nt, 2, nx, generator=g, dtype=torch.float64)
        x = torch.linspace(0.0, 16.0, nx, dtype=torch.float64).repeat(B, 1)
        variables = {"a": 0.1 + 0.9 * torch.rand(B, generator=g, dtype=torch.float64),
                     "b": 1.0 + 9.0 * torch.rand(B, generator=g, dtype=torch.float64)}
        loader.append((traj, traj, x, variables))
    model = mg2.MP_PDE_Solver2DLEMLinGated(pde, time_window=tw, hidden_features=128, hidden_layer=6,
                                            eq_variables={"a": 1.0, "b": 1.0})
    formula_weights_(model)
    gc = cu.GraphCreator(pde=pde, neighbors=3, time_window=tw, t_resolution=nt, x_resolution=nx)
    l2, l2_rel = th.compute_L2_norms(model, B, 1, loader, gc, "cpu")
    print("compute_L2_norms", l2, l2_rel)
    g2 = torch.Generator().manual_seed(78)
    losses = torch.rand(5, 7, 2, 11, generator=g2, dtype=torch.float64)
    norms = 0.5 + torch.rand(5, 7, 2, 11, generator=g2, dtype=torch.float64)
    st, st_rel = th.compute_spacetime_L2_norms(losses, norms)
    sp, sp_rel = th.compute_space_L2_norms(losses, norms)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "l2_norms_ad.npz"),
                        l2=np.array([l2, l2_rel]), losses=losses.numpy(), norms=norms.numpy(),
                        spacetime=np.array([float(st), float(st_rel)]), space=sp.numpy(), space_rel=sp_rel.numpy())


def time_window_cases(mg, mg2):
    """The other constructible time windows (models_gnn.py:176,208-224: 20 / 25 / 50; models_gnn2D.py:322,382-391:
    25 / 50): decoder geometries 128 -> 29 -> 20 and 128 -> 59 -> 50, F_u = 20 / 50 / 100 input columns, LEM T = 50."""
    model_case(mg.MP_PDE_Solver, synth.config_c1, dict(B=2, nx=30, tw=20, seed=60), "tw20_MP_PDE_Solver.npz")
    model_case(mg.MP_PDE_Solver, synth.config_c1, dict(B=2, nx=30, tw=50, seed=61), "tw50_MP_PDE_Solver.npz")
    model_case(mg2.MP_PDE_Solver2DLEMLinGated, synth.config_c2, dict(B=2, nx=30, tw=50, seed=62),
               "tw50_MP_PDE_Solver2DLEMLinGated.npz")


def glu_cases(mg, mg2):
    """The two GLU variants at their own width (hidden_features = 164: models_gnn.py:1379-1523, models_gnn2D.py:1198-1366)."""
    model_case(mg.MP_PDE_SolverLEMLinGatedGLU, synth.config_c1, dict(B=2, nx=30, seed=60), "var_MP_PDE_SolverLEMLinGatedGLU.npz",
               extra_params={"alpha": 3.0}, model_kwargs=dict(hidden_features=164))
    model_case(mg2.MP_PDE_Solver2DLEMLinGatedGLU, synth.config_c2, dict(B=2, nx=30, seed=61),
               "var_MP_PDE_Solver2DLEMLinGatedGLU.npz", model_kwargs=dict(hidden_features=164))
    import json
    pde1, _, m1 = synth.config_c1(B=1, nx=10)
    pde2, _, m2 = synth.config_c2(B=1, nx=10)
    path = os.path.join(os.path.dirname(__file__), "state_dict_tables.json")
    tables = json.load(open(path))
    for name, model in {"MP_PDE_SolverLEMLinGatedGLU": mg.MP_PDE_SolverLEMLinGatedGLU(pde1, 25, 164, 6, {}),
                        "MP_PDE_Solver2DLEMLinGatedGLU": mg2.MP_PDE_Solver2DLEMLinGatedGLU(pde2, 25, 164, 6, {"a": 1.0, "b": 1.0})}.items():
        tables[name] = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(path, "w") as f:
        json.dump(tables, f, indent=0, sort_keys=True)


def main():
    mg, mg2 = _load_reference()
    if "--glu-only" in sys.argv:
        glu_cases(mg, mg2)
        return
    if "--l2-only" in sys.argv:
        l2_norms_case(mg2)
        return
    if "--tw-only" in sys.argv:
        time_window_cases(mg, mg2)
        return
    graph_cases()
    training_loop_case(mg2)
    l2_norms_case(mg2)
    layer_case(mg, "GNN_Layer", 11, 25, 1, "layer_gnn.npz")
    layer_case(mg, "GNN_LayerLin", 12, 50, 3, "layer_gnnlin.npz")
    model_case(mg.MP_PDE_Solver, synth.config_c1, dict(B=3, nx=40), "mp_pde_c1.npz")
    model_case(mg.MP_PDE_SolverLEMLinGated, synth.config_c1, dict(B=2, nx=40, seed=3), "msmp_pde_1f.npz",
               extra_params={"alpha": 3.0, "beta": 0.4, "gamma": 1.0})
    model_case(mg2.MP_PDE_Solver2DLEMLinGated, synth.config_c2, dict(B=3, nx=40, seed=4), "msmp_pde2d_c2.npz")
    model_case(mg2.MP_PDE_Solver2DLEMLinGated, synth.config_c3, dict(B=2, nx=100, seed=5, neighbors=3),
               "msmp_pde2d_c3.npz")
    # the variant classes (same layers, different encoder / gating; SURVEY.md 8a "variants")
    for i, name in enumerate(VARIANTS_1F):
        model_case(getattr(mg, name), synth.config_c1, dict(B=2, nx=30, seed=20 + i), f"var_{name}.npz",
                   extra_params={"alpha": 3.0} if i % 2 else None,
                   second_call=name.endswith("Save"))
    for i, name in enumerate(VARIANTS_2F):
        model_case(getattr(mg2, name), synth.config_c2, dict(B=2, nx=30, seed=40 + i), f"var_{name}.npz")
    time_window_cases(mg, mg2)
    # structural fixtures: state_dict key/shape tables
    import json
    tables = {}
    pde1, _, m1 = synth.config_c1(B=1, nx=10)
    pde2, _, m2 = synth.config_c2(B=1, nx=10)
    models = {
        "MP_PDE_Solver": mg.MP_PDE_Solver(pde1, 25, 128, 6, {}),
        "MP_PDE_SolverLEMLinGated": mg.MP_PDE_SolverLEMLinGated(pde1, 25, 128, 6, {}),
        "MP_PDE_Solver2DLEMLinGated": mg2.MP_PDE_Solver2DLEMLinGated(pde2, 25, 128, 6, {"a": 1.0, "b": 1.0}),
    }
    models.update({n: getattr(mg, n)(pde1, 25, 128, 6, {}) for n in VARIANTS_1F})
    models.update({n: getattr(mg2, n)(pde2, 25, 128, 6, {"a": 1.0, "b": 1.0}) for n in VARIANTS_2F})
    for name, model in models.items():
        tables[name] = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(os.path.dirname(__file__), "state_dict_tables.json"), "w") as f:
        json.dump(tables, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in tables.items()})
    glu_cases(mg, mg2)


if __name__ == "__main__":
    main()
