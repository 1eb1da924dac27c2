"""msmp_pde_b200.train_helper.training_loop against the losses returned by the reference's own training_loop
(experiments/train_helper.py:66-148) run with the reference's GraphCreator and model classes in float64
(tests/golden/make_golden.py::training_loop_case)."""
import random

import pytest
import torch

from tests import golden_io
from tests.util import formula_weights_

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,unrolling", [("u0", [0]), ("u01", [0, 1])])
@pytest.mark.parametrize("fused", [False, True])
def test_training_loop_matches_reference(tag, unrolling, fused):
    from msmp_pde_b200 import models_gnn2D
    from msmp_pde_b200.graph_creator import GraphCreator
    from msmp_pde_b200.synth import SyntheticPDE
    from msmp_pde_b200.train_helper import training_loop
    torch.set_default_dtype(torch.float64)
    dev = torch.device("cuda:0")
    g = golden_io.load("training_loop_ad.npz")
    nt, nx, tw, B = 120, 40, 25, 4
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(nt, nx))
    loader = []
    for i in range(3):
        traj = torch.from_numpy(g[f"traj{i}"])
        loader.append((traj, traj, torch.from_numpy(g[f"x{i}"]),
                       {"a": torch.from_numpy(g[f"a{i}"]), "b": torch.from_numpy(g[f"b{i}"])}))
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, time_window=tw, hidden_features=128, hidden_layer=6,
                                                    eq_variables={"a": 1.0, "b": 1.0})
    formula_weights_(model)
    model = model.to(dev)
    gc = GraphCreator(pde=pde, neighbors=3, time_window=tw, t_resolution=nt, x_resolution=nx)
    if fused:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    random.seed(5)
    losses = training_loop(model, unrolling, B, opt, loader, gc, torch.nn.MSELoss(reduction="sum"), dev)
    want = torch.from_numpy(g["losses_" + tag])
    got = losses.detach().double().cpu()
    assert got.shape == want.shape
    # rtol 1e-5 on the first step (same weights); later steps add the fp32-vs-fp64 difference of two AdamW updates
    assert abs(float(got[0] - want[0])) < 1e-5 * float(want[0])
    assert float((got - want).abs().max()) < 5e-5 * float(want.abs().max()), (got, want)


def test_unrolled_losses_match_reference():
    """Autoregressive rollout (create_next_graph chain) against the reference's own test_unrolled_losses."""
    from msmp_pde_b200 import models_gnn2D
    from msmp_pde_b200.graph_creator import GraphCreator
    from msmp_pde_b200.synth import SyntheticPDE
    from msmp_pde_b200.train_helper import test_unrolled_losses
    torch.set_default_dtype(torch.float64)
    dev = torch.device("cuda:0")
    g = golden_io.load("training_loop_ad.npz")
    nt, nx, tw, B = 120, 40, 25, 4
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(nt, nx))
    loader = []
    for i in range(3):
        traj = torch.from_numpy(g[f"traj{i}"])
        loader.append((traj, traj, torch.from_numpy(g[f"x{i}"]),
                       {"a": torch.from_numpy(g[f"a{i}"]), "b": torch.from_numpy(g[f"b{i}"])}))
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, time_window=tw, hidden_features=128, hidden_layer=6,
                                                    eq_variables={"a": 1.0, "b": 1.0})
    formula_weights_(model)
    model = model.to(dev)
    gc = GraphCreator(pde=pde, neighbors=3, time_window=tw, t_resolution=nt, x_resolution=nx)
    losses = test_unrolled_losses(model, [], B, 1, nx, loader, gc, torch.nn.MSELoss(reduction="sum"), dev)
    want = torch.from_numpy(g["unrolled"])
    got = losses.detach().double().cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 2e-5 * float(want.abs().max()), (got, want)      # three chained windows


def test_compute_L2_norms_match_reference():
    """The reference's reported metric (experiments/train_helper.py:362-471): full rollout + space-time L2 norms, against
    the values its own compute_L2_norms returns on the same trajectories (float64 CPU fixture)."""
    from msmp_pde_b200 import models_gnn2D
    from msmp_pde_b200.graph_creator import GraphCreator
    from msmp_pde_b200.synth import SyntheticPDE
    from msmp_pde_b200.train_helper import compute_L2_norms
    torch.set_default_dtype(torch.float64)
    dev = torch.device("cuda:0")
    g = golden_io.load("training_loop_ad.npz")
    want = golden_io.load("l2_norms_ad.npz")["l2"]
    nt, nx, tw, B = 120, 40, 25, 4
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(nt, nx))
    loader = []
    for i in range(3):
        traj = torch.from_numpy(g[f"traj{i}"])
        loader.append((traj, traj, torch.from_numpy(g[f"x{i}"]),
                       {"a": torch.from_numpy(g[f"a{i}"]), "b": torch.from_numpy(g[f"b{i}"])}))
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, time_window=tw, hidden_features=128, hidden_layer=6,
                                                    eq_variables={"a": 1.0, "b": 1.0})
    formula_weights_(model)
    model = model.to(dev)
    gc = GraphCreator(pde=pde, neighbors=3, time_window=tw, t_resolution=nt, x_resolution=nx)
    l2, l2_rel = compute_L2_norms(model, B, 1, loader, gc, dev)
    assert abs(l2 - want[0]) < 2e-5 * want[0], (l2, want)
    assert abs(l2_rel - want[1]) < 2e-5 * want[1], (l2_rel, want)
