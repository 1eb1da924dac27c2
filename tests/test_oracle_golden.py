"""The oracle against the fixtures produced by the reference's own model code (CPU only)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import models as om
from oracle import variants as ov
from tests import golden_io
from tests.util import formula_weights_, grads_digest, rel_err

TOL = 1e-11      # float64 restatement vs float64 reference: summation-order noise only


def _check_digests(model, gold):
    mine = grads_digest(model)
    want = golden_io.grad_digests(gold)
    assert set(mine) == set(want)
    scale = max(float(np.abs(v[0])) for v in want.values())
    for k in want:
        # update_net_2 bias of a GNN_LayerLin has a structurally zero gradient
        assert np.allclose(mine[k], want[k], rtol=1e-8, atol=1e-10 * max(scale, 1.0)), k


@pytest.mark.parametrize("cls,fname,F_u,V", [("GNN_Layer", "layer_gnn.npz", 25, 1),
                                             ("GNN_LayerLin", "layer_gnnlin.npz", 50, 3)])
def test_layer_matches_reference(cls, fname, F_u, V):
    torch.set_default_dtype(torch.float64)
    g = golden_io.load(fname)
    layer = getattr(om, cls)(128, 128, 128, F_u, V)
    formula_weights_(layer)
    x = torch.from_numpy(g["in_x"]).requires_grad_(True)
    out = layer(x, torch.from_numpy(g["in_u"]), torch.from_numpy(g["in_pos"]), torch.from_numpy(g["in_variables"]),
                torch.from_numpy(g["in_edge_index"]), torch.from_numpy(g["in_batch"]))
    (out * torch.from_numpy(g["in_wout"])).sum().backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < TOL
    assert rel_err(x.grad, torch.from_numpy(g["grad_x"])) < TOL
    _check_digests(layer, g)


TW_CASES = [          # (class, fixture, pde name, eq_variables, time_window) -- the other constructible time windows
    ("MP_PDE_Solver", "tw20_MP_PDE_Solver.npz", "CE", {}, 20),
    ("MP_PDE_Solver", "tw50_MP_PDE_Solver.npz", "CE", {}, 50),
    ("MP_PDE_Solver2DLEMLinGated", "tw50_MP_PDE_Solver2DLEMLinGated.npz", "AD", {"a": 1.0, "b": 1.0}, 50),
]

CASES = [
    ("MP_PDE_Solver", "mp_pde_c1.npz", "CE", {}, {}),
    ("MP_PDE_SolverLEMLinGated", "msmp_pde_1f.npz", "CE", {"alpha": 3.0, "beta": 0.4, "gamma": 1.0}, {}),
    ("MP_PDE_Solver2DLEMLinGated", "msmp_pde2d_c2.npz", "AD", {"a": 1.0, "b": 1.0}, {}),
    ("MP_PDE_Solver2DLEMLinGated", "msmp_pde2d_c3.npz", "AD", {"a": 1.0, "b": 1.0}, {}),
]


@pytest.mark.parametrize("cls,fname,pde_name,eq,kw", CASES)
def test_model_matches_reference(cls, fname, pde_name, eq, kw):
    torch.set_default_dtype(torch.float64)
    g = golden_io.load(fname)
    pde, data = golden_io.model_inputs(g, pde_name)
    model = getattr(om, cls)(pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables=eq, **kw)
    formula_weights_(model)
    out = model(data)
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, data.y, reduction="sum"))
    loss.backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-9 * float(g["loss"])
    _check_digests(model, g)


@pytest.mark.parametrize("cls,fname,pde_name,eq,tw", TW_CASES)
def test_time_windows_match_reference(cls, fname, pde_name, eq, tw):
    """time_window 20 / 50 (models_gnn.py:176,208-224; models_gnn2D.py:322,382-391): other decoder geometries,
    F_u = 20 / 50 / 100, LEM over 50 steps."""
    torch.set_default_dtype(torch.float64)
    g = golden_io.load(fname)
    pde, data = golden_io.model_inputs(g, pde_name)
    model = getattr(om, cls)(pde, time_window=tw, hidden_features=128, hidden_layer=6, eq_variables=eq)
    formula_weights_(model)
    out = model(data)
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, data.y, reduction="sum"))
    loss.backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-9 * float(g["loss"])
    _check_digests(model, g)


VARIANTS_1F = ["MP_PDE_SolverLEM", "MP_PDE_SolverLEMLin", "MP_PDE_SolverLSTMLin", "MP_PDE_SolverLSTMLinGated",
               "MP_PDE_SolverGated", "MP_PDE_SolverLEMLinGatedSave", "MSSMP_PDE_Solver"]
VARIANTS_2F = ["MP_PDE_Solver2D", "MP_PDE_Solver2DGated", "MP_PDE_Solver2DLEMLinG2", "MP_PDE_Solver2DLSTMLinGated",
               "MP_PDE_Solver2DLSTMLin", "MP_PDE_Solver2DLEMLin"]


GLU_1F, GLU_2F = "MP_PDE_SolverLEMLinGatedGLU", "MP_PDE_Solver2DLEMLinGatedGLU"      # hidden_features = 164


def variant_hidden(name):
    return 164 if name.endswith("GLU") else 128


def variant_eq(name):
    """eq_variables each var_*.npz was generated with (tests/golden/make_golden.py)."""
    if name == GLU_1F:
        return "CE", {"alpha": 3.0}
    if name in VARIANTS_2F or name == GLU_2F:
        return "AD", {"a": 1.0, "b": 1.0}
    return "CE", ({"alpha": 3.0} if VARIANTS_1F.index(name) % 2 else {})


@pytest.mark.parametrize("name", VARIANTS_1F + VARIANTS_2F + [GLU_1F, GLU_2F])
def test_variant_matches_reference(name):
    torch.set_default_dtype(torch.float64)
    g = golden_io.load(f"var_{name}.npz")
    pde_name, eq = variant_eq(name)
    pde, data = golden_io.model_inputs(g, pde_name)
    model = getattr(ov, name)(pde, time_window=25, hidden_features=variant_hidden(name), hidden_layer=6, eq_variables=eq)
    formula_weights_(model)
    out = model(data)
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, data.y, reduction="sum"))
    loss.backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-9 * float(g["loss"])
    _check_digests(model, g)
    if "out2" in g:            # LEMS: the second call starts from the stored states
        with torch.no_grad():
            assert rel_err(model(data), torch.from_numpy(g["out2"])) < TOL


def test_state_dict_tables():
    torch.set_default_dtype(torch.float64)
    with open(os.path.join(golden_io.GOLDEN_DIR, "state_dict_tables.json")) as f:
        tables = json.load(f)
    from msmp_pde_b200.synth import config_c1, config_c2
    pde1 = config_c1(B=1, nx=10)[0]
    pde2 = config_c2(B=1, nx=10)[0]
    models = {
        "MP_PDE_Solver": om.MP_PDE_Solver(pde1, 25, 128, 6, {}),
        "MP_PDE_SolverLEMLinGated": om.MP_PDE_SolverLEMLinGated(pde1, 25, 128, 6, {}),
        "MP_PDE_Solver2DLEMLinGated": om.MP_PDE_Solver2DLEMLinGated(pde2, 25, 128, 6, {"a": 1.0, "b": 1.0}),
    }
    models.update({n: getattr(ov, n)(pde1, 25, 128, 6, {}) for n in VARIANTS_1F})
    models.update({n: getattr(ov, n)(pde2, 25, 128, 6, {"a": 1.0, "b": 1.0}) for n in VARIANTS_2F})
    models[GLU_1F] = ov.MP_PDE_SolverLEMLinGatedGLU(pde1, 25, 164, 6, {})
    models[GLU_2F] = ov.MP_PDE_Solver2DLEMLinGatedGLU(pde2, 25, 164, 6, {"a": 1.0, "b": 1.0})
    assert set(models) == set(tables)
    for name, m in models.items():
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == tables[name], name
        assert repr(m) == "GNN"
