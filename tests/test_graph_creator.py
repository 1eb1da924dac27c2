"""The vectorised GraphCreator drop-in against fixtures written by the reference's own class
(common/utils.py:267-471, tests/golden/make_golden.py::graph_cases): every field bit for bit."""
import numpy as np
import pytest
import torch

from tests import golden_io

CASES = [("ce", "CE", False, 2), ("ad", "AD", False, 3), ("adu", "AD", True, 3), ("we", "WE", False, 2)]
KEYS = {"CE": ("alpha", "beta", "gamma"), "AD": ("a", "b"), "WE": ("bc_left", "bc_right", "c")}


def _run(tag, name, unstructured, n, device):
    from msmp_pde_b200.graph_creator import GraphCreator
    from msmp_pde_b200.synth import SyntheticPDE
    torch.set_default_dtype(torch.float64)              # the reference runs under a float64 default
    g = golden_io.load(f"graph_{tag}.npz")
    nt, nx = g["traj"].shape[1], g["traj"].shape[-1]
    pde = SyntheticPDE(name, L=16.0, tmax=4.0, grid_size=(nt, nx), untructured_grid=unstructured)
    gc = GraphCreator(pde=pde, neighbors=n, time_window=10, t_resolution=nt, x_resolution=nx)
    t = lambda k: torch.from_numpy(g[k]).to(device)
    variables = {k: t("var_" + k) for k in KEYS[name]}
    steps, steps2 = g["steps"].tolist(), g["steps2"].tolist()
    data, labels = gc.create_data(t("traj"), steps)
    graph = gc.create_graph(data, labels, t("x"), variables, steps)
    return g, gc, data, labels, graph, t, steps2


def _same(a: torch.Tensor, ref: np.ndarray, what):
    a = a.detach().cpu()
    r = torch.from_numpy(ref)
    assert a.dtype == r.dtype, (what, a.dtype, r.dtype)
    assert a.shape == r.shape, (what, tuple(a.shape), tuple(r.shape))
    assert torch.equal(a, r), what


@pytest.mark.parametrize("tag,name,unstructured,n", CASES)
def test_graph_creator_matches_reference(tag, name, unstructured, n):
    g, gc, data, labels, graph, t, steps2 = _run(tag, name, unstructured, n, torch.device("cpu"))
    _same(data, g["data"], "data")
    _same(labels, g["labels"], "labels")
    ref_keys = sorted(k[2:] for k in g if k.startswith("g_"))
    assert sorted(graph.keys()) == ref_keys
    for k in ref_keys:
        _same(graph[k], g["g_" + k], k)
    # topology is built once: the same tensors are handed out again (model-side CSR cache hits by identity)
    graph_b = gc.create_graph(data, labels, t("x"), {k: t("var_" + k) for k in KEYS[name]}, g["steps"].tolist())
    assert graph_b.edge_index is graph.edge_index and graph_b.batch is graph.batch
    _, labels2 = gc.create_data(t("traj"), steps2)
    nxt = gc.create_next_graph(graph, t("pred"), labels2, steps2)
    for k in ("x", "y", "pos"):
        _same(nxt[k], g["n_" + k], "next " + k)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,name,unstructured,n", CASES)
def test_graph_creator_on_device(tag, name, unstructured, n):
    g, gc, data, labels, graph, t, steps2 = _run(tag, name, unstructured, n, torch.device("cuda:0"))
    assert graph.x.is_cuda and graph.edge_index.is_cuda
    for k in sorted(k[2:] for k in g if k.startswith("g_")):
        _same(graph[k], g["g_" + k], k)
    _, labels2 = gc.create_data(t("traj"), steps2)
    nxt = gc.create_next_graph(graph, t("pred"), labels2, steps2)
    for k in ("x", "y", "pos"):
        _same(nxt[k], g["n_" + k], "next " + k)
