"""CPU-side checks: the C-ABI library loads and exports every declared symbol, graph construction and
topology indexing are bit-exact against independent restatements, the host glue fails loudly without CUDA."""
import numpy as np
import pytest
import torch

from oracle import pyg_semantics as pg


def test_library_exports_every_declared_symbol():
    from msmp_pde_b200 import _lib
    syms = _lib.declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), s
    assert set(syms) == set(_lib._SIGS), set(syms) ^ set(_lib._SIGS)
    assert _lib.lib.msmp_abi_version() == 1          # pure host call, no GPU needed


def test_workspace_queries_are_host_only():
    from msmp_pde_b200 import _lib
    assert _lib.lib.msmp_edge_tiles(9408) == 74
    assert _lib.lib.msmp_edge_fwd_workspace(9408) == 74 * 2 * 128 * 4
    assert _lib.lib.msmp_linear_wgrad_splits(100000, 128, 128) >= 148
    assert _lib.lib.msmp_linear_wgrad_splits(6400, 128, 256) == 25          # at least 256 rows per split
    assert _lib.lib.msmp_edge_ws_workspace(9408) == 74 * 2 * 2 * 128 * 4      # two 64-edge units per tile


def test_radius_graph_matches_bruteforce_and_closed_form():
    from msmp_pde_b200.compat.torch_cluster import radius_graph
    B, nx = 4, 100
    x = torch.linspace(0, 16, nx, dtype=torch.float64).repeat(B)
    batch = torch.arange(B).repeat_interleave(nx)
    r = 3 * (x[1] - x[0]) + 1e-4                       # common/utils.py:366-367
    e1 = radius_graph(x, r=r, batch=batch, loop=False)
    e2 = pg.radius_graph(x, r=float(r), batch=batch)
    assert torch.equal(e1, e2)
    assert e1.shape[1] == B * 588                      # 588 directed edges per graph
    assert bool((e1[1][1:] >= e1[1][:-1]).all())       # destination-sorted => CSR for free
    deg = torch.bincount(e1[1], minlength=B * nx)[:nx]
    assert deg[:3].tolist() == [3, 4, 5] and int(deg[50]) == 6


def test_knn_graph_on_pseudo_random_grid():
    from msmp_pde_b200.compat.torch_cluster import knn_graph
    from msmp_pde_b200.synth import pseudo_random_grid
    g = pseudo_random_grid(0, 16, 100)
    assert np.array_equal(g, pg.pseudo_random_grid(0, 16, 100))
    assert g[0] == 0 and g[-1] == 16 and np.all(np.diff(g) > 0)
    xg = torch.tensor(g)
    X = 2 * np.pi * xg / (xg.max() - 1e-3)
    xp = torch.stack([torch.cos(X), torch.sin(X)], 1).repeat(2, 1)
    batch = torch.arange(2).repeat_interleave(100)
    for k in (3, 6):
        e1 = knn_graph(xp, k, batch=batch)
        assert torch.equal(e1, pg.knn_graph(xp, k, batch=batch))
        assert e1.shape[1] == 2 * 100 * k
    # periodic seam: nodes 0 and 99 are neighbours on the circle
    e = knn_graph(xp[:100], 3)
    assert 99 in e[0][e[1] == 0].tolist()


def test_topology_bit_exact_vs_numpy():
    from msmp_pde_b200.graph import build_topology
    from tests.util import random_directed_graph
    ei, batch = random_directed_graph([23, 300, 9, 129], 4.0, seed=3)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    for edges in (ei, ei[:, perm]):                    # sorted and unsorted input
        N = batch.numel()
        t = build_topology(edges, batch, N)
        src, dst = edges[0].numpy(), edges[1].numpy()
        order = np.argsort(dst, kind="stable")
        s_sorted, d_sorted = src[order], dst[order]
        assert np.array_equal(t.src.numpy(), s_sorted) and np.array_equal(t.dst.numpy(), d_sorted)
        rowptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=N))])
        assert np.array_equal(t.rowptr.numpy(), rowptr)
        colptr = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=N))])
        assert np.array_equal(t.colptr.numpy(), colptr)
        assert np.array_equal(t.csc_perm.numpy(), np.argsort(s_sorted, kind="stable"))
        deg = np.maximum(np.bincount(dst, minlength=N), 1)
        assert np.array_equal(t.inv_deg.numpy(), (1.0 / deg).astype(np.float32))
        # chunks: graph aligned, <= 128 rows, cover every node exactly once
        cb, ce = t.chunk_begin.numpy(), t.chunk_end.numpy()
        assert np.all(ce - cb <= 128) and np.all(ce > cb)
        covered = np.concatenate([np.arange(a, b) for a, b in zip(cb, ce)])
        assert np.array_equal(covered, np.arange(N))
        for c0, c1 in zip(cb, ce):
            assert len(set(batch[c0:c1].tolist())) == 1
        assert t.graph_chunk_ptr.tolist() == [0, 1, 4, 5, 7]
        # per-edge scale streamed by the backward edge kernel; the one-launch InstanceNorm needs <= 128-node graphs
        assert np.array_equal(t.inv_deg_e.numpy(), (1.0 / deg).astype(np.float32)[d_sorted])
        assert not t.one_chunk_per_graph
    small = build_topology(ei[:, ei[1] < 23], batch[:23], 23)
    assert small.one_chunk_per_graph and small.nchunks == small.B == 1


def test_new_entry_points_host_side():
    """Host-only checks of the round-2 additions: workspace query of the criterion kernel, argument validation of the
    input-assembly / criterion / G^2 wrappers (CPU tensors are rejected: there is no CPU path behind them)."""
    import torch
    from msmp_pde_b200 import _lib, ops
    assert _lib.lib.msmp_sse_workspace(0) == 8 and _lib.lib.msmp_sse_workspace(4096) == 16 and _lib.lib.msmp_sse_workspace(4097) == 24
    u, px, v = torch.zeros(4, 50), torch.zeros(4, 1), torch.zeros(4, 2)
    with pytest.raises(ValueError):
        ops.node_features(u, px, v, 64)
    with pytest.raises(ValueError):
        ops.lem_inputs(25, 4, [("static", px, 0), ("time", u, 0)])
    with pytest.raises(ValueError):
        ops.lem_inputs(25, 4, [])
    with pytest.raises(ValueError):
        ops.sse_fwd(torch.zeros(4, 25), torch.zeros(4, 25, dtype=torch.float64))
    # the C entry points validate their arguments before touching the device
    assert _lib.lib.msmp_sse_fwd(0, 0, 0, 0, 0, 0, 0, 0) != 0
    assert _lib.lib.msmp_lem_inputs(0, 0, 0, 0, 1, 1, 0, 0) != 0
    assert _lib.lib.msmp_node_features(0, 0, 0, 0, 0, 1, 0, 0, 0, 0) != 0
    assert _lib.lib.msmp_g2_fwd(0, 0, 0, 0, 0, 0, 1, 0) != 0


def test_no_cpu_fallback():
    from msmp_pde_b200 import models_gnn, synth
    pde, data, meta = synth.config_c1(B=1, nx=20)
    model = models_gnn.MP_PDE_Solver(pde, 25, 128, 6, {})
    with pytest.raises(RuntimeError, match="CUDA"):
        model(data)


def test_install_makes_reference_imports_resolve():
    import sys
    import msmp_pde_b200
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith(("experiments", "torch_geometric", "torch_cluster", "torch_scatter"))}
    try:
        msmp_pde_b200.install()
        from experiments.models_gnn import MP_PDE_Solver, MP_PDE_SolverLEMLinGated, GNN_Layer, LEM, LEMS, LSTM, Swish  # noqa: F401
        from experiments.models_gnn2D import MP_PDE_Solver2DLEMLinGated, unflatten_u  # noqa: F401
        from torch_geometric.data import Data
        from torch_cluster import radius_graph, knn_graph  # noqa: F401
        from torch_scatter import scatter
        d = Data(x=torch.zeros(3, 2), edge_index=torch.zeros(2, 0, dtype=torch.long))
        d.y = torch.ones(3)
        assert d.to("cpu").y.sum() == 3
        out = scatter(torch.ones(4, 2), torch.tensor([0, 0, 2, 2]), dim=0, dim_size=3, reduce="mean")
        assert out.tolist() == [[1, 1], [0, 0], [1, 1]]
    finally:
        for k in list(sys.modules):
            if k.startswith(("experiments", "torch_geometric", "torch_cluster", "torch_scatter")) and k not in saved:
                del sys.modules[k]


def test_drop_in_state_dict_layout_matches_reference_tables():
    """Every drop-in class has the reference's state_dict keys and shapes (tables written from the reference's
    own classes by tests/golden/make_golden.py), so reference checkpoints load unchanged."""
    import json
    import os
    from msmp_pde_b200 import models_gnn, models_gnn2D
    from msmp_pde_b200.synth import config_c1, config_c2
    from tests import golden_io
    with open(os.path.join(golden_io.GOLDEN_DIR, "state_dict_tables.json")) as f:
        tables = json.load(f)
    pde1, pde2 = config_c1(B=1, nx=10)[0], config_c2(B=1, nx=10)[0]
    for name, want in tables.items():
        H = 164 if name.endswith("GLU") else 128          # the GLU variants' own width (torch-operator classes, glu.py)
        if hasattr(models_gnn, name):
            m = getattr(models_gnn, name)(pde1, 25, H, 6, {})
        else:
            m = getattr(models_gnn2D, name)(pde2, 25, H, 6, {"a": 1.0, "b": 1.0})
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == want, name
        assert repr(m) == "GNN"


def test_unsupported_variants_fail_loudly():
    import pytest
    from msmp_pde_b200 import models_gnn2D
    from msmp_pde_b200.synth import config_c1
    pde = config_c1(B=1, nx=10)[0]
    with pytest.raises(NotImplementedError):
        models_gnn2D.G_PDE_Solver2DLEMLinGated(pde, 25, 164, 6, {})


def test_l2_norm_helpers_match_reference_functions():
    """compute_spacetime_L2_norms / compute_space_L2_norms (experiments/train_helper.py:299-360) against the outputs
    of the reference's own functions on seeded [B, n_t, d, n_x] inputs (tests/golden/l2_norms_ad.npz)."""
    import numpy as np
    import torch
    from tests import golden_io
    from msmp_pde_b200.train_helper import compute_space_L2_norms, compute_spacetime_L2_norms
    g = golden_io.load("l2_norms_ad.npz")
    losses, norms = torch.from_numpy(g["losses"]), torch.from_numpy(g["norms"])
    a, r = compute_spacetime_L2_norms(losses, norms)
    np.testing.assert_allclose([float(a), float(r)], g["spacetime"], rtol=1e-13)
    a, r = compute_space_L2_norms(losses, norms)
    np.testing.assert_allclose(a.numpy(), g["space"], rtol=1e-13)
    np.testing.assert_allclose(r.numpy(), g["space_rel"], rtol=1e-13)


def test_custom_ops_are_registered_with_fake_implementations():
    """torch.ops.msmp.* exist after importing torch_ops, trace with fake tensors (shape inference without a GPU) and have
    no CPU kernels (a CPU tensor fails in the dispatcher: no fallback)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from msmp_pde_b200 import torch_ops
    for name in torch_ops.names():
        assert hasattr(torch.ops.msmp, name), name
    with FakeTensorMode():
        src = torch.empty(10, 128, device="cuda")
        ptr = torch.empty(5, dtype=torch.int32, device="cuda")
        assert tuple(torch.ops.msmp.scatter_mean(src, ptr, None, True).shape) == (4, 128)
        x, w = torch.empty(7, 64, device="cuda"), torch.empty(128, 64, device="cuda")
        assert tuple(torch.ops.msmp.linear(x, w, None, True).shape) == (7, 128)
        dW, db = torch.ops.msmp.linear_wgrad(x, torch.empty(7, 128, device="cuda"))
        assert tuple(dW.shape) == (128, 64) and tuple(db.shape) == (128,)
    with pytest.raises(NotImplementedError):
        torch.ops.msmp.linear(torch.zeros(4, 32), torch.zeros(128, 32), None, False)


def test_install_routes_train_helper_and_lem_cuda():
    """install() puts this package's loops at experiments.train_helper (train.py:21 star-imports it) and the lem_cuda shim
    at the name the reference's LEMFunction imports (models_gnn.py:287-302)."""
    import sys
    import msmp_pde_b200
    keep = ("experiments", "torch_geometric", "torch_cluster", "torch_scatter", "lem_cuda")
    saved = {k: v for k, v in sys.modules.items() if k.startswith(keep)}
    try:
        sys.modules.pop("lem_cuda", None)
        msmp_pde_b200.install()
        import lem_cuda
        from msmp_pde_b200.compat import lem_cuda as shim
        assert lem_cuda is shim and callable(lem_cuda.forward) and callable(lem_cuda.backward)
        from experiments.train_helper import training_loop, test_unrolled_losses, compute_L2_norms, unflatten_u  # noqa: F401
        from msmp_pde_b200 import train_helper as ours
        assert training_loop.__doc__ == ours.training_loop.__doc__

        class NotAGnn:
            def __repr__(self):
                return "CNN"
        with pytest.raises((NotImplementedError, TypeError, AttributeError)):
            training_loop(NotAGnn(), [0], 1, None, [], None, None)       # not routed to the GNN loop
    finally:
        for k in [k for k in sys.modules if k.startswith(keep)]:
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("name", ["MP_PDE_SolverLEMLinGatedGLU", "MP_PDE_Solver2DLEMLinGatedGLU"])
def test_glu_variants_match_reference_fixtures(name):
    """The GLU variants (hidden_features = 164; torch-operator classes of msmp_pde_b200/glu.py, no CUDA kernels) in float64
    against the outputs, loss and gradient digests written from the reference's own classes (make_golden.py --glu-only)."""
    import torch
    from msmp_pde_b200 import models_gnn, models_gnn2D
    from tests import golden_io
    from tests.test_oracle_golden import _check_digests, variant_eq
    from tests.util import formula_weights_, rel_err
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        g = golden_io.load(f"var_{name}.npz")
        pde_name, eq = variant_eq(name)
        pde, data = golden_io.model_inputs(g, pde_name)
        cls = getattr(models_gnn, name, None) or getattr(models_gnn2D, name)
        model = cls(pde, time_window=25, hidden_features=164, hidden_layer=6, eq_variables=eq)
        formula_weights_(model)
        out = model(data)
        loss = torch.sqrt(torch.nn.functional.mse_loss(out, data.y, reduction="sum"))
        loss.backward()
        assert rel_err(out, torch.from_numpy(g["out"])) < 1e-10
        assert abs(float(loss) - float(g["loss"])) < 1e-9 * float(g["loss"])
        _check_digests(model, g)
    finally:
        torch.set_default_dtype(prev)
