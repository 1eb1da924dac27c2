"""Per-entry-point parity of the C ABI against float64 torch math on the same inputs (B200 only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import random_directed_graph, rel_err  # noqa: E402

TOL = 1e-5          # max|delta| <= TOL * max|ref|  (north_star: rtol 1e-5 in fp32, SURVEY section 4 metric)


def _sw(x):
    return x * torch.sigmoid(x)


def _dsw(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def _graph(sizes, deg, seed, hub=None):
    ei, batch = random_directed_graph(sizes, deg, seed)
    if hub is not None:      # one destination with > 128 in-edges: segment cut by several tile boundaries
        n = int(batch.numel())
        extra_src = torch.arange(0, min(n, hub))
        extra = torch.stack([extra_src, torch.full_like(extra_src, 5)])
        extra = extra[:, extra[0] != 5]
        ei = torch.unique(torch.cat([ei, extra], 1), dim=1)
        order = torch.argsort(ei[1] * n + ei[0], stable=True)
        ei = ei[:, order]
    return ei, batch


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(params=["tc", "ffma"])
def gemm_mode(request):
    from msmp_pde_b200 import ops
    prev = ops.GEMM_MODE
    ops.GEMM_MODE = request.param
    yield request.param
    ops.GEMM_MODE = prev


@pytest.fixture(params=["tc", "ws", "ffma"])
def edge_mode(request):
    """tc: single-role tensor-core edge kernels (edge_tc.cu); ws: warp-specialised ones (edge_ws.cu) forced for every
    size (the default dispatch only uses them from four tiles per SM on); ffma: exact-fp32 CUDA-core kernels."""
    from msmp_pde_b200 import ops
    prev = ops.GEMM_MODE, ops.EDGE_WS, ops.EDGE_WS_MIN_TILES_FWD, ops.EDGE_WS_MIN_TILES_BWD
    ops.GEMM_MODE = "ffma" if request.param == "ffma" else "tc"
    ops.EDGE_WS = request.param == "ws"
    ops.EDGE_WS_MIN_TILES_FWD = ops.EDGE_WS_MIN_TILES_BWD = 0
    yield request.param
    ops.GEMM_MODE, ops.EDGE_WS, ops.EDGE_WS_MIN_TILES_FWD, ops.EDGE_WS_MIN_TILES_BWD = prev


def test_abi_loaded():
    from msmp_pde_b200 import _lib
    assert _lib.lib.msmp_abi_version() == 1


@pytest.mark.parametrize("M", [1, 300, 1024])
def test_linear_fwd(dev, gemm_mode, M):
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(M)
    A0 = torch.randn(M, 128, generator=g)
    A1 = torch.randn(M, 64, generator=g)
    Wt = torch.randn(192, 256, generator=g) / 14
    bias = torch.randn(256, generator=g)
    side = torch.randn(M, 8, generator=g)
    Ws = torch.randn(8, 256, generator=g)
    R = torch.randn(M, 256, generator=g)
    Z = torch.randn(M, 256, generator=g)
    c = lambda t: t.to(dev)
    ypre = torch.empty(M, 256, device=dev)
    y = ops.linear_fwd([c(A0), c(A1)], c(Wt), bias=c(bias), side=c(side), r=3, Wside=c(Ws), Zmul=c(Z), Ypre=ypre,
                       act=True, R=c(R), aswish=[0, 1])
    d = lambda t: t.double()
    z = torch.cat([d(A0), _sw(d(A1))], 1) @ d(Wt) + d(bias) + d(side)[:, :3] @ d(Ws)[:3]
    z = z * _dsw(d(Z))
    ref = _sw(z) + d(R)
    assert rel_err(ypre, z) < TOL
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize("M,Nout", [(1, 128), (300, 256), (1024, 384), (6400, 128)])
def test_linear_tc_fwd(dev, M, Nout):
    """tcgen05 3xTF32 path: same contract and the same 1e-5 bar as the FFMA path."""
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(M + Nout)
    A0 = torch.randn(M, 128, generator=g)
    A1 = torch.randn(M, 64, generator=g)
    Wt = torch.randn(192, Nout, generator=g) / 14
    bias = torch.randn(Nout, generator=g)
    side = torch.randn(M, 8, generator=g)
    Ws = torch.randn(8, Nout, generator=g)
    R = torch.randn(M, Nout, generator=g)
    Z = torch.randn(M, Nout, generator=g)
    c = lambda t: t.to(dev)
    img = ops.tc_images(c(Wt))
    ypre = torch.empty(M, Nout, device=dev)
    y = ops.linear_tc_fwd([c(A0), c(A1)], img, Nout, bias=c(bias), side=c(side), r=3, Wside=c(Ws), Zmul=c(Z),
                          Ypre=ypre, act=True, R=c(R), aswish=[0, 1])
    d = lambda t: t.double()
    z = torch.cat([d(A0), _sw(d(A1))], 1) @ d(Wt) + d(bias) + d(side)[:, :3] @ d(Ws)[:3]
    z = z * _dsw(d(Z))
    ref = _sw(z) + d(R)
    e1, e2 = rel_err(ypre, z), rel_err(y, ref)
    print(f"linear_tc M={M} Nout={Nout}: pre {e1:.2e} out {e2:.2e}")
    assert e1 < TOL and e2 < TOL
    # plain single-segment use, bit-stable
    y1 = ops.linear_tc_fwd([c(A0)], ops.tc_images(c(Wt[:128].contiguous())), Nout)
    y2 = ops.linear_tc_fwd([c(A0)], ops.tc_images(c(Wt[:128].contiguous())), Nout)
    assert torch.equal(y1, y2)
    assert rel_err(y1, d(A0) @ d(Wt[:128])) < TOL


def test_linear_fwd_plain_and_strided(dev):
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(3)
    M = 777
    big = torch.randn(M, 256, generator=g).to(dev)
    W = (torch.randn(128, 128, generator=g) / 11).to(dev)
    y = ops.linear_fwd([big[:, 128:]], W)             # strided view as A; W[n][k] used as Wt[k=n][n=k]
    ref = big[:, 128:].double() @ W.double()
    assert rel_err(y, ref) < TOL


def test_linear_ws_equals_linear_tc(dev):
    """The persistent warp-specialised linear kernel (used from two tiles per SM on) runs the same MMAs and the same
    epilogue as k_linear_tc: bit-identical outputs.  M is chosen so that the default dispatch picks it (320 row tiles
    x 2 column tiles, ragged last tile)."""
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(21)
    M = 128 * 320 - 37
    A0, A1 = torch.randn(M, 128, generator=g).to(dev), torch.randn(M, 64, generator=g).to(dev)
    Wt = (torch.randn(192, 256, generator=g) / 14).to(dev)
    bias, side, Ws = torch.randn(256, generator=g).to(dev), torch.randn(M, 8, generator=g).to(dev), torch.randn(8, 256, generator=g).to(dev)
    R, Z = torch.randn(M, 256, generator=g).to(dev), torch.randn(M, 256, generator=g).to(dev)
    ypre = torch.empty(M, 256, device=dev)
    y_ws = ops.linear_fwd([A0, A1], Wt, bias=bias, side=side, r=3, Wside=Ws, Zmul=Z, Ypre=ypre, act=True, R=R, aswish=[1, 0])
    ref = (torch.cat([_sw(A0.double()), A1.double()], 1) @ Wt.double() + bias.double() + side.double()[:, :3] @ Ws.double()[:3]) \
        * _dsw(Z.double())
    assert rel_err(ypre, ref) < TOL
    assert rel_err(y_ws, _sw(ref) + R.double()) < TOL
    # the single-tile kernel on the same rows, 100 tiles at a time (below the dispatch threshold)
    for lo in range(0, M, 128 * 50):
        hi = min(M, lo + 128 * 50)
        yp = torch.empty(hi - lo, 256, device=dev)
        y_tc = ops.linear_fwd([A0[lo:hi], A1[lo:hi]], Wt, bias=bias, side=side[lo:hi], r=3, Wside=Ws, Zmul=Z[lo:hi], Ypre=yp,
                              act=True, R=R[lo:hi], aswish=[1, 0])
        assert torch.equal(y_tc, y_ws[lo:hi]) and torch.equal(yp, ypre[lo:hi])


@pytest.mark.parametrize("M,K,Nout", [(1000, 192, 256), (130, 128, 128), (5000, 160, 384), (40, 64, 128)])
def test_linear_wgrad(dev, gemm_mode, M, K, Nout):
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(K)
    X = torch.randn(M, K, generator=g)
    dY = torch.randn(M, Nout, generator=g)
    side = torch.randn(M, 8, generator=g)
    dWt, dWs = ops.linear_wgrad(X.to(dev), dY.to(dev), xswish=True, side=side.to(dev), r=2, has_bias=True)
    refW = _sw(X.double()).t() @ dY.double()
    refS = torch.cat([side.double()[:, :2], torch.ones(M, 1, dtype=torch.float64)], 1).t() @ dY.double()
    assert rel_err(dWt, refW) < TOL
    assert rel_err(dWs, refS) < TOL
    # deterministic
    dWt2, _ = ops.linear_wgrad(X.to(dev), dY.to(dev), xswish=True, side=side.to(dev), r=2, has_bias=True)
    assert torch.equal(dWt, dWt2)


def test_linear_wgrad_two_blocks(dev, gemm_mode):
    """[X | X1]^T dY without materialising the concatenation (update_net_1: [h | agg]; P|Q projection: [h | u])."""
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(5)
    M = 3000
    X, X1, dY = torch.randn(M, 128, generator=g), torch.randn(M, 64, generator=g), torch.randn(M, 256, generator=g)
    side = torch.randn(M, 8, generator=g)
    dWt, dWs = ops.linear_wgrad(X.to(dev), dY.to(dev), X1=X1.to(dev), side=side.to(dev), r=4, has_bias=True)
    ref = torch.cat([X, X1], 1).double().t() @ dY.double()
    refS = torch.cat([side.double()[:, :4], torch.ones(M, 1, dtype=torch.float64)], 1).t() @ dY.double()
    assert dWt.shape == (192, 256)
    assert rel_err(dWt, ref) < TOL and rel_err(dWs, refS) < TOL


def _edge_inputs(sizes, deg, seed, hub, dev):
    from msmp_pde_b200.graph import build_topology
    ei, batch = _graph(sizes, deg, seed, hub)
    N = batch.numel()
    g = torch.Generator().manual_seed(seed + 100)
    PQ = torch.randn(N, 256, generator=g)
    W2 = torch.randn(128, 128, generator=g) / 11
    b2 = torch.randn(128, generator=g) * 0.1
    topo = build_topology(ei.to(dev), batch.to(dev), N)
    return ei, batch, N, PQ, W2, b2, topo


def _edge_ref(ei, N, PQ, W2, b2):
    j, i = ei[0], ei[1]
    P, Q = PQ[:, :128].double(), PQ[:, 128:].double()
    z1 = P[i] + Q[j]
    a1 = _sw(z1)
    z2 = a1 @ W2.double().t() + b2.double()
    m = _sw(z2)
    agg = torch.zeros(N, 128, dtype=torch.float64).index_add_(0, i, m)
    deg = torch.bincount(i, minlength=N).clamp(min=1).double()
    return z1, a1, z2, agg / deg[:, None], deg


@pytest.mark.parametrize("sizes,deg,hub", [([50, 37, 64], 5.0, None), ([300, 500], 7.0, 450), ([20], 1.5, None),
                                           ([2000, 1000], 16.0, None)])
def test_edge_fwd(dev, edge_mode, sizes, deg, hub):
    from msmp_pde_b200 import ops
    ei, batch, N, PQ, W2, b2, topo = _edge_inputs(sizes, deg, 1, hub, dev)
    PQd = PQ.to(dev)
    agg, z2 = ops.edge_fwd(PQd[:, :128], PQd[:, 128:], topo, W2.t().contiguous().to(dev), b2.to(dev))
    _, _, z2_ref, agg_ref, _ = _edge_ref(ei, N, PQ, W2, b2)
    assert rel_err(z2, z2_ref) < TOL
    assert rel_err(agg, agg_ref) < TOL
    agg2, _ = ops.edge_fwd(PQd[:, :128], PQd[:, 128:], topo, W2.t().contiguous().to(dev), b2.to(dev))
    assert torch.equal(agg, agg2)


@pytest.mark.parametrize("sizes,deg,hub", [([50, 37, 64], 5.0, None), ([300, 500], 7.0, 450), ([2000, 1000], 16.0, None)])
def test_edge_bwd(dev, edge_mode, sizes, deg, hub):
    from msmp_pde_b200 import ops
    ei, batch, N, PQ, W2, b2, topo = _edge_inputs(sizes, deg, 2, hub, dev)
    g = torch.Generator().manual_seed(9)
    dagg = torch.randn(N, 128, generator=g)
    z1, a1, z2, _, degc = _edge_ref(ei, N, PQ, W2, b2)
    j, i = ei[0], ei[1]
    dm = dagg.double()[i] / degc[i][:, None]
    dz2 = dm * _dsw(z2)
    dW2_ref = dz2.t() @ a1
    db2_ref = dz2.sum(0)
    dz1_ref = (dz2 @ W2.double()) * _dsw(z1)
    dP_ref = torch.zeros(N, 128, dtype=torch.float64).index_add_(0, i, dz1_ref)
    dQ_ref = torch.zeros(N, 128, dtype=torch.float64).index_add_(0, j, dz1_ref)
    PQd = PQ.to(dev)
    dPQ = torch.full((N, 256), float("nan"), device=dev)
    dz1, dW2, db2 = ops.edge_bwd(PQd[:, :128], PQd[:, 128:], topo, W2.to(dev), z2.float().to(dev), dagg.to(dev), dPQ[:, :128])
    ops.segment_reduce(dz1, topo.colptr, perm=topo.csc_perm, out=dPQ[:, 128:], N=N)
    assert rel_err(dz1, dz1_ref) < TOL
    assert rel_err(dPQ[:, :128], dP_ref) < TOL
    assert rel_err(dPQ[:, 128:], dQ_ref) < TOL
    assert rel_err(dW2, dW2_ref) < TOL
    assert rel_err(db2, db2_ref) < TOL


@pytest.mark.parametrize("sizes,deg,hub", [([300, 500], 7.0, 450), ([3000, 2500, 4000], 9.0, None)])
def test_edge_ws_matches_edge_tc(dev, sizes, deg, hub):
    """The warp-specialised kernels read the [n][k] parameter itself (weights resident in tensor memory) and must give
    the single-role kernels' results: z2 / dz2 / a1 / dz1 identical up to the MMA's accumulation order, the
    destination-segment sums in the same fixed order."""
    from msmp_pde_b200 import ops
    ei, batch, N, PQ, W2, b2, topo = _edge_inputs(sizes, deg, 3, hub, dev)
    g = torch.Generator().manual_seed(11)
    dagg = torch.randn(N, 128, generator=g).to(dev)
    PQd, W2d, b2d = PQ.to(dev), W2.to(dev).contiguous(), b2.to(dev)
    prev = ops.EDGE_WS, ops.EDGE_WS_MIN_TILES_FWD, ops.EDGE_WS_MIN_TILES_BWD
    res = {}
    try:
        ops.EDGE_WS_MIN_TILES_FWD = ops.EDGE_WS_MIN_TILES_BWD = 0
        for mode in ("tc", "ws"):
            ops.EDGE_WS = mode == "ws"
            agg, z2 = ops.edge_fwd(PQd[:, :128], PQd[:, 128:], topo, W2d.t().contiguous(), b2d, W2raw=W2d)
            dP = torch.full((N, 128), float("nan"), device=dev)
            dz1, a1, dz2 = ops.edge_bwd(PQd[:, :128], PQd[:, 128:], topo, W2d, z2, dagg, dP, defer_wgrad=True, W2raw=W2d)
            res[mode] = (agg, z2, dP, dz1, a1, dz2)
    finally:
        ops.EDGE_WS, ops.EDGE_WS_MIN_TILES_FWD, ops.EDGE_WS_MIN_TILES_BWD = prev
    _, _, z2_ref, agg_ref, _ = _edge_ref(ei, N, PQ, W2, b2)
    assert rel_err(res["ws"][0], agg_ref) < TOL and rel_err(res["ws"][1], z2_ref) < TOL
    for a, b in zip(res["tc"], res["ws"]):
        assert not torch.isnan(b).any()
        assert rel_err(b, a.double().cpu()) < 2e-6
    # a1 = sw(P[dst] + Q[src]) involves no GEMM: only the activation differs (MUFU ex2 / rcp against expf)
    assert rel_err(res["ws"][4], res["tc"][4]) < 5e-7


def test_segment_mean_standalone(dev):
    """torch_scatter.scatter(reduce='mean') semantics, keyed on the SOURCE index (models_gnn2D.py:600-601)."""
    from msmp_pde_b200 import ops
    from msmp_pde_b200.graph import build_topology
    ei, batch = _graph([200, 100], 6.0, 5)
    N = batch.numel()
    src = torch.randn(ei.shape[1], 128, generator=torch.Generator().manual_seed(1))
    topo = build_topology(ei.to(dev), batch.to(dev), N)
    outdeg = torch.bincount(ei[0], minlength=N).clamp(min=1)
    out = ops.segment_reduce(src.to(dev), topo.colptr, perm=topo.csc_perm, scale=(1.0 / outdeg.float()).to(dev))
    ref = torch.zeros(N, 128, dtype=torch.float64).index_add_(0, ei[0], src.double()) / outdeg[:, None]
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("sizes", [[100, 100, 100], [37, 300, 1, 129]])
@pytest.mark.parametrize("mode", [0, 1])
def test_instnorm(dev, sizes, mode):
    from oracle.pyg_semantics import instance_norm
    from msmp_pde_b200 import ops
    from msmp_pde_b200.graph import build_topology
    batch = torch.cat([torch.full((n,), b) for b, n in enumerate(sizes)])
    N = batch.numel()
    ei = torch.stack([torch.arange(N), torch.arange(N)])       # topology irrelevant here
    topo = build_topology(ei.to(dev), batch.to(dev), N)
    g = torch.Generator().manual_seed(4)
    y0 = (torch.randn(N, 128, generator=g) * 3 + 50).double().requires_grad_(True)     # |mean| >> sigma
    y1 = torch.randn(N, 128, generator=g).double().requires_grad_(True)
    h = torch.randn(N, 128, generator=g).double().requires_grad_(True)
    w = torch.randn(N, 128, generator=g).double()
    if mode == 0:
        ref = instance_norm(y0, batch)
    else:
        tau = torch.sigmoid(instance_norm(y0, batch))
        ref = (1 - tau) * h + tau * _sw(instance_norm(y1, batch))
    (ref * w).sum().backward()
    f = lambda t: t.detach().float().to(dev)
    if mode == 0:
        out, stat = ops.instnorm_fwd(f(y0), topo)
        dy0 = ops.instnorm_bwd(f(w), f(y0), topo, stat)
    else:
        out, stat = ops.instnorm_fwd(f(y0), topo, y1=f(y1), h=f(h))
        dy0, dy1, dh = ops.instnorm_bwd(f(w), f(y0), topo, stat, y1=f(y1), h=f(h))
        assert rel_err(dy1, y1.grad) < TOL
        assert rel_err(dh, h.grad) < TOL
    # y0 ~ 50 +- 3 in fp32: the input itself carries 4e-6 relative rounding, amplified by 1/sigma
    assert rel_err(out, ref) < 5e-5
    assert rel_err(dy0, y0.grad) < 2e-4


@pytest.mark.parametrize("mode", [0, 1])
def test_instnorm_single_launch_equals_three_kernel_path(dev, mode):
    """Batches of <= 128-node graphs take one launch per direction (msmp_instnorm1_*): same arithmetic in the same
    order as the chunked three-kernel path, so the results must be bit-identical."""
    from msmp_pde_b200 import ops
    from msmp_pde_b200.graph import build_topology
    sizes = [100, 128, 1, 37, 100]
    batch = torch.cat([torch.full((n,), b) for b, n in enumerate(sizes)])
    N = batch.numel()
    topo = build_topology(torch.stack([torch.arange(N), torch.arange(N)]).to(dev), batch.to(dev), N)
    assert topo.one_chunk_per_graph
    g = torch.Generator().manual_seed(8)
    y0, y1, h, w = [torch.randn(N, 128, generator=g).to(dev) for _ in range(4)]
    kw = dict(y1=y1, h=h) if mode else {}
    res = []
    prev = ops.INSTNORM_FUSED
    try:
        for fused in (False, True):
            ops.INSTNORM_FUSED = fused
            out, stat = ops.instnorm_fwd(y0, topo, **kw)
            grads = ops.instnorm_bwd(w, y0, topo, stat, **kw)
            res.append([out, stat] + (list(grads) if mode else [grads]))
    finally:
        ops.INSTNORM_FUSED = prev
    for a, b in zip(*res):
        assert torch.equal(a, b)
    big = build_topology(torch.stack([torch.arange(300), torch.arange(300)]).to(dev),
                         torch.cat([torch.zeros(200), torch.ones(100)]).long().to(dev), 300)
    assert not big.one_chunk_per_graph


@pytest.mark.parametrize("persistent", [False, True])
def test_lem_matches_oracle(dev, gemm_mode, persistent):
    from oracle.models import lem_forward
    from msmp_pde_b200 import ops
    from msmp_pde_b200.lem import LEMcuda
    if persistent and gemm_mode != "tc":
        pytest.skip("persistent LEM kernels are tensor-core only")
    prev = ops.LEM_PERSISTENT
    ops.LEM_PERSISTENT = persistent
    try:
        _lem_check(dev, lem_forward, LEMcuda)
    finally:
        ops.LEM_PERSISTENT = prev


def _lem_check(dev, lem_forward, LEMcuda):
    torch.manual_seed(0)
    T, N, ninp = 7, 333, 6
    rnn = LEMcuda(ninp, 128, 1.0).to(dev)
    x = torch.randn(T, N, ninp)
    wy, wz = torch.randn(T, N, 128), torch.randn(T, N, 128)
    ys, zs = rnn(x.to(dev))
    ((ys * wy.to(dev)).sum() + (zs * wz.to(dev)).sum()).backward()
    ps = [p.detach().double().cpu().requires_grad_(True) for p in (rnn.weights, rnn.weights_lin_z, rnn.bias, rnn.bias_lin_z)]
    z0 = torch.zeros(N, 128, dtype=torch.float64)
    yr, zr = lem_forward(x.double(), *ps, z0, z0, 1.0)
    ((yr * wy.double()).sum() + (zr * wz.double()).sum()).backward()
    assert rel_err(ys, yr) < TOL
    assert rel_err(zs, zr) < TOL
    for p, q in zip((rnn.weights, rnn.weights_lin_z, rnn.bias, rnn.bias_lin_z), ps):
        assert rel_err(p.grad, q.grad) < TOL


@pytest.mark.parametrize("persistent", [False, True])
def test_lem_module_last_state(dev, persistent):
    """LEM / LEMS path (only y_T, z_T leave the op) incl. a carried initial state, vs the float64 oracle."""
    from oracle import models as om
    from msmp_pde_b200 import ops
    from msmp_pde_b200.lem import LEMS
    prev = ops.LEM_PERSISTENT
    ops.LEM_PERSISTENT = persistent
    try:
        torch.manual_seed(1)
        T, N, ninp = 5, 300, 4
        mod = LEMS(ninp, 128).to(dev)
        torch.set_default_dtype(torch.float64)
        ref = om.LEMS(ninp, 128)
        ref.load_state_dict({k: v.double().cpu() for k, v in mod.state_dict().items()})
        x1, x2 = torch.randn(T, N, ninp), torch.randn(T, N, ninp)
        w = torch.randn(N, 128)
        out = mod(x2.float().to(dev)) if False else None
        mod.reset_states()
        h1 = mod(x1.float().to(dev))
        h2 = mod(x2.float().to(dev))          # second call starts from the carried (y, z)
        ((h1 + h2) * w.float().to(dev)).sum().backward()
        r1 = ref(x1)
        r2 = ref(x2)
        ((r1 + r2) * w).sum().backward()
        assert rel_err(h2, r2) < TOL
        for (n, p_), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
            assert rel_err(p_.grad, q.grad) < TOL, n
    finally:
        ops.LEM_PERSISTENT = prev


def test_large_graph_properties(dev):
    """BASELINE config 5 scale (1 Mi nodes x 6 neighbours): size-independent properties of the aggregation kernels --
    linearity, mean-of-constant, sum conservation (checksum of checksums), CSR-vs-CSC consistency, bit-stability."""
    from msmp_pde_b200 import ops, synth
    from msmp_pde_b200.graph import build_topology
    n, deg = 1 << 20, 6
    g = synth.large_graph(n, deg, topology="random", seed=1)
    topo = build_topology(g["edge_index"].to(dev), g["batch"].to(dev), n)
    E = topo.E
    gen = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(E, 128, device=dev, generator=gen)
    b = torch.randn(E, 128, device=dev, generator=gen)
    mean = lambda t: ops.segment_reduce(t, topo.rowptr, scale=topo.inv_deg)
    ma, mb, mab = mean(a), mean(b), mean(a + 2 * b)
    assert rel_err(mab, ma.double() + 2 * mb.double()) < 1e-5                     # linearity
    ones = mean(torch.ones(E, 128, device=dev))
    assert torch.equal(ones, torch.ones_like(ones))                               # every node has in-degree 6
    s_dst = ops.segment_reduce(a, topo.rowptr)                                    # sums by destination
    s_src = ops.segment_reduce(a, topo.colptr, perm=topo.csc_perm)                # sums by source
    tot = a.double().sum(0)
    assert rel_err(s_dst.double().sum(0), tot) < 1e-5 and rel_err(s_src.double().sum(0), tot) < 1e-5
    assert torch.equal(mean(a), ma)                                               # bit-stable
    # by-source result against an index_add on a slice (spot check of 4096 nodes)
    src_idx = g["edge_index"][0].to(dev)
    pick = torch.arange(0, n, n // 4096, device=dev)[:4096]
    mask = torch.isin(src_idx, pick)
    ref = torch.zeros(n, 128, dtype=torch.float64, device=dev).index_add_(0, src_idx[mask], a[mask].double())
    assert rel_err(s_src[pick], ref[pick]) < 1e-5


def test_large_graph_layer_forward_backward_properties(dev):
    """One GNN_Layer(128,128,128,25,1) on 256 Ki nodes x 6 neighbours (config 5 shape, reduced): permutation of the
    graphs inside the batch permutes the outputs (per-graph independence), and the run is bit-reproducible."""
    from msmp_pde_b200 import layers, synth
    n, npg = 1 << 18, 1 << 12
    g = synth.large_graph(n, 6, topology="band", nodes_per_graph=npg, seed=2)
    torch.manual_seed(0)
    layer = layers.GNN_Layer(128, 128, 128, 25, 1).to(dev)
    t = {k: v.to(dev) for k, v in g.items()}
    x = t["x"].clone().requires_grad_(True)
    out = layer(x, t["u"], t["pos"], t["variables"], t["edge_index"], t["batch"])
    out.square().mean().backward()
    g1 = x.grad.clone()
    # swap graph 0 and graph 1 (nodes and edges), everything else untouched
    perm = torch.arange(n, device=dev)
    perm[:npg], perm[npg:2 * npg] = torch.arange(npg, 2 * npg, device=dev), torch.arange(0, npg, device=dev)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=dev)
    ei = inv[t["edge_index"]]
    order = torch.argsort(ei[1] * n + ei[0])
    x2 = t["x"][perm].clone().requires_grad_(True)
    out2 = layer(x2, t["u"][perm], t["pos"][perm], t["variables"][perm], ei[:, order], t["batch"])
    assert rel_err(out2, out[perm]) < 1e-5
    out3 = layer(x, t["u"], t["pos"], t["variables"], t["edge_index"], t["batch"])
    assert torch.equal(out3, out)
    assert bool(torch.isfinite(g1).all())


@pytest.mark.parametrize("C", [1, 2])
@pytest.mark.parametrize("N", [1, 37, 4100])
def test_decoder_fwd_bwd_vs_conv1d(dev, C, N):
    """msmp_decoder_fwd / _bwd on the tw = 25 geometry (register-tiled kernels of decoder_rt.cu: persistent 32-node tiles,
    ragged last tile) against float64 torch Conv1d -> Swish -> Conv1d (models_gnn.py:208-224,275-279; models_gnn2D.py:382-391,
    448-458): outputs, dh and all four weight gradients."""
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(100 + 7 * C + N)
    tw = 25
    h = torch.randn(N, C * 128, generator=g)
    w1, b1 = torch.randn(8, C, 16, generator=g) / 4, torch.randn(8, generator=g) / 4
    w2, b2 = torch.randn(C, 8, 14, generator=g) / 6, torch.randn(C, generator=g) / 4
    u = torch.randn(N, C * tw + 3, generator=g)
    dt = torch.cumsum(torch.full((tw,), 0.04), 0)
    dout = torch.randn(N, C * tw, generator=g)
    geom = (C, 16, 3, 38, 14, tw)
    # float64 reference
    hd = h.double().view(N, C, 128).requires_grad_(True)
    p64 = [t.double().requires_grad_(True) for t in (w1, b1, w2, b2)]
    za_ref = torch.nn.functional.conv1d(hd, p64[0], p64[1], stride=3)
    diff = torch.nn.functional.conv1d(za_ref * torch.sigmoid(za_ref), p64[2], p64[3])
    base = u.double()[:, tw - 1:tw, None].expand(N, 1, tw) if C == 1 else u.double()[:, :C * tw].view(N, C, tw)
    out_ref = (base + dt.double() * diff).reshape(N, C * tw)
    out_ref.backward(dout.double())
    # kernels
    d = lambda t: t.to(dev).contiguous()
    out, za = ops.decoder_fwd(d(h), d(w1), d(b1), d(w2), d(b2), d(u), d(dt), geom)
    dh, dW = ops.decoder_bwd(d(dout), d(h), za, d(w1), d(w2), d(dt), geom)
    assert rel_err(out, out_ref) < TOL
    assert rel_err(za, za_ref.reshape(N, 8 * 38)) < TOL
    assert rel_err(dh, hd.grad.reshape(N, C * 128)) < TOL
    dW_ref = torch.cat([p.grad.reshape(-1) for p in p64])
    assert rel_err(dW, dW_ref) < 2 * TOL


@pytest.mark.parametrize("M,segs,Nout,opts", [
    (38000, [128], 128, dict(act=True, res=True, aswish=[1])),
    (38001, [128, 128], 128, dict(bias=True, side=3)),
    (37999, [64, 32, 32], 256, dict(zmul=True, bias=True)),
    (128 * 296, [32], 128, dict()),
])
def test_linear_large_m_paths_vs_float64(dev, M, segs, Nout, opts):
    """The persistent large-M node GEMM (k_linear_ts: >= 296 tiles; activation operand in tensor memory, ragged last row
    tile, one to three A segments, every epilogue option) against float64 math."""
    from msmp_pde_b200 import ops
    g = torch.Generator().manual_seed(M + Nout)
    A = [torch.randn(M, k, generator=g).to(dev) for k in segs]
    K = sum(segs)
    Wt = (torch.randn(K, Nout, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(Nout, generator=g).to(dev) if opts.get("bias") else None
    r = opts.get("side", 0)
    side = torch.randn(M, 8, generator=g).to(dev) if r else None
    Ws = torch.randn(8, Nout, generator=g).to(dev) if r else None
    Z = torch.randn(M, Nout, generator=g).to(dev) if opts.get("zmul") else None
    R = torch.randn(M, Nout, generator=g).to(dev) if opts.get("res") else None
    asw = opts.get("aswish", [0] * len(segs))
    y = ops.linear_fwd(A, Wt, bias=bias, side=side, r=r, Wside=Ws, Zmul=Z, act=bool(opts.get("act")), R=R, aswish=asw)
    Ad = torch.cat([_sw(a.double()) if s else a.double() for a, s in zip(A, asw)], 1)
    ref = Ad @ Wt.double()
    if bias is not None:
        ref = ref + bias.double()
    if r:
        ref = ref + side.double()[:, :r] @ Ws.double()[:r]
    if Z is not None:
        ref = ref * _dsw(Z.double())
    if opts.get("act"):
        ref = _sw(ref)
    if R is not None:
        ref = ref + R.double()
    assert rel_err(y, ref) < TOL
