"""torch.ops.msmp.* (msmp_pde_b200/torch_ops.py): the registered custom ops give the results of the ops.py calls they wrap
and pass torch.library.opcheck (schema, fake tensor, dispatcher registration)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import rel_err  # noqa: E402


def test_custom_ops_match_reference_math():
    from msmp_pde_b200 import synth, torch_ops  # noqa: F401  (registers torch.ops.msmp)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    _, data, _ = synth.config_c1(B=3, nx=50, seed=0)
    ei, batch = data.edge_index.to(dev), data.batch.to(dev)
    N, E = data.x.shape[0], ei.shape[1]
    src = torch.randn(E, 128, generator=g).to(dev)
    deg = torch.bincount(ei[1], minlength=N)
    rowptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    rowptr[1:] = torch.cumsum(deg, 0)
    out = torch.ops.msmp.scatter_mean(src, rowptr, None, True)
    ref = torch.zeros(N, 128, dtype=torch.float64, device=dev).index_add_(0, ei[1], src.double()) / deg.clamp(min=1)[:, None]
    assert rel_err(out, ref) < 1e-6
    torch.library.opcheck(torch.ops.msmp.scatter_mean.default, (src, rowptr, None, True))
    # linear + its weight gradient
    x, W, b = torch.randn(N, 64, generator=g).to(dev), (torch.randn(128, 64, generator=g) / 8).to(dev), torch.randn(128, generator=g).to(dev)
    y = torch.ops.msmp.linear(x, W, b, True)
    z = x.double() @ W.double().t() + b.double()
    assert rel_err(y, z * torch.sigmoid(z)) < 1e-5
    dy = torch.randn(N, 128, generator=g).to(dev)
    dW, db = torch.ops.msmp.linear_wgrad(x, dy)
    assert rel_err(dW, dy.double().t() @ x.double()) < 1e-5 and rel_err(db, dy.double().sum(0)) < 1e-5
    # fused message kernel
    P, Q = torch.randn(N, 128, generator=g).to(dev), torch.randn(N, 128, generator=g).to(dev)
    W2, b2 = (torch.randn(128, 128, generator=g) / 11).to(dev), torch.randn(128, generator=g).to(dev) * 0.1
    agg, z2 = torch.ops.msmp.edge_mlp_scatter(P, Q, ei, batch, W2, b2)
    z1 = P.double()[ei[1]] + Q.double()[ei[0]]
    z2r = (z1 * torch.sigmoid(z1)) @ W2.double().t() + b2.double()
    m = z2r * torch.sigmoid(z2r)
    aggr = torch.zeros(N, 128, dtype=torch.float64, device=dev).index_add_(0, ei[1], m) / deg.clamp(min=1)[:, None]
    assert rel_err(z2, z2r) < 1e-5 and rel_err(agg, aggr) < 1e-5
    # InstanceNorm
    xin = torch.randn(N, 128, generator=g).to(dev)
    o = torch.ops.msmp.instance_norm(xin, ei, batch)
    xr = xin.double().view(3, 50, 128)
    refn = (xr - xr.mean(1, keepdim=True)) / torch.sqrt(xr.var(1, unbiased=False, keepdim=True) + 1e-5)
    assert rel_err(o, refn.view(N, 128)) < 1e-5
    # training criterion (float32 predictions, float64 labels) with its registered autograd formula
    pred = torch.randn(N, 25, generator=g).to(dev).requires_grad_(True)
    lab = torch.randn(N, 25, generator=g, dtype=torch.float64).to(dev)
    s = torch.ops.msmp.sse(pred, lab)
    torch.sqrt(s).backward()
    d = pred.detach().double() - lab
    sref = (d ** 2).sum()
    assert s.dtype == torch.float64 and abs(float(s.detach()) - float(sref)) < 1e-13 * float(sref)
    assert rel_err(pred.grad, d / torch.sqrt(sref)) < 1e-7
    torch.library.opcheck(torch.ops.msmp.sse.default, (pred.detach(), lab), test_utils=("test_schema", "test_faketensor"))
