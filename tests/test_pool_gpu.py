"""Pool / unpool extension (msmp_pde_b200/pool.py; SURVEY.md row a13: no reference semantics, parity unpinned): the CUDA
path against a pure-torch float64 restatement of the same definitions, forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import rel_err  # noqa: E402


def test_avg_pool_and_gather_unpool():
    from msmp_pde_b200 import pool
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    N, C = 1000, 137
    cluster = torch.randint(0, C - 3, (N,), generator=g)          # a few empty clusters at the end
    x = torch.randn(N, 128, generator=g)
    w = torch.randn(C, 128, generator=g)
    cm = pool.ClusterMap(cluster.to(dev), num_clusters=C)
    xd = x.to(dev).requires_grad_(True)
    out = pool.avg_pool_x(xd, cm)
    (out * w.to(dev)).sum().backward()
    xr = x.double().requires_grad_(True)
    cnt = torch.bincount(cluster, minlength=C).clamp(min=1).double()
    ref = torch.zeros(C, 128, dtype=torch.float64).index_add_(0, cluster, xr) / cnt[:, None]
    (ref * w.double()).sum().backward()
    assert rel_err(out, ref) < 1e-6 and rel_err(xd.grad, xr.grad) < 1e-6
    assert float(out[C - 3:].abs().max()) == 0.0
    # unpool = gather; its backward = segmented sum
    cd = out.detach().clone().requires_grad_(True)
    up = pool.unpool_gather(cd, cm)
    v = torch.randn(N, 128, generator=g)
    (up * v.to(dev)).sum().backward()
    assert torch.equal(up, cd.detach()[cluster.to(dev)])
    gref = torch.zeros(C, 128, dtype=torch.float64).index_add_(0, cluster, v.double())
    assert rel_err(cd.grad, gref) < 1e-6


def test_linear_interpolation_unpool():
    from msmp_pde_b200 import pool
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    xc = torch.sort(torch.rand(40, generator=g, dtype=torch.float64) * 16).values
    xf = torch.sort(torch.rand(300, generator=g, dtype=torch.float64) * 18 - 1).values      # some outside the coarse range
    vals = torch.randn(40, 128, generator=g)
    im = pool.InterpMap(xf.to(dev), xc.to(dev))
    vd = vals.to(dev).requires_grad_(True)
    out = pool.unpool_interp(vd, im)
    wgt = torch.randn(300, 128, generator=g)
    (out * wgt.to(dev)).sum().backward()
    # float64 restatement: clamped piecewise-linear interpolation
    vr = vals.double().requires_grad_(True)
    i1 = torch.searchsorted(xc, xf, right=True).clamp(1, 39)
    i0 = i1 - 1
    t = ((xf - xc[i0]) / (xc[i1] - xc[i0])).clamp(0, 1)[:, None]
    ref = (1 - t) * vr[i0] + t * vr[i1]
    (ref * wgt.double()).sum().backward()
    assert rel_err(out, ref) < 1e-6 and rel_err(vd.grad, vr.grad) < 1e-6
