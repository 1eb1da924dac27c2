"""Restated third-party semantics used by the hot path (oracle; test infrastructure only).

None of these packages is vendored under /root/reference or installed in this image, so the
functions below restate their published behaviour ("parity unpinned" at this boundary):

* PyG ``MessagePassing(node_dim=-2, aggr='mean')``  -- call sites models_gnn.py:42,65,107,128
* PyG ``InstanceNorm(affine=False)``                -- call sites models_gnn.py:59,66,122,129
* ``torch_scatter.scatter(reduce='mean')``          -- call site  models_gnn2D.py:600-601
* ``torch_cluster.radius_graph / knn_graph``        -- call sites common/utils.py:368,377,380
"""
from __future__ import annotations

import numpy as np
import torch


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter(src, index, dim=0, dim_size=..., reduce='mean'): sum / clamp(count, 1)."""
    s = scatter_sum(src, index, dim_size)
    cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
    return s / cnt.view(-1, *([1] * (src.dim() - 1)))


def propagate_mean(layer, edge_index: torch.Tensor, x, u, pos, variables) -> torch.Tensor:
    """PyG propagate for flow='source_to_target': j = edge_index[0] (source), i = edge_index[1]
    (target); ``*_i`` tensors are indexed by i, ``*_j`` by j; messages are mean-aggregated at i
    (isolated nodes receive 0); then ``update(aggr, x=x, variables=variables)``."""
    j, i = edge_index[0], edge_index[1]
    msg = layer.message(x[i], x[j], u[i], u[j], pos[i], pos[j], variables[i])
    aggr = scatter_mean(msg, i, x.shape[0])
    return layer.update(aggr, x, variables)


def instance_norm(x: torch.Tensor, batch: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """PyG InstanceNorm(affine=False, track_running_stats=False): per graph, per channel, biased
    variance of the centred values."""
    nb = int(batch.max()) + 1 if batch.numel() else 1
    cnt = torch.bincount(batch, minlength=nb).clamp(min=1).to(x.dtype).view(-1, 1)
    mean = scatter_sum(x, batch, nb) / cnt
    xc = x - mean[batch]
    var = scatter_sum(xc * xc, batch, nb) / cnt
    return xc / (var + eps).sqrt()[batch]


# --------------------------------------------------------------------------------------------
# graph construction (torch_cluster semantics, brute force in float arithmetic of the input dtype)
# --------------------------------------------------------------------------------------------

def _as_2d(x: torch.Tensor) -> torch.Tensor:
    return x.view(-1, 1) if x.dim() == 1 else x


def radius_graph(x: torch.Tensor, r: float, batch: torch.Tensor | None = None, loop: bool = False,
                 max_num_neighbors: int = 32) -> torch.Tensor:
    """torch_cluster.radius_graph: for every target node y (ascending), all sources x in the same
    batch element with ||x - y||^2 < r^2 (strict; squared distances, as upstream does), sources
    ascending, capped at max_num_neighbors (self excluded when loop=False).  Returns int64
    [2, E] = [source, target], targets ascending."""
    x = _as_2d(x)
    n = x.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    xs = x.detach().cpu()
    bs = batch.detach().cpu().numpy()
    rows, cols = [], []
    r2 = torch.tensor(r, dtype=xs.dtype) ** 2
    # per batch element brute force (batch is sorted in every caller)
    start = 0
    order = np.argsort(bs, kind="stable")
    assert (order == np.arange(n)).all(), "batch vector must be sorted"
    bounds = np.flatnonzero(np.diff(bs)) + 1
    bounds = np.concatenate(([0], bounds, [n]))
    for a, b in zip(bounds[:-1], bounds[1:]):
        xb = xs[a:b]
        d2 = ((xb[:, None, :] - xb[None, :, :]) ** 2).sum(-1)      # [target?, source?] symmetric
        within = d2 < r2
        if not loop:
            within.fill_diagonal_(False)
        for t in range(b - a):
            src = torch.nonzero(within[t], as_tuple=False).view(-1)[:max_num_neighbors]
            rows.append(src + a)
            cols.append(torch.full_like(src, t + a))
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(rows), torch.cat(cols)]).long()


def knn_graph(x: torch.Tensor, k: int, batch: torch.Tensor | None = None, loop: bool = False) -> torch.Tensor:
    """torch_cluster.knn_graph: for every target node (ascending) its k nearest sources in the same
    batch element (self excluded when loop=False); ties -> lowest index; sources listed by increasing
    distance.  Returns [2, E] = [source, target]."""
    x = _as_2d(x)
    n = x.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    xs = x.detach().cpu()
    bs = batch.detach().cpu().numpy()
    bounds = np.flatnonzero(np.diff(bs)) + 1
    bounds = np.concatenate(([0], bounds, [n]))
    rows, cols = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        xb = xs[a:b]
        d2 = ((xb[:, None, :] - xb[None, :, :]) ** 2).sum(-1)
        if not loop:
            d2 = d2.clone()
            d2.fill_diagonal_(float("inf"))
        kk = min(k, (b - a) - (0 if loop else 1))
        # stable sort => ties resolved to the lowest index
        idx = torch.sort(d2, dim=1, stable=True).indices[:, :kk]
        tgt = torch.arange(a, b).view(-1, 1).expand(-1, kk)
        rows.append((idx + a).reshape(-1))
        cols.append(tgt.reshape(-1))
    return torch.stack([torch.cat(rows), torch.cat(cols)]).long()


def pseudo_random_grid(xmin: float, xmax: float, n: int) -> np.ndarray:
    """generate/generate_data.py:80-113 -- LCG (a=75, c=74, p=65537, x0=0), normalised by its max,
    mapped to [xmin, xmax], sorted, end points forced."""
    c, p, a = 74, 2 ** 16 + 1, 75
    ns = [(a * 0 + c) % p]
    for _ in range(n - 1):
        ns.append((a * ns[-1] + c) % p)
    arr = np.array(ns)
    arr = arr / max(arr)
    arr = arr * (xmax - xmin) + xmin
    out = sorted(arr)
    out[0] = xmin
    out[-1] = xmax
    return np.asarray(out, dtype=np.float64)
