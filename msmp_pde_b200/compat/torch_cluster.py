"""`torch_cluster.radius_graph / knn_graph` stand-ins (call sites: common/utils.py:368,377,380).

Vectorised, deterministic, device-agnostic restatement of the upstream semantics
(SURVEY.md section 8c):

* radius_graph: for each target node (ascending) every source in the same batch element with
  squared distance < r^2 (strict), sources ascending, at most ``max_num_neighbors`` per target,
  self-loops removed; returns int64 ``[2, E] = [source, target]`` -- i.e. the edge list is already
  sorted by destination, which is what makes the product's CSR layout free.
* knn_graph: the k nearest sources per target (ties -> lowest index), listed by increasing
  distance, same orientation.

The edge *set* is the bit-exactness contract; it is cross-checked against the independent
brute-force oracle in tests/test_graph_creator.py.
"""
from __future__ import annotations

import torch

_MAX_PAIR_BLOCK = 1 << 24      # pairwise-distance entries materialised at once


def _segments(batch: torch.Tensor, n: int):
    """[(start, end)] of the sorted batch vector."""
    if batch is None:
        return [(0, n)]
    if n == 0:
        return []
    b = batch.detach()
    if b.numel() > 1 and bool((b[1:] < b[:-1]).any()):
        raise ValueError("batch vector must be sorted")
    change = torch.nonzero(b[1:] != b[:-1]).view(-1) + 1
    bounds = [0] + change.cpu().tolist() + [n]
    return list(zip(bounds[:-1], bounds[1:]))


def _groups(segs):
    """Group consecutive equally sized segments so they can be processed as one dense batch."""
    groups, i = [], 0
    while i < len(segs):
        size = segs[i][1] - segs[i][0]
        j = i
        while j + 1 < len(segs) and segs[j + 1][1] - segs[j + 1][0] == size:
            j += 1
        groups.append((segs[i][0], size, j - i + 1))
        i = j + 1
    return groups


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target",
                 num_workers=1, batch_size=None):
    assert flow == "source_to_target"
    x = x.view(-1, 1) if x.dim() == 1 else x
    n = x.size(0)
    r2 = torch.as_tensor(r, dtype=x.dtype, device=x.device) ** 2
    src_out, dst_out = [], []
    for start, size, count in _groups(_segments(batch, n)):
        xb = x[start:start + size * count].view(count, size, -1)
        rows_per = max(1, min(size, _MAX_PAIR_BLOCK // max(1, size * count)))
        s_parts, d_parts = [], []
        for t0 in range(0, size, rows_per):
            t1 = min(size, t0 + rows_per)
            d2 = ((xb[:, t0:t1, None, :] - xb[:, None, :, :]) ** 2).sum(-1)      # [count, T, size]
            within = d2 < r2
            if not loop:
                ar = torch.arange(t0, t1, device=x.device)
                within[:, ar - t0, ar] = False
            within &= within.cumsum(-1) <= max_num_neighbors
            g, t, s = torch.nonzero(within, as_tuple=True)
            base = start + g * size
            s_parts.append(base + s)
            d_parts.append(base + t + t0)
        s_cat, d_cat = torch.cat(s_parts), torch.cat(d_parts)
        if len(s_parts) > 1:       # target chunks were emitted graph-interleaved: restore dst order
            order = torch.argsort(d_cat, stable=True)
            s_cat, d_cat = s_cat[order], d_cat[order]
        src_out.append(s_cat)
        dst_out.append(d_cat)
    if not src_out:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(src_out), torch.cat(dst_out)]).long()


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", cosine=False, num_workers=1,
              batch_size=None):
    assert flow == "source_to_target" and not cosine
    x = x.view(-1, 1) if x.dim() == 1 else x
    n = x.size(0)
    src_out, dst_out = [], []
    for start, size, count in _groups(_segments(batch, n)):
        xb = x[start:start + size * count].view(count, size, -1)
        kk = min(k, size - (0 if loop else 1))
        if kk <= 0:
            continue
        rows_per = max(1, min(size, _MAX_PAIR_BLOCK // max(1, size * count)))
        s_parts, d_parts = [], []
        for t0 in range(0, size, rows_per):
            t1 = min(size, t0 + rows_per)
            d2 = ((xb[:, t0:t1, None, :] - xb[:, None, :, :]) ** 2).sum(-1)
            if not loop:
                ar = torch.arange(t0, t1, device=x.device)
                d2[:, ar - t0, ar] = float("inf")
            idx = torch.sort(d2, dim=-1, stable=True).indices[..., :kk]          # [count, T, kk]
            base = (start + torch.arange(count, device=x.device) * size).view(-1, 1, 1)
            tgt = torch.arange(t0, t1, device=x.device).view(1, -1, 1).expand(count, -1, kk)
            s_parts.append((base + idx))
            d_parts.append((base + tgt))
        s_all = torch.cat(s_parts, dim=1).reshape(-1)
        d_all = torch.cat(d_parts, dim=1).reshape(-1)
        src_out.append(s_all)
        dst_out.append(d_all)
    if not src_out:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(src_out), torch.cat(dst_out)]).long()
