"""Data-parallel sharding of a batched graph.

A batch is a disjoint union of per-trajectory graphs (common/utils.py:349-362): nodes and destination-sorted
edges of one graph are contiguous and no edge crosses graphs, so a shard is a slice of every node tensor plus an
edge range with indices rebased.  Graphs are never split (InstanceNorm is per graph).  Pure index arithmetic.
"""
from __future__ import annotations

import torch

from .compat.torch_geometric.data import Data


def shard_graph(data, rank: int, world: int):
    """Contiguous block of whole graphs for ``rank`` out of ``world`` (the first B % world ranks get one more)."""
    batch = data.batch
    B = int(batch.max()) + 1 if batch.numel() else 0
    per, extra = divmod(B, world)
    g0 = rank * per + min(rank, extra)
    g1 = g0 + per + (1 if rank < extra else 0)
    counts = torch.bincount(batch, minlength=B)
    ptr = torch.zeros(B + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(counts.cpu(), 0)
    n0, n1 = int(ptr[g0]), int(ptr[g1])
    ei = data.edge_index
    keep = (ei[1] >= n0) & (ei[1] < n1)
    if bool(((ei[0][keep] < n0) | (ei[0][keep] >= n1)).any()):
        raise ValueError("an edge crosses graphs; cannot shard")
    out = Data()
    N = batch.numel()
    for k in data.keys():
        v = getattr(data, k)
        if k == "edge_index":
            out.edge_index = ei[:, keep] - n0
        elif k == "batch":
            out.batch = batch[n0:n1] - g0
        elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == N:
            setattr(out, k, v[n0:n1])
        else:
            setattr(out, k, v)
    return out
