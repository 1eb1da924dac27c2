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


def dp_selfcheck(make_model, make_optimizer, data_global, device, group=None, use_graph: bool = True) -> dict:
    """One training step of the SAME global batch two ways on every rank -- data parallel over the ranks' graph shards
    (one NCCL all-reduce of the gradient bucket) and as a single-process full-batch step -- and how far they differ:
    the loss, the gradient of the loss (``param.grad`` after the step) and the updated weights.  The two differ only by
    the order of the floating-point sums (per-rank partial sums, then the all-reduce).  ``make_model()`` must build
    identically initialised models (seed inside).  Returns maxima over all ranks.

    AdamW's first update is ``lr * g / (|g| + eps)``: an entry whose gradient is numerically zero may flip its sign
    between the two evaluations, so the weights are reported as the largest difference in units of ``lr`` (<= 2 by
    construction) and as the fraction of entries that moved differently by more than ``1e-3 lr``."""
    import torch.distributed as dist

    from .train_step import GraphedTrainStep
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard = shard_graph(data_global, rank, world).to(device)
    full = data_global.clone().to(device)
    out = {}
    res = []
    for graph, grp in ((shard, group), (full, False)):
        model = make_model().to(device)
        opt = make_optimizer(model)
        step = GraphedTrainStep(model, opt, graph, group=grp, warmup=1, use_graph=use_graph, preserve_state=True)
        loss = step(graph)
        torch.cuda.synchronize()
        res.append((float(loss), step.bucket.flat[:step.bucket.n].detach().double().clone(),
                    torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double(),
                    float(opt.param_groups[0]["lr"])))
        del step, model, opt
    (l_dp, g_dp, w_dp, lr), (l_1, g_1, w_1, _) = res
    dw = (w_dp - w_1).abs()
    vals = torch.tensor([abs(l_dp - l_1) / abs(l_1), float((g_dp - g_1).abs().max() / g_1.abs().max()),
                         float(dw.max()) / lr, float((dw > 1e-3 * lr).double().mean())], dtype=torch.float64,
                        device=device)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX, group=group)
    out = {"loss_rel_err": float(vals[0]), "grad_max_rel_err": float(vals[1]), "weight_max_abs_err_over_lr": float(vals[2]),
           "weight_mismatch_frac": float(vals[3]), "world": world, "graphs_global": int(data_global.batch.max()) + 1}
    return out
