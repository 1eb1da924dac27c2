"""Fine-to-coarse pooling and coarse-to-fine unpooling / interpolation gathers (north_star kernel 4).

EXTENSION, PARITY UNPINNED: the reference imports ``avg_pool_x`` / ``global_mean_pool`` (experiments/models_gnn.py:7) but
never calls them, so there is no reference behaviour to match.  The semantics restated
here are PyG's ``avg_pool_x`` (mean of the node features of every cluster id), the plain index gather as its adjoint
direction, and 1-D linear interpolation between the two nearest coarse nodes with the convention of the one
interpolation routine in the reference tree, ``interp1d_single`` (common/utils.py:15-33: clamped at both ends).  Tests
compare against a pure-torch restatement of the same definitions (tests/test_pool_gpu.py).

Everything runs on the deterministic segmented-reduction kernel of the hot path (``msmp_segment_reduce``: one warp per
output row, fixed order, no atomics): pooling is a segmented mean over the nodes sorted by cluster, a gather is a
"segmented sum" of one-element segments, and the backward of either is the other.  Rows are 128 floats (the hidden width).
"""
from __future__ import annotations

import torch

from . import ops

H = 128


class ClusterMap:
    """Sorted view of a cluster assignment ``cluster[n] in [0, C)`` (int64 / int32, any order): ``perm`` = node ids sorted
    (stably) by cluster, ``ptr`` [C+1] = CSR offsets, ``inv_count`` [C] = 1 / max(size, 1).  Built once per assignment."""

    def __init__(self, cluster: torch.Tensor, num_clusters: int | None = None):
        c = cluster.long()
        C = int(c.max()) + 1 if num_clusters is None else int(num_clusters)
        self.N, self.C = int(c.numel()), C
        self.cluster = c.to(torch.int32).contiguous()
        self.perm = torch.argsort(c, stable=True).to(torch.int32).contiguous()
        cnt = torch.bincount(c, minlength=C)
        ptr = torch.zeros(C + 1, dtype=torch.int64, device=c.device)
        ptr[1:] = torch.cumsum(cnt, 0)
        self.ptr = ptr.to(torch.int32).contiguous()
        self.inv_count = (1.0 / cnt.clamp(min=1).to(torch.float32)).contiguous()
        self.unit_ptr = torch.arange(self.N + 1, dtype=torch.int32, device=c.device)


def _gather_rows(src, idx, unit_ptr, scale=None):
    """out[n] = scale[n] * src[idx[n]]  (a segment reduction over one-element segments)."""
    return ops.segment_reduce(src, unit_ptr, perm=idx, scale=scale, N=idx.numel())


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cm: ClusterMap):
        ctx.cm = cm
        return ops.segment_reduce(x.contiguous(), cm.ptr, perm=cm.perm, scale=cm.inv_count, N=cm.C)

    @staticmethod
    def backward(ctx, dout):
        cm = ctx.cm
        per_node_scale = _gather_scalar(cm.inv_count, cm.cluster)
        return _gather_rows(dout.contiguous(), cm.cluster, cm.unit_ptr, per_node_scale), None


def _gather_scalar(v, idx):
    return v[idx.long()].contiguous()


class _UnpoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xc, cm: ClusterMap):
        ctx.cm = cm
        return _gather_rows(xc.contiguous(), cm.cluster, cm.unit_ptr)

    @staticmethod
    def backward(ctx, dout):
        cm = ctx.cm
        return ops.segment_reduce(dout.contiguous(), cm.ptr, perm=cm.perm, N=cm.C), None


def avg_pool_x(x: torch.Tensor, cm: ClusterMap) -> torch.Tensor:
    """fine [N, 128] -> coarse [C, 128]: mean over the nodes of every cluster (empty clusters -> 0)."""
    return _PoolFn.apply(x, cm)


def unpool_gather(xc: torch.Tensor, cm: ClusterMap) -> torch.Tensor:
    """coarse [C, 128] -> fine [N, 128]: every node takes its cluster's row."""
    return _UnpoolFn.apply(xc, cm)


class InterpMap:
    """1-D linear interpolation from coarse positions ``xc`` (sorted ascending) to fine positions ``xf``: node n reads the
    coarse nodes i0[n], i0[n] + 1 with weights (1 - w[n], w[n]); positions outside [xc[0], xc[-1]] are clamped to the end
    values (interp1d_single, common/utils.py:15-33)."""

    def __init__(self, xf: torch.Tensor, xc: torch.Tensor):
        C = xc.numel()
        i1 = torch.searchsorted(xc.contiguous(), xf.contiguous(), right=True).clamp(1, C - 1)
        i0 = i1 - 1
        w = ((xf - xc[i0]) / (xc[i1] - xc[i0])).clamp(0.0, 1.0).to(torch.float32)
        self.N, self.C = int(xf.numel()), int(C)
        self.i0, self.i1 = i0.to(torch.int32).contiguous(), i1.to(torch.int32).contiguous()
        self.w0, self.w1 = (1.0 - w).contiguous(), w.contiguous()
        self.unit_ptr = torch.arange(self.N + 1, dtype=torch.int32, device=xf.device)
        # adjoint: for every coarse node the fine nodes that read it, with their weights (sorted by coarse id)
        idx = torch.cat([i0, i1]).long()
        wt = torch.cat([self.w0, self.w1])
        node = torch.cat([torch.arange(self.N, device=xf.device)] * 2)
        order = torch.argsort(idx, stable=True)
        self.adj_node = node[order].to(torch.int32).contiguous()
        self.adj_w = wt[order].contiguous()
        cnt = torch.bincount(idx, minlength=C)
        ptr = torch.zeros(C + 1, dtype=torch.int64, device=xf.device)
        ptr[1:] = torch.cumsum(cnt, 0)
        self.adj_ptr = ptr.to(torch.int32).contiguous()
        self.adj_unit_ptr = torch.arange(2 * self.N + 1, dtype=torch.int32, device=xf.device)


class _InterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xc, im: InterpMap):
        ctx.im = im
        xc = xc.contiguous()
        return _gather_rows(xc, im.i0, im.unit_ptr, im.w0) + _gather_rows(xc, im.i1, im.unit_ptr, im.w1)

    @staticmethod
    def backward(ctx, dout):
        im = ctx.im
        # weighted rows in adjoint order, then a segmented sum per coarse node
        rows = _gather_rows(dout.contiguous(), im.adj_node, im.adj_unit_ptr, im.adj_w)
        return ops.segment_reduce(rows, im.adj_ptr, N=im.C), None


def unpool_interp(xc: torch.Tensor, im: InterpMap) -> torch.Tensor:
    """coarse [C, 128] -> fine [N, 128] by 1-D linear interpolation."""
    return _InterpFn.apply(xc, im)
