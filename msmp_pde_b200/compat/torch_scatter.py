"""`torch_scatter.scatter` stand-in (call site: models_gnn2D.py:600-601).  Deterministic (index_add
on CPU; on CUDA the product routes mean-aggregation through its own segmented kernels instead)."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
        return res / cnt.view(-1, *([1] * (src.dim() - 1)))
    raise NotImplementedError(reduce)
