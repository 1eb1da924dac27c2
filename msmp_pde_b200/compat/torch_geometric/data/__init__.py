"""`torch_geometric.data.Data` stand-in: an attribute bag of tensors with `.to()`.

Used by common/utils.py:388-391 (`Data(x=u, edge_index=edge_index)`, then attribute assignment) and
read by the models (`data.x, .pos, .edge_index, .batch, .a ...`)."""
from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __contains__(self, key):
        return key in self.keys()

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def apply(self, fn):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, fn(v))
        return self

    def to(self, device, *args, **kwargs):
        return self.apply(lambda t: t.to(device, *args, **kwargs))

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def clone(self):
        out = Data()
        for k in self.keys():
            v = getattr(self, k)
            setattr(out, k, v.clone() if torch.is_tensor(v) else v)
        return out

    @property
    def num_nodes(self):
        return None if self.x is None else self.x.size(0)

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else self.edge_index.size(1)

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}")
        return "Data(" + ", ".join(parts) + ")"
