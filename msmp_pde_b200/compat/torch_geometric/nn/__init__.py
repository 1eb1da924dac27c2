"""Names the reference imports from torch_geometric.nn (models_gnn.py:7, models_gnn2D.py:6).

The product's model classes do not subclass any of these; they exist so unchanged reference files
that merely *import* them keep importing.  Constructing one raises."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise ImportError(f"torch_geometric.nn.{type(self).__name__} is not provided by the msmp_pde_b200 shim")


class MessagePassing(_Unavailable): pass
class InstanceNorm(_Unavailable): pass
class BatchNorm(_Unavailable): pass
class GCNConv(_Unavailable): pass
class GATConv(_Unavailable): pass
class SAGEConv(_Unavailable): pass
class TransformerConv(_Unavailable): pass
class RGATConv(_Unavailable): pass


def global_mean_pool(*a, **k):
    raise ImportError("global_mean_pool is not provided by the msmp_pde_b200 shim (unused by the reference)")


def avg_pool_x(*a, **k):
    raise ImportError("avg_pool_x is not provided by the msmp_pde_b200 shim (unused by the reference)")
