import torch


def erdos_renyi_graph(num_nodes, edge_prob, directed=False):
    """Only reached when GraphCreator.random_probability > 0 (common/utils.py:369-373; default 0)."""
    idx = torch.combinations(torch.arange(num_nodes), r=2)
    mask = torch.rand(idx.size(0)) < edge_prob
    idx = idx[mask].t()
    if not directed:
        idx = torch.cat([idx, idx.flip(0)], dim=1)
    return idx
