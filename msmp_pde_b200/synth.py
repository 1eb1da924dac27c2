"""Synthetic inputs of the BASELINE.json shapes.  Host-side, CPU tensors.

Every generator returns ``(pde, data, meta)``: ``pde`` is an attribute bag with what the models read
(``L, tmax, dt, grid_size`` -- equations/PDEs.py:74-85,274-285), ``data`` a torch_geometric-style
``Data`` with ``x, y, pos, edge_index, batch`` (+ per-experiment parameters), exactly the fields
``GraphCreator.create_graph`` (common/utils.py:320-426) produces.  dtype defaults to float64 -- the
reference feeds float64 (temporal/solvers.py:10).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from .compat.torch_cluster import knn_graph, radius_graph
from .compat.torch_geometric.data import Data


@dataclass
class SyntheticPDE:
    """Stand-in for equations/PDEs.py CE / AD objects: only the attributes the hot path reads."""
    name: str = "CE"
    L: float = 16.0
    tmin: float = 0.0
    tmax: float = 4.0
    grid_size: tuple = (250, 100)
    dt: float = field(default=0.0)
    untructured_grid: bool = False       # sic (equations/PDEs.py:296)

    def __post_init__(self):
        if not self.dt:
            self.dt = self.tmax / (self.grid_size[0] - 1)

    def __repr__(self):
        return self.name


def pseudo_random_grid(xmin: float, xmax: float, n: int) -> np.ndarray:
    """generate/generate_data.py:80-113 (deterministic LCG grid of the RPU experiments)."""
    c, p, a = 74, 2 ** 16 + 1, 75
    ns = [c % p]
    for _ in range(n - 1):
        ns.append((a * ns[-1] + c) % p)
    arr = np.array(ns) / max(ns)
    arr = sorted(arr * (xmax - xmin) + xmin)
    arr[0], arr[-1] = xmin, xmax
    return np.asarray(arr, dtype=np.float64)


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _finish(pde, x_grid, B, F_u, edge_index, seed, dtype, extra):
    g = _gen(seed)
    nx = x_grid.numel()
    N = B * nx
    nt = pde.grid_size[0]
    t = torch.linspace(pde.tmin, pde.tmax, nt, dtype=dtype)
    steps = torch.randint(25, max(26, nt - 49), (B,), generator=g)
    data = Data(x=torch.randn(N, F_u, generator=g, dtype=dtype), edge_index=edge_index)
    data.y = torch.randn(N, F_u, generator=g, dtype=dtype)
    t_pos = t[steps].repeat_interleave(nx)
    data.pos = torch.stack([t_pos, x_grid.repeat(B)], 1)
    data.batch = torch.arange(B).repeat_interleave(nx)
    for k, (lo, hi) in extra.items():
        vals = lo + (hi - lo) * torch.rand(B, generator=g, dtype=dtype)
        setattr(data, k, vals.repeat_interleave(nx)[:, None])
    return data


def config_c1(B=16, nx=100, tw=25, neighbors=3, seed=0, dtype=torch.float64):
    """C1: MP-PDE on E1 -- uniform grid, radius graph r = n*dx + 1e-4 (common/utils.py:365-368)."""
    pde = SyntheticPDE("CE", L=16.0, tmax=4.0, grid_size=(250, nx))
    xg = torch.linspace(0, pde.L, nx, dtype=dtype)
    batch = torch.arange(B).repeat_interleave(nx)
    radius = neighbors * (xg[1] - xg[0]) + 0.0001
    ei = radius_graph(xg.repeat(B), r=radius, batch=batch, loop=False)
    data = _finish(pde, xg, B, tw, ei, seed, dtype, {})
    return pde, data, dict(name="C1", model="MP-PDE", eq_variables={}, tw=tw, B=B, nx=nx)


def config_c2(B=64, nx=100, tw=25, neighbors=3, seed=0, dtype=torch.float64):
    """C2: MSMP-PDE2D on RP -- two fields, uniform grid, radius graph; a~U(.1,1), b~U(1,10)."""
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(250, nx))
    xg = torch.linspace(0, pde.L, nx, dtype=dtype)
    batch = torch.arange(B).repeat_interleave(nx)
    radius = neighbors * (xg[1] - xg[0]) + 0.0001
    ei = radius_graph(xg.repeat(B), r=radius, batch=batch, loop=False)
    data = _finish(pde, xg, B, 2 * tw, ei, seed, dtype, {"a": (0.1, 1.0), "b": (1.0, 10.0)})
    return pde, data, dict(name="C2", model="MSMP-PDE2D", eq_variables={"a": 1.0, "b": 1.0}, tw=tw, B=B, nx=nx)


def config_c3(B=64, nx=100, tw=25, neighbors=3, seed=0, dtype=torch.float64, radius=None):
    """C3: MSMP-PDE2D on RPU -- pseudo-random grid, kNN on the periodic embedding
    (common/utils.py:343-346,376-377); ``radius`` switches to the irregular in-degree variant."""
    pde = SyntheticPDE("AD", L=16.0, tmax=4.0, grid_size=(250, nx), untructured_grid=True)
    xg = torch.tensor(pseudo_random_grid(0.0, pde.L, nx), dtype=dtype)
    batch = torch.arange(B).repeat_interleave(nx)
    if radius is None:
        X = 2 * np.pi * xg / (torch.max(xg) - 1e-3)
        x_per = torch.stack([torch.cos(X), torch.sin(X)], 1)
        ei = knn_graph(x_per.repeat(B, 1), k=neighbors, batch=batch, loop=False)
    else:
        ei = radius_graph(xg.repeat(B), r=radius, batch=batch, loop=False)
    data = _finish(pde, xg, B, 2 * tw, ei, seed, dtype, {"a": (0.1, 1.0), "b": (1.0, 10.0)})
    return pde, data, dict(name="C3", model="MSMP-PDE2D", eq_variables={"a": 1.0, "b": 1.0}, tw=tw, B=B, nx=nx)


def lattice_edges(side: int, eight: bool = False) -> torch.Tensor:
    """Directed edges of a side x side lattice (4- or 8-neighbour), sorted by destination."""
    idx = torch.arange(side * side).view(side, side)
    offs = [(-1, 0), (0, -1), (0, 1), (1, 0)]
    if eight:
        offs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    src, dst = [], []
    for di, dj in offs:
        i0, i1 = max(0, -di), side - max(0, di)
        j0, j1 = max(0, -dj), side - max(0, dj)
        dst.append(idx[i0:i1, j0:j1].reshape(-1))
        src.append(idx[i0 + di:i1 + di, j0 + dj:j1 + dj].reshape(-1))
    src, dst = torch.cat(src), torch.cat(dst)
    order = torch.argsort(dst * (side * side) + src)
    return torch.stack([src[order], dst[order]])


def config_c4(B=8, side=128, tw=25, seed=0, dtype=torch.float64, eight=False):
    """C4: MSMP-PDE2D (two-field model) on a synthetic side x side lattice graph per sample
    (the reference has no 2-D grid; topology is synthetic, math unchanged)."""
    n = side * side
    pde = SyntheticPDE("AD", L=2 * math.pi, tmax=1.0, grid_size=(250, n))
    e1 = lattice_edges(side, eight)
    ei = torch.cat([e1 + b * n for b in range(B)], 1)
    xg = torch.linspace(0, pde.L, n, dtype=dtype)
    data = _finish(pde, xg, B, 2 * tw, ei, seed, dtype, {"a": (0.1, 1.0), "b": (1.0, 10.0)})
    return pde, data, dict(name="C4", model="MSMP-PDE2D", eq_variables={"a": 1.0, "b": 1.0}, tw=tw, B=B, nx=n)


def large_graph(n_nodes: int, degree: int, topology: str = "band", nodes_per_graph: int = 0, seed: int = 0,
                tw: int = 25, dtype=torch.float32):
    """C5: inputs of a single GNN_Layer(128,128,128,tw,1) on a large synthetic graph.

    topology: 'band' (sources i-d/2..i+d/2 on a ring: gather friendly), 'random' (uniform random
    sources: gather hostile).  Every node has in-degree ``degree``; edges sorted by destination."""
    g = _gen(seed)
    dst = torch.arange(n_nodes).repeat_interleave(degree)
    if topology == "band":
        offs = torch.tensor([o for o in range(-(degree // 2), degree - degree // 2 + 1) if o != 0][:degree])
        src = (torch.arange(n_nodes).view(-1, 1) + offs.view(1, -1)) % n_nodes
        src = src.reshape(-1)
    elif topology == "random":
        src = torch.randint(0, n_nodes, (n_nodes * degree,), generator=g)
    else:
        raise ValueError(topology)
    npg = nodes_per_graph or n_nodes
    batch = torch.arange(n_nodes) // npg
    if nodes_per_graph:           # keep edges inside their graph (InstanceNorm is per graph)
        g0 = (dst // npg) * npg
        gsize = torch.clamp(n_nodes - g0, max=npg)         # the last graph may be smaller
        src = g0 + (src % gsize)
    return dict(
        x=torch.randn(n_nodes, 128, generator=g, dtype=dtype),
        u=torch.randn(n_nodes, tw, generator=g, dtype=dtype),
        pos=torch.rand(n_nodes, 1, generator=g, dtype=dtype),
        variables=torch.rand(n_nodes, 1, generator=g, dtype=dtype),
        edge_index=torch.stack([src, dst]), batch=batch)
