"""Helpers shared by the golden-fixture generator and the tests (no GPU, no product imports)."""
from __future__ import annotations

import math

import numpy as np
import torch


def _hash_uniform(n: int, t: int) -> torch.Tensor:
    """Exact integer hash (int64 arithmetic mod 2^32) of (index, tensor id) -> float64 uniform in [-1, 1)."""
    M = 0xFFFFFFFF
    x = (torch.arange(1, n + 1, dtype=torch.int64) * 2654435761 + (t + 1) * 40503) & M
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & M
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & M
    x = x ^ (x >> 16)
    return x.double() / 2.0 ** 31 - 1.0


def formula_weights_(module: torch.nn.Module) -> None:
    """Fill every parameter with a closed-form, RNG-independent pattern (so fixtures need not store weights):
    w = U(-1, 1) / sqrt(fan_in) from an exact integer hash -- the scale of torch's default Linear/Conv init."""
    with torch.no_grad():
        fan = {}
        for name, p in module.state_dict().items():
            if p.dim() > 1:
                fan[name.rsplit(".", 1)[0]] = max(1, p[0].numel())
        for t, (name, p) in enumerate(module.state_dict().items()):
            owner = name.rsplit(".", 1)[0]
            f = fan.get(owner, p.numel())
            if "rnn" in name:                 # LEMcuda: every tensor is U(-1/sqrt(nhid), 1/sqrt(nhid))
                f = 128
            vals = _hash_uniform(p.numel(), t) / math.sqrt(f)
            p.copy_(vals.view(p.shape).to(p.dtype))


def tensor_digest(x: torch.Tensor) -> np.ndarray:
    """[norm, sum, projection on cos(0.37 i), 32 strided samples] of a tensor, float64."""
    v = x.detach().double().cpu().reshape(-1)
    n = v.numel()
    i = torch.arange(n, dtype=torch.float64)
    idx = torch.linspace(0, n - 1, 32).long()
    head = torch.stack([v.norm(), v.sum(), (v * torch.cos(0.37 * i)).sum()])
    return torch.cat([head, v[idx]]).numpy()


def grads_digest(module: torch.nn.Module) -> dict:
    out = {}
    for name, p in module.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[name] = tensor_digest(g)
    return out


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max|a - ref| / max|ref| (the parity metric of SURVEY.md section 4)."""
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    denom = float(ref.abs().max())
    return float((a - ref).abs().max()) / (denom if denom > 0 else 1.0)


def random_directed_graph(sizes, avg_deg, seed, isolated_frac=0.1):
    """Directed, asymmetric, dst-sorted edge list over graphs of unequal size with some nodes that
    receive no edge (and some that send none)."""
    g = torch.Generator().manual_seed(seed)
    src_all, dst_all, batch = [], [], []
    off = 0
    for b, n in enumerate(sizes):
        e = int(n * avg_deg)
        dst = torch.randint(0, n, (e,), generator=g)
        src = torch.randint(0, n, (e,), generator=g)
        keep = src != dst
        iso = torch.rand(n, generator=g) < isolated_frac
        keep &= ~iso[dst]
        src_all.append(src[keep] + off)
        dst_all.append(dst[keep] + off)
        batch.append(torch.full((n,), b, dtype=torch.long))
        off += n
    src, dst = torch.cat(src_all), torch.cat(dst_all)
    order = torch.argsort(dst * off + src, stable=True)
    ei = torch.stack([src[order], dst[order]])
    ei = torch.unique(ei, dim=1)              # dedupe; unique sorts by (src,dst) -> resort by dst
    order = torch.argsort(ei[1] * off + ei[0], stable=True)
    return ei[:, order], torch.cat(batch)
