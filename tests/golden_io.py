"""Load the committed golden fixtures (tests/golden/*.npz, written by make_golden.py)."""
from __future__ import annotations

import os

import numpy as np
import torch

from msmp_pde_b200.compat.torch_geometric.data import Data
from msmp_pde_b200.synth import SyntheticPDE

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name))
    return {k: z[k] for k in z.files}


def model_inputs(g: dict, pde_name: str):
    data = Data()
    for k, v in g.items():
        if k.startswith("in_"):
            setattr(data, k[3:], torch.from_numpy(v))
    nx = int((data.batch == 0).sum())
    pde = SyntheticPDE(pde_name, L=float(g["pde_L"]), tmax=float(g["pde_tmax"]), grid_size=(250, nx),
                       dt=float(g["pde_dt"]))
    return pde, data


def grad_digests(g: dict) -> dict:
    return {k[5:]: v for k, v in g.items() if k.startswith("gdig_")}
