"""Shared test helpers."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(d[k]) if d[k].ndim else d[k].item() for k in d.files}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b) -> float:
    return float((a.double() - b.double()).abs().max())
