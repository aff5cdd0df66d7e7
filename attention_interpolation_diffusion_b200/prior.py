"""Coefficient schedule of the interpolation frames (host side, once per sequence).

Mirrors ``generate_beta_tensor`` of the reference (prior.py:481-502): the i-th
coefficient is the Beta(alpha, beta) quantile of i/(size-1).  Only this function
of prior.py is on the hot path's boundary (interpolation.py:7, :21); the
exploration / Bayesian search of prior.py is out of scope.
"""
import numpy as np
import torch
from scipy.stats import beta as _beta_distribution


def generate_beta_tensor(size: int, alpha: float = 3, beta: float = 3) -> torch.FloatTensor:
    quantile_levels = np.arange(size, dtype=np.float64) / max(size - 1, 1)
    return torch.tensor(_beta_distribution.ppf(quantile_levels, alpha, beta), dtype=torch.float32)
