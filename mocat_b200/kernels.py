"""Kernels: host mirror of mocat/src/kernels.py for the Gaussian RBF (kernels.py:82-116) and the
bandwidth heuristics (kernels.py:220-229).  Values are computed by the CUDA kernels."""
import numpy as np

from . import _lib, engine
from .core import cdict


def _torch():
    import torch
    return torch


class Kernel:
    def __init__(self, **kwargs):
        self.parameters = cdict(**kwargs)


class Gaussian(Kernel):
    """k(x, y) = exp(-0.5 |(x - y)/bandwidth|^2)  (kernels.py:90-95)."""

    def __init__(self, bandwidth=1.):
        super().__init__(bandwidth=bandwidth)

    def _pair(self, x, y, bandwidth):
        # evaluate through the interaction kernel: phi with G = 0, X = [x; y] gives grad terms; for the
        # scalar API a two-point problem is tiny, so use the closed form on the host scalars returned by
        # the device distance kernel.
        torch = _torch()
        X = torch.as_tensor(np.stack([np.asarray(x, np.float32), np.asarray(y, np.float32)]), device="cuda")
        # mean distance over the 2x2 matrix (two zeros on the diagonal) * sqrt(2 log 2) = |x - y| / 2
        h = engine.pairdist_bandwidth(X, "mean").item()
        dist = h * np.sqrt(2.0 * np.log(2.0)) * 2.0
        return float(np.exp(-0.5 * (dist / bandwidth) ** 2))

    def __call__(self, x, y, bandwidth=None):
        return self._pair(x, y, self.parameters.bandwidth if bandwidth is None else bandwidth)

    def grad_x(self, x, y, bandwidth=None):                              # kernels.py:97-102
        b = self.parameters.bandwidth if bandwidth is None else bandwidth
        return (np.asarray(y, np.float64) - np.asarray(x, np.float64)) * self._pair(x, y, b) / b ** 2

    def grad_y(self, x, y, bandwidth=None):                              # kernels.py:104-109
        b = self.parameters.bandwidth if bandwidth is None else bandwidth
        return (np.asarray(x, np.float64) - np.asarray(y, np.float64)) * self._pair(x, y, b) / b ** 2


def _bandwidth(vals, mode, variant=None):
    torch = _torch()
    if isinstance(vals, torch.Tensor):
        return engine.pairdist_bandwidth(vals.contiguous(), mode, variant)   # stays on the device
    X = torch.as_tensor(np.asarray(vals, np.float32), device="cuda").contiguous()
    return float(engine.pairdist_bandwidth(X, mode, variant).item())


def median_bandwidth_update(vals, variant=None):
    """kernels.py:220-224: median of the full n x n distance matrix / sqrt(2 log n).
    variant: engine.interaction_variant (None = tensor cores for n >= 2048, exact fp32 below)."""
    return _bandwidth(vals, "median", variant)


def mean_bandwidth_update(vals, variant=None):
    """kernels.py:227-229."""
    return _bandwidth(vals, "mean", variant)
