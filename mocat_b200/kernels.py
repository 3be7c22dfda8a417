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
        """k(x, y) of one pair, evaluated on the device (mb_gaussian_kernel)"""
        torch = _torch()
        L = _lib.get()
        xd = torch.as_tensor(np.atleast_1d(np.asarray(x, np.float32)), device="cuda").contiguous()
        yd = torch.as_tensor(np.atleast_1d(np.asarray(y, np.float32)), device="cuda").contiguous()
        out = torch.empty(1, dtype=torch.float32, device="cuda")
        L.call("mb_gaussian_kernel", L.ctx(), _lib.ptr(xd), _lib.ptr(yd), xd.numel(), float(bandwidth), _lib.ptr(out),
               _lib.stream())
        return float(out.item())

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
