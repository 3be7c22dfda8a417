"""Kernelised Stein discrepancy -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/:
  metrics.py:88-130      ksd (k_0_inds :116-124, weights exp(log_weight) :108-111, normalisation :126-130)
  kernels.py:82-116      Gaussian kernel: _call :90-95, _grad_x :97-102, _grad_y :104-109, _diag_grad_xy :111-116
`reference_sign=True` contracts the kernel gradients with grad_potential exactly as metrics.py:119-121 does;
False uses the score -grad_potential (the Stein kernel whose discrepancy vanishes for an exact sample).
"""
import numpy as np


def gaussian_call(x, y, bandwidth):                                   # kernels.py:90-95
    diff = (x - y) / bandwidth
    return np.exp(-0.5 * np.sum(np.square(diff), axis=-1))


def gaussian_grad_x(x, y, bandwidth):                                 # kernels.py:97-102
    return (y - x) * gaussian_call(x, y, bandwidth)[..., None] / bandwidth ** 2


def gaussian_grad_y(x, y, bandwidth):                                 # kernels.py:104-109
    return (x - y) * gaussian_call(x, y, bandwidth)[..., None] / bandwidth ** 2


def gaussian_diag_grad_xy(x, y, bandwidth):                           # kernels.py:111-116
    return (bandwidth ** 2 - (x - y) ** 2) * gaussian_call(x, y, bandwidth)[..., None] / bandwidth ** 4


def ksd(vals, grad_potential, bandwidth, log_weight=None, reference_sign=True):
    """metrics.py:88-130 without ensemble minibatching, O(n^2 d) in fp64, one row of the pair matrix at a time"""
    x = np.asarray(vals, np.float64)
    g = np.asarray(grad_potential, np.float64)
    n = len(x)
    w = np.exp(np.asarray(log_weight, np.float64)) if log_weight is not None else np.ones(n)     # :108-111
    sgn = 1.0 if reference_sign else -1.0
    total = 0.0
    for i in range(n):
        xi, gi = x[i][None], g[i][None]
        k0 = (np.sum(gaussian_diag_grad_xy(xi, x, bandwidth), axis=-1)                             # :117
              + sgn * np.sum(gaussian_grad_x(xi, x, bandwidth) * g, axis=-1)                       # :118
              + sgn * np.sum(gi * gaussian_grad_y(xi, x, bandwidth), axis=-1)                      # :119
              + gaussian_call(xi, x, bandwidth) * np.sum(gi * g, axis=-1))                         # :120-121
        total += np.sum(k0 * w[i] * w) / (w.sum() * w.sum())                                       # :123-127
    return float(np.sqrt(total))                                                                   # :130
