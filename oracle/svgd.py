"""SVGD interaction, bandwidth heuristics and adagrad -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/:
  transport/svgd.py:18-32     kernelised_grad_matrix (double vmap)
  transport/svgd.py:77-107    SVGD.startup ;  :122-146 SVGD.update
  kernels.py:90-102           Gaussian._call / _grad_x
  kernels.py:220-229          median_bandwidth_update / mean_bandwidth_update
  utils.py:437-439            l2_distance_matrix
and jax.example_libraries.optimizers.adagrad (third party, unpinned; restated from its
published source: g_sq += g^2 ; m = (1-momentum) * g * rsqrt(g_sq) (0 where g_sq == 0)
+ momentum * m ; x -= step(i) * m ; momentum = 0.9).
"""
import numpy as np


def gaussian_kernel(x, y, bandwidth=1.0):                               # kernels.py:90-95
    diff = (np.asarray(x, np.float64) - np.asarray(y, np.float64)) / bandwidth
    return np.exp(-0.5 * np.sum(diff * diff, axis=-1))


def gaussian_kernel_grad_x(x, y, bandwidth=1.0):                        # kernels.py:97-102
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    return (y - x) * gaussian_kernel(x, y, bandwidth)[..., None] / bandwidth ** 2


def phi_double_loop(X, G, bandwidth):
    """svgd.py:25-31 literally: phi_i = mean_j[ -k(x_j, x_i) g_j + grad_x k(x_j, x_i) ]  (full batch)."""
    X = np.asarray(X, np.float64)
    G = np.asarray(G, np.float64)
    n = X.shape[0]
    out = np.empty_like(X)
    for i in range(n):
        k = gaussian_kernel(X, X[i], bandwidth)                         # k(x_j, x_i) for all j
        out[i] = (-(k[:, None] * G) + gaussian_kernel_grad_x(X, X[i], bandwidth)).mean(axis=0)
    return out


def phi(X, G, bandwidth):
    """GEMM form: phi = [ -K G + (X o rowsum(K) - K X) / h^2 ] / n,
    K = exp(-(|x_i|^2 + |x_j|^2 - 2 X X^T) / (2 h^2)); equals phi_double_loop to 1e-15."""
    X = np.asarray(X, np.float64)
    G = np.asarray(G, np.float64)
    n = X.shape[0]
    sq = np.sum(X * X, axis=1)
    D2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * X @ X.T, 0.0)
    K = np.exp(-D2 / (2.0 * bandwidth ** 2))
    return (-(K @ G) + (X * K.sum(axis=1)[:, None] - K @ X) / bandwidth ** 2) / n


def l2_distance_matrix(X):                                              # utils.py:437-439
    X = np.asarray(X, np.float64)
    sq = np.sum(X * X, axis=1)
    return np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2.0 * X @ X.T, 0.0))


def l2_distance_matrix_exact(X):
    X = np.asarray(X, np.float64)
    return np.sqrt(np.sum((X[:, None, :] - X[None, :, :]) ** 2, axis=-1))


def median_bandwidth(X):                                                # kernels.py:220-224
    """median over the FULL n x n matrix incl. the zero diagonal, / sqrt(2 log n)."""
    D = l2_distance_matrix_exact(X) if len(X) <= 4096 else l2_distance_matrix(X)
    return float(np.median(D) / np.sqrt(2.0 * np.log(D.shape[0])))


def mean_bandwidth(X):                                                  # kernels.py:227-229
    D = l2_distance_matrix_exact(X) if len(X) <= 4096 else l2_distance_matrix(X)
    return float(np.mean(D) / np.sqrt(2.0 * np.log(D.shape[0])))


class Adagrad:
    """jax.example_libraries.optimizers.adagrad(step_size, momentum=0.9), used svgd.py:104-106,138-139."""

    def __init__(self, x0, stepsize, momentum=0.9):
        self.x = np.asarray(x0, np.float64).copy()
        self.g_sq = np.zeros_like(self.x)
        self.m = np.zeros_like(self.x)
        self.stepsize, self.momentum = stepsize, float(momentum)

    def update(self, i, g):
        g = np.asarray(g, np.float64)
        self.g_sq = self.g_sq + g * g
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = np.where(self.g_sq > 0, 1.0 / np.sqrt(self.g_sq), 0.0)
        self.m = (1.0 - self.momentum) * (g * inv) + self.momentum * self.m
        step = self.stepsize(i) if callable(self.stepsize) else self.stepsize
        self.x = self.x - step * self.m
        return self.x


class SVGD:
    """svgd.py:77-146 with full-batch interaction.  `bandwidth` = float (fixed) | 'mean' | 'median'."""

    def __init__(self, potential_and_grad, x0, stepsize, bandwidth='mean', max_iter=1000):
        self.pg = potential_and_grad
        self.opt = Adagrad(x0, stepsize)
        self.bandwidth_mode, self.max_iter = bandwidth, int(max_iter)
        self.x = self.opt.x
        self.U, self.G = self.pg(self.x)                                # :99
        self.h = self._adapt()                                          # :102
        self.iter = 0

    def _adapt(self):
        if self.bandwidth_mode == 'mean':
            return mean_bandwidth(self.x)
        if self.bandwidth_mode == 'median':
            return median_bandwidth(self.x)
        return float(self.bandwidth_mode)

    def update(self):                                                   # :122-146
        self.iter += 1
        p = phi(self.x, self.G, self.h)
        self.x = self.opt.update(self.iter, -p)                         # :138-139
        self.U, self.G = self.pg(self.x)
        self.h = self._adapt()
        return self.x

    def run(self):
        while self.iter < self.max_iter:
            self.update()
        return self.x
