"""Potentials and simulators of the built-in scenarios -- NumPy fp64.  TEST INFRASTRUCTURE.

Follows (under /root/reference/mocat/src/):
  utils.py:49-105                       gaussian_potential
  scenarios/toy_examples.py:17-51       Gaussian.likelihood_potential
  scenarios/toy_examples.py:135-149     Rastrigin.likelihood_potential
  core.py:190-194                       tempered_potential = prior + T * likelihood
  ssm/linear_gaussian/linear_gaussian.py:86-94,118-128   LG transition_sample / likelihood_potential
  ssm/nonlinear_gaussian.py:107-121     NonLinearGaussian transition_sample / likelihood_potential
  ssm/scenarios/lorenz96.py:14-26       lorenz96_dynamics / integrator
  abc/scenarios/gk.py:68-96             GKTransformedUniformPrior
All functions are vectorised over a leading particle axis (what jax.vmap does upstream).
"""
import numpy as np
from scipy.special import ndtr, ndtri

LOG_2PI = float(np.log(2.0 * np.pi))


# ----------------------------------------------------------------------------- utils.py:49-105
def gaussian_potential(x, mean=0.0, prec=None, sqrt_prec=None, det_prec=None):
    """0.5 * quad + (d*log(2 pi) - log det_prec)/2 exactly as utils.py:49-105 (incl. the
    'no det_prec for a matrix -> no normalising constant' branch, :73-77)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    d = x.shape[-1]
    if prec is None and sqrt_prec is None:
        prec = 1.0
    if prec is not None and np.ndim(prec) == 0:
        prec = np.ones(d) * prec
    if sqrt_prec is not None and np.ndim(sqrt_prec) == 0:
        sqrt_prec = np.ones(d) * sqrt_prec
    if det_prec is None:
        if prec is not None and np.ndim(prec) < 2:
            det_prec = np.prod(prec)
        elif sqrt_prec is not None and np.ndim(sqrt_prec) < 2:
            det_prec = np.prod(sqrt_prec) ** 2
    neg_log_z = 0.0 if det_prec is None else (d * LOG_2PI - np.log(det_prec)) / 2.0
    diff = x - mean
    if sqrt_prec is None:
        prec = np.asarray(prec, dtype=np.float64)
        if prec.ndim < 2:
            out = 0.5 * np.sum(diff ** 2 * prec, axis=-1)                      # :42-46
        elif x.ndim == 1:
            out = 0.5 * diff @ prec @ diff                                     # :34-38
        else:
            out = 0.5 * np.sum((diff @ np.linalg.cholesky(prec)) ** 2, axis=-1)  # :93-96
    else:
        sqrt_prec = np.asarray(sqrt_prec, dtype=np.float64)
        if sqrt_prec.ndim < 2:
            out = 0.5 * np.sum(diff ** 2 * sqrt_prec ** 2, axis=-1)            # :102
        else:
            out = 0.5 * np.sum((diff @ sqrt_prec) ** 2, axis=-1)               # :26-30,104
    return out + neg_log_z


# ----------------------------------------------------------------------------- static targets
class IsoGaussianPrior:
    """prior_sample = mean + std * z ;  prior_potential = 0.5 * sum(((x - mean) * pscale)^2).

    pscale is free so that the reference's fixture quirk `0.5*square(x/7**2)` with
    `prior_sample = 7 z` (tests/test_transport.py:27-28) is representable (std=7, pscale=1/49);
    the consistent choice is pscale = 1/std."""

    def __init__(self, dim, mean=0.0, std=1.0, pscale=None):
        self.dim, self.mean, self.std = dim, float(mean), float(std)
        self.pscale = (1.0 / self.std) if pscale is None else float(pscale)

    def sample(self, z):
        return self.mean + self.std * np.asarray(z, dtype=np.float64)

    def potential_and_grad(self, x):
        r = (np.asarray(x, np.float64) - self.mean) * self.pscale
        return 0.5 * np.sum(r * r, axis=-1), r * self.pscale


class Rastrigin:
    """toy_examples.py:146-149: a*d + sum(x^2 - a cos(2 pi x)); gradient 2x + 2 pi a sin(2 pi x)."""

    def __init__(self, dim, a=1.0):
        self.dim, self.a = dim, float(a)

    def potential_and_grad(self, x):
        x = np.asarray(x, np.float64)
        u = self.a * self.dim + np.sum(x ** 2 - self.a * np.cos(2 * np.pi * x), axis=-1)
        g = 2 * x + 2 * np.pi * self.a * np.sin(2 * np.pi * x)
        return u, g


class GaussianTarget:
    """toy_examples.py:40-44: x_diff = (x - mean) @ precision_sqrt.T ; 0.5 * sum(x_diff^2)
    with precision_sqrt = inv(chol(cov)) (utils.py:258-260)."""

    def __init__(self, mean, covariance):
        self.mean = np.atleast_1d(np.asarray(mean, np.float64))
        cov = np.asarray(covariance, np.float64)
        self.dim = self.mean.shape[0]
        if cov.ndim < 2:
            self.precision_sqrt = np.diag(1.0 / np.sqrt(cov * np.ones(self.dim)))
        else:
            self.precision_sqrt = np.linalg.inv(np.linalg.cholesky(cov))
        self.covariance = cov

    def potential_and_grad(self, x):
        diff = np.asarray(x, np.float64) - self.mean
        y = diff @ self.precision_sqrt.T
        return 0.5 * np.sum(y * y, axis=-1), y @ self.precision_sqrt


class LogisticRegression:
    """Config C4 scenario (none exists upstream): labels t in {0,1}, features A (N,d):
    U_lik(w) = sum_k softplus(a_k.w) - t_k a_k.w ; prior N(0, I) handled by IsoGaussianPrior."""

    def __init__(self, features, labels):
        self.A = np.asarray(features, np.float64)
        self.t = np.asarray(labels, np.float64)
        self.dim = self.A.shape[1]

    def potential_and_grad(self, x):
        x = np.asarray(x, np.float64)
        s = x @ self.A.T                                   # (n, N)
        u = np.sum(np.logaddexp(0.0, s) - self.t * s, axis=-1)
        p = 0.5 * (1.0 + np.tanh(0.5 * s))                     # sigmoid without overflow
        return u, (p - self.t) @ self.A


# ----------------------------------------------------------------------------- state-space models
class LinearGaussianSSM:
    """TimeHomogenousLinearGaussian (linear_gaussian.py:142-261).  Square-root factors follow
    reset_covariance (utils.py:247-262): L = chol(cov), precision_sqrt = inv(L),
    precision_det = 1/det(cov)."""

    def __init__(self, initial_mean, initial_covariance, transition_matrix, transition_covariance,
                 likelihood_matrix, likelihood_covariance):
        f = lambda a: np.atleast_2d(np.asarray(a, np.float64))
        self.m0 = np.atleast_1d(np.asarray(initial_mean, np.float64))
        self.P0, self.F, self.Q, self.H, self.R = map(f, (initial_covariance, transition_matrix,
                                                          transition_covariance, likelihood_matrix,
                                                          likelihood_covariance))
        self.dim, self.dim_obs = self.F.shape[0], self.H.shape[0]
        self.L0, self.LQ, self.LR = (np.linalg.cholesky(a) for a in (self.P0, self.Q, self.R))
        self.R_prec_sqrt = np.linalg.inv(self.LR)
        self.R_prec_det = 1.0 / np.linalg.det(self.R)

    def initial_sample(self, z):                                    # linear_gaussian.py:45-50
        return np.asarray(z, np.float64) @ self.L0.T + self.m0

    def transition_sample(self, x, z):                              # :86-94  F x + L_Q z
        return np.asarray(x, np.float64) @ self.F.T + np.asarray(z, np.float64) @ self.LQ.T

    def likelihood_potential(self, x, y):                           # :118-128
        return gaussian_potential(np.asarray(y, np.float64), np.asarray(x, np.float64) @ self.H.T,
                                  sqrt_prec=self.R_prec_sqrt, det_prec=self.R_prec_det)

    def simulate(self, T, rng):
        """ssm.py:138-163 with numpy Generator randomness (host helper, not parity-critical)."""
        x = np.empty((T, self.dim))
        y = np.empty((T, self.dim_obs))
        x[0] = self.initial_sample(rng.standard_normal(self.dim))
        for t in range(1, T):
            x[t] = self.transition_sample(x[t - 1], rng.standard_normal(self.dim))
        for t in range(T):
            y[t] = x[t] @ self.H.T + rng.standard_normal(self.dim_obs) @ self.LR.T
        return x, y


def lorenz96_rhs(x, forcing=8.0):
    """lorenz96.py:14-20: (x[k+1] - x[k-2]) * x[k-1] - x[k] + F, cyclic."""
    return (np.roll(x, -1, axis=-1) - np.roll(x, 2, axis=-1)) * np.roll(x, 1, axis=-1) - x + forcing


def lorenz96_rk4(x, dt, forcing=8.0, substeps=1):
    """Device definition of the L96 transition map: `substeps` classical RK4 steps of size
    dt/substeps (the reference uses adaptive odeint, lorenz96.py:23-26)."""
    x = np.asarray(x, np.float64)
    h = dt / substeps
    for _ in range(substeps):
        k1 = lorenz96_rhs(x, forcing)
        k2 = lorenz96_rhs(x + 0.5 * h * k1, forcing)
        k3 = lorenz96_rhs(x + 0.5 * h * k2, forcing)
        k4 = lorenz96_rhs(x + h * k3, forcing)
        x = x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def lorenz96_dopri(x, dt, forcing=8.0, rtol=1.4e-8, atol=1.4e-8):
    """Secondary oracle: the reference's flow (jax odeint = adaptive Dormand-Prince with
    rtol=atol=1.4e-8) via SciPy RK45 in fp64, one particle at a time."""
    from scipy.integrate import solve_ivp
    x = np.atleast_2d(np.asarray(x, np.float64))
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        sol = solve_ivp(lambda t, s: lorenz96_rhs(s, forcing), (0.0, dt), x[i], method='RK45',
                        rtol=rtol, atol=atol)
        out[i] = sol.y[:, -1]
    return out


class Lorenz96SSM:
    """Lorenz96(NonLinearGaussian) with diagonal noise: x' = Phi(x) + z * sq ; y ~ N(x, sr^2)
    (nonlinear_gaussian.py:107-121 with H = I and diagonal Q, R -- config C3 uses Q=R=I)."""

    pairwise_normals = True      # csrc/pf_l96.cu draws the normals of a particle PAIR from one Philox stream

    def __init__(self, dim=40, forcing=8.0, dt=0.05, substeps=1, q_std=1.0, r_std=1.0,
                 init_mean=0.0, init_std=1.0):
        self.dim = self.dim_obs = dim
        self.forcing, self.dt, self.substeps = float(forcing), float(dt), int(substeps)
        self.q_std, self.r_std = float(q_std), float(r_std)
        self.init_mean, self.init_std = float(init_mean), float(init_std)

    def initial_sample(self, z):
        return self.init_mean + self.init_std * np.asarray(z, np.float64)

    def transition_function(self, x):
        return lorenz96_rk4(x, self.dt, self.forcing, self.substeps)

    def transition_sample(self, x, z):
        return self.transition_function(x) + self.q_std * np.asarray(z, np.float64)

    def likelihood_potential(self, x, y):
        d = self.dim
        r = (np.asarray(y, np.float64) - np.asarray(x, np.float64)) / self.r_std
        return 0.5 * np.sum(r * r, axis=-1) + 0.5 * (d * LOG_2PI + 2.0 * d * np.log(self.r_std))

    def simulate(self, T, rng, spinup=1000):
        x = self.initial_sample(rng.standard_normal(self.dim)) + self.forcing
        for _ in range(spinup):
            x = self.transition_function(x)
        xs = np.empty((T, self.dim))
        ys = np.empty((T, self.dim))
        for t in range(T):
            if t > 0:
                x = self.transition_sample(x, rng.standard_normal(self.dim))
            xs[t] = x
            ys[t] = x + self.r_std * rng.standard_normal(self.dim)
        return xs, ys


# ----------------------------------------------------------------------------- g-and-k (ABC)
class GKTransformed:
    """GKTransformedUniformPrior (gk.py:68-96): theta = min + Phi(x)(max-min); prior N(0,I);
    u ~ U(buffer, 1-buffer)^m ; z = Phi^-1(u) ; y = A + B(1 + c(1-e^{-gz})/(1+e^{-gz})) z (1+z^2)^k.
    Summary: the m draws sorted ascending; distance = L2 to data
    (abc/abc.py:36-38)."""

    def __init__(self, data, c=0.8, prior_min=0.0, prior_max=10.0, buffer=1e-5):
        self.data = np.asarray(data, np.float64)
        self.m = self.data.shape[0]
        self.c, self.lo, self.hi, self.buffer = float(c), float(prior_min), float(prior_max), float(buffer)
        self.dim = 4

    def constrain(self, x):
        return self.lo + ndtr(np.asarray(x, np.float64)) * (self.hi - self.lo)

    def prior_potential(self, x):
        return 0.5 * np.sum(np.asarray(x, np.float64) ** 2, axis=-1)

    def simulate(self, x, u01):
        """u01: (n, m) uniforms in [0,1) -> sorted simulated data (n, m)."""
        th = self.constrain(x)
        u = self.buffer + np.asarray(u01, np.float64) * (1.0 - 2.0 * self.buffer)
        z = ndtri(u)
        e = np.exp(-th[:, 2:3] * z)
        y = th[:, 0:1] + th[:, 1:2] * (1 + self.c * (1 - e) / (1 + e)) * z * (1 + z * z) ** th[:, 3:4]
        return np.sort(y, axis=-1)

    def distance(self, sim):
        return np.sqrt(np.sum((sim - self.data) ** 2, axis=-1))
