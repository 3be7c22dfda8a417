"""Builders of the POD scenario descriptions consumed by the kernels (include/mocat_b200.h).

The reference describes a target by per-particle Python callables traced under jax.vmap
(mocat/src/core.py:146-261).  Device execution needs closed-form families compiled into the kernels,
so the host API flattens the built-in scenarios into these structs; anything else raises (there is no
CPU fallback by design).
"""
import math

import numpy as np

from . import _lib


def make_target(kind, dim, prior_mean=0.0, prior_std=1.0, prior_pscale=None, a=1.0, mean=None, covariance=None):
    t = _lib.Target()
    t.kind, t.dim = int(kind), int(dim)
    t.prior_mean, t.prior_std = float(prior_mean), float(prior_std)
    t.prior_pscale = float(1.0 / prior_std if prior_pscale is None else prior_pscale)
    t.a = float(a)
    if kind == _lib.LIK_GAUSSIAN:
        if dim > _lib.MB_MAX_SMALL_DIM:
            raise _lib.MocatB200Error(f"Gaussian target supports dim <= {_lib.MB_MAX_SMALL_DIM} on the device")
        m = np.zeros(dim) if mean is None else np.atleast_1d(np.asarray(mean, np.float64))
        cov = np.ones(dim) if covariance is None else np.asarray(covariance, np.float64)
        # reset_covariance (mocat/src/utils.py:247-262): precision_sqrt = inv(chol(cov)) or 1/sqrt(diag)
        ps = np.diag(1.0 / np.sqrt(cov * np.ones(dim))) if cov.ndim < 2 else np.linalg.inv(np.linalg.cholesky(cov))
        for k in range(dim):
            t.mean[k] = float(m[k])
        _lib.fill_matrix(t.prec_sqrt, ps)
    return t


def make_move(kind, stepsize, mcmc_steps=1, leapfrog_steps=1):
    m = _lib.Move()
    m.kind, m.mcmc_steps, m.leapfrog_steps, m.stepsize = int(kind), int(mcmc_steps), int(leapfrog_steps), float(stepsize)
    return m


def make_temper(max_temperature=1.0, ess_retain=0.9, ess_resample=0.5, tol=1e-5, max_search_iter=1000,
                max_iter=10000):
    p = _lib.Temper()
    p.max_temperature, p.ess_retain, p.ess_resample, p.tol = float(max_temperature), float(ess_retain), \
        float(ess_resample), float(tol)
    p.max_search_iter, p.max_iter = int(max_search_iter), int(max_iter)
    p.schedule, p.schedule_len = None, 0
    return p


def make_lg_ssm(initial_mean, initial_covariance, transition_matrix, transition_covariance, likelihood_matrix,
                likelihood_covariance):
    """TimeHomogenousLinearGaussian (mocat/src/ssm/linear_gaussian/linear_gaussian.py:142-261)."""
    f2 = lambda a: np.atleast_2d(np.asarray(a, np.float64))
    m0 = np.atleast_1d(np.asarray(initial_mean, np.float64))
    P0, F, Q, H, R = map(f2, (initial_covariance, transition_matrix, transition_covariance, likelihood_matrix,
                              likelihood_covariance))
    d, dy = F.shape[0], H.shape[0]
    if d > _lib.MB_MAX_SMALL_DIM or dy > d:
        raise _lib.MocatB200Error(f"linear-Gaussian SSM on the device needs dim_obs <= dim <= {_lib.MB_MAX_SMALL_DIM}")
    s = _lib.SSM()
    s.kind, s.dim, s.dim_obs, s.substeps = _lib.SSM_LINEAR_GAUSSIAN, d, dy, 1
    for k in range(d):
        s.m0[k] = float(m0[k])
    _lib.fill_matrix(s.L0, np.linalg.cholesky(P0))
    _lib.fill_matrix(s.F, F)
    _lib.fill_matrix(s.LQ, np.linalg.cholesky(Q))
    _lib.fill_matrix(s.H, H)                      # rows >= dim_obs stay zero
    Rps = np.eye(d)
    Rps[:dy, :dy] = np.linalg.inv(np.linalg.cholesky(R))     # reset_covariance: precision_sqrt = inv(chol)
    _lib.fill_matrix(s.Rps, Rps)
    # gaussian_potential normaliser (utils.py:79): (d_y log 2pi - log det R^-1)/2
    s.lik_const = float(0.5 * (dy * math.log(2 * math.pi) + math.log(np.linalg.det(R))))
    return s


def make_lorenz96(dim=40, forcing=8.0, dt=0.05, substeps=1, q_std=1.0, r_std=1.0, init_mean=0.0, init_std=1.0):
    """Lorenz96 (mocat/src/ssm/scenarios/lorenz96.py:29-44) with diagonal noise, H = I, RK4 flow."""
    s = _lib.SSM()
    s.kind, s.dim, s.dim_obs, s.substeps = _lib.SSM_LORENZ96, int(dim), int(dim), int(substeps)
    s.forcing, s.dt, s.q_std, s.r_std = float(forcing), float(dt), float(q_std), float(r_std)
    s.init_mean, s.init_std = float(init_mean), float(init_std)
    s.lik_const = float(0.5 * (dim * math.log(2 * math.pi) + 2.0 * dim * math.log(r_std)))
    return s


def make_gk(data, c=0.8, prior_min=0.0, prior_max=10.0, buffer=1e-5):
    data = np.asarray(data, np.float64).ravel()
    if data.shape[0] > 16:
        raise _lib.MocatB200Error("g-and-k summary on the device supports m <= 16 draws")
    g = _lib.GK()
    g.m, g.c, g.prior_min, g.prior_max, g.buffer = data.shape[0], float(c), float(prior_min), float(prior_max), float(buffer)
    for k in range(data.shape[0]):
        g.data[k] = float(data[k])
    return g
