"""MCMC moves used inside the SMC samplers -- NumPy, vectorised over particles.  TEST INFRASTRUCTURE.

Follows (under /root/reference/mocat/src/):
  utils.py:108-146                 _leapfrog
  mcmc/sampler.py:92-111           MCMCSampler.update  (always -> proposal -> correction)
  mcmc/metropolis.py:48-70         Metropolis.correct  (alpha, NaN -> 0, accept iff u < alpha)
  mcmc/standard_mcmc.py:21-65      RandomWalk
  mcmc/standard_mcmc.py:72-153     Underdamped (friction=inf: MALA for 1 leapfrog step, HMC for L)
  abc/mcmc.py:40-76                RandomWalkABC
Randomness (normals z, uniforms u) is passed in explicitly so the CUDA kernels can be fed the
same numbers.
"""
import numpy as np


def leapfrog(potential_and_grad, x, p, g, stepsize, steps, dtype=np.float64):
    """utils.py:117-134.  potential_and_grad(x) -> (U, grad) of the *tempered* potential.
    Returns (x', p', U', g').  `dtype=np.float32` reproduces the reference's fp32 arithmetic
    (used for the known-answer test tests/test_utils.py:137-163)."""
    eps = dtype(stepsize)
    half = dtype(stepsize) / dtype(2.0)
    x = np.asarray(x, dtype)
    p = np.asarray(p, dtype)
    g = np.asarray(g, dtype)
    u = None
    for _ in range(steps):
        p_half = p - half * g                    # :120
        x = x + eps * p_half                     # :122
        u, g = potential_and_grad(x)             # :124-131
        g = np.asarray(g, dtype)
        p = p_half - half * g                    # :133
    return x, p, u, g


def metropolis_accept(alpha, u):
    """metropolis.py:55-66: NaN -> 0; accept iff u < alpha."""
    alpha = np.where(np.isnan(alpha), 0.0, alpha)
    return alpha, (u < alpha)


def hmc_step(potential_and_grad, x, u_cur, g_cur, z, u, stepsize, leapfrog_steps=1):
    """One Underdamped(friction=inf) update (standard_mcmc.py:107-153, sampler.py:92-111).

    always():  p = -p ; p = p*exp(-inf*eps) + sqrt(1 - exp(-2 inf eps)) z  ==  z        (:116-122)
    proposal(): leapfrog, then p' = -p'                                              (:125-143)
    alpha = min(1, exp(-U' + U - 0.5|p'|^2 + 0.5|p|^2))                               (:145-153)
    Returns (x_new, U_new, g_new, alpha, accepted)."""
    p = np.asarray(z, np.float64)
    xp, pp, up, gp = leapfrog(potential_and_grad, x, p, g_cur, stepsize, leapfrog_steps)
    pp = -pp
    with np.errstate(over='ignore', invalid='ignore'):
        alpha = np.minimum(1.0, np.exp(-up + u_cur - 0.5 * np.sum(pp * pp, -1) + 0.5 * np.sum(p * p, -1)))
    alpha, acc = metropolis_accept(alpha, np.asarray(u, np.float64))
    a = acc[:, None]
    return (np.where(a, xp, x), np.where(acc, up, u_cur), np.where(a, gp, g_cur), alpha, acc)


def rw_step(potential, x, u_cur, z, u, stepsize):
    """RandomWalk (standard_mcmc.py:43-65): x' = x + sqrt(stepsize) z ; alpha = min(1, exp(-U'+U))."""
    xp = np.asarray(x, np.float64) + np.sqrt(stepsize) * np.asarray(z, np.float64)
    up = potential(xp)
    with np.errstate(over='ignore', invalid='ignore'):
        alpha = np.minimum(1.0, np.exp(-up + u_cur))
    alpha, acc = metropolis_accept(alpha, np.asarray(u, np.float64))
    return np.where(acc[:, None], xp, x), np.where(acc, up, u_cur), alpha, acc


def rw_abc_step(prior_potential, simulate_distance, x, up_cur, dist_cur, z, u, stepsize, threshold):
    """RandomWalkABC (abc/mcmc.py:56-76): x' = x + sqrt(stepsize) (.) z (stepsize may be a d-vector);
    alpha = min(1, exp(-Up' + Up)) * 1[dist' < threshold]."""
    xp = np.asarray(x, np.float64) + np.sqrt(np.asarray(stepsize, np.float64)) * np.asarray(z, np.float64)
    upp = prior_potential(xp)
    distp = simulate_distance(xp)
    with np.errstate(over='ignore', invalid='ignore'):
        alpha = np.minimum(1.0, np.exp(-upp + up_cur) * (distp < threshold))
    alpha, acc = metropolis_accept(alpha, np.asarray(u, np.float64))
    return (np.where(acc[:, None], xp, x), np.where(acc, upp, up_cur), np.where(acc, distp, dist_cur),
            alpha, acc)
