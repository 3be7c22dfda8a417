"""Backward simulation (FFBSi) -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/ssm/backward.py:
  :20-40    full_resample_single / full_resampling: log_weight = x0_log_weight - transition_potential(x0_all -> x1_single);
            x0_all[random.categorical(key, log_weight)]   (jax.random.categorical = arg-max of logits + Gumbel noise)
  :241-272  backward_simulation_full: final particles ~ Cat(log_weight[-1]); scan back over ind = T-2 .. 0
Randomness (convention of csrc/backward.cu, the reference's threefry streams are unpinned): Gumbel noise for the pair
(backward sample j, filter particle i) at time index `step` = -log(-log(u_open(word i % 4 of Philox(seed, gid = j, step,
P_BACKWARD, slot i // 4)))).
"""
import numpy as np
from . import philox

P_BACKWARD = 4


def gumbel(seed, n_s, n_pf, step):
    """(n_s, n_pf) Gumbel(0, 1) noise of one backward step"""
    j = np.arange(n_s, dtype=np.uint64)
    out = np.empty((n_s, 4 * ((n_pf + 3) // 4)))
    for s in range((n_pf + 3) // 4):
        words = philox.raw(seed, j, step, P_BACKWARD, s)
        for c in range(4):
            with np.errstate(divide='ignore'):
                out[:, 4 * s + c] = -np.log(-np.log(philox.u_open(words[c]).astype(np.float64)))
    return out[:, :n_pf]


def full_resampling(ssm, x0, lw0, x1, seed, step):
    """backward.py:20-40; ssm: oracle.models.LinearGaussianSSM or Lorenz96SSM.  x1 None: no transition term.
    Returns (indices (n_s,), x0[indices])"""
    x0 = np.asarray(x0, np.float64)
    lw0 = np.asarray(lw0, np.float64)
    n_pf = x0.shape[0]
    if x1 is None:
        raise ValueError("use final_draw for the final-time categorical")
    x1 = np.asarray(x1, np.float64)
    if hasattr(ssm, 'LQ'):                                              # linear Gaussian: mean F x0, whiten with chol(Q)
        mean = x0 @ ssm.F.T
        # the reference evaluates 0.5 |(x1 - mean) @ inv(chol(Q))|^2 (utils.py:26-30, reset_covariance :257-258): row
        # vector times L^-1, i.e. the precision (L^T L)^-1 -- equal to Q^-1 for diagonal Q only.  Mirrored exactly.
        a = np.linalg.solve(ssm.LQ.T, mean.T).T
        b = np.linalg.solve(ssm.LQ.T, x1.T).T
    else:                                                               # Lorenz-96: RK4 flow, isotropic process noise
        mean = ssm.transition_function(x0)
        a, b = mean / ssm.q_std, x1 / ssm.q_std
    d2 = np.sum((a[None, :, :] - b[:, None, :]) ** 2, axis=-1)          # (n_s, n_pf)
    logits = lw0[None, :] - 0.5 * d2 + gumbel(seed, x1.shape[0], n_pf, step)
    idx = np.argmax(logits, axis=1)
    return idx, x0[idx]


def final_draw(x, lw, n_s, seed, step):
    """backward.py:254-256: marg_particles_vals[-1, random.categorical(key, marginal_log_weight[-1], shape=(n_samps,))]"""
    lw = np.asarray(lw, np.float64)
    idx = np.argmax(lw[None, :] + gumbel(seed, n_s, lw.shape[0], step), axis=1)
    return idx, np.asarray(x)[idx]


def backward_simulation(ssm, values, log_weights, n_s, seed):
    """backward.py:241-272 for stacked filter output values (T, n_pf, d), log_weights (T, n_pf) -> (T, n_s, d)"""
    T = len(values)
    out = np.empty((T, n_s, np.asarray(values).shape[2]))
    _, out[T - 1] = final_draw(values[T - 1], log_weights[T - 1], n_s, seed, T - 1)
    for ind in range(T - 2, -1, -1):
        _, out[ind] = full_resampling(ssm, values[ind], log_weights[ind], out[ind + 1], seed, ind)
    return out
