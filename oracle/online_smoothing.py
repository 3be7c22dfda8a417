"""Fixed-lag online particle smoothing -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/ssm/online_smoothing.py:
  :21-44    full_stitch_single / full_stitch: log_weight = x1_log_weight - transition_potential(x0_single -> x1_all);
            random.categorical(key, log_weight)   (arg-max of logits + Gumbel noise)
  :167-207  fixed_lag_stitching: x0_fixed = early_block[-1]; (x0_vary, x1_vary) = recent_block[0:2];
            non_interacting_log_weight = recent_log_weight + transition_potential(x0_vary_i -> x1_vary_i);
            stitched = append(early_block, recent_block[1:, inds])
Transition potentials with the normalising constant of utils.py:49-79 (linear_gaussian.py:73-84,
nonlinear_gaussian.py:98-105).  The rejection variant (:60-164) samples the same law and is not restated.
Randomness (convention of csrc/backward.cu): Gumbel noise for the pair (fixed end i, candidate j) at time index `step` =
-log(-log(u_open(word j % 4 of Philox(seed, gid = i, step, P_STITCH, slot j // 4)))).
"""
import numpy as np
from . import philox

P_STITCH = 5


def gumbel(seed, n_s, n_c, step, purpose=P_STITCH):
    """(n_s, n_c) Gumbel(0, 1) noise of one stitching step"""
    i = np.arange(n_s, dtype=np.uint64)
    out = np.empty((n_s, 4 * ((n_c + 3) // 4)))
    for s in range((n_c + 3) // 4):
        words = philox.raw(seed, i, step, purpose, s)
        for c in range(4):
            with np.errstate(divide='ignore'):
                out[:, 4 * s + c] = -np.log(-np.log(philox.u_open(words[c]).astype(np.float64)))
    return out[:, :n_c]


def _whitened(ssm, x0, x1):
    """(transition_mean(x0) @ L_Q^-1, x1 @ L_Q^-1, log det L_Q) -- the reference's row-vector convention"""
    x0, x1 = np.asarray(x0, np.float64), np.asarray(x1, np.float64)
    if hasattr(ssm, 'LQ'):
        # (x - mean) @ inv(chol(Q)) as the reference (utils.py:26-30, :257-258): see oracle/backward.py
        a = np.linalg.solve(ssm.LQ.T, (x0 @ ssm.F.T).T).T
        b = np.linalg.solve(ssm.LQ.T, x1.T).T
        logdet = float(np.sum(np.log(np.diag(ssm.LQ))))
    else:
        a, b = ssm.transition_function(x0) / ssm.q_std, x1 / ssm.q_std
        logdet = x0.shape[-1] * float(np.log(ssm.q_std))
    return a, b, logdet


def transition_potential(ssm, x0, x1):
    """matched pairs (x0_i -> x1_i): |(x1 - mean(x0)) @ L_Q^-1|^2 / 2 + (d log 2 pi - log det prec) / 2 (utils.py:26-30,79)"""
    a, b, logdet = _whitened(ssm, x0, x1)
    d = a.shape[-1]
    return 0.5 * np.sum((a - b) ** 2, axis=-1) + 0.5 * d * np.log(2 * np.pi) + logdet


def full_stitch(ssm, x0_fixed, x1_cand, lw1, seed, step):
    """online_smoothing.py:21-44: one index into the candidates for every fixed trajectory end"""
    a, b, _ = _whitened(ssm, x0_fixed, x1_cand)
    d2 = np.sum((a[:, None, :] - b[None, :, :]) ** 2, axis=-1)          # (n_s, n_c)
    logits = np.asarray(lw1, np.float64)[None, :] - 0.5 * d2 + gumbel(seed, a.shape[0], b.shape[0], step)
    return np.argmax(logits, axis=1)


def fixed_lag_stitching(ssm, early_block, recent_block, recent_log_weight, seed, step):
    """online_smoothing.py:167-207 (maximum_rejections = 0 branch) -> (stitched block, indices)"""
    early_block, recent_block = np.asarray(early_block), np.asarray(recent_block)
    x0_fixed = early_block[-1]
    x0_vary, x1_vary = recent_block[0], recent_block[1]
    lw = np.asarray(recent_log_weight, np.float64) + transition_potential(ssm, x0_vary, x1_vary)
    inds = full_stitch(ssm, x0_fixed, x1_vary, lw, seed, step)
    return np.append(early_block, recent_block[1:, inds], axis=0), inds
