"""Bootstrap particle filter and Kalman filter -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/:
  ssm/filtering.py:154-170   BootstrapFilter (proposal = transition_sample; incr = -likelihood_potential)
  ssm/filtering.py:173-193   initiate_particles  (w0 = -likelihood_potential(x0, y0))
  ssm/filtering.py:255-324   run_particle_filter_for_marginals (resample iff ess_prev < thr*n, strict)
  ssm/linear_gaussian/kalman.py:16-57   run_kalman_filter_for_marginals
The reference accumulates no log-evidence in the filter and none in the Kalman filter; the
convention adopted is the SMC-sampler one already in the reference
(transport/smc.py:160,212-215):  log Z_0 = LSE(w_0) - log n ;
log Z_t = log Z_{t-1} + LSE(w_t) - LSE(w_{t-1} after resampling), checked against the Kalman
innovation log-likelihood added here.
Randomness: oracle/philox.py streams; step index = time index (0 = initial sample).
"""
import numpy as np
from . import core, philox


class BootstrapPF:
    def __init__(self, ssm, n, seed, ess_threshold=0.5, resampling='multinomial', normal_dtype=np.float64):
        self.ssm, self.n, self.seed = ssm, int(n), int(seed)
        self.thr, self.resampling = float(ess_threshold), resampling
        self.gid = np.arange(self.n, dtype=np.uint64)
        self.normal_dtype = normal_dtype
        # Lorenz-96 runs through the lane-split kernel (csrc/pf_l96.cu): pairwise Philox streams
        self.normals = philox.normals_pairwise if getattr(ssm, 'pairwise_normals', False) else philox.normals

    def init(self, y0):                                                # filtering.py:173-193
        z = self.normals(self.seed, self.gid, 0, philox.P_INIT, self.ssm.dim, dtype=self.normal_dtype)
        x = self.ssm.initial_sample(z)
        lw = -self.ssm.likelihood_potential(x, y0)                     # :49-54
        return dict(x=x, lw=lw, ess=core.ess_log_weight(lw), t=0,
                    log_z=core.logsumexp(lw) - np.log(self.n), resampled=False, ancestors=None)

    def step(self, st, y):                                             # filtering.py:280-311
        n = self.n
        t = st['t'] + 1
        x, lw = st['x'], st['lw']
        resample = st['ess'] < self.thr * n                            # :287 (strict)
        anc = None
        if resample:                                                   # :196-199
            if self.resampling == 'systematic':                        # exact-rational convention (resample_fused.cu)
                k0 = int(philox.uniform32(self.seed, np.zeros(1, np.uint64), t, philox.P_RESAMPLE)[0])
                anc = core.ancestors_systematic_exact(core.integer_weights_log(lw), k0)
            else:
                anc = core.ancestors_multinomial_stratified(core.cdf_from_log_weights(lw), self.seed, t)[0]
            x = x[anc]
            lw = np.zeros(n)                                           # :292
        z = self.normals(self.seed, self.gid, t, philox.P_MOVE, self.ssm.dim, dtype=self.normal_dtype)
        x_new = self.ssm.transition_sample(x, z)                       # :154-161
        lw_new = lw - self.ssm.likelihood_potential(x_new, y)          # :163-170, :303
        return dict(x=x_new, lw=lw_new, ess=core.ess_log_weight(lw_new), t=t,
                    log_z=st['log_z'] + core.logsumexp(lw_new) - core.logsumexp(lw),
                    resampled=bool(resample), ancestors=anc)

    def run(self, y):
        y = np.asarray(y, np.float64)
        if y.ndim == 1:
            y = y[:, None]                                             # :263-264
        st = self.init(y[0])
        out = [st]
        for t in range(1, len(y)):
            st = self.step(st, y[t])
            out.append(st)
        return out


class OptimalPF(BootstrapPF):
    """OptimalNonLinearGaussianParticleFilter (ssm/nonlinear_gaussian.py:134-276) for the device family H = I with
    diagonal Q = q^2 I, R = r^2 I, P0 = p0^2 I (the Lorenz-96 scenario of config C3), where every matrix of `startup`
    (:152-186) is a scalar:
      initial_kalman_gain K0 = p0^2 / (p0^2 + r^2)                      (:161, utils kalman_gain)
      initial conditioned covariance 1 / (1/p0^2 + 1/r^2)               (:163-166)
      proposal_kalman_gain Kp = q^2 / (q^2 + r^2)                       (:169-172)
      proposal covariance q^2 - Kp q^2 = q^2 r^2 / (q^2 + r^2)          (:174-177)
      weight precision 1 / (q^2 + r^2)                                  (:170-171, 181-182)
    Initial weights are zero (:209-214); the step samples x' = mx + Kp (y - mx) + sd_p z and adds
    log N(y; mx, (q^2 + r^2) I) to the log-weight (:257-271), mx = transition_function(x)."""

    def init(self, y0):
        s = self.ssm
        y0 = np.asarray(y0, np.float64)
        z = self.normals(self.seed, self.gid, 0, philox.P_INIT, s.dim, dtype=self.normal_dtype)
        k0 = s.init_std ** 2 / (s.init_std ** 2 + s.r_std ** 2)
        sd0 = np.sqrt(1.0 / (1.0 / s.init_std ** 2 + 1.0 / s.r_std ** 2))
        x = s.init_mean + k0 * (y0 - s.init_mean) + sd0 * np.asarray(z, np.float64)     # :191-203
        lw = np.zeros(self.n)
        return dict(x=x, lw=lw, ess=core.ess_log_weight(lw), t=0,
                    log_z=core.logsumexp(lw) - np.log(self.n), resampled=False, ancestors=None)

    def step(self, st, y):
        s, n = self.ssm, self.n
        y = np.asarray(y, np.float64)
        t = st['t'] + 1
        x, lw = st['x'], st['lw']
        resample = st['ess'] < self.thr * n
        anc = None
        if resample:
            if self.resampling == 'systematic':
                k0 = int(philox.uniform32(self.seed, np.zeros(1, np.uint64), t, philox.P_RESAMPLE)[0])
                anc = core.ancestors_systematic_exact(core.integer_weights_log(lw), k0)
            else:
                anc = core.ancestors_multinomial_stratified(core.cdf_from_log_weights(lw), self.seed, t)[0]
            x = x[anc]
            lw = np.zeros(n)
        z = self.normals(self.seed, self.gid, t, philox.P_MOVE, s.dim, dtype=self.normal_dtype)
        mx = s.transition_function(x)
        v = s.q_std ** 2 + s.r_std ** 2
        kp = s.q_std ** 2 / v
        sdp = s.q_std * s.r_std / np.sqrt(v)
        x_new = mx + kp * (y - mx) + sdp * np.asarray(z, np.float64)                   # :262-266
        lw_new = lw - (0.5 * np.sum((y - mx) ** 2, axis=-1) / v + 0.5 * s.dim * np.log(2.0 * np.pi * v))   # :268-271
        return dict(x=x_new, lw=lw_new, ess=core.ess_log_weight(lw_new), t=t,
                    log_z=st['log_z'] + core.logsumexp(lw_new) - core.logsumexp(lw),
                    resampled=bool(resample), ancestors=anc)


class EnKF(OptimalPF):
    """EnsembleKalmanFilter (ssm/nonlinear_gaussian.py:279-350) for H = I, R = r^2 I: initial ensemble as the optimal
    filter (:313-323); step: mx = f(x) + q z1 (:335-337), P = cov(mx) with 1/(n-1) (:339), K = P (P + r^2 I)^-1
    (:341-343, utils.py:477-484), y_prop = mx + r z2 (:345-346), x' = mx + (y - y_prop) K^T, zero log-weights (:348-350).
    z1: the step kernel's pairwise stream (purpose P_MOVE); z2: the particle's own stream (purpose P_SIM)."""

    def step(self, st, y):
        s, n = self.ssm, self.n
        y = np.asarray(y, np.float64)
        t = st['t'] + 1
        z1 = self.normals(self.seed, self.gid, t, philox.P_MOVE, s.dim, dtype=self.normal_dtype)
        mx = s.transition_sample(st['x'], z1)
        P = np.atleast_2d(np.cov(mx.T, ddof=1))
        K = P @ np.linalg.inv(P + s.r_std ** 2 * np.eye(s.dim))
        z2 = philox.normals(self.seed, self.gid, t, philox.P_SIM, s.dim, dtype=self.normal_dtype)
        x_new = mx + (y - mx - s.r_std * np.asarray(z2, np.float64)) @ K.T
        lw = np.zeros(n)
        return dict(x=x_new, lw=lw, ess=float(n), t=t, log_z=0.0, resampled=False, ancestors=None,
                    forecast=mx, gain=K, cov=P, mean=mx.mean(0))


def weighted_moments(x, lw):
    w = np.exp(lw - np.max(lw))
    w = w / w.sum()
    mean = w @ x
    var = w @ (x - mean) ** 2
    return mean, var


def kalman_filter(lg, y, reproduce_cov0_bug=False):
    """kalman.py:16-57 plus the innovation log-likelihood the reference lacks.

    reproduce_cov0_bug=True uses `get_initial_covariance_sqrt` as cov_0 exactly as kalman.py:20
    does (correct only when P0 = I, as in config C1); False uses L0 L0^T.
    Returns (means (T,d), covs (T,d,d), loglik)."""
    y = np.asarray(y, np.float64)
    if y.ndim == 1:
        y = y[:, None]
    T = len(y)
    F, H = lg.F, lg.H
    Q = lg.LQ @ lg.LQ.T
    R = lg.LR @ lg.LR.T
    mu = lg.m0.copy()
    cov = lg.L0.copy() if reproduce_cov0_bug else lg.L0 @ lg.L0.T
    means, covs = np.empty((T, lg.dim)), np.empty((T, lg.dim, lg.dim))
    ll = 0.0
    for t in range(T):
        if t > 0:                                                      # predict :34-35
            mu = F @ mu
            cov = F @ cov @ F.T + Q
        S = H @ cov @ H.T + R
        innov = y[t] - H @ mu
        ll += -0.5 * (innov @ np.linalg.solve(S, innov) + np.linalg.slogdet(S)[1]
                      + lg.dim_obs * np.log(2 * np.pi))
        K = cov @ H.T @ np.linalg.inv(S)                               # :42 / :26
        mu = mu + K @ innov
        cov = cov - K @ H @ cov
        means[t], covs[t] = mu, cov
    return means, covs, float(ll)
