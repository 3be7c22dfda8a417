"""SMC-ABC with random-walk ABC moves -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/abc/:
  smc.py:44-79     ABCSMCSampler.startup (prior potential, simulate, distance, threshold=inf)
  smc.py:128-150   MetropolisedABCSMCSampler.startup (... then one adapt())
  smc.py:152-155   resample_criterion: ess[0] < thr * n  (strict)
  smc.py:157-161   termination: alpha_mean <= termination_alpha or iter >= max_iter
  smc.py:163-166   next_threshold_adaptive = quantile(distance, retain * ess[0] / n)
  smc.py:168-173   log_weight = where(distance > threshold, -inf, 0)   (not cumulative)
  smc.py:210-219   forward_proposal: only particles with log_weight > -inf are moved
  smc.py:228-245   adapt ; smc.py:94-98 adapt_stepsize_scaled_diag_cov
  mcmc.py:56-76    RandomWalkABC ; abc.py:36-38 distance_function
Randomness: oracle/philox.py; simulator uniforms are purpose P_SIM (m per particle, u24).
"""
import numpy as np
from . import core, philox, mcmc


class SMCABC:
    def __init__(self, scenario, n, seed, mcmc_steps=1, max_iter=10000, ess_threshold_retain=0.9,
                 ess_threshold_resample=0.5, termination_alpha=0.01, threshold_schedule=None,
                 resampling='multinomial', normal_dtype=np.float64):
        self.sc, self.n, self.seed = scenario, int(n), int(seed)
        self.d, self.m = scenario.dim, scenario.m
        self.mcmc_steps, self.max_iter = int(mcmc_steps), int(max_iter)
        self.retain, self.resample_thr = float(ess_threshold_retain), float(ess_threshold_resample)
        self.termination_alpha = float(termination_alpha)
        self.schedule = None if threshold_schedule is None else np.asarray(threshold_schedule, np.float64)
        if self.schedule is not None:
            self.max_iter = len(self.schedule)                         # smc.py:39-40
        self.resampling = resampling
        self.normal_dtype = normal_dtype
        self.gid = np.arange(self.n, dtype=np.uint64)

    def _sim_dist(self, x, step, s, gids):
        ms = (self.m + 3) // 4
        u = philox.uniforms24(self.seed, gids, step, philox.P_SIM, self.m, index0=s * ms)
        return self.sc.distance(self.sc.simulate(x, u))

    def _adapt(self, prev_lw, st, it):                                 # smc.py:228-245
        n = self.n
        if self.schedule is None:
            thr = core.quantile_linear(st['dist'], self.retain * st['ess'] / n)   # :163-166
        else:
            thr = float(self.schedule[it])
        st['threshold'] = thr
        st['lw'] = np.where(st['dist'] > thr, -np.inf, 0.0)            # :168-173
        st['ess'] = core.ess_log_weight(st['lw'])
        alive_prev = prev_lw > -np.inf
        st['alpha_mean'] = float((st['alpha'] * alive_prev).sum() / alive_prev.sum())   # :240-241
        _, var = core.colstats(st['x'])
        st['stepsize'] = var / self.d * 2.38 ** 2                      # :94-98
        return st

    def startup(self, x0=None):
        n = self.n
        if x0 is None:
            x = philox.normals(self.seed, self.gid, 0, philox.P_INIT, self.d, dtype=self.normal_dtype)
        else:
            x = np.asarray(x0, np.float64).copy()
        st = dict(x=x, up=self.sc.prior_potential(x), dist=self._sim_dist(x, 0, 0, self.gid),
                  lw=np.zeros(n), ess=float(n), alpha=np.ones(n), iter=0, threshold=np.inf, resampled=False)
        return self._adapt(np.zeros(n), st, 0)

    def terminated(self, st):                                          # smc.py:157-161
        return st['alpha_mean'] <= self.termination_alpha or st['iter'] >= self.max_iter

    def update(self, st):
        n, d = self.n, self.d
        it = st['iter'] + 1
        x, up, dist, lw, ess = st['x'], st['up'], st['dist'], st['lw'], st['ess']
        resample = ess < self.resample_thr * n                         # :152-155 strict
        anc = None
        if resample:
            if self.resampling == 'systematic':                        # exact-rational convention (resample_fused.cu)
                k0 = int(philox.uniform32(self.seed, np.zeros(1, np.uint64), it, philox.P_RESAMPLE)[0])
                anc = core.ancestors_systematic_exact(core.integer_weights_log(lw), k0)
            else:
                anc = core.ancestors_multinomial_stratified(core.cdf_from_log_weights(lw), self.seed, it)[0]
            x, up, dist = x[anc], up[anc], dist[anc]
            lw, ess = np.zeros(n), float(n)
        alive = lw > -np.inf                                           # :210-219
        nz = (d + 3) // 4
        S = nz + 1
        alphas = np.zeros(n)
        x, up, dist = x.copy(), up.copy(), dist.copy()
        for s in range(self.mcmc_steps):
            z = philox.normals(self.seed, self.gid, it, philox.P_MOVE, d, index0=s * S, dtype=self.normal_dtype)
            u = philox.u24(philox.raw(self.seed, self.gid, it, philox.P_MOVE, s * S + nz)[0]).astype(np.float64)
            simd = lambda xx: self._sim_dist(xx, it, s, self.gid)
            xn, upn, dn, alpha, _ = mcmc.rw_abc_step(self.sc.prior_potential, simd, x, up, dist, z, u,
                                                     st['stepsize'], st['threshold'])
            a2 = alive[:, None]
            x, up, dist = np.where(a2, xn, x), np.where(alive, upn, up), np.where(alive, dn, dist)
            alphas += np.where(alive, alpha, 0.0)
        # particles not moved keep their previous alpha in the reference (state passes through
        # unchanged, smc.py:216-218); it is masked out by alive_prev in alpha_mean anyway.
        alpha = np.where(alive, alphas / self.mcmc_steps, st['alpha'][anc] if anc is not None else st['alpha'])
        new = dict(x=x, up=up, dist=dist, lw=lw, ess=ess, alpha=alpha, iter=it,
                   threshold=st['threshold'], resampled=bool(resample), ancestors=anc)
        return self._adapt(lw, new, it)

    def run(self, x0=None):
        st = self.startup(x0)
        chain = [st]
        while not self.terminated(st):
            st = self.update(st)
            chain.append(st)
        return chain
