"""Tempered SMC sampler with Metropolised moves -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/transport/smc.py:
  :26-39    SMCSampler.startup            :61-71   resample          :73-99   update
  :128-164  TemperedSMCSampler.startup    :171-175 termination       :193-217 adapt
  :267-296  MetropolisedSMCSampler.startup :298-301 resample_criterion (ess <= thr*n)
  :311-326  next_temperature_adaptive     :337-365 forward_proposal   :367-373 log_weight
and transport/sampler.py:24-30 (particles initialised from vmap(prior_sample)).

Randomness: Philox streams of oracle/philox.py (the reference's threefry streams are
unpinned, see oracle/__init__.py).  Per iteration `it` and particle `gid`:
  move normals   purpose P_MOVE, step it, slots s*S .. s*S+nz-1   (nz = ceil(d/4))
  accept uniform purpose P_MOVE, step it: d mod 4 in {1, 2} leaves two unused words in the last normal slot -> word 2 of
                 slot s*S+nz-1 (u24), S = nz; otherwise word 0 of slot s*S+nz, S = nz+1  (csrc/rng.cuh MB_MOVE_SLOTS)
  resampling     purpose P_RESAMPLE, step it: systematic u0 = uniform32(gid 0) / 2^32;
                 multinomial u_i = uniform53(gid i)
Deviation noted in DESIGN.md: the reference recomputes lik = (U - U_prior)/beta after the
move (smc.py:362-364); here (and on the device) the likelihood potential of the accepted
state is carried directly -- identical in exact arithmetic, more accurate in fp32.
"""
import numpy as np
from . import core, philox, mcmc


class TemperedSMC:
    def __init__(self, prior, likelihood, n, seed, move='mala', stepsize=0.1, leapfrog_steps=1,
                 mcmc_steps=1, temperature_schedule=None, max_temperature=1.0, max_iter=10000,
                 ess_threshold_retain=0.9, ess_threshold_resample=0.5, bisection_tol=1e-5,
                 max_bisection_iter=1000, resampling='multinomial', normal_dtype=np.float64,
                 rm_stepsize=None, rm_target=0.651):
        self.prior, self.lik, self.n, self.seed = prior, likelihood, int(n), int(seed)
        self.d = prior.dim
        self.move, self.stepsize, self.L, self.mcmc_steps = move, float(stepsize), int(leapfrog_steps), int(mcmc_steps)
        sched = None if temperature_schedule is None else np.asarray(temperature_schedule, np.float64)
        if sched is not None and sched[0] == 0.0:                      # smc.py:243-245
            sched = sched[1:]
        self.schedule = sched
        if sched is not None:                                          # smc.py:123-125
            max_temperature, max_iter = float(sched[-1]), len(sched)
        self.max_temperature, self.max_iter = float(max_temperature), int(max_iter)
        self.retain, self.resample_thr = float(ess_threshold_retain), float(ess_threshold_resample)
        self.tol, self.max_bis = float(bisection_tol), int(max_bisection_iter)
        self.resampling = resampling
        self.normal_dtype = normal_dtype
        self.gid = np.arange(self.n, dtype=np.uint64)
        self.bisect_iters = []
        self.rm_stepsize, self.rm_target = rm_stepsize, float(rm_target)   # smc.py:376-421 (RMMetropolisedSMCSampler)

    # -- potentials ---------------------------------------------------------------------------
    def _eval(self, x, beta):
        up, gp = self.prior.potential_and_grad(x)
        ul, gl = self.lik.potential_and_grad(x)
        return up, ul, up + beta * ul, gp + beta * gl

    def _next_temperature(self, lw, lik, beta, ess0, it):
        if self.schedule is not None:                                  # smc.py:123
            return float(self.schedule[it]), 0
        return core.next_temperature_adaptive(lw, lik, beta, self.max_temperature, ess0,
                                              self.retain, self.tol, self.max_bis)

    # -- startup (smc.py:128-164, 267-296) ------------------------------------------------------
    def startup(self, x0=None):
        n, d = self.n, self.d
        if x0 is None:
            z = philox.normals(self.seed, self.gid, 0, philox.P_INIT, d, dtype=self.normal_dtype)
            x = self.prior.sample(z)
        else:
            x = np.asarray(x0, np.float64).copy()
        up, _ = self.prior.potential_and_grad(x)
        lik, _ = self.lik.potential_and_grad(x)
        beta, its = self._next_temperature(np.zeros(n), lik, 0.0, float(n), 0)
        self.bisect_iters.append(its)
        lw = -beta * lik
        st = dict(x=x, up=up, lik=lik, lw=lw, beta=beta, iter=0, alpha=np.ones(n),
                  ess=core.ess_log_weight(lw),
                  log_norm_constant=core.logsumexp(lw, b=1.0 / n), resampled=False)
        return st

    def terminated(self, st):                                          # smc.py:171-175
        return (st['beta'] >= self.max_temperature or st['iter'] >= self.max_iter
                or np.isnan(st['x']).mean() > 0.1)

    # -- vmap(forward_proposal) (smc.py:91-95, 337-365) ---------------------------------------------
    def _move(self, x, gid, it, beta):
        """MCMC startup re-evaluates the potentials at the current temperature (standard_mcmc.py:94-102),
        then mcmc_steps Metropolised moves; returns (x, U_prior, U_lik, mean alpha)."""
        d = self.d
        n = x.shape[0]
        nz = (d + 3) // 4
        spare = d % 4 in (1, 2)                                         # unused words in the last normal slot
        S = nz if spare else nz + 1
        alphas = np.zeros(n)
        up_c, lik_c, U, g = self._eval(x, beta)
        for s in range(self.mcmc_steps):
            z = philox.normals(self.seed, gid, it, philox.P_MOVE, d, index0=s * S, dtype=self.normal_dtype)
            if spare:
                u = philox.u24(philox.raw(self.seed, gid, it, philox.P_MOVE, s * S + nz - 1)[2]).astype(np.float64)
            else:
                u = philox.u24(philox.raw(self.seed, gid, it, philox.P_MOVE, s * S + nz)[0]).astype(np.float64)
            if self.move == 'mala':
                pg = lambda xx: self._eval(xx, beta)[2:]
                x, U, g, alpha, _ = mcmc.hmc_step(pg, x, U, g, z, u, self.stepsize, self.L)
            else:
                pot = lambda xx: self._eval(xx, beta)[2]
                x, U, alpha, _ = mcmc.rw_step(pot, x, U, z, u, self.stepsize)
            alphas += alpha
        up, _ = self.prior.potential_and_grad(x)                       # :362
        lik, _ = self.lik.potential_and_grad(x)                        # carried directly (see header)
        return x, up, lik, alphas

    # -- one population step (smc.py:73-99 + 193-217) ---------------------------------------------
    def update(self, st):
        n, d = self.n, self.d
        it = st['iter'] + 1
        beta = st['beta']
        x, up, lik, lw = st['x'], st['up'], st['lik'], st['lw']
        ess = st['ess']
        resample = ess <= self.resample_thr * n                        # :298-301
        anc = None
        if resample:                                                   # :61-71
            if self.resampling == 'systematic':                        # exact-rational convention (resample_fused.cu)
                k0 = int(philox.uniform32(self.seed, np.zeros(1, np.uint64), it, philox.P_RESAMPLE)[0])
                anc = core.ancestors_systematic_exact(core.integer_weights_log(lw), k0)
            else:
                anc = core.ancestors_multinomial_stratified(core.cdf_from_log_weights(lw), self.seed, it)[0]
            x, up, lik = x[anc], up[anc], lik[anc]
            lw = np.zeros(n)
            ess = float(n)
        x, up, lik, alphas = self._move(x, self.gid, it, beta)
        # adapt (:193-217)
        beta_new, its = self._next_temperature(lw, lik, beta, ess, it)
        self.bisect_iters.append(its)
        lw_new = lw - (beta_new - beta) * lik                          # :367-373
        new = dict(x=x, up=up, lik=lik, lw=lw_new, beta=beta_new, iter=it,
                   alpha=alphas / self.mcmc_steps, ess=core.ess_log_weight(lw_new),
                   log_norm_constant=st['log_norm_constant'] + core.logsumexp(lw_new) - core.logsumexp(lw),
                   resampled=bool(resample), ancestors=anc)
        if self.rm_stepsize is not None:                               # smc.py:406-421: Robbins-Monro on log stepsize
            w = np.exp(lw_new - np.max(lw_new))
            alpha_mean = float(np.sum(w * new['alpha']) / np.sum(w))
            self.stepsize = float(np.exp(np.log(self.stepsize) + self.rm_stepsize * (alpha_mean - self.rm_target)))
        new['stepsize'] = self.stepsize
        return new

    def run(self, x0=None):
        st = self.startup(x0)
        chain = [st]
        while not self.terminated(st):
            st = self.update(st)
            chain.append(st)
        return chain
