"""Multi-core driver for the oracle's tempered SMC step.  TEST / BASELINE INFRASTRUCTURE.

The reference evaluates the per-particle move under one big jax.vmap (transport/smc.py:91-95), which XLA-CPU
spreads over the host cores.  This module does the same for the NumPy restatement: the move of oracle.smc
is mapped over contiguous particle shards in a fork-ed process pool; reductions, the temperature search and
the inverse-CDF resampling stay in the parent.  Used by bench.py (`cpu_baseline`, `--impl reference`) so the
CPU number uses all host cores.  Results are identical to the serial oracle (same Philox streams per gid)."""
import multiprocessing as mp
import os

import numpy as np

from . import smc as osmc

_WORKER = {}


def _init(obj):
    _WORKER['obj'] = obj


def _move_shard(args):
    lo, hi, x, it, beta = args
    o = _WORKER['obj']
    return osmc.TemperedSMC._move(o, x, o.gid[lo:hi], it, beta)


class ParallelTemperedSMC(osmc.TemperedSMC):
    def __init__(self, *a, workers=None, **k):
        super().__init__(*a, **k)
        self.workers = workers or os.cpu_count() or 1
        self.pool = None
        if self.workers > 1:
            self.pool = mp.get_context('fork').Pool(self.workers, initializer=_init, initargs=(self,))

    def _move(self, x, gid, it, beta):
        if self.pool is None:
            return super()._move(x, gid, it, beta)
        n = x.shape[0]
        edges = np.linspace(0, n, self.workers * 4 + 1).astype(int)
        jobs = [(int(lo), int(hi), x[lo:hi], it, beta) for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]
        parts = self.pool.map(_move_shard, jobs)
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(4))

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None


# ----------------------------------------------------------------------------------------------------------------
# Bootstrap particle filter on Lorenz-96 over all host cores (config C3's CPU baseline).  The reference runs the
# filter body as one vmap over the particles (ssm/filtering.py:124-139, 280-311): propagate + weight are mapped over
# contiguous particle shards in fork-ed workers that share the two population buffers; the inverse-CDF resampling
# (n^2 Gumbel-max `random.categorical` in the reference -- infeasible beyond n ~ 2e4) stays in the parent.
# Arithmetic is fp32 like the reference (JAX default); normals come from NumPy's generator (one stream per worker),
# i.e. the CPU side is NOT charged for the oracle's slow NumPy Philox.
_PF = {}


def _pf_init(shared):
    _PF.update(shared)


def _pf_shard(args):
    lo, hi, src, dst, resample, y, seed, t = args
    P = _PF
    n, d = P['n'], P['d']
    xs = np.frombuffer(P['x'][src], dtype=np.float32).reshape(n, d)
    xd = np.frombuffer(P['x'][dst], dtype=np.float32).reshape(n, d)
    lw = np.frombuffer(P['lw'], dtype=np.float32)
    anc = np.frombuffer(P['anc'], dtype=np.int64)
    x = xs[anc[lo:hi]] if resample else xs[lo:hi].copy()
    h = np.float32(P['dt'] / P['substeps'])
    F = np.float32(P['forcing'])

    def rhs(v):
        return (np.roll(v, -1, 1) - np.roll(v, 2, 1)) * np.roll(v, 1, 1) - v + F

    for _ in range(P['substeps']):
        k1 = rhs(x)
        k2 = rhs(x + np.float32(0.5) * h * k1)
        k3 = rhs(x + np.float32(0.5) * h * k2)
        k4 = rhs(x + h * k3)
        x = x + (h / np.float32(6.0)) * (k1 + 2 * k2 + 2 * k3 + k4)
    rng = np.random.default_rng([seed, t, lo])
    x = x + np.float32(P['q_std']) * rng.standard_normal(x.shape, dtype=np.float32)
    r = (y[None, :] - x) * np.float32(1.0 / P['r_std'])
    incr = -(np.float32(0.5) * np.einsum('ij,ij->i', r, r) + np.float32(P['lik_const']))
    xd[lo:hi] = x
    lw[lo:hi] = incr if resample else lw[lo:hi] + incr
    return float(incr.max())


class ParallelBootstrapPF:
    def __init__(self, ssm, n, seed, ess_threshold=2.0, workers=None):
        from . import core
        self.core = core
        self.ssm, self.n, self.seed, self.thr = ssm, int(n), int(seed), float(ess_threshold)
        self.workers = workers or os.cpu_count() or 1
        d = ssm.dim
        shared = dict(n=self.n, d=d, dt=ssm.dt, substeps=ssm.substeps, forcing=ssm.forcing, q_std=ssm.q_std,
                      r_std=ssm.r_std, lik_const=0.5 * (d * np.log(2 * np.pi) + 2.0 * d * np.log(ssm.r_std)),
                      x=[mp.RawArray('f', self.n * d), mp.RawArray('f', self.n * d)], lw=mp.RawArray('f', self.n),
                      anc=mp.RawArray('q', self.n))
        self.sh = shared
        self.cur, self.t = 0, 0
        self.pool = mp.get_context('fork').Pool(self.workers, initializer=_pf_init, initargs=(shared,))
        self.lw = np.frombuffer(shared['lw'], dtype=np.float32)
        self.anc = np.frombuffer(shared['anc'], dtype=np.int64)
        self.ess = float(self.n)

    def x(self):
        return np.frombuffer(self.sh['x'][self.cur], dtype=np.float32).reshape(self.n, self.ssm.dim)

    def init(self, y0):
        rng = np.random.default_rng(self.seed)
        self.x()[:] = self.ssm.init_mean + self.ssm.init_std * rng.standard_normal((self.n, self.ssm.dim), dtype=np.float32)
        r = (np.asarray(y0, np.float32)[None, :] - self.x()) / np.float32(self.ssm.r_std)
        self.lw[:] = -(0.5 * np.einsum('ij,ij->i', r, r) + self.sh['lik_const'])
        self.ess = self.core.ess_log_weight(self.lw)

    def step(self, y):
        self.t += 1
        resample = self.ess < self.thr * self.n                        # filtering.py:287
        if resample:                                                   # inverse-CDF systematic resampling in the parent
            w = np.exp(self.lw - self.lw.max(), dtype=np.float64)
            cdf = np.cumsum(w)
            u = (np.arange(self.n) + np.random.default_rng([self.seed, self.t]).random()) * (cdf[-1] / self.n)
            self.anc[:] = np.minimum(np.searchsorted(cdf, u, side='right'), self.n - 1)
        edges = np.linspace(0, self.n, self.workers * 4 + 1).astype(int)
        y32 = np.asarray(y, np.float32)
        jobs = [(int(lo), int(hi), self.cur, self.cur ^ 1, bool(resample), y32, self.seed, self.t)
                for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]
        self.pool.map(_pf_shard, jobs)
        self.cur ^= 1
        self.ess = self.core.ess_log_weight(self.lw)
        return self.ess

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None
