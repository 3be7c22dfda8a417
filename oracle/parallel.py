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
