"""MCMC kernels used inside the SMC samplers: parameter holders mirroring mocat/src/mcmc/standard_mcmc.py.
The moves themselves run inside `mb_smc_move` (csrc/propagate.cu).  Standalone serial MCMC chains
(`run` with an MCMCSampler) are not a data-parallel path and are out of scope."""
import numpy as np

from . import _lib
from .core import cdict


class MCMCSampler:
    name = "MCMC Sampler"
    move_kind = None

    def __init__(self, **kwargs):
        self.parameters = cdict(**kwargs)
        self.tuning = cdict()


class Metropolis:
    """mcmc/metropolis.py:34-70 -- the only correction compiled into the device moves."""


class RandomWalk(MCMCSampler):
    """mcmc/standard_mcmc.py:21-65: x' = x + sqrt(stepsize) z, alpha = min(1, exp(-U' + U))."""
    name = 'Random Walk'
    correction = Metropolis
    move_kind = _lib.MOVE_RW

    def __init__(self, stepsize=None):
        super().__init__(stepsize=stepsize)
        self.tuning.target = 0.234


class Underdamped(MCMCSampler):
    """mcmc/standard_mcmc.py:72-153 with friction = inf: HMC with `leapfrog_steps` steps (MALA for 1)."""
    name = 'Underdamped'
    correction = Metropolis
    move_kind = _lib.MOVE_MALA

    def __init__(self, stepsize=None, leapfrog_steps=1, friction=np.inf):
        if np.isfinite(friction):
            raise _lib.MocatB200Error("Underdamped with finite friction keeps momenta between iterations; only "
                                      "friction=inf (MALA / HMC) is compiled into the device move")
        super().__init__(stepsize=stepsize, leapfrog_steps=int(leapfrog_steps), friction=friction)
        self.tuning.target = 0.651


Overdamped = Underdamped
