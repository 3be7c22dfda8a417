"""SMC-ABC: host mirror of mocat/src/abc/{abc,smc,mcmc}.py and abc/scenarios/gk.py for the population path.

`MetropolisedABCSMCSampler(mcmc_sampler=None, mcmc_correction, mcmc_steps=1, threshold_schedule=None,
max_iter, ess_threshold_retain=.9, ess_threshold_resample=.5, termination_alpha=.01)` (abc/smc.py:103-126)
with the default RandomWalkABC move and diagonal-covariance step adaptation (abc/smc.py:94-98,116-118).
"""
import numpy as np

from . import _lib, engine, models
from .core import cdict, key_to_seed
from .sample import Sampler
from .transport import _RESAMPLING, HISTORY_AUTO_BYTES


class ABCScenario:
    """abc/abc.py:16-38 (interface).  Device scenarios implement `_gk()`."""
    name = None
    dim = None
    data = None

    def _device(self):
        raise _lib.MocatB200Error(f"{type(self).__name__}: only built-in device ABC scenarios (GKTransformedUniformPrior "
                                  "with m in {4, 8, 16} sorted draws) can be simulated; no CPU fallback")


class GKTransformedUniformPrior(ABCScenario):
    """abc/scenarios/gk.py:68-96.  Summary statistic (user code upstream, gk.py:34-36): the
    `n_unsummarised_data` simulated draws sorted ascending."""
    name = 'GK_fewN'
    dim = 4
    c = 0.8
    prior_mins = 0.
    prior_maxs = 10.
    buffer = 1e-5

    def __init__(self, data=None, n_unsummarised_data=8, name=None, **kwargs):
        if name is not None:
            self.name = name
        self.n_unsummarised_data = int(n_unsummarised_data if data is None else len(data))
        self.data = None if data is None else np.sort(np.asarray(data, np.float64))
        for k, v in kwargs.items():
            if hasattr(self, k):
                setattr(self, k, v)

    def constrain(self, unconstrained_x):                               # gk.py:70-72
        from scipy.special import ndtr
        return self.prior_mins + ndtr(np.asarray(unconstrained_x, np.float64)) * (self.prior_maxs - self.prior_mins)

    def unconstrain(self, constrained_x):                               # gk.py:74-76
        from scipy.special import ndtri
        return ndtri((np.asarray(constrained_x, np.float64) - self.prior_mins) / (self.prior_maxs - self.prior_mins))

    def _device(self):
        if self.data is None:
            raise _lib.MocatB200Error("GKTransformedUniformPrior.data (the observed summary) must be set")
        return models.make_gk(self.data, self.c, self.prior_mins, self.prior_maxs, self.buffer)


class RandomWalkABC:
    """abc/mcmc.py:40-76 (parameter holder; the move runs in mb_abc_move)."""
    name = 'Random Walk ABC'

    def __init__(self, threshold=None, stepsize=None):
        self.parameters = cdict(threshold=threshold, stepsize=stepsize)
        self.tuning = cdict(parameter='threshold', target=0.1, metric='alpha', monotonicity='increasing')


class ABCSMCSampler(Sampler):
    """abc/smc.py:22-91."""
    name = "ABC SMC"

    def __init__(self, threshold_schedule=None, max_iter=int(1e4), **kwargs):
        self.max_iter = max_iter
        self.threshold_schedule = threshold_schedule
        super().__init__(**kwargs)

    def __setattr__(self, key, value):                                   # abc/smc.py:33-42
        if key == 'threshold_schedule':
            if value is None:
                if getattr(self, 'threshold_schedule', None) is not None \
                        and self.max_iter == len(self.threshold_schedule):
                    self.max_iter = int(1e4)
            else:
                self.max_iter = len(value)
        super().__setattr__(key, value)


class MetropolisedABCSMCSampler(ABCSMCSampler):
    """abc/smc.py:101-245."""

    _FIELDS = ('value', 'log_weight', 'prior_potential', 'distance', 'alpha')

    def __init__(self, mcmc_sampler=None, mcmc_correction='sampler_default', mcmc_steps=1, threshold_schedule=None,
                 max_iter=int(1e4), ess_threshold_retain=0.9, ess_threshold_resample=0.5, termination_alpha=0.01,
                 resampling='multinomial', keep_history=None, check_every=8, **kwargs):
        super().__init__(max_iter=max_iter, threshold_schedule=threshold_schedule, **kwargs)
        if mcmc_sampler is None:
            mcmc_sampler = RandomWalkABC()
        if isinstance(mcmc_sampler, type):
            mcmc_sampler = mcmc_sampler()
        if not isinstance(mcmc_sampler, RandomWalkABC):
            raise _lib.MocatB200Error("only RandomWalkABC (the reference default) is compiled into the device move")
        self.mcmc_sampler = mcmc_sampler
        self.parameters.mcmc_steps = mcmc_steps
        self.parameters.ess_threshold_retain = ess_threshold_retain
        self.parameters.ess_threshold_resample = ess_threshold_resample
        self.parameters.termination_alpha = termination_alpha
        self.resampling, self.keep_history, self.check_every = resampling, keep_history, check_every

    def _snapshot(self, eng):
        return dict(value=eng.values().clone(memory_format=__import__('torch').contiguous_format), log_weight=eng.lw.clone(), prior_potential=eng.up.clone(),
                    distance=eng.dist.clone(), alpha=eng.alpha.clone())

    def startup(self, abc_scenario, n, initial_state, initial_extra, **kwargs):
        initial_state, initial_extra = super().startup(abc_scenario, n, initial_state, initial_extra, **kwargs)
        P = self.parameters
        kw = dict(mcmc_steps=P.mcmc_steps, max_iter=self.max_iter, ess_retain=P.ess_threshold_retain,
                  ess_resample=P.ess_threshold_resample, termination_alpha=P.termination_alpha,
                  threshold_schedule=self.threshold_schedule, resampling=_RESAMPLING[self.resampling])
        seed = key_to_seed(getattr(initial_extra, 'random_key', None))
        world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size()
        except Exception:
            world = 1
        if world > 1:
            # under torchrun `n` is the GLOBAL population size: every rank holds n / world particles and the returned
            # chain carries this rank's shard of the per-particle fields (thresholds, ESS, ... are global)
            from . import parallel
            _, n_local = parallel.shard_range(n, dist.get_rank(), world)
            eng = parallel.ShardedABCEngine(parallel.shard_context(), abc_scenario._device(), n_local, seed, **kw)
            if initial_state is not None and getattr(initial_state, 'value', None) is not None:
                r0 = dist.get_rank() * n_local                     # a global initial population: keep this rank's rows
                initial_state = initial_state.copy()
                initial_state.value = np.asarray(initial_state.value)[r0:r0 + n_local]
        else:
            eng = engine.ABCEngine(abc_scenario._device(), n, seed, **kw)
        x0 = None if initial_state is None else getattr(initial_state, 'value', None)
        eng.startup(x0)                                                  # abc/smc.py:44-79,128-150
        if initial_state is None:
            initial_state = cdict()
        initial_state.engine = eng
        initial_extra.engine = eng
        return initial_state, initial_extra

    def update(self, abc_scenario, ensemble_state, extra):
        extra.engine.update()
        extra.iter = extra.iter + 1
        return ensemble_state, extra

    def termination_criterion(self, ensemble_state, extra):              # abc/smc.py:157-161 (on device)
        return bool(extra.engine.ctl.read()['done'])

    def _run_device(self, abc_scenario, initial_state, initial_extra):
        import torch
        eng = initial_extra.engine
        keep = self.keep_history
        auto = keep is None
        if auto:
            keep = eng.n * 12 * 4 * 2 <= HISTORY_AUTO_BYTES
        if self.max_iter > engine.MB_HIST_MAX - 1:
            raise _lib.MocatB200Error(f"max_iter <= {engine.MB_HIST_MAX - 1} (device history ring)")
        # snapshots go to the host at the end of every burst; an automatic history that outgrows the budget is dropped
        host_snaps, pending, kept_bytes = [], ([self._snapshot(eng)] if keep else []), 0
        it = 0
        while it < self.max_iter:
            for _ in range(min(self.check_every, self.max_iter - it)):
                eng.update()
                it += 1
                if keep:
                    pending.append(self._snapshot(eng))
            done = bool(eng.ctl.read()['done'])
            if keep:
                for sn in pending:
                    host_snaps.append({k: v.cpu().numpy() for k, v in sn.items()})
                    kept_bytes += sum(v.nbytes for v in host_snaps[-1].values())
                pending = []
                if auto and kept_bytes > HISTORY_AUTO_BYTES:
                    keep, host_snaps = False, []
            if done:
                break
        c = eng.settle()                                                 # valid ping-pong buffers of the last real step
        iters = int(c['iter'])
        hist = eng.ctl.read_hist(iters + 1)
        chain = cdict()
        if keep:
            for k in self._FIELDS:
                setattr(chain, k, np.stack([sn[k] for sn in host_snaps[:iters + 1]]))
        else:
            for k, v in self._snapshot(eng).items():
                setattr(chain, k, v.cpu().numpy()[None])
        chain.threshold = hist['beta'].copy()                            # clean_chain, abc/smc.py:86-91
        chain.ess = hist['ess'].copy()
        chain.alpha_mean = hist['alpha_mean'].copy()
        chain.resampled = hist['resampled'].copy()
        chain.stepsize = eng.stepsize.cpu().numpy()
        initial_extra.iter = iters
        return chain
