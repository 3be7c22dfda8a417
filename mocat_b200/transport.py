"""Population (transport) samplers: host mirror of mocat/src/transport/{sampler,smc,svgd}.py.

Constructor signatures, parameter names, criteria (`<=` vs `<`), schedules and returned fields follow the
reference; the per-iteration work is enqueued on the device through mocat_b200.engine.  Extra keyword
arguments (ours): `resampling` ('multinomial' as in the reference | 'systematic'), `keep_history`
(None = auto: the full (iters+1, n, d) stack the reference returns is kept when it is small, otherwise only
the final population is returned with a leading axis of length 1), `check_every` (iterations between
termination polls).
"""
import numpy as np

from . import _lib, engine, models
from .core import cdict, key_to_seed
from .mcmc import MCMCSampler
from .sample import Sampler

_RESAMPLING = {'multinomial': _lib.RESAMPLE_MULTINOMIAL, 'systematic': _lib.RESAMPLE_SYSTEMATIC}
HISTORY_AUTO_BYTES = 1 << 30


def _torch():
    import torch
    return torch


class TransportSampler(Sampler):
    """transport/sampler.py:16-51."""

    def startup(self, scenario, n, initial_state, initial_extra, **kwargs):
        initial_state, initial_extra = super().startup(scenario, n, initial_state, initial_extra, **kwargs)
        return initial_state, initial_extra

    def termination_criterion(self, ensemble_state, extra):
        return extra.iter >= self.max_iter


class SMCSampler(TransportSampler):
    """transport/smc.py:23-99."""
    name = 'SMC Sampler'


class TemperedSMCSampler(SMCSampler):
    """transport/smc.py:102-225."""
    name = 'Tempered SMC Sampler'

    def __init__(self, temperature_schedule=None, max_temperature=1., max_iter=int(1e4), **kwargs):
        self.max_iter = max_iter
        self.max_temperature = max_temperature
        self.temperature_schedule = temperature_schedule
        super().__init__(**kwargs)

    def __setattr__(self, key, value):                                 # smc.py:115-126
        if key == 'temperature_schedule':
            if value is None:
                if getattr(self, 'temperature_schedule', None) is not None \
                        and self.max_iter == len(self.temperature_schedule):
                    self.max_iter = int(1e4)
            else:
                value = np.asarray(value, np.float64)
                self.max_temperature = float(value[-1])
                self.max_iter = len(value)
        super().__setattr__(key, value)

    def clean_chain(self, scenario, chain_ensemble_state):             # smc.py:177-184
        scenario.temperature = float(chain_ensemble_state.temperature[-1])
        return chain_ensemble_state


class MetropolisedSMCSampler(TemperedSMCSampler):
    """transport/smc.py:228-373."""
    name = "Metropolised SMC Sampler"

    def __init__(self, mcmc_sampler, mcmc_correction='sampler_default', mcmc_steps=1, max_iter=int(1e4),
                 temperature_schedule=None, max_temperature=1., ess_threshold_retain=0.9,
                 ess_threshold_resample=0.5, bisection_tol=1e-5, max_bisection_iter=1000,
                 resampling='multinomial', keep_history=None, check_every=8, **kwargs):
        if temperature_schedule is not None and temperature_schedule[0] == 0.:      # smc.py:243-245
            temperature_schedule = temperature_schedule[1:]
        super().__init__(max_iter=max_iter, temperature_schedule=temperature_schedule,
                         max_temperature=max_temperature, **kwargs)
        if isinstance(mcmc_sampler, type):
            mcmc_sampler = mcmc_sampler()
        if not isinstance(mcmc_sampler, MCMCSampler) or mcmc_sampler.move_kind is None:
            raise _lib.MocatB200Error("mcmc_sampler must be mocat_b200.RandomWalk or mocat_b200.Underdamped "
                                      "(the moves compiled into the device kernel)")
        self.mcmc_sampler = mcmc_sampler
        self.parameters.mcmc_steps = mcmc_steps
        self.parameters.ess_threshold_retain = ess_threshold_retain
        self.parameters.ess_threshold_resample = ess_threshold_resample
        self.parameters.bisection_tol = bisection_tol
        self.parameters.max_bisection_iter = max_bisection_iter
        self.resampling = resampling
        self.keep_history = keep_history
        self.check_every = check_every

    def __setattr__(self, key, value):                                 # smc.py:261-265
        if key == 'temperature_schedule' and value is not None and value[0] == 0.:
            value = value[1:]
        super().__setattr__(key, value)

    # -- state <-> host ---------------------------------------------------------------------------------
    _FIELDS = ('value', 'log_weight', 'prior_potential', 'likelihood_potential', 'alpha')

    def _snapshot(self, eng):
        """device clones of the per-particle fields (Appendix B of SURVEY.md; gradients and momenta are
        recomputed inside the move and not stored)"""
        return dict(value=eng.values().clone(memory_format=_torch().contiguous_format), log_weight=eng.lw.clone(), prior_potential=eng.up.clone(),
                    likelihood_potential=eng.lik.clone(), alpha=eng.alpha.clone())

    def startup(self, scenario, n, initial_state, initial_extra, **kwargs):
        initial_state, initial_extra = super().startup(scenario, n, initial_state, initial_extra, **kwargs)
        P = self.parameters
        ms = self.mcmc_sampler
        stepsize = getattr(initial_extra.parameters, 'stepsize', None)
        if stepsize is None:
            stepsize = ms.parameters.stepsize
        if stepsize is None:
            raise _lib.MocatB200Error("mcmc_sampler.parameters.stepsize must be set")
        target = scenario._target()
        move = models.make_move(ms.move_kind, stepsize, P.mcmc_steps, getattr(ms.parameters, 'leapfrog_steps', 1))
        temper = models.make_temper(self.max_temperature, P.ess_threshold_retain, P.ess_threshold_resample,
                                    P.bisection_tol, P.max_bisection_iter, self.max_iter)
        seed = key_to_seed(getattr(initial_extra, 'random_key', None))
        world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size()
        except Exception:
            world = 1
        if world > 1 and getattr(self, '_private_engine', False):
            raise _lib.MocatB200Error("RMMetropolisedSMCSampler is single-GPU for now (the acceptance mean is not sharded)")
        if world > 1 and getattr(self, 'sharded', True):
            # under torchrun `n` is the GLOBAL population size; every rank holds n/world particles, passes its
            # own shard of initial_state.value and gets its own shard back (mocat_b200/parallel.py)
            from . import parallel
            _, n_local = parallel.shard_range(n, dist.get_rank(), world)
            eng = parallel.acquire_sharded_smc(target, move, temper, n_local, seed, _RESAMPLING[self.resampling],
                                               schedule=self.temperature_schedule)
        else:
            if getattr(self, '_private_engine', False):                  # engine with sampler-specific device state
                eng = engine.SMCEngine(target, move, temper, n, seed, resampling=_RESAMPLING[self.resampling],
                                       schedule=self.temperature_schedule)
            else:
                eng = engine.SMCEngine.acquire(target, move, temper, n, seed, resampling=_RESAMPLING[self.resampling],
                                               schedule=self.temperature_schedule)
        x0 = None if initial_state is None else getattr(initial_state, 'value', None)
        self._configure_engine(eng)
        eng.startup(x0)                                                 # smc.py:128-164, 267-296 on the device
        scenario.temperature = 0.
        initial_extra.engine = eng
        initial_extra.resample_bool = True
        if initial_state is None:
            initial_state = cdict()
        initial_state.engine = eng
        return initial_state, initial_extra

    def update(self, scenario, ensemble_state, extra):                 # smc.py:219-225 / 73-99
        extra.engine.update()
        extra.iter = extra.iter + 1
        return ensemble_state, extra

    def termination_criterion(self, ensemble_state, extra):            # smc.py:171-175 (evaluated on device)
        return bool(extra.engine.ctl.read()['done'])

    def fetch(self, ensemble_state):
        """materialise the current device population as a cdict of NumPy arrays"""
        eng = ensemble_state.engine
        c = eng.settle()
        snap = {k: v.cpu().numpy() for k, v in self._snapshot(eng).items()}
        out = cdict(**snap)
        out.potential = out.prior_potential + c['beta'] * out.likelihood_potential
        out.temperature, out.ess, out.log_norm_constant = c['beta'], c['ess'], c['log_z']
        return out

    def _configure_engine(self, eng):
        pass

    def _post_update(self, eng, extra):
        """hook between population steps (the reference's `adapt` extensions); nothing for the plain sampler"""

    def _run_device(self, scenario, initial_state, initial_extra):
        torch = _torch()
        eng = initial_extra.engine
        n, d = eng.n, eng.d
        keep = self.keep_history
        auto = keep is None
        if auto:
            keep = n * (d + 4) * 4 * 2 <= HISTORY_AUTO_BYTES
        # per-iteration snapshots are moved to the host at the end of every burst (the loop synchronises there anyway),
        # so the device never holds more than `check_every` of them; an automatic history that outgrows
        # HISTORY_AUTO_BYTES is dropped in favour of the final population
        host_snaps, pending, kept_bytes = [], ([self._snapshot(eng)] if keep else []), 0
        it = 0
        while it < self.max_iter:
            if self.max_iter > engine.MB_HIST_MAX - 1:
                raise _lib.MocatB200Error(f"max_iter <= {engine.MB_HIST_MAX - 1} (device history ring)")
            burst = min(self.check_every, self.max_iter - it)
            for _ in range(burst):
                eng.update()
                it += 1
                self._post_update(eng, initial_extra)
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
        c = eng.settle()
        iters = int(c['iter'])
        hist = eng.ctl.read_hist(iters + 1)
        chain = cdict()
        if keep:
            host_snaps = host_snaps[:iters + 1]
            for k in self._FIELDS:
                setattr(chain, k, np.stack([sn[k] for sn in host_snaps]))
            betas = hist['beta'][:, None]
        else:
            # no device clones: only `value` needs a kernel (SoA -> row-major), the rest is copied from the engine buffers
            host = _to_host(dict(value=eng.values().contiguous(), log_weight=eng.lw, prior_potential=eng.up,
                                 likelihood_potential=eng.lik, alpha=eng.alpha))
            for k in self._FIELDS:
                setattr(chain, k, host[k][None])
            betas = hist['beta'][-1:, None]
        chain.potential = chain.prior_potential + betas * chain.likelihood_potential
        chain.temperature = hist['beta'].copy()
        chain.ess = hist['ess'].copy()
        chain.log_norm_constant = hist['log_z'].copy()
        chain.alpha_mean = hist['alpha_mean'].copy()
        chain.resampled = hist['resampled'].copy()
        chain.bisection_iters = hist['search_iters'].copy()
        initial_extra.iter = iters
        return chain


def _to_host(tensors):
    """device tensors -> NumPy arrays through pinned host memory (torch's caching host allocator hands the blocks of
    a dropped result back to the next run): all copies are enqueued, ONE synchronisation, PCIe at DMA speed instead
    of the staged pageable path.  The arrays own their (pinned) memory."""
    torch = _torch()
    outs = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in tensors.items()}
    for k, v in tensors.items():
        outs[k].copy_(v, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return {k: o.numpy() for k, o in outs.items()}


class RMMetropolisedSMCSampler(MetropolisedSMCSampler):
    """transport/smc.py:376-428: Robbins-Monro adaptation of the MCMC stepsize between iterations,
    log eps += rm_stepsize * (alpha_mean - target), alpha_mean = average acceptance weighted by exp(w - max w).
    Runs entirely on the device (mb_rm_adapt): the stepsize lives in the control block, the adaptation is part of the
    captured step, and the chain of stepsizes is read back once at the end."""
    _private_engine = True

    def __init__(self, *args, rm_stepsize=1., **kwargs):
        super().__init__(*args, **kwargs)
        self.parameters.rm_stepsize = rm_stepsize

    def _configure_engine(self, eng):
        eng.enable_rm(self.parameters.rm_stepsize, self.mcmc_sampler.tuning.target)

    def _run_device(self, scenario, initial_state, initial_extra):
        chain = super()._run_device(scenario, initial_state, initial_extra)
        eng = initial_extra.engine
        chain.stepsize = eng.stepsize_hist[:len(chain.temperature)].cpu().numpy()  # clean_chain :423-428
        initial_extra.parameters.stepsize = float(chain.stepsize[-1])
        return chain


# ======================================================================================================
def adagrad(step_size, momentum=0.9):
    """marker mirroring jax.example_libraries.optimizers.adagrad (the optimiser SVGD defaults to)."""
    return ('adagrad', step_size, momentum)


class SVGD(TransportSampler):
    """transport/svgd.py:42-146.  `adapt(ensemble_state, extra)` may be overridden exactly as in the
    reference (tests/test_transport.py:76-89); `ensemble_state.value` is then a device tensor and
    mocat_b200.kernels.median_bandwidth_update / mean_bandwidth_update run on the device."""
    name = 'SVGD'

    def __init__(self, stepsize, max_iter=1000, kernel=None, kernel_params=None, ensemble_batchsize=None,
                 optimiser=adagrad, keep_history=None, phi_variant=None, sharded=True, **optim_params):
        super().__init__(max_iter=max_iter)
        from .kernels import Gaussian
        self._default_adapt = False
        if kernel is None:
            kernel = Gaussian()
            if type(self).adapt is SVGD.adapt:
                self._default_adapt = True                              # svgd.py:59-62: mean bandwidth
        if not isinstance(kernel, Gaussian):
            raise _lib.MocatB200Error("only the Gaussian kernel is compiled into the device interaction")
        if kernel_params is None:
            kernel_params = kernel.parameters
        if optimiser is not adagrad:
            raise _lib.MocatB200Error("only adagrad (the reference default) is compiled for the device")
        self.parameters.stepsize = stepsize
        self.kernel = kernel
        self.parameters.kernel_params = kernel_params
        self.parameters.ensemble_batchsize = ensemble_batchsize
        self.parameters.optim_params = optim_params
        self.keep_history = keep_history
        self.phi_variant = phi_variant
        self.sharded = bool(sharded)             # under torchrun: split the interaction over the ranks (False: replicate)

    def adapt(self, ensemble_state, extra):                             # svgd.py:109-113
        if self._default_adapt:
            from .kernels import mean_bandwidth_update
            extra.parameters.kernel_params.bandwidth = mean_bandwidth_update(ensemble_state.value, self.phi_variant)
        return ensemble_state, extra

    def _bandwidth_tensor(self, extra, device):
        torch = _torch()
        h = extra.parameters.kernel_params.bandwidth
        if isinstance(h, torch.Tensor):
            return h.to(device=device, dtype=torch.float32).reshape(1)
        return torch.tensor([float(h)], dtype=torch.float32, device=device)

    def startup(self, scenario, n, initial_state, initial_extra, **kwargs):
        torch = _torch()
        initial_state, initial_extra = super().startup(scenario, n, initial_state, initial_extra, **kwargs)
        P = self.parameters
        if P.ensemble_batchsize is not None and P.ensemble_batchsize != n:
            raise _lib.MocatB200Error("ensemble minibatching (ensemble_batchsize < n) is not built for the device")
        dev = torch.device("cuda", torch.cuda.current_device())
        seed = key_to_seed(getattr(initial_extra, 'random_key', None))
        if initial_state is None or getattr(initial_state, 'value', None) is None:
            # transport/sampler.py:24-30: vmap(prior_sample); same Philox stream as the SMC init kernel
            L = _lib.get()
            X = torch.empty((n, scenario.dim), dtype=torch.float32, device=dev)
            L.call("mb_prior_sample", L.ctx(), scenario.prior_mean, scenario.prior_std, scenario.dim, n, seed, 0,
                   _lib.ptr(X), _lib.stream())
            initial_state = cdict()
        else:
            X = torch.as_tensor(np.asarray(initial_state.value, np.float32), device=dev).contiguous()
        st = cdict(value=X)
        # under torchrun the ensemble is sharded by rows over the GPUs (every rank holds all of X; engine.ensemble_shard)
        self._shard = engine.ensemble_shard(n, scenario.dim, self.phi_variant) if self.sharded else None
        st.potential, st.grad_potential = self._potential_grad(scenario, X)
        initial_extra.parameters.kernel_params = self.parameters.kernel_params
        st, initial_extra = self.adapt(st, initial_extra)                # svgd.py:102
        initial_extra.gsq = torch.zeros_like(X)
        initial_extra.mom = torch.zeros_like(X)
        return st, initial_extra

    def _potential_grad(self, scenario, X):
        """vmap(scenario.potential_and_grad) (svgd.py:99,141); sharded: own rows, then one all-gather of (U, G)"""
        sh = getattr(self, '_shard', None)
        if sh is None:
            return scenario._potential_grad_device(X, scenario.temperature)
        torch = _torch()
        import torch.distributed as dist
        _, _, r0, r1 = sh
        U, G = scenario._potential_grad_device(X[r0:r1], scenario.temperature)
        n, d = X.shape
        UG = torch.empty((n, d + 1), dtype=torch.float32, device=X.device)
        dist.all_gather_into_tensor(UG, torch.cat([G, U[:, None]], dim=1).contiguous())
        return UG[:, d].contiguous(), UG[:, :d].contiguous()

    def update(self, scenario, ensemble_state, extra):                  # svgd.py:122-146
        extra.iter = extra.iter + 1
        X = ensemble_state.value
        h = self._bandwidth_tensor(extra, X.device)
        step = self.parameters.stepsize
        step = float(step(extra.iter)) if callable(step) else float(step)
        momentum = self.parameters.optim_params.get('momentum', 0.9)
        sh = getattr(self, '_shard', None)
        if sh is None:
            phi = engine.svgd_phi(X, ensemble_state.grad_potential, h, self.phi_variant)
            engine.adagrad_step(X, extra.gsq, extra.mom, phi, step, momentum)
        else:
            # this rank's rows of the interaction and of the optimiser step, then one all-gather of the new ensemble
            import torch.distributed as dist
            _, _, r0, r1 = sh
            phi = engine.svgd_phi_rows(X, ensemble_state.grad_potential, h, r0, r1)
            mine = X[r0:r1].clone()
            engine.adagrad_step(mine, extra.gsq[r0:r1], extra.mom[r0:r1], phi.contiguous(), step, momentum)
            dist.all_gather_into_tensor(X, mine)
        ensemble_state.potential, ensemble_state.grad_potential = self._potential_grad(scenario, X)
        ensemble_state, extra = self.adapt(ensemble_state, extra)
        return ensemble_state, extra

    def _run_device(self, scenario, state, extra):
        torch = _torch()
        n, d = state.value.shape
        keep = self.keep_history
        if keep is None:
            keep = n * d * 4 * (self.max_iter + 1) <= HISTORY_AUTO_BYTES
        vals = [state.value.clone()] if keep else None
        pots = [state.potential.clone()] if keep else None
        while not self.termination_criterion(state, extra):
            state, extra = self.update(scenario, state, extra)
            if keep:
                vals.append(state.value.clone())
                pots.append(state.potential.clone())
        chain = cdict()
        if keep:
            chain.value = torch.stack(vals).cpu().numpy()
            chain.potential = torch.stack(pots).cpu().numpy()
        else:
            chain.value = state.value.cpu().numpy()[None]
            chain.potential = state.potential.cpu().numpy()[None]
        chain.grad_potential = state.grad_potential.cpu().numpy()[None]
        chain.bandwidth = float(self._bandwidth_tensor(extra, state.value.device).item())
        return chain
