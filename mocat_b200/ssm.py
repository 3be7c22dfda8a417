"""State-space models and bootstrap particle filtering: host mirror of mocat/src/ssm/{ssm,filtering}.py,
ssm/linear_gaussian/{linear_gaussian,kalman}.py, ssm/nonlinear_gaussian.py and ssm/scenarios/lorenz96.py.

The filtering functions keep the reference signatures (filtering.py:173,202,220,255).  Differences, all
forced by scale: the particle cdict carries a device engine; the stacked (T, n, d) history the
reference returns is kept only when `keep_history` is set / small (otherwise `value` holds the latest
population with a leading axis of 1), and per-step weighted means/variances, ESS and the running
log-evidence (`log_norm_constant`, convention of transport/smc.py:160,212-215) are always returned.
The transition of Lorenz-96 is `substeps` fixed RK4 steps (the reference uses adaptive odeint,
lorenz96.py:23-26); see DESIGN.md.
"""
import numpy as np

from . import _lib, engine, models
from .core import cdict, key_to_seed

_RESAMPLING = {'multinomial': _lib.RESAMPLE_MULTINOMIAL, 'systematic': _lib.RESAMPLE_SYSTEMATIC}
HISTORY_AUTO_BYTES = 1 << 30


def _torch():
    import torch
    return torch


class StateSpaceModel:
    """ssm/ssm.py:18-81 (interface); device models implement `_ssm(dt)`."""
    name = None
    dim = None
    dim_obs = None

    def __repr__(self):
        return f"mocat.StateSpaceModel.{self.__class__.__name__}"

    def _ssm(self, dt=None):
        raise _lib.MocatB200Error(f"{type(self).__name__}: only built-in device state-space models "
                                  "(TimeHomogenousLinearGaussian, Lorenz96) can be filtered; no CPU fallback")


class TimeHomogenousLinearGaussian(StateSpaceModel):
    """ssm/linear_gaussian/linear_gaussian.py:142-261."""
    name = 'Time-homogenous Linear Gaussian'

    def __init__(self, initial_mean=None, initial_covariance=None, transition_matrix=None,
                 transition_covariance=None, likelihood_matrix=None, likelihood_covariance=None, name=None, dim=None):
        if name is not None:
            self.name = name
        if dim is None:
            for a in (initial_mean, initial_covariance, transition_covariance, likelihood_matrix, likelihood_covariance):
                if a is not None and np.ndim(a) > 0:
                    dim = np.asarray(a).shape[-1]
                    break
        if dim is None:
            raise AttributeError(f'Could not find dimension for {self.__class__.__name__}')
        self.dim = int(dim)
        eye = np.eye(self.dim)
        self.initial_mean = np.zeros(self.dim) if initial_mean is None else np.asarray(initial_mean, np.float64)
        self.initial_covariance = eye if initial_covariance is None else np.atleast_2d(initial_covariance)
        self.transition_matrix = eye if transition_matrix is None else np.atleast_2d(transition_matrix)
        self.transition_covariance = eye if transition_covariance is None else np.atleast_2d(transition_covariance)
        self.likelihood_matrix = eye if likelihood_matrix is None else np.atleast_2d(likelihood_matrix)
        self.dim_obs = self.likelihood_matrix.shape[0]
        self.likelihood_covariance = np.eye(self.dim_obs) if likelihood_covariance is None \
            else np.atleast_2d(likelihood_covariance)

    def _ssm(self, dt=None):
        return models.make_lg_ssm(self.initial_mean, self.initial_covariance, self.transition_matrix,
                                  self.transition_covariance, self.likelihood_matrix, self.likelihood_covariance)

    def simulate(self, t_all, random_key):
        """ssm/ssm.py:138-163 (host; data generation is not on the hot path)."""
        rng = np.random.default_rng(key_to_seed(random_key))
        T = len(t_all)
        L0, LQ, LR = (np.linalg.cholesky(a) for a in (self.initial_covariance, self.transition_covariance,
                                                     self.likelihood_covariance))
        x = np.empty((T, self.dim))
        x[0] = self.initial_mean + L0 @ rng.standard_normal(self.dim)
        for i in range(1, T):
            x[i] = self.transition_matrix @ x[i - 1] + LQ @ rng.standard_normal(self.dim)
        y = x @ self.likelihood_matrix.T + rng.standard_normal((T, self.dim_obs)) @ LR.T
        return cdict(x=x, y=y, t=np.asarray(t_all), name=f'{self.name} simulation')


class Lorenz96(StateSpaceModel):
    """ssm/scenarios/lorenz96.py:29-44 on NonLinearGaussian (ssm/nonlinear_gaussian.py:19-131) with the
    reference defaults Q = R = P0 = I, H = I, m0 = 0 (isotropic std's may be changed)."""
    name = 'Lorenz 96'

    def __init__(self, dim=40, forcing_constant=8., substeps=1, transition_std=1.0, likelihood_std=1.0,
                 initial_mean=0.0, initial_std=1.0, name=None):
        if name is not None:
            self.name = name
        self.dim = self.dim_obs = int(dim)
        self.forcing_constant, self.substeps = float(forcing_constant), int(substeps)
        self.transition_std, self.likelihood_std = float(transition_std), float(likelihood_std)
        self.initial_mean, self.initial_std = float(initial_mean), float(initial_std)

    def _ssm(self, dt=None):
        return models.make_lorenz96(self.dim, self.forcing_constant, 0.05 if dt is None else dt, self.substeps,
                                    self.transition_std, self.likelihood_std, self.initial_mean, self.initial_std)

    def _flow(self, x, dt):
        """host fp64 version of the device transition map (`substeps` RK4 steps); data generation only"""
        F, h = self.forcing_constant, dt / self.substeps
        rhs = lambda v: (np.roll(v, -1) - np.roll(v, 2)) * np.roll(v, 1) - v + F          # lorenz96.py:14-20
        for _ in range(self.substeps):
            k1 = rhs(x); k2 = rhs(x + 0.5 * h * k1); k3 = rhs(x + 0.5 * h * k2); k4 = rhs(x + h * k3)
            x = x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        return x

    def simulate(self, t_all, random_key, spinup=1000):
        """ssm/ssm.py:138-163 (host; data generation is not on the hot path): a trajectory started on the attractor
        (`spinup` noise-free steps) with process noise, observed through y = x + likelihood_std * noise"""
        rng = np.random.default_rng(key_to_seed(random_key))
        t_all = np.asarray(t_all, np.float64)
        dt = float(t_all[1] - t_all[0]) if len(t_all) > 1 else 0.05
        x = self.initial_mean + self.initial_std * rng.standard_normal(self.dim) + self.forcing_constant
        for _ in range(spinup):
            x = self._flow(x, dt)
        xs, ys = np.empty((len(t_all), self.dim)), np.empty((len(t_all), self.dim))
        for i in range(len(t_all)):
            if i > 0:
                x = self._flow(x, dt) + self.transition_std * rng.standard_normal(self.dim)
            xs[i] = x
            ys[i] = x + self.likelihood_std * rng.standard_normal(self.dim)
        return cdict(x=xs, y=ys, t=t_all, name=f'{self.name} simulation')


class ParticleFilter:
    """ssm/filtering.py:20-139 (interface)."""
    name = 'Particle Filter'

    def __init__(self, name=None):
        if name is not None:
            self.name = name

    def __repr__(self):
        return f"mocat.ParticleFilter.{self.__class__.__name__}"

    def startup(self, ssm_scenario):
        pass


class BootstrapFilter(ParticleFilter):
    """ssm/filtering.py:142-170: proposal = transition, weight increment = -likelihood_potential."""
    name = 'Bootstrap Filter'


class OptimalNonLinearGaussianParticleFilter(ParticleFilter):
    """ssm/nonlinear_gaussian.py:134-276: the locally optimal proposal p(x_t | x_{t-1}, y_t) of a non-linear Gaussian
    model with a linear-Gaussian observation, weight increment log N(y_t; H f(x_{t-1}), H Q H^T + R).  Compiled into
    the Lorenz-96 step kernel (H = I, isotropic Q, R, P0: every matrix of `startup` :152-186 is a scalar, kept here
    under the reference's attribute names)."""
    name = 'Optimal Non-linear Gaussian Particle Filter'

    def startup(self, ssm_scenario):
        if not isinstance(ssm_scenario, Lorenz96):
            raise _lib.MocatB200Error("OptimalNonLinearGaussianParticleFilter is compiled for Lorenz96 (H = I, isotropic "
                                      "noise) only; no CPU fallback")
        q2, r2, p2 = ssm_scenario.transition_std ** 2, ssm_scenario.likelihood_std ** 2, ssm_scenario.initial_std ** 2
        eye = np.eye(ssm_scenario.dim)
        self.initial_kalman_gain = p2 / (p2 + r2) * eye
        self.initial_conditioned_covariance_sqrt = np.sqrt(1.0 / (1.0 / p2 + 1.0 / r2)) * eye
        self.initial_conditioned_precision_sqrt = eye / np.sqrt(1.0 / (1.0 / p2 + 1.0 / r2))
        self.proposal_kalman_gain = q2 / (q2 + r2) * eye
        self.proposal_covariance_sqrt = np.sqrt(q2 * r2 / (q2 + r2)) * eye
        self.proposal_precision_sqrt = eye / np.sqrt(q2 * r2 / (q2 + r2))
        self.weight_precision_sqrt = eye / np.sqrt(q2 + r2)


class EnsembleKalmanFilter(ParticleFilter):
    """ssm/nonlinear_gaussian.py:279-350: conditioned initial ensemble (as the optimal filter), forecast through the
    transition, ensemble covariance -> Kalman gain -> perturbed-observation update; log-weights stay zero (no
    resampling, no evidence).  Device path: Lorenz96 (H = I, isotropic R), one GPU; the forecast is the Lorenz-96 step
    kernel, the analysis mb_enkf_analysis (csrc/enkf.cu)."""
    name = 'Ensemble Kalman Filter'

    def startup(self, ssm_scenario):
        if not isinstance(ssm_scenario, Lorenz96):
            raise _lib.MocatB200Error("EnsembleKalmanFilter is compiled for Lorenz96 (H = I, isotropic noise) only; "
                                      "no CPU fallback")
        r2, p2 = ssm_scenario.likelihood_std ** 2, ssm_scenario.initial_std ** 2
        eye = np.eye(ssm_scenario.dim)
        self.initial_kalman_gain = p2 / (p2 + r2) * eye
        self.initial_conditioned_covariance_sqrt = np.sqrt(1.0 / (1.0 / p2 + 1.0 / r2)) * eye


def _check_filter(pf):
    if not isinstance(pf, (BootstrapFilter, OptimalNonLinearGaussianParticleFilter, EnsembleKalmanFilter)):
        raise _lib.MocatB200Error("only BootstrapFilter, OptimalNonLinearGaussianParticleFilter and EnsembleKalmanFilter "
                                  "are compiled into the device step (no CPU fallback)")


def _device_ssm(ssm_scenario, particle_filter, dt=None):
    """POD model of the scenario with the filter's proposal folded in"""
    s = ssm_scenario._ssm() if dt is None else ssm_scenario._ssm(dt)
    if isinstance(particle_filter, (OptimalNonLinearGaussianParticleFilter, EnsembleKalmanFilter)):
        particle_filter.startup(ssm_scenario)                         # raises for anything but Lorenz96
        s.proposal = _lib.PROPOSAL_ENKF if isinstance(particle_filter, EnsembleKalmanFilter) else _lib.PROPOSAL_OPTIMAL
    return s


def _moments(eng):
    return eng.moments()


def _engine_of(particles):
    eng = getattr(particles, 'engine', None)
    if eng is None:
        raise _lib.MocatB200Error("this particle cdict has no live device engine (it was loaded from disk or created "
                                  "elsewhere); start again with initiate_particles / run_particle_filter_for_marginals")
    if getattr(particles, 'engine_generation', eng.generation) != eng.generation:
        raise _lib.MocatB200Error("the device engine of this particle cdict has since been re-used by another filter run "
                                  "of the same configuration (engines are pooled); continue from the latest result or "
                                  "set MOCAT_B200_ENGINE_POOL=0")
    return eng


def _world():
    """(rank, world) of the torch.distributed job this process belongs to ((0, 1) outside torchrun)"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _make_engine(ssm, n, seed, ess_threshold, resampling):
    """single-GPU engine, or -- under torchrun, where `n` is the GLOBAL population size -- this rank's shard of one
    population spread over the GPUs of the node (mocat_b200/parallel.py)"""
    rank, world = _world()
    if world > 1:
        from . import parallel
        _, n_local = parallel.shard_range(n, rank, world)
        return parallel.acquire_sharded_pf(ssm, n_local, seed, ess_threshold, resampling)
    return engine.PFEngine.acquire(ssm, n, seed, ess_threshold=ess_threshold, resampling=resampling)


def _host_value(eng):
    """latest population as a host array, or None when it is too large to be worth a device->host copy by default
    (n = 1e8, d = 40 is 16 GB); the engine keeps it on the device (`particles.engine.values()`)"""
    if eng.n * (eng.d + 1) * 4 > HISTORY_AUTO_BYTES:
        return None, None
    return eng.values().cpu().numpy()[None], eng.lw.cpu().numpy()[None]


def initiate_particles(ssm_scenario, particle_filter, n, random_key, y=None, t=None, ess_threshold=0.5,
                       resampling='multinomial'):
    """ssm/filtering.py:173-193.  The returned cdict carries the live device engine (`engine`, dropped on save): the
    online API is STATEFUL -- propagate_particle_filter advances that engine in place."""
    torch = _torch()
    _check_filter(particle_filter)
    particle_filter.startup(ssm_scenario)
    if y is None:
        raise _lib.MocatB200Error("initiate_particles needs the first observation y")
    y = np.atleast_1d(np.asarray(y, np.float32))
    eng = engine.PFEngine.acquire(_device_ssm(ssm_scenario, particle_filter), n, key_to_seed(random_key), ess_threshold=ess_threshold,
                                  resampling=_RESAMPLING[resampling])
    eng.init(torch.as_tensor(y, device="cuda"))
    c = eng.ctl.read()
    mean, var = _moments(eng)
    out = cdict(t=np.atleast_1d(t) if t is not None else np.zeros(1), y=y[None],
                ess=np.atleast_1d(c['ess']), log_norm_constant=np.atleast_1d(c['log_z']),
                mean=mean.cpu().numpy()[None], var=var.cpu().numpy()[None], engine=eng,
                engine_generation=eng.generation)
    out.value, out.log_weight = _host_value(eng)
    return out


def resample_particles(particles, random_key=None, resample_full=True):
    """ssm/filtering.py:202-217: unconditional resampling of the latest population, log-weights reset to zero.
    Runs the engine's resampler (ancestors) and a device gather (mb_gather_state / mb_gather_rows); only the latest
    time slice lives on the device, so `resample_full` (re-indexing the stored trajectories) applies to the host copy."""
    torch = _torch()
    eng = _engine_of(particles)
    c = eng.ctl.read()
    c['resample'] = 1
    eng.ctl.write(c)
    eng._resample_kernels(_lib.stream())
    eng.gather_current()
    anc = eng.anc.cpu().numpy() if particles.value is not None else None
    eng._lw_full.zero_()
    logn = float(np.log(eng.n_total))
    c['wmax'], c['s1'], c['s2'] = 0.0, float(eng.n_total), float(eng.n_total)
    c['lse'] = c['lse2'] = c['log_ess'] = logn
    c['ess'], c['resample'], c['resampled'] = float(eng.n_total), 0, 1
    eng.ctl.write(c)
    out = particles.copy()
    if particles.value is not None:
        out.value = particles.value[:, anc] if resample_full else particles.value.copy()
        out.value[-1] = eng.values().cpu().numpy()
        lw = np.asarray(particles.log_weight)
        if lw.ndim == 1:                                            # the smoother keeps the latest weights only
            out.log_weight = np.zeros_like(lw)
        else:
            out.log_weight = lw.copy()
            out.log_weight[-1] = 0.0
    if np.ndim(particles.ess) == 0:
        out.ess = float(eng.n_total)
    else:
        out.ess = np.array(particles.ess, dtype=np.float64)
        out.ess[-1] = float(eng.n_total)
    return out


def propagate_particle_filter(ssm_scenario, particle_filter, particles, y_new, t_new, random_key=None,
                              ess_threshold=0.5, resample_full=True):
    """ssm/filtering.py:220-252: resample iff ess[-1] < ess_threshold*n, propose, weight, append."""
    torch = _torch()
    _check_filter(particle_filter)
    eng = _engine_of(particles)
    eng.ess_threshold = float(ess_threshold)
    y_new = np.atleast_1d(np.asarray(y_new, np.float32))
    t_prev = float(particles.t[-1])
    eng.ssm = _device_ssm(ssm_scenario, particle_filter, float(t_new) - t_prev)
    # the resample decision for this step was taken on the device with the threshold stored in the control
    # block by the previous step; refresh it if the caller changed ess_threshold
    c = eng.ctl.read()
    want = 1 if c['ess'] < ess_threshold * eng.n_total else 0
    if want != c['resample']:
        c['resample'] = want
        eng.ctl.write(c)
    eng.step(torch.as_tensor(y_new, device="cuda"))
    c = eng.ctl.read()
    mean, var = _moments(eng)
    out = particles.copy()
    if particles.value is not None:
        out.value = np.append(particles.value, eng.values().cpu().numpy()[None], axis=0)
        out.log_weight = np.append(particles.log_weight, eng.lw.cpu().numpy()[None], axis=0)
    out.y = np.append(particles.y, y_new[None], axis=0)
    out.t = np.append(particles.t, t_new)
    out.ess = np.append(particles.ess, c['ess'])
    out.log_norm_constant = np.append(particles.log_norm_constant, c['log_z'])
    out.mean = np.append(particles.mean, mean.cpu().numpy()[None], axis=0)
    out.var = np.append(particles.var, var.cpu().numpy()[None], axis=0)
    return out


def run_particle_filter_for_marginals(ssm_scenario, particle_filter, y, t, random_key, n=None, initial_sample=None,
                                      ess_threshold=0.5, resampling='multinomial', keep_history=None,
                                      moments=True, history_every=None):
    """ssm/filtering.py:255-324.  The time loop is enqueued without any host synchronisation; per-step ESS,
    log-evidence and (moments=True) weighted means / variances are read back once at the end.  The stacked
    (T, n, d) history of the reference is returned when it fits `HISTORY_AUTO_BYTES` (or keep_history=True);
    otherwise `value` / `log_weight` hold the final population only, or None when even that exceeds the budget
    (it stays on the device: `out.engine.values()`).  initial_sample: a cdict from initiate_particles / a previous
    call -- every (y, t) is then a propagation step of its engine (filtering.py:266-276).
    history_every=k: a THINNED history -- the populations of steps 0, k, 2k, ... and of the last step -- streamed to
    pinned host memory while the filter runs (mocat_b200.history.HistoryStream); `value` / `log_weight` then hold those
    records and `history_index` their positions in `t`."""
    torch = _torch()
    _check_filter(particle_filter)
    y = np.asarray(y, np.float32)
    if y.ndim == 1:
        y = y[..., np.newaxis]
    t = np.asarray(t, np.float64)
    T = len(y)
    if T > engine.MB_HIST_MAX:
        raise _lib.MocatB200Error(f"at most {engine.MB_HIST_MAX} time steps per call (device history ring)")
    if initial_sample is None:
        dt = float(t[1] - t[0]) if T > 1 else None
        t_all = t
    else:
        dt = float(t[0] - initial_sample.t[-1])
        t_all = np.concatenate([[float(initial_sample.t[-1])], t])
    if len(t_all) > 2 and not np.allclose(np.diff(t_all), dt):
        raise _lib.MocatB200Error("run_particle_filter_for_marginals needs equally spaced t (time-homogeneous model)")
    if initial_sample is None:
        eng = _make_engine(_device_ssm(ssm_scenario, particle_filter, dt), n, key_to_seed(random_key), ess_threshold, _RESAMPLING[resampling])
    else:
        eng = _engine_of(initial_sample)
        eng.ess_threshold = float(ess_threshold)
        eng.ssm = _device_ssm(ssm_scenario, particle_filter, dt)
        c = eng.ctl.read()
        want = 1 if c['ess'] < ess_threshold * eng.n_total else 0
        if want != c['resample']:
            c['resample'] = want
            eng.ctl.write(c)
    n, d = eng.n, eng.d
    if keep_history is None:
        keep_history = T * n * (d + 1) * 4 <= HISTORY_AUTO_BYTES
    yd = torch.as_tensor(y, device="cuda")
    sharded = _world()[1] > 1
    hs = None
    if history_every is not None or keep_history:
        from .history import HistoryStream
        every = 1 if history_every is None else max(1, int(history_every))
        hs = HistoryStream((n, d), every, (T + every - 1) // every + 1)
    if sharded and not eng.rowmajor:
        moments = False                                   # per-shard moment sums exist for the row-major layout only
    mom = torch.empty((T, 2, d), dtype=torch.float64, device="cuda") if moments and not sharded else None
    msum = torch.empty((T, 1 + 2 * d), dtype=torch.float64, device="cuda") if moments and sharded else None

    def record(i):
        if hs is not None and (hs.wants(i) or i == T - 1):                # streamed to pinned host memory, no host sync
            v = eng.values()
            hs.push(i, v if v.is_contiguous() else v.contiguous(), eng.lw, force=True)
        if msum is not None:                              # this shard's raw sums, shifted by the observation (H = I)
            eng.moment_sums(yd[i], out=msum[i])
        elif moments:
            eng.moments(out=mom[i])                                    # written in place: no per-step copies

    t0 = eng.t + 1 if initial_sample is not None else 0
    for i in range(T):
        if hs is not None:
            hs.before_overwrite(i)
        if initial_sample is None and i == 0:
            eng.init(yd[0])
        else:
            eng.step(yd[i])
        record(i)
    hist = eng.ctl.read_hist(t0 + T)[t0:]
    out = cdict(t=t, y=y, ess=hist['ess'].copy(), log_norm_constant=hist['log_z'].copy(),
                resampled=hist['resampled'].copy(), engine=eng, engine_generation=eng.generation)
    if hs is not None:
        out.value, out.log_weight, kept = hs.finish()
        if history_every is not None:                                  # else: the full stacked history of the reference
            out.history_index = kept
    else:
        out.value, out.log_weight = _host_value(eng)
    if msum is not None:                                  # one all-reduce of T (1 + 2d) doubles for the whole run
        import torch.distributed as dist
        dist.all_reduce(msum)
        sh = msum.cpu().numpy()
        m1 = sh[:, 1:1 + d] / sh[:, :1]
        out.mean, out.var = m1 + y.astype(np.float64), sh[:, 1 + d:] / sh[:, :1] - m1 * m1
    elif moments:
        mh = mom.cpu().numpy()
        out.mean, out.var = mh[:, 0], mh[:, 1]
    if initial_sample is not None:
        for k in ('t', 'y', 'ess', 'log_norm_constant', 'mean', 'var', 'value', 'log_weight'):
            prev, new = getattr(initial_sample, k, None), getattr(out, k, None)
            if prev is not None and new is not None and (k not in ('value', 'log_weight') or (keep_history and history_every is None)):
                setattr(out, k, np.concatenate([prev, new], axis=0))
    return out


def backward_simulation(ssm_scenario, marginal_particles, random_key, n_samps=None, maximum_rejections=0,
                        transition_dens_bound_parameter=0., bound_inflation=1.01):
    """ssm/backward.py:275-300: backward simulation (FFBSi) from stacked filter output `marginal_particles` (value
    (T, n_pf, d), log_weight (T, n_pf), t (T,): run the filter with keep_history=True).  Every time step is one device
    contraction of the n_samps backward samples against the n_pf filter particles with a Gumbel-max categorical draw
    (mb_backward_sample, csrc/backward.cu) -- `backward_simulation_full`, backward.py:241-272.  The rejection variant
    (maximum_rejections > 0, :170-238) draws from the SAME law with data-dependent work; it is served by the full
    contraction here and `num_transition_evals` reports the n_pf * n_samps evaluations actually made."""
    torch = _torch()
    import ctypes as C
    vals = getattr(marginal_particles, 'value', None)
    lws = getattr(marginal_particles, 'log_weight', None)
    if vals is None or lws is None or np.ndim(vals) != 3:
        raise _lib.MocatB200Error("backward_simulation needs the stacked filter history: value (T, n_pf, d) and log_weight "
                                  "(T, n_pf) -- run the filter with keep_history=True")
    vals, lws = np.asarray(vals, np.float32), np.asarray(lws, np.float32)
    T, n_pf, d = vals.shape
    n_s = n_pf if n_samps is None else int(n_samps)
    times = np.asarray(marginal_particles.t, np.float64)
    seed = key_to_seed(random_key)
    L = _lib.get()
    dev = torch.device("cuda", torch.cuda.current_device())
    work = torch.empty((n_pf, d), dtype=torch.float32, device=dev)
    idx = torch.empty(n_s, dtype=torch.int32, device=dev)
    out = torch.empty((T, n_s, d), dtype=torch.float32, device=dev)

    def draw(ind, x1):
        x0 = torch.as_tensor(vals[ind], device=dev).contiguous()
        lw0 = torch.as_tensor(lws[ind], device=dev).contiguous()
        dt = float(times[ind + 1] - times[ind]) if x1 is not None else 0.0
        s = ssm_scenario._ssm(dt if x1 is not None else None)
        L.call("mb_backward_sample", L.ctx(), C.byref(s), dt, _lib.ptr(x0), _lib.ptr(lw0), n_pf, _lib.ptr(x1), n_s,
               _lib.ptr(work), seed, ind, _lib.ptr(idx), _lib.ptr(out[ind]), _lib.stream())

    draw(T - 1, None)                                                   # backward.py:254-256
    for ind in range(T - 2, -1, -1):                                    # :258-266
        draw(ind, out[ind + 1])
    res = marginal_particles.copy()
    res.value = out.cpu().numpy()
    res.num_transition_evals = np.append(0, np.ones(T - 1) * n_pf * n_s)
    if hasattr(res, 'log_weight'):
        del res.log_weight
    return res


def forward_filtering_backward_simulation(ssm_scenario, particle_filter, y, t, n_samps, random_key, n_pf=None,
                                          ess_threshold=0.5, maximum_rejections=0, transition_dens_bound_parameter=0.,
                                          bound_inflation=1.01):
    """ssm/backward.py:303-350"""
    import time
    if n_pf is None:
        n_pf = n_samps
    seed = key_to_seed(random_key)
    t0 = time.time()
    pf_samps = run_particle_filter_for_marginals(ssm_scenario, particle_filter, y, t, seed, n=n_pf,
                                                 ess_threshold=ess_threshold, keep_history=True)
    _torch().cuda.synchronize()
    t1 = time.time()
    out = backward_simulation(ssm_scenario, pf_samps, seed + 1, n_samps, maximum_rejections,
                              transition_dens_bound_parameter, bound_inflation)
    t2 = time.time()
    out.time, out.pf_time, out.bsi_time = t2 - t0, t1 - t0, t2 - t1
    return out


def run_kalman_filter_for_marginals(lgssm_scenario, y, t, return_log_likelihood=False):
    """ssm/linear_gaussian/kalman.py:16-57: exact filtering means / covariances of a TimeHomogenousLinearGaussian model,
    evaluated on the device (mb_kalman_filter: one warp, fp64, matrices of at most 8 x 8) so that the cross-check of
    the particle filter runs where the filter runs.  Fixes kalman.py:20 (cov_0 is L0 L0^T here; identical when P0 = I)
    and can also return the innovation log-likelihood."""
    torch = _torch()
    y = np.asarray(y, np.float32)
    if y.ndim == 1:
        y = y[:, None]
    s = lgssm_scenario._ssm()
    if int(s.kind) != _lib.SSM_LINEAR_GAUSSIAN:
        raise _lib.MocatB200Error("run_kalman_filter_for_marginals needs a TimeHomogenousLinearGaussian model")
    L = _lib.get()
    T, d = len(y), lgssm_scenario.dim
    yd = torch.as_tensor(np.ascontiguousarray(y), device="cuda")
    means = torch.empty((T, d), dtype=torch.float64, device="cuda")
    covs = torch.empty((T, d, d), dtype=torch.float64, device="cuda")
    ll = torch.empty(1, dtype=torch.float64, device="cuda")
    import ctypes as C
    L.call("mb_kalman_filter", L.ctx(), C.byref(s), _lib.ptr(yd), T, _lib.ptr(means), _lib.ptr(covs), _lib.ptr(ll),
           _lib.stream())
    mus, cv = means.cpu().numpy(), covs.cpu().numpy()
    return (mus, cv, float(ll.item())) if return_log_likelihood else (mus, cv)


def kalman_filter_host(lgssm_scenario, y, t=None, return_log_likelihood=False):
    """the same recursion in NumPy fp64 on the host (used by the CPU-only tests of the host API; the product path is
    run_kalman_filter_for_marginals on the device)"""
    y = np.asarray(y, np.float64)
    if y.ndim == 1:
        y = y[:, None]
    s = lgssm_scenario
    F, H, Q, R = s.transition_matrix, s.likelihood_matrix, s.transition_covariance, s.likelihood_covariance
    mu, cov = np.asarray(s.initial_mean, np.float64).copy(), np.asarray(s.initial_covariance, np.float64).copy()
    T = len(y)
    mus, covs, ll = np.empty((T, s.dim)), np.empty((T, s.dim, s.dim)), 0.0
    for i in range(T):
        if i > 0:
            mu, cov = F @ mu, F @ cov @ F.T + Q
        S = H @ cov @ H.T + R
        innov = y[i] - H @ mu
        ll += -0.5 * (innov @ np.linalg.solve(S, innov) + np.linalg.slogdet(S)[1] + s.dim_obs * np.log(2 * np.pi))
        K = cov @ H.T @ np.linalg.inv(S)
        mu, cov = mu + K @ innov, cov - K @ H @ cov
        mus[i], covs[i] = mu, cov
    return (mus, covs, ll) if return_log_likelihood else (mus, covs)


def propagate_particle_smoother(*args, **kwargs):
    """ssm/online_smoothing.py:364-386 (mocat_b200/online_smoothing.py)"""
    from .online_smoothing import propagate_particle_smoother as f
    return f(*args, **kwargs)


def propagate_particle_smoother_pf(*args, **kwargs):
    """ssm/online_smoothing.py:211-282"""
    from .online_smoothing import propagate_particle_smoother_pf as f
    return f(*args, **kwargs)


def propagate_particle_smoother_bs(*args, **kwargs):
    """ssm/online_smoothing.py:285-361"""
    from .online_smoothing import propagate_particle_smoother_bs as f
    return f(*args, **kwargs)
