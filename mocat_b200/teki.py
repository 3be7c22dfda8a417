"""Tempered ensemble Kalman inversion: host mirror of mocat/src/transport/teki.py.

`TemperedEKI` / `AdaptiveTemperedEKI` keep the reference constructors and the startup / update /
termination_criterion protocol (teki.py:38-185).  The ensemble (value (n, 4), simulated_data (n, m)) and everything
the reference keeps in ensemble_state / extra (temperature, covariances, Kalman gain, ...) live on the device
(mb_teki, csrc/teki.cu); `run` enqueues updates and polls the device record every `check_every` iterations.
Device scenario family: GKTransformedUniformPrior (the simulator of abc/scenarios/gk.py:68-96 with m in {4, 8, 16}
sorted draws as summary) -- arbitrary Python simulators cannot run on the device and raise; no CPU fallback.
A user-supplied `next_temperature` callable is a per-iteration host function upstream; here the schedule, the
default geometric rule and the adaptive ESS rule are compiled, anything else raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from .core import cdict, key_to_seed
from .transport import TransportSampler, HISTORY_AUTO_BYTES


def _torch():
    import torch
    return torch


class TEKIEngine:
    """device state of one tempered-EKI ensemble"""
    _mocat_transient = True

    def __init__(self, gk, n, seed, mode, max_temperature, max_iter, term_std, nugget, schedule=None, ess_threshold=0.9,
                 tol=1e-5, max_search_iter=1000):
        torch = _torch()
        self.L = _lib.get()
        self.gk, self.n, self.m, self.seed = gk, int(n), int(gk.m), int(seed)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.x = torch.empty((self.n, 4), dtype=torch.float32, device=dev)
        self.sim = torch.empty((self.n, self.m), dtype=torch.float32, device=dev)
        self.partials = torch.zeros(self.L.dll.mb_teki_workspace_doubles(self.m), dtype=torch.float64, device=dev)
        self.scratch = torch.zeros(2 * ((self.n + 31) // 32 * 32), dtype=torch.float32, device=dev) if mode == 2 else None
        self.search_ctl = torch.zeros(C.sizeof(_lib.Control), dtype=torch.uint8, device=dev) if mode == 2 else None
        self.state = torch.zeros(C.sizeof(_lib.Teki), dtype=torch.uint8, device=dev)
        self.max_iter = int(max_iter)
        self.temp_hist = torch.zeros(self.max_iter + 2, dtype=torch.float64, device=dev)
        self.sched = None if schedule is None else torch.as_tensor(np.asarray(schedule, np.float64), device=dev)
        p = _lib.TekiPrm()
        p.max_temperature, p.nugget, p.term_std = float(max_temperature), float(nugget), float(term_std)
        p.ess_threshold, p.tol, p.max_search_iter = float(ess_threshold), float(tol), int(max_search_iter)
        p.max_iter, p.mode = self.max_iter, int(mode)
        p.schedule_len = 0 if self.sched is None else int(self.sched.numel())
        p.schedule = None if self.sched is None else self.sched.data_ptr()
        self.prm = p

    def startup(self, x0=None):
        torch = _torch()
        if x0 is not None:
            x0 = np.ascontiguousarray(x0, np.float32)
            if x0.shape != (self.n, 4):
                raise _lib.MocatB200Error(f"initial value must have shape ({self.n}, 4)")
            self.x.copy_(torch.as_tensor(x0))
        self.L.call("mb_teki_init", self.L.ctx(), C.byref(self.gk), _lib.ptr(self.x), _lib.ptr(self.sim), self.n,
                    0 if x0 is not None else 1, self.seed, 0, _lib.ptr(self.partials), _lib.ptr(self.temp_hist),
                    _lib.ptr(self.state), _lib.stream())

    def update(self):
        self.L.call("mb_teki_update", self.L.ctx(), C.byref(self.gk), C.byref(self.prm), _lib.ptr(self.x), _lib.ptr(self.sim),
                    self.n, self.seed, 0, _lib.ptr(self.partials), _lib.ptr(self.scratch), _lib.ptr(self.search_ctl),
                    _lib.ptr(self.temp_hist), _lib.ptr(self.state), _lib.stream())

    def read(self):
        """host copy of the device record (synchronises)"""
        raw = self.state.cpu().numpy().tobytes()
        return _lib.Teki.from_buffer_copy(raw)


class TemperedEKI(TransportSampler):
    """transport/teki.py:38-150"""
    name = "Tempered EKI"

    def __init__(self, temperature_schedule=None, next_temperature=None, max_temperature=1., max_iter=int(1e4),
                 term_std=0., nugget=1e-5, keep_history=None, check_every=8, **kwargs):
        self.max_iter = max_iter
        self.max_temperature = max_temperature
        self.temperature_schedule = temperature_schedule
        self.keep_history, self.check_every = keep_history, int(check_every)
        super().__init__(**kwargs)
        self.parameters.nugget = nugget
        self.parameters.term_std = term_std
        if next_temperature is not None:
            raise _lib.MocatB200Error("TemperedEKI: a Python next_temperature callable cannot run on the device; use "
                                      "temperature_schedule, the default geometric rule or AdaptiveTemperedEKI")

    def __setattr__(self, key, value):                                   # teki.py:60-66
        super().__setattr__(key, value)
        if key == 'temperature_schedule' and value is not None:
            super().__setattr__('max_temperature', float(np.asarray(value)[-1]))
            super().__setattr__('max_iter', len(value))

    def _mode(self):
        return 0 if self.temperature_schedule is not None else 1

    def _engine_kwargs(self):
        return {}

    def startup(self, scenario, n, initial_state, initial_extra, **kwargs):
        initial_state, initial_extra = super().startup(scenario, n, initial_state, initial_extra, **kwargs)
        if not hasattr(scenario, 'data') or scenario.data is None:       # teki.py:77-78
            raise AttributeError(f'{self.name} requires scenario to have data attribute != None')
        if not hasattr(scenario, '_device'):
            raise _lib.MocatB200Error(f"{self.name}: only the built-in device simulator (GKTransformedUniformPrior) can be "
                                      "inverted; no CPU fallback")
        P = self.parameters
        eng = TEKIEngine(scenario._device(), n, key_to_seed(getattr(initial_extra, 'random_key', None)), self._mode(),
                         self.max_temperature, self.max_iter, P.term_std, P.nugget, schedule=self.temperature_schedule,
                         **self._engine_kwargs())
        x0 = None if initial_state is None else getattr(initial_state, 'value', None)
        eng.startup(x0)
        if initial_state is None:
            initial_state = cdict()
        initial_state.temperature = 0.
        initial_state.engine = eng
        initial_extra.engine = eng
        return initial_state, initial_extra

    def update(self, scenario, ensemble_state, extra):
        extra.engine.update()
        extra.iter = extra.iter + 1
        return ensemble_state, extra

    def termination_criterion(self, ensemble_state, extra):              # teki.py:104-111, evaluated by the next update
        s = extra.engine.read()
        P = self.parameters
        return bool(s.done or s.temperature >= self.max_temperature or s.iter >= self.max_iter)

    def _snapshot(self, eng):
        return dict(value=eng.x.clone(), simulated_data=eng.sim.clone())

    def _run_device(self, scenario, initial_state, initial_extra):
        eng = initial_extra.engine
        keep = self.keep_history
        auto = keep is None
        if auto:
            keep = eng.n * (4 + eng.m) * 4 * 2 <= HISTORY_AUTO_BYTES
        host_snaps, pending, kept_bytes = [], ([self._snapshot(eng)] if keep else []), 0
        it = 0
        while it <= self.max_iter:                                       # one extra call lets the device see termination
            for _ in range(min(self.check_every, self.max_iter + 1 - it)):
                eng.update()
                it += 1
                if keep:
                    pending.append(self._snapshot(eng))
            s = eng.read()
            if keep:
                for sn in pending:
                    host_snaps.append({k: v.cpu().numpy() for k, v in sn.items()})
                    kept_bytes += sum(v.nbytes for v in host_snaps[-1].values())
                pending = []
                if auto and kept_bytes > HISTORY_AUTO_BYTES:
                    keep, host_snaps = False, []
            if s.done:
                break
        s = eng.read()
        iters = int(s.iter)
        chain = cdict()
        if keep:                                                         # updates after termination changed nothing
            for k in ('value', 'simulated_data'):
                setattr(chain, k, np.stack([sn[k] for sn in host_snaps[:iters + 1]]))
        else:
            for k, v in self._snapshot(eng).items():
                setattr(chain, k, v.cpu().numpy()[None])
        chain.temperature = eng.temp_hist[:iters + 1].cpu().numpy()
        chain.perturb_nan = int(s.perturb_nan)
        chain.kalman_gain = np.ctypeslib.as_array(s.gain).reshape(4, 16)[:, :eng.m].copy()
        chain.prec_y_given_x = np.ctypeslib.as_array(s.prec).reshape(16, 16)[:eng.m, :eng.m].copy()
        initial_extra.iter = iters
        return chain


class AdaptiveTemperedEKI(TemperedEKI):
    """transport/teki.py:153-185"""

    def __init__(self, max_temperature=1., max_iter=int(1e4), nugget=1e-5, term_std=0., ess_threshold=0.9,
                 bisection_tol=1e-5, max_bisection_iter=1000, **kwargs):
        super().__init__(temperature_schedule=None, max_temperature=max_temperature, max_iter=max_iter, nugget=nugget,
                         term_std=term_std, **kwargs)
        self.parameters.ess_threshold = ess_threshold
        self.parameters.bisection_tol = bisection_tol
        self.parameters.max_bisection_iter = max_bisection_iter

    def _mode(self):
        return 2

    def _engine_kwargs(self):
        P = self.parameters
        return dict(ess_threshold=P.ess_threshold, tol=P.bisection_tol, max_search_iter=P.max_bisection_iter)
