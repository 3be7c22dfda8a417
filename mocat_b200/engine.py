"""Device engines: torch tensors own the HBM, the C-ABI kernels do the work.

Layout in HBM: particle values are SoA `x[d][ld]` (a torch (d, ld) float32 tensor, ld = n
rounded up to 32 elements so every column starts 128 B aligned), double buffered for the fused
ancestor-gather; per-particle scalars (`lw`, `lik`, `up`, `alpha`, `dist`) are (n,) float32; the fp64 CDF
(n,) (multinomial only) and int32 ancestors (n,); all per-iteration scalars live in one
144-byte control block on the device plus a history ring of 48-byte records.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream


MB_HIST_MAX = _lib.MB_HIST_MAX


def _pad(n):
    return (int(n) + 31) // 32 * 32


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


class ControlBlock:
    """Device mb_control + history ring."""

    def __init__(self, hist_len=_lib.MB_HIST_MAX):
        self.t = torch.zeros(_lib.CONTROL_DTYPE.itemsize, dtype=torch.uint8, device=_dev())
        self.hist = torch.zeros(hist_len * _lib.HIST_DTYPE.itemsize, dtype=torch.uint8, device=_dev())
        self._pin = torch.empty(_lib.CONTROL_DTYPE.itemsize, dtype=torch.uint8).pin_memory()

    def read(self):
        """synchronising read of the control block -> numpy record"""
        self._pin.copy_(self.t, non_blocking=False)
        return self._pin.numpy().view(_lib.CONTROL_DTYPE)[0].copy()

    def write(self, rec):
        buf = np.zeros(1, dtype=_lib.CONTROL_DTYPE)
        buf[0] = rec
        self.t.copy_(torch.from_numpy(buf.view(np.uint8)))

    def read_hist(self, count):
        if count > self.hist.numel() // _lib.HIST_DTYPE.itemsize:
            raise _lib.MocatB200Error(f"history ring holds {self.hist.numel() // _lib.HIST_DTYPE.itemsize} records, "
                                      f"{count} requested (MB_HIST_MAX)")
        nbytes = count * _lib.HIST_DTYPE.itemsize
        return self.hist[:nbytes].cpu().numpy().view(_lib.HIST_DTYPE).copy()


# ------------------------------------------------------------------------------------------- primitives
def lse_ess(lw, lik=None, dbeta=0.0):
    """(wmax, s1, s2, lse, lse2, log_ess) of lw - dbeta*lik as a (6,) float64 device tensor."""
    L = _lib.get()
    out = torch.empty(6, dtype=torch.float64, device=lw.device)
    L.call("mb_lse_ess", L.ctx(), ptr(lw), ptr(lik), float(dbeta), lw.numel(), ptr(out), stream())
    return out


def cumsum_f32(w, scale):
    L = _lib.get()
    cdf = torch.empty(w.numel(), dtype=torch.float64, device=w.device)
    L.call("mb_cumsum_f32", L.ctx(), ptr(w), w.numel(), float(scale), ptr(cdf), stream())
    return cdf


def cumsum_lw(lw, ctl, force=True, out=None):
    L = _lib.get()
    cdf = out if out is not None else torch.empty(lw.numel(), dtype=torch.float64, device=lw.device)
    L.call("mb_cumsum_lw", L.ctx(), ptr(lw), lw.numel(), ptr(ctl.t), 1 if force else 0, ptr(cdf), stream())
    return cdf


def ancestors(cdf, mode, n_out=None, u=None, seed=0, step=0, gid0=0, ctl=None, out=None):
    L = _lib.get()
    n_out = cdf.numel() if n_out is None else int(n_out)
    anc = out if out is not None else torch.empty(n_out, dtype=torch.int32, device=cdf.device)
    L.call("mb_ancestors", L.ctx(), ptr(cdf), cdf.numel(), int(mode), ptr(u), int(seed), int(step), int(gid0),
           ptr(anc), n_out, ptr(ctl.t) if ctl is not None else None, stream())
    return anc


def gather_state(anc, src, n_out=None):
    """src: (ncols, ld) float32 SoA block -> (ncols, pad(n_out))"""
    L = _lib.get()
    n_out = anc.numel() if n_out is None else n_out
    dst = torch.empty((src.shape[0], _pad(n_out)), dtype=torch.float32, device=src.device)
    L.call("mb_gather_state", L.ctx(), ptr(anc), n_out, src.shape[0], ptr(src), src.stride(0), ptr(dst), dst.stride(0),
           stream())
    return dst


def quantile(v, q):
    L = _lib.get()
    out = torch.empty(3, dtype=torch.float64, device=v.device)
    L.call("mb_quantile", L.ctx(), ptr(v), v.numel(), float(q), ptr(out), stream())
    return out


def colstats(x, n):
    """x: (d, ld) SoA -> (mean (d,), var ddof=1 (d,)) float64 device tensors"""
    L = _lib.get()
    d = x.shape[0]
    mean = torch.empty(d, dtype=torch.float64, device=x.device)
    var = torch.empty(d, dtype=torch.float64, device=x.device)
    L.call("mb_colstats", L.ctx(), ptr(x), x.stride(0), int(n), d, ptr(mean), ptr(var), stream())
    return mean, var


def weighted_moments(x, n, lw, ctl):
    L = _lib.get()
    d = x.shape[0]
    mean = torch.empty(d, dtype=torch.float64, device=x.device)
    var = torch.empty(d, dtype=torch.float64, device=x.device)
    L.call("mb_weighted_moments", L.ctx(), ptr(x), x.stride(0), int(n), d, ptr(lw), ptr(ctl.t), ptr(mean), ptr(var),
           stream())
    return mean, var


def target_potential_grad(target, beta, X):
    """X: (n, d) row-major float32 -> (U (n,), G (n, d))"""
    L = _lib.get()
    n, d = X.shape
    U = torch.empty(n, dtype=torch.float32, device=X.device)
    G = torch.empty((n, d), dtype=torch.float32, device=X.device)
    L.call("mb_target_potential_grad", L.ctx(), C.byref(target), float(beta), ptr(X), n, ptr(U), ptr(G), stream())
    return U, G


def logistic_potential_grad(features, labels, prior_mean, prior_pscale, beta, X, variant=0):
    """Bayesian logistic regression (config C4): features (N, d), labels (N,), X (n, d) -> (U (n,), G (n, d))"""
    L = _lib.get()
    n, d = X.shape
    U = torch.empty(n, dtype=torch.float32, device=X.device)
    G = torch.empty((n, d), dtype=torch.float32, device=X.device)
    L.call("mb_logistic_potential_grad", L.ctx(), ptr(features), ptr(labels), features.shape[0], d, float(prior_mean),
           float(prior_pscale), float(beta), ptr(X), n, ptr(U), ptr(G), int(variant), stream())
    return U, G


class _Resampler:
    """scan -> (strata histogram) -> sorted-uniform ancestor search; every kernel is predicated on the device-side
    resample flag of the control block, so the host enqueues them unconditionally."""

    def _alloc_resampler(self):
        dev = _dev()
        self.anc = torch.zeros(self.n, dtype=torch.int32, device=dev)
        # systematic: fused exact-integer resampler (csrc/resample_fused.cu), no CDF in memory, 8-byte-aligned workspace
        self.rs_ws = torch.zeros((int(self.L.dll.mb_rs_workspace_bytes(self.n)) + 7) // 8, dtype=torch.int64, device=dev)
        self.cdf = None
        if self.resampling == _lib.RESAMPLE_MULTINOMIAL:
            self.cdf = torch.empty(self.n, dtype=torch.float64, device=dev)
            self.B = int(self.L.dll.mb_strata_count(self.n_total))
            self.hist = torch.zeros(self.B, dtype=torch.int32, device=dev)
            self.offsets = torch.zeros(self.B + 1, dtype=torch.int32, device=dev)

    def _cond_begin(self, st):
        """open the graph-conditional section (no-op outside capture); returns the stream to launch the body on"""
        if os.environ.get("MOCAT_B200_COND", "0") != "1":      # measured: the IF node costs more than 6 early-exit kernels
            return st
        body = C.c_void_p()
        self.L.call("mb_cond_begin", self.ctx, ptr(self.ctl.t), st, C.byref(body))
        return body

    def _cond_end(self, st):
        self.L.call("mb_cond_end", self.ctx, st)

    def _resample_kernels(self, st):
        L = self.L
        if self.resampling == _lib.RESAMPLE_SYSTEMATIC:
            L.call("mb_rs_tile_sums", self.ctx, ptr(self.rs_ws), ptr(self.lw), self.n, self.n, 1, ptr(self.ctl.t), 0, st)
            L.call("mb_rs_ancestors", self.ctx, ptr(self.rs_ws), ptr(self.lw), self.n, self.n, 1, ptr(self.ctl.t), 0, -1,
                   None, None, ptr(self.anc), st)
            return
        L.call("mb_cumsum_lw", self.ctx, ptr(self.lw), self.n, ptr(self.ctl.t), 0, ptr(self.cdf), st)
        if self.resampling == _lib.RESAMPLE_MULTINOMIAL:
            # clear = 0: allocated zeroed, and mb_ancestors_sorted zeroes the counts it consumed (no memset node per step)
            L.call("mb_strata_hist", self.ctx, self.n, self.gid0, self.B, self.seed, 0, ptr(self.ctl.t), ptr(self.hist), 0, st)
        L.call("mb_ancestors_sorted", self.ctx, ptr(self.cdf), self.n, None, self.resampling, ptr(self.hist),
               ptr(self.offsets), self.B, self.seed, 0, self.gid0, self.n_total, ptr(self.anc), self.n, ptr(self.ctl.t), st)


# ------------------------------------------------------------------------------------------- tempered SMC
class SMCEngine(_Resampler):
    """Device state + kernel sequence of one tempered-SMC population (transport/smc.py)."""

    def __init__(self, target, move, temper, n, seed, resampling=_lib.RESAMPLE_MULTINOMIAL, gid0=0, n_total=None,
                 keep_alpha=True, keep_prior_potential=True, schedule=None):
        self.L = _lib.get()
        self.ctx = self.L.ctx()
        self.target, self.move, self.temper = target, move, temper
        self.n, self.d, self.ld = int(n), int(target.dim), _pad(n)
        self.n_total = self.n if n_total is None else int(n_total)
        self.seed, self.gid0, self.resampling = int(seed), int(gid0), int(resampling)
        dev = _dev()
        self.xbuf = [torch.zeros((self.d, self.ld), dtype=torch.float32, device=dev) for _ in range(2)]
        self.cur = 0
        f = lambda: torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.lw, self.lik = f(), f()
        self.up = f() if keep_prior_potential else None
        self.alpha = f() if keep_alpha else None
        self._alloc_resampler()
        self.ctl = ControlBlock()
        self._schedule = None
        if schedule is not None:
            self._schedule = torch.as_tensor(np.asarray(schedule, np.float64), device=dev)
            self.temper.schedule = self._schedule.data_ptr()
            self.temper.schedule_len = self._schedule.numel()
        self.enqueued = 0
        self.use_graphs = os.environ.get("MOCAT_B200_GRAPHS", "1") != "0"
        self._graphs = [None, None]
        self._ws_gen = None
        self.comm = None          # mb_comm* when the population is sharded over several GPUs (parallel.py)
        self.rm = None            # (rm_stepsize, target, initial stepsize): Robbins-Monro adaptation on the device
        self.stepsize_hist = None

    def enable_rm(self, rm_stepsize, target):
        """RMMetropolisedSMCSampler (transport/smc.py:376-428): the stepsize becomes device-resident (ctl->aux1) and is
        adapted by mb_rm_adapt at the end of every update -- part of the captured step, no host round trip"""
        if self.comm is not None:
            raise _lib.MocatB200Error("the Robbins-Monro adaptation runs on one GPU")
        self.rm = (float(rm_stepsize), float(target), float(self.move.stepsize))
        self.stepsize_hist = torch.zeros(MB_HIST_MAX, dtype=torch.float64, device=self.lw.device)
        self.move.stepsize = -1.0                                     # the move kernel reads ctl->aux1

    def _rm_adapt(self, init):
        self.L.call("mb_rm_adapt", self.ctx, ptr(self.alpha), ptr(self.lw), self.n, self.rm[0], self.rm[1],
                    self.rm[2] if init else 0.0, ptr(self.ctl.t), ptr(self.stepsize_hist), stream())

    def _shard_ref(self):
        return None

    @property
    def x(self):
        return self.xbuf[self.cur]

    def _temper(self, advance):
        self.L.call("mb_temper_adapt", self.ctx, ptr(self.lw), ptr(self.lik), self.n, C.byref(self.temper),
                    1 if advance else 0, self.n_total * self.d, self.n_total, ptr(self.ctl.t), ptr(self.ctl.hist), self.comm,
                    stream())

    def startup(self, x0=None):
        """transport/smc.py:128-164 + 267-296.  x0: optional (n, d) host/device array of initial values."""
        if x0 is not None:
            if not isinstance(x0, torch.Tensor):
                # from_numpy keeps the caller's memory: a pinned host buffer is then copied by DMA (torch.as_tensor(...,
                # device=) would first clone it into pageable memory: 5.2 ms instead of 0.8 ms for 20 MB)
                x0 = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float32))
            x0 = x0.to(device=self.x.device, dtype=torch.float32, non_blocking=True)
            self.x[:, :self.n].copy_(x0.t())
        self.L.call("mb_smc_init", self.ctx, C.byref(self.target), ptr(self.x), self.ld, self.n, self.n_total,
                    0 if x0 is not None else 1, ptr(self.up), ptr(self.lik), ptr(self.lw), self.seed, self.gid0,
                    ptr(self.ctl.t), stream())
        self._temper(advance=False)
        if self.rm is not None:
            self._rm_adapt(init=True)                                   # smc.py:403: the stepsize of iteration 0
        self.enqueued = 0
        self._cur0 = self.cur

    def settle(self):
        """Synchronise and point `cur` at the buffer of the LAST REAL step.  update() flips the ping-pong index on every
        call, but once the device set `done` every kernel of a step exits early and writes nothing; the valid buffer is
        fixed by the device iteration count.  Returns the control block."""
        c = self.ctl.read()
        self.cur = (self._cur0 + int(c['iter'])) & 1
        return c

    def _enqueue(self, events=None):
        st = stream()
        if events:
            events[0].record()
        body = self._cond_begin(st)                                 # captured: body of a device-evaluated IF node
        try:
            self._resample_kernels(body)                            # RNG step comes from ctl->iter + 1 on the device
        finally:
            self._cond_end(st)
        if events:
            events[1].record()
        src, dst = self.xbuf[self.cur], self.xbuf[self.cur ^ 1]
        self.L.call("mb_smc_move", self.ctx, C.byref(self.target), C.byref(self.move), ptr(src), ptr(dst), self.ld,
                    self.n, ptr(self.anc), ptr(self.lw), ptr(self.up), ptr(self.lik), ptr(self.alpha), self.seed,
                    self.gid0, ptr(self.ctl.t), self._shard_ref(), st)
        if events:
            events[2].record()
        self._temper(advance=True)
        if self.rm is not None:
            self._rm_adapt(init=False)
        if events:
            events[3].record()

    def update(self, events=None):
        """enqueue one SMCSampler.update (transport/smc.py:73-99); fully asynchronous, predicated on device.
        After the first (plain) call the four kernels are replayed from a CUDA graph (one per ping-pong
        parity) so that a population step costs one host call.
        events: optional list of 4 torch.cuda.Event recorded before / between / after the kernel groups
        (forces the plain path)."""
        if events is None and self.use_graphs and self.enqueued >= 1:
            # graphs embed the addresses of the context's workspaces: another engine (a larger population, an SVGD
            # ensemble) may have made them grow since the capture
            gen = int(self.L.dll.mb_workspace_generation(self.ctx))
            if gen != self._ws_gen:
                self._graphs = [None, None]
                self._ws_gen = gen
            g = self._graphs[self.cur]
            if g is None:
                g = self._capture()
            if g is not None:
                g.replay()
                self.cur ^= 1
                self.enqueued += 1
                return
        self._enqueue(events)
        self.cur ^= 1
        self.enqueued += 1

    def _capture(self):
        try:
            # capture on a side stream without torch.cuda.graph()'s synchronize/gc/empty_cache preamble: the
            # step allocates nothing, so the plain begin/end pair is enough and costs ~100 us instead of ~10 ms
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            side = self._side_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                g.capture_begin(capture_error_mode="thread_local")
                try:
                    self._enqueue()
                finally:
                    g.capture_end()
            cur.wait_stream(side)
            self._graphs[self.cur] = g
            return g
        except Exception as exc:                         # capture unsupported: keep the plain launch path
            import warnings
            warnings.warn(f"mocat_b200: CUDA graph capture of the SMC step failed ({exc}); using plain launches")
            self.use_graphs = False
            torch.cuda.synchronize()
            return None

    def values(self):
        """(n, d) float32 device tensor view of the current particle values"""
        return self.x[:, :self.n].t()

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream()
        return self._side

    # -- engine pool: repeated runs of the same configuration reuse HBM buffers and captured graphs ----------
    _POOL = {}
    _POOL_MAX = 2

    @classmethod
    def acquire(cls, target, move, temper, n, seed, resampling=_lib.RESAMPLE_MULTINOMIAL, schedule=None):
        sched = None if schedule is None else np.asarray(schedule, np.float64)
        key = (bytes(target), bytes(move), temper.max_temperature, temper.ess_retain, temper.ess_resample, temper.tol,
               temper.max_search_iter, temper.max_iter, int(n), int(resampling),
               None if sched is None else sched.tobytes(), torch.cuda.current_device())
        if os.environ.get("MOCAT_B200_ENGINE_POOL", "1") == "0":
            return cls(target, move, temper, n, seed, resampling=resampling, schedule=schedule)
        eng = cls._POOL.pop(key, None)
        if eng is None:
            eng = cls(target, move, temper, n, seed, resampling=resampling, schedule=schedule)
        eng.seed = int(seed)
        cls._POOL[key] = eng                                       # most recently used last
        while len(cls._POOL) > cls._POOL_MAX:
            cls._POOL.pop(next(iter(cls._POOL)))
        return eng


# ------------------------------------------------------------------------------------------- bootstrap PF
class PFEngine(_Resampler):
    """Device state + kernel sequence of a bootstrap particle filter (ssm/filtering.py)."""
    _mocat_transient = True          # cdict: live-session member, not pickled

    def __init__(self, ssm, n, seed, ess_threshold=0.5, resampling=_lib.RESAMPLE_MULTINOMIAL, gid0=0, n_total=None):
        self.L = _lib.get()
        self.ctx = self.L.ctx()
        self.ssm = ssm
        self.n, self.d, self.ld = int(n), int(ssm.dim), _pad(n)
        self.n_total = self.n if n_total is None else int(n_total)
        self.seed, self.gid0, self.resampling = int(seed), int(gid0), int(resampling)
        self.ess_threshold = float(ess_threshold)
        dev = _dev()
        # Lorenz-96 runs on the reference's own ROW-MAJOR (n, d) layout (whole rows move through the TMA engine,
        # csrc/pf_l96.cu); the small dense models keep SoA columns
        self.rowmajor = int(ssm.kind) == _lib.SSM_LORENZ96
        self.xbuf = [self._alloc_x(dev) for _ in range(2)]
        self.cur = 0
        self._lw_full = torch.zeros(self.ld, dtype=torch.float32, device=dev)     # padded to a multiple of 32
        self.lw = self._lw_full[:self.n]
        self._alloc_resampler()
        self.ctl = ControlBlock()
        self.t = 0

    def _x_shape(self):
        return (self.ld, self.d) if self.rowmajor else (self.d, self.ld)

    def _alloc_x(self, dev):
        # row-major layout: the init kernel writes every row the step kernels ever touch, no memset of 16 GB needed
        return (torch.empty if self.rowmajor else torch.zeros)(self._x_shape(), dtype=torch.float32, device=dev)

    # -- engine pool: repeated filters of one configuration reuse the HBM buffers (2 x 16 GB at config C3) -------
    _POOL = {}
    generation = 0

    @classmethod
    def acquire(cls, ssm, n, seed, ess_threshold=0.5, resampling=_lib.RESAMPLE_MULTINOMIAL):
        """pooled engine; `generation` is bumped on every hand-out so that a result cdict of an EARLIER run can tell
        that its engine has been re-used (ssm._engine_of raises instead of continuing from foreign state)"""
        key = (int(ssm.kind), int(ssm.dim), int(n), int(resampling), torch.cuda.current_device())
        if os.environ.get("MOCAT_B200_ENGINE_POOL", "1") == "0":
            eng = cls(ssm, n, seed, ess_threshold=ess_threshold, resampling=resampling)
        else:
            eng = cls._POOL.get(key)
            if eng is None:
                cls._POOL.clear()                                  # one pooled filter: populations are large
                eng = cls(ssm, n, seed, ess_threshold=ess_threshold, resampling=resampling)
                cls._POOL[key] = eng
        eng.ssm, eng.seed, eng.ess_threshold = ssm, int(seed), float(ess_threshold)
        eng.generation += 1
        return eng

    @property
    def x(self):
        return self.xbuf[self.cur]

    def _comm(self):
        return None

    def _shard_ref(self):
        return None

    def init(self, y0):
        """initiate_particles (ssm/filtering.py:173-193).  y0: device float32 (dim_obs,)"""
        if self.rowmajor:
            self.L.call("mb_pf_l96_init", self.ctx, C.byref(self.ssm), ptr(self.x), self.n, self.n_total, ptr(y0),
                        ptr(self.lw), self.seed, self.gid0, self.ess_threshold, ptr(self.ctl.t), ptr(self.ctl.hist),
                        self._comm(), stream())
        else:
            self.L.call("mb_pf_init", self.ctx, C.byref(self.ssm), ptr(self.x), self.ld, self.n, self.n_total, ptr(y0),
                        ptr(self.lw), self.seed, self.gid0, self.ess_threshold, ptr(self.ctl.t), ptr(self.ctl.hist),
                        self._comm(), stream())
        self.t = 0

    def _step_kernel(self, y, st):
        src, dst = self.xbuf[self.cur], self.xbuf[self.cur ^ 1]
        if self.rowmajor:
            self.L.call("mb_pf_l96_step", self.ctx, C.byref(self.ssm), ptr(src), ptr(dst), self.n, self.n_total,
                        ptr(self.anc), ptr(y), ptr(self.lw), self.seed, self.t, self.gid0, self.ess_threshold,
                        ptr(self.ctl.t), ptr(self.ctl.hist), self._shard_ref(), self._comm(), st)
        else:
            self.L.call("mb_pf_step", self.ctx, C.byref(self.ssm), ptr(src), ptr(dst), self.ld, self.n, self.n_total,
                        ptr(self.anc), ptr(y), ptr(self.lw), self.seed, self.t, self.gid0, self.ess_threshold,
                        ptr(self.ctl.t), ptr(self.ctl.hist), self._shard_ref(), self._comm(), st)
        self.cur ^= 1

    def step(self, y):
        """one body of the scan in run_particle_filter_for_marginals (ssm/filtering.py:280-311)"""
        st = stream()
        self.t += 1
        if int(getattr(self.ssm, 'proposal', 0)) == _lib.PROPOSAL_ENKF:
            return self._enkf_step(y, st)
        self._resample_kernels(st)
        self._step_kernel(y, st)

    def _enkf_step(self, y, st):
        """EnsembleKalmanFilter.propose_and_intermediate_weight_vectorised (ssm/nonlinear_gaussian.py:325-350): forecast
        f(x) + q z by the step kernel (weights are equal, so it never resamples), then the analysis update in place"""
        if self._comm() is not None:
            raise _lib.MocatB200Error("the ensemble Kalman filter runs on one GPU (the ensemble covariance is not sharded)")
        if not hasattr(self, '_enkf_gain'):
            dev = _dev()
            self._enkf_gain = torch.empty((self.d, self.d), dtype=torch.float32, device=dev)
            self._enkf_mean = torch.empty(self.d, dtype=torch.float64, device=dev)
            self._enkf_cov = torch.empty((self.d, self.d), dtype=torch.float64, device=dev)
        self._step_kernel(y, st)
        self.L.call("mb_enkf_analysis", self.ctx, C.byref(self.ssm), ptr(self.x), self.n, ptr(y), ptr(self.lw), self.seed,
                    self.t, self.gid0, ptr(self.ctl.t), ptr(self.ctl.hist), ptr(self._enkf_gain), ptr(self._enkf_mean),
                    ptr(self._enkf_cov), st)

    def values(self):
        """(n, d) float32 device tensor view of the current particle values"""
        if self.rowmajor:
            return self.x[:self.n]
        return self.x[:, :self.n].t()

    def gather_current(self):
        """x <- x[anc] into the other buffer (resample_particles, ssm/filtering.py:202-217)"""
        src, dst = self.xbuf[self.cur], self.xbuf[self.cur ^ 1]
        if self.rowmajor:
            self.L.call("mb_gather_rows", self.ctx, ptr(self.anc), self.n, self.d, ptr(src), self.n, ptr(dst), 1, stream())
        else:
            self.L.call("mb_gather_state", self.ctx, ptr(self.anc), self.n, self.d, ptr(src), self.ld, ptr(dst), self.ld,
                        stream())
        self.cur ^= 1

    def moment_sums(self, shift, out=None):
        """un-normalised weighted sums of THIS shard, (1 + 2d,) float64: sum e, sum e (x - shift), sum e (x - shift)^2 with
        e = exp(lw - global max); the ranks' records add up to the moments of the whole population"""
        if not self.rowmajor:
            raise _lib.MocatB200Error("moment_sums: row-major (Lorenz-96) populations only")
        sums = out if out is not None else torch.empty(1 + 2 * self.d, dtype=torch.float64, device=self.x.device)
        self.L.call("mb_weighted_moment_sums_rows", self.ctx, ptr(self.x), self.n, self.d, ptr(self.lw), ptr(self.ctl.t),
                    ptr(shift), ptr(sums), stream())
        return sums

    def moments(self, out=None):
        """weighted mean / variance of every coordinate under the current weights (device float64 (d,) tensors, or the
        two rows of `out` (2, d))"""
        if self.rowmajor:
            if out is not None:
                mean, var = out[0], out[1]
            else:
                mean = torch.empty(self.d, dtype=torch.float64, device=self.x.device)
                var = torch.empty(self.d, dtype=torch.float64, device=self.x.device)
            self.L.call("mb_weighted_moments_rows", self.ctx, ptr(self.x), self.n, self.d, ptr(self.lw), ptr(self.ctl.t),
                        ptr(mean), ptr(var), stream())
            return mean, var
        mean, var = weighted_moments(self.x, self.n, self.lw, self.ctl)
        if out is not None:
            out[0], out[1] = mean, var
        return mean, var


# ------------------------------------------------------------------------------------------- SMC-ABC
class ABCEngine(_Resampler):
    """Device state + kernel sequence of SMC-ABC on the g-and-k model (abc/smc.py, abc/mcmc.py)."""

    def __init__(self, gk, n, seed, mcmc_steps=1, max_iter=10000, ess_retain=0.9, ess_resample=0.5,
                 termination_alpha=0.01, threshold_schedule=None, resampling=_lib.RESAMPLE_MULTINOMIAL, gid0=0,
                 n_total=None):
        self.L = _lib.get()
        self.ctx = self.L.ctx()
        self.gk = gk
        self.n, self.d, self.ld = int(n), 4, _pad(n)
        self.n_total = self.n if n_total is None else int(n_total)
        self.seed, self.gid0, self.resampling = int(seed), int(gid0), int(resampling)
        self.mcmc_steps, self.max_iter = int(mcmc_steps), int(max_iter)
        self.ess_retain, self.ess_resample, self.termination_alpha = float(ess_retain), float(ess_resample), \
            float(termination_alpha)
        dev = _dev()
        f = lambda: torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.xbuf = [torch.zeros((self.d, self.ld), dtype=torch.float32, device=dev) for _ in range(2)]
        self.upbuf, self.distbuf, self.alphabuf = [f(), f()], [f(), f()], [f(), f()]
        self.cur = 0
        self.lw = f()
        self.stepsize = torch.ones(self.d, dtype=torch.float32, device=dev)
        self._alloc_resampler()
        self.ctl = ControlBlock()
        self._schedule = None
        if threshold_schedule is not None:
            self._schedule = torch.as_tensor(np.asarray(threshold_schedule, np.float64), device=dev)
            self.max_iter = self._schedule.numel()
        self.enqueued = 0

    x = property(lambda self: self.xbuf[self.cur])
    up = property(lambda self: self.upbuf[self.cur])
    dist = property(lambda self: self.distbuf[self.cur])
    alpha = property(lambda self: self.alphabuf[self.cur])

    def _adapt(self, advance):
        self.L.call("mb_abc_adapt", self.ctx, ptr(self.x), self.ld, self.n, self.n_total, self.d, ptr(self.dist),
                    ptr(self.lw), ptr(self.alpha), ptr(self.stepsize), self.ess_retain, self.ess_resample,
                    self.termination_alpha, self.max_iter, ptr(self._schedule), 1 if advance else 0, ptr(self.ctl.t),
                    ptr(self.ctl.hist), stream())

    def startup(self, x0=None):
        if x0 is not None:
            if not isinstance(x0, torch.Tensor):
                # from_numpy keeps the caller's memory: a pinned host buffer is then copied by DMA (torch.as_tensor(...,
                # device=) would first clone it into pageable memory: 5.2 ms instead of 0.8 ms for 20 MB)
                x0 = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float32))
            x0 = x0.to(device=self.x.device, dtype=torch.float32, non_blocking=True)
            self.x[:, :self.n].copy_(x0.t())
        self.L.call("mb_abc_init", self.ctx, C.byref(self.gk), ptr(self.x), self.ld, self.n, self.n_total,
                    0 if x0 is not None else 1, ptr(self.up), ptr(self.dist), ptr(self.lw), ptr(self.alpha), self.seed,
                    self.gid0, ptr(self.ctl.t), stream())
        self._adapt(advance=False)
        self.enqueued = 0
        self._cur0 = self.cur

    def settle(self):
        """see SMCEngine.settle: x / up / dist / alpha are all double buffered"""
        c = self.ctl.read()
        self.cur = (self._cur0 + int(c['iter'])) & 1
        return c

    def update(self):
        st = stream()
        self._resample_kernels(st)
        c, o = self.cur, self.cur ^ 1
        self.L.call("mb_abc_move", self.ctx, C.byref(self.gk), self.mcmc_steps, ptr(self.xbuf[c]), ptr(self.xbuf[o]),
                    self.ld, self.n, ptr(self.anc), ptr(self.upbuf[c]), ptr(self.upbuf[o]), ptr(self.distbuf[c]),
                    ptr(self.distbuf[o]), ptr(self.lw), ptr(self.alphabuf[c]), ptr(self.alphabuf[o]),
                    ptr(self.stepsize), self.seed, self.gid0, ptr(self.ctl.t), st)
        self.cur = o
        self._adapt(advance=True)
        self.enqueued += 1

    def values(self):
        return self.x[:, :self.n].t()


# ------------------------------------------------------------------------------------------- SVGD
TENSOR_CORE_MIN_N = 2048


def interaction_variant(n, d, variant=None):
    """0 = exact fp32 SIMT kernels, 1 = tcgen05 kernels (bf16 operands, fp32 accumulation; |err| ~ 3e-3 of max|phi|).
    None picks the tensor cores for ensembles that are large enough to fill them (n >= 2048, d <= 60)."""
    if variant is None:
        return 1 if (n >= TENSOR_CORE_MIN_N and d <= 60) else 0
    return int(variant)


def svgd_phi_rows(X, G, bandwidth, row_begin, row_end):
    """rows [row_begin, row_end) of phi (tensor-core variant; X, G the whole ensemble) -> (row_end - row_begin, d)"""
    L = _lib.get()
    n, d = X.shape
    phi = torch.empty_like(X)
    L.call("mb_svgd_phi_rows", L.ctx(), ptr(X), ptr(G), n, d, ptr(bandwidth), ptr(phi), int(row_begin),
           int(row_end - row_begin), stream())
    return phi[row_begin:row_end]


def svgd_phi(X, G, bandwidth, variant=None):
    """X, G: (n, d) row-major float32; bandwidth: device float32 scalar tensor -> phi (n, d)"""
    L = _lib.get()
    n, d = X.shape
    variant = interaction_variant(n, d, variant)
    phi = torch.empty_like(X)
    L.call("mb_svgd_phi", L.ctx(), ptr(X), ptr(G), n, d, ptr(bandwidth), ptr(phi), int(variant), stream())
    return phi


def ensemble_shard(n, d, variant=None):
    """(rank, world, row_begin, row_end) of this process's share of an SVGD ensemble, or None when the ensemble is not
    sharded: under torchrun the tensor-core kernels split the n x n interaction by rows of 128 over the ranks
    (SURVEY 8e item 5); every rank holds the whole (n, d) ensemble, all-gathered once per iteration."""
    try:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
    except Exception:
        return None
    world, rank = dist.get_world_size(), dist.get_rank()
    if interaction_variant(n, d, variant) != 1 or n % (128 * world):
        return None
    per = n // world
    return rank, world, rank * per, (rank + 1) * per


def pairdist_bandwidth(X, mode, variant=None):
    """mode 'median' | 'mean' -> device float32 (1,) tensor h = stat(D)/sqrt(2 log n)  (kernels.py:220-229)"""
    L = _lib.get()
    n, d = X.shape
    h = torch.empty(1, dtype=torch.float32, device=X.device)
    sh = ensemble_shard(n, d, variant) if os.environ.get("MOCAT_B200_SVGD_SHARD", "1") != "0" else None
    if sh is not None:
        # this rank's share of the (symmetric) distance matrix, all-reduce of the counters, the same finish everywhere
        import torch.distributed as dist
        rank, world = sh[0], sh[1]
        m = 0 if mode == "median" else 1
        acc = torch.zeros(_lib.MB_PAIRDIST_ACC_BYTES // 4, dtype=torch.int32, device=X.device)
        L.call("mb_pairdist_partial", L.ctx(), ptr(X), n, d, m, rank, world, ptr(acc), stream())
        if m == 0:
            dist.all_reduce(acc[256:])                                 # 2048 histogram counters (< 2^31 in total)
            dist.all_reduce(acc[16:18].view(torch.int64))              # entries below the bracket
        else:
            dist.all_reduce(acc[18:20].view(torch.float64))            # sum of distances
        L.call("mb_pairdist_finish", L.ctx(), m, n, ptr(acc), ptr(h), stream())
        return h
    L.call("mb_pairdist_bandwidth", L.ctx(), ptr(X), n, d, 0 if mode == "median" else 1, ptr(h),
           interaction_variant(n, d, variant), stream())
    return h


def adagrad_step(X, gsq, mom, phi, step, momentum=0.9):
    L = _lib.get()
    L.call("mb_adagrad", L.ctx(), ptr(X), ptr(gsq), ptr(mom), ptr(phi), X.numel(), float(step), float(momentum),
           stream())
