"""Online (fixed-lag) particle smoothing: host mirror of mocat/src/ssm/online_smoothing.py.

`propagate_particle_smoother` keeps the reference signature (online_smoothing.py:364-386) and both of its branches:
the particle-filter smoother (`backward_sim=False`, :211-282) and the backward-simulation smoother (`backward_sim=True`,
:285-361).  The n x n work of a step -- the stitching contraction `full_stitch` (:21-44) and, for the second branch, the
backward simulation over the lag window -- runs on the device (mb_stitch_sample, mb_transition_potential,
mb_backward_sample: csrc/backward.cu); proposals and weights are the filter engine's step kernel.  The stored
trajectories (T, n, d) live in host memory as in the rest of this API (ssm.py): a step re-indexes only the `lag` most
recent slices.  The rejection variant of the stitching (maximum_rejections > 0, :60-164) draws from the same law with
data-dependent work; it is served by the full contraction and `num_transition_evals` reports the n^2 evaluations made.
The smoother is STATEFUL like the filter API: the particle cdict carries the live device engine.
"""
import ctypes as C

import numpy as np

from . import _lib
from .core import cdict, key_to_seed
from . import ssm as _ssm


def _torch():
    import torch
    return torch


def _dev():
    torch = _torch()
    return torch.device("cuda", torch.cuda.current_device())


def transition_potential(ssm_scenario, x_previous, t_previous, x_new, t_new):
    """StateSpaceModel.transition_potential for matched rows (linear_gaussian.py:73-84, nonlinear_gaussian.py:98-105)
    evaluated on the device -> host float32 (n,)"""
    torch = _torch()
    dev = _dev()
    x0 = torch.as_tensor(np.ascontiguousarray(np.atleast_2d(x_previous), np.float32), device=dev)
    x1 = torch.as_tensor(np.ascontiguousarray(np.atleast_2d(x_new), np.float32), device=dev)
    n, d = x0.shape
    work = torch.empty((n, d), dtype=torch.float32, device=dev)
    pot = torch.empty(n, dtype=torch.float32, device=dev)
    dt = float(t_new) - float(t_previous)
    L = _lib.get()
    L.call("mb_transition_potential", L.ctx(), C.byref(ssm_scenario._ssm(dt)), dt, _lib.ptr(x0), _lib.ptr(x1), n,
           _lib.ptr(work), _lib.ptr(pot), _lib.stream())
    return pot.cpu().numpy()


def _stitch_indices(ssm_scenario, x0_fixed, t, x1_vary, tplus1, log_weight, seed, step):
    """full_stitch (:34-44) on device tensors -> device int32 indices"""
    torch = _torch()
    n_s, d = x0_fixed.shape
    n_c = x1_vary.shape[0]
    work = torch.empty(((n_s + n_c), d), dtype=torch.float32, device=x0_fixed.device)
    idx = torch.empty(n_s, dtype=torch.int32, device=x0_fixed.device)
    dt = float(tplus1) - float(t)
    L = _lib.get()
    L.call("mb_stitch_sample", L.ctx(), C.byref(ssm_scenario._ssm(dt)), dt, _lib.ptr(x0_fixed), n_s, _lib.ptr(x1_vary),
           _lib.ptr(log_weight), n_c, _lib.ptr(work), int(seed), int(step), _lib.ptr(idx), _lib.stream())
    return idx


def fixed_lag_stitching(ssm_scenario, early_block, t, recent_block, recent_block_log_weight, tplus1, random_key,
                        maximum_rejections=0, init_bound_param=0., bound_inflation=1.01, step=None, return_indices=False):
    """online_smoothing.py:167-207.  early_block (s + 1, n, d), recent_block (lag + 1, n, d) host arrays; returns the
    stitched (s + 1 + lag, n, d) block and the number of transition evaluations (n^2: the full contraction)."""
    torch = _torch()
    dev = _dev()
    early_block, recent_block = np.asarray(early_block, np.float32), np.asarray(recent_block, np.float32)
    n, d = recent_block.shape[1:]
    x0_fixed = torch.as_tensor(np.ascontiguousarray(early_block[-1]), device=dev)
    x0_vary = torch.as_tensor(np.ascontiguousarray(recent_block[0]), device=dev)
    x1_vary = torch.as_tensor(np.ascontiguousarray(recent_block[1]), device=dev)
    lw = torch.as_tensor(np.ascontiguousarray(recent_block_log_weight, np.float32), device=dev)
    dt = float(tplus1) - float(t)
    L = _lib.get()
    work = torch.empty((n, d), dtype=torch.float32, device=dev)
    pot = torch.empty(n, dtype=torch.float32, device=dev)
    L.call("mb_transition_potential", L.ctx(), C.byref(ssm_scenario._ssm(dt)), dt, _lib.ptr(x0_vary), _lib.ptr(x1_vary), n,
           _lib.ptr(work), _lib.ptr(pot), _lib.stream())
    non_interacting = lw + pot                                                        # :182-184
    step = len(early_block) - 1 if step is None else step
    idx = _stitch_indices(ssm_scenario, x0_fixed, t, x1_vary, tplus1, non_interacting, key_to_seed(random_key), step)
    inds = idx.cpu().numpy().astype(np.int64)
    stitched = np.append(early_block, recent_block[1:, inds], axis=0)                 # :207
    if return_indices:
        return stitched, early_block.shape[1] ** 2, idx
    return stitched, early_block.shape[1] ** 2


def _zero_weights(eng):
    """log_weight = zeros(n) on the device with the control block's weight summaries to match"""
    eng._lw_full.zero_()
    c = eng.ctl.read()
    logn = float(np.log(eng.n_total))
    c['wmax'], c['s1'], c['s2'] = 0.0, float(eng.n_total), float(eng.n_total)
    c['lse'] = c['lse2'] = c['log_ess'] = logn
    c['ess'], c['resample'] = float(eng.n_total), 0
    eng.ctl.write(c)


def _ess(log_weight):
    lw = np.asarray(log_weight, np.float64)
    m = np.max(lw)
    w = np.exp(lw - m)
    return float(w.sum() ** 2 / np.sum(w * w))


def propagate_particle_smoother_pf(ssm_scenario, particle_filter, particles, y_new, t_new, random_key, lag,
                                   maximum_rejections=0, init_bound_param=0., bound_inflation=1.01):
    """online_smoothing.py:211-282: resample the stored trajectories if they are weighted, propose and weight the new
    time slice (the filter engine's step kernel without resampling), and -- once more than `lag` slices are stored --
    stitch the `lag` most recent slices onto the fixed early block."""
    torch = _torch()
    _ssm._check_filter(particle_filter)
    eng = _ssm._engine_of(particles)
    if particles.value is None:
        raise _lib.MocatB200Error("propagate_particle_smoother needs the stored trajectories (value (T, n, d) on the host)")
    n = particles.value.shape[1]
    seed = key_to_seed(random_key)
    lw_last = np.atleast_2d(particles.log_weight)[-1]
    if _ess(lw_last) < n - 1e-3:                                                      # :227-231
        out = _ssm.resample_particles(particles, random_key, True)
    else:
        out = particles.copy()
    _zero_weights(eng)
    y_new = np.atleast_1d(np.asarray(y_new, np.float32))
    t_prev = float(np.atleast_1d(out.t)[-1])
    eng.ssm = _ssm._device_ssm(ssm_scenario, particle_filter, float(t_new) - t_prev)
    eng.step(torch.as_tensor(y_new, device="cuda"))                                   # :238-241 (resample flag is 0)
    c = eng.ctl.read()
    x_new = eng.values().cpu().numpy()
    out.log_weight = eng.lw.cpu().numpy()
    out.value = np.append(out.value, x_new[None], axis=0)
    out.y = np.append(np.atleast_2d(out.y), y_new[None], axis=0)
    out.t = np.append(out.t, t_new)
    out.ess = c['ess']
    nte = getattr(particles, 'num_transition_evals', np.array(0))
    len_t = len(out.t)
    stitch_ind_min_1, stitch_ind = len_t - lag - 1, len_t - lag
    num_transition_evals = 0
    if stitch_ind_min_1 >= 0:                                                         # :266-277
        out.value, num_transition_evals, idx = fixed_lag_stitching(
            ssm_scenario, out.value[:stitch_ind_min_1 + 1], out.t[stitch_ind_min_1], out.value[stitch_ind_min_1:],
            out.log_weight, out.t[stitch_ind], seed, maximum_rejections, init_bound_param, bound_inflation,
            step=len_t - 1, return_indices=True)
        eng.anc[:n].copy_(idx)                                                        # the device population follows
        eng.gather_current()
        _zero_weights(eng)
        out.log_weight = np.zeros(n, np.float32)                                      # :281
    out.num_transition_evals = np.append(nte, num_transition_evals)
    return out


def propagate_particle_smoother_bs(ssm_scenario, particle_filter, particles, y_new, t_new, random_key, lag,
                                   ess_threshold=0.5, maximum_rejections=0, init_bound_param=0., bound_inflation=1.01):
    """online_smoothing.py:285-361: advance the marginal filter, backward-simulate over the lag window (or over
    everything while fewer than lag + 1 slices exist) and stitch the window onto the fixed early block."""
    if particles.value is None:
        raise _lib.MocatB200Error("propagate_particle_smoother needs the stored trajectories (value (T, n, d) on the host)")
    n = particles.value.shape[1]
    seed = key_to_seed(random_key)
    out = particles.copy()
    nte = getattr(particles, 'num_transition_evals', np.array(0))
    if not hasattr(particles, 'marginal_filter'):                                     # :303-308
        out.marginal_filter = cdict(value=particles.value, log_weight=np.atleast_2d(particles.log_weight),
                                    y=np.atleast_2d(particles.y), t=np.atleast_1d(particles.t),
                                    ess=np.atleast_1d(particles.ess),
                                    log_norm_constant=np.atleast_1d(getattr(particles, 'log_norm_constant', 0.0)),
                                    mean=particles.mean, var=particles.var,
                                    engine=_ssm._engine_of(particles), engine_generation=particles.engine_generation)
    out.marginal_filter = _ssm.propagate_particle_filter(ssm_scenario, particle_filter, out.marginal_filter, y_new, t_new,
                                                         seed + 1, ess_threshold, False)                    # :315-316
    y_new = np.atleast_1d(np.asarray(y_new, np.float32))
    out.y = np.append(np.atleast_2d(particles.y), y_new[None], axis=0)
    out.t = np.append(particles.t, t_new)
    out.log_weight = np.zeros(n, np.float32)
    out.ess = out.marginal_filter.ess[-1]
    len_t = len(out.t)
    stitch_ind_min_1, stitch_ind = len_t - lag - 1, len_t - lag
    mf = out.marginal_filter
    if stitch_ind_min_1 >= 0:                                                         # back_sim_and_stitch, :336-355
        window = cdict(value=mf.value[stitch_ind_min_1:], log_weight=mf.log_weight[stitch_ind_min_1:],
                       t=mf.t[stitch_ind_min_1:])
        bs = _ssm.backward_simulation(ssm_scenario, window, seed + 2, n, maximum_rejections, init_bound_param,
                                      bound_inflation)
        out.value, stitch_nte = fixed_lag_stitching(ssm_scenario, particles.value[:stitch_ind_min_1 + 1],
                                                    out.t[stitch_ind_min_1], bs.value, np.zeros(n, np.float32),
                                                    out.t[stitch_ind], seed, maximum_rejections, init_bound_param,
                                                    bound_inflation, step=len_t - 1)
        num_transition_evals = stitch_nte + bs.num_transition_evals.sum()
    else:                                                                             # back_sim_only, :326-334
        bs = _ssm.backward_simulation(ssm_scenario, cdict(value=mf.value, log_weight=mf.log_weight, t=mf.t), seed + 2, n,
                                      maximum_rejections, init_bound_param, bound_inflation)
        out.value, num_transition_evals = bs.value, bs.num_transition_evals.sum()
    out.num_transition_evals = np.append(nte, num_transition_evals)
    return out


def propagate_particle_smoother(ssm_scenario, particle_filter, particles, y_new, t_new, random_key, lag, backward_sim=True,
                                ess_threshold=0.5, maximum_rejections=0, init_bound_param=0., bound_inflation=1.01):
    """online_smoothing.py:364-386"""
    if backward_sim:
        return propagate_particle_smoother_bs(ssm_scenario, particle_filter, particles, y_new, t_new, random_key, lag,
                                              ess_threshold, maximum_rejections, init_bound_param, bound_inflation)
    return propagate_particle_smoother_pf(ssm_scenario, particle_filter, particles, y_new, t_new, random_key, lag,
                                          maximum_rejections, init_bound_param, bound_inflation)
