"""Streaming history and checkpoints of a device population.

The reference returns the whole stacked run -- `(T, n, d)` for a filter (ssm/filtering.py:317-322), `(iterations, n, d)`
for a sampler (sample.py:138-146 via utils.while_loop_stacked) -- and persists it with pickle (core.py:91-121).  At
device scale that history does not fit anywhere, so the device path keeps per-step moments / ESS / evidence always and
offers two opt-in tools:

* `HistoryStream`: a THINNED history streamed to pinned host memory while the run continues.  Every kept step is one
  device -> host copy on a side stream, ordered after the step that produced the population (event) and before the step
  that overwrites its ping-pong buffer (event the other way) -- no host synchronisation inside the run.
* `save_checkpoint` / `load_checkpoint`: the state a filter needs to continue (population, log-weights, control
  block, time index, seed, model POD) in one `.cdict` pickle.  Philox counters are keyed on (particle, time index), so a
  run continued from a checkpoint is bit-identical to the uninterrupted run.
"""
import ctypes as C

import numpy as np

from . import _lib
from .core import cdict, save_cdict, load_cdict


class HistoryStream:
    """Thinned population history in pinned host memory.

    every: keep steps 0, every, 2 every, ... (and whatever step `push(..., force=True)` is called on)
    capacity: number of records the pinned arrays hold; shape: (n, d) of one record"""

    def __init__(self, shape, every, capacity, with_weights=True):
        import torch
        self.torch = torch
        self.every, self.capacity = max(1, int(every)), int(capacity)
        self.values = torch.empty((self.capacity,) + tuple(shape), dtype=torch.float32).pin_memory()
        self.log_weight = torch.empty((self.capacity, shape[0]), dtype=torch.float32).pin_memory() if with_weights else None
        self.index = []                                   # step index of every record
        self.side = torch.cuda.Stream()
        self._pending = []                                # (record event, step) of copies whose source buffer is still live
        self.bytes_streamed = 0

    def wants(self, step):
        return step % self.every == 0 and len(self.index) < self.capacity

    def push(self, step, values, log_weight=None, force=False):
        """enqueue the copy of this step's population (device tensors) if the step is kept; returns True if it was"""
        torch = self.torch
        if not (force or self.wants(step)) or len(self.index) >= self.capacity or (self.index and self.index[-1] == step):
            return False
        k = len(self.index)
        done = torch.cuda.Event()
        ready = torch.cuda.current_stream().record_event()          # the step that produced `values` has been enqueued
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            self.values[k].copy_(values, non_blocking=True)
            if self.log_weight is not None and log_weight is not None:
                self.log_weight[k].copy_(log_weight, non_blocking=True)
            done.record(self.side)
        self._pending.append((done, step))
        self.index.append(int(step))
        self.bytes_streamed += values.numel() * 4 + (log_weight.numel() * 4 if log_weight is not None else 0)
        return True

    def before_overwrite(self, step):
        """call before enqueuing step `step`: a ping-pong engine writes the buffer that held step - 2 (and, in place,
        the log-weights of step - 1), so the compute stream waits for the copies of every earlier step"""
        cur = self.torch.cuda.current_stream()
        keep = []
        for ev, s in self._pending:
            if s < step:
                cur.wait_event(ev)
            else:
                keep.append((ev, s))
        self._pending = keep

    def finish(self):
        """synchronise the side stream; returns (values (K, n, d), log_weight (K, n) or None, index (K,)) as NumPy views
        of the pinned buffers"""
        self.side.synchronize()
        k = len(self.index)
        lw = self.log_weight[:k].numpy() if self.log_weight is not None else None
        return self.values[:k].numpy(), lw, np.asarray(self.index, dtype=np.int64)


def save_checkpoint(particles, path, overwrite=False):
    """Persist what a filter needs to CONTINUE (not the history): the latest population, log-weights, control block, time
    index, seed, thresholds and the model POD, as a `.cdict` pickle (core.py:91-108).  `particles`: a result of
    initiate_particles / propagate_particle_filter / run_particle_filter_for_marginals on ONE GPU."""
    eng = getattr(particles, 'engine', None)
    if eng is None:
        raise _lib.MocatB200Error("save_checkpoint needs a particle cdict with a live device engine")
    if eng._comm() is not None:
        raise _lib.MocatB200Error("save_checkpoint: sharded populations are not checkpointed (save per rank with values())")
    ck = cdict(kind='mocat_b200.pf_checkpoint', version=1,
               value=eng.values().cpu().numpy(), log_weight=eng.lw.cpu().numpy(),
               control=np.frombuffer(eng.ctl.read().tobytes(), dtype=np.uint8).copy(),
               t_index=int(eng.t), seed=int(eng.seed), gid0=int(eng.gid0), n=int(eng.n), n_total=int(eng.n_total),
               ess_threshold=float(eng.ess_threshold), resampling=int(eng.resampling),
               ssm=np.frombuffer(bytes(eng.ssm), dtype=np.uint8).copy(),
               t=np.asarray(particles.t), y=np.asarray(particles.y),
               ess=np.asarray(particles.ess), log_norm_constant=np.asarray(particles.log_norm_constant))
    for k in ('mean', 'var', 'resampled'):
        if getattr(particles, k, None) is not None:
            setattr(ck, k, np.asarray(getattr(particles, k)))
    save_cdict(ck, path, overwrite)


def load_checkpoint(path):
    """Rebuild the device engine from a checkpoint; returns a particle cdict that propagate_particle_filter /
    run_particle_filter_for_marginals(initial_sample=...) continue exactly where the saved run stopped."""
    import torch
    from . import engine
    ck = load_cdict(path)
    if getattr(ck, 'kind', None) != 'mocat_b200.pf_checkpoint':
        raise _lib.MocatB200Error(f"{path}: not a particle-filter checkpoint")
    ssm = _lib.SSM.from_buffer_copy(np.asarray(ck.ssm, np.uint8).tobytes())
    eng = engine.PFEngine.acquire(ssm, ck.n, ck.seed, ess_threshold=ck.ess_threshold, resampling=ck.resampling)
    if ck.n_total != ck.n or ck.gid0 != 0:
        raise _lib.MocatB200Error("load_checkpoint: single-GPU checkpoints only")
    x = torch.as_tensor(np.asarray(ck.value, np.float32), device="cuda")
    if eng.rowmajor:
        eng.x[:eng.n].copy_(x)
    else:
        eng.x[:, :eng.n].copy_(x.t())
    eng._lw_full.fill_(float('-inf'))                    # padding beyond n carries no weight
    eng.lw.copy_(torch.as_tensor(np.asarray(ck.log_weight, np.float32), device="cuda"))
    eng.ctl.write(np.frombuffer(np.asarray(ck.control, np.uint8).tobytes(), dtype=_lib.CONTROL_DTYPE)[0])
    eng.t = int(ck.t_index)
    out = cdict(t=np.asarray(ck.t), y=np.asarray(ck.y), ess=np.asarray(ck.ess),
                log_norm_constant=np.asarray(ck.log_norm_constant), engine=eng, engine_generation=eng.generation)
    for k in ('mean', 'var', 'resampled'):
        if getattr(ck, k, None) is not None:
            setattr(out, k, np.asarray(getattr(ck, k)))
    out.value, out.log_weight = ck.value[None], ck.log_weight[None]
    return out
