"""Sharded populations over the GPUs of one NVSwitch domain.

One process per GPU (torchrun).  Rank r owns the contiguous global particle range [r*n_local, (r+1)*n_local);
Philox is keyed on the global particle id, the fp64 CDF arithmetic is exact, so a population sharded over P GPUs
produces the same ancestors and the same particles as on one GPU.  The only communication of a population step:

  * the (max, sum, sumsq) LSE/ESS triple of every rank        -> exchanged INSIDE the kernels (tempering search:
    once per regula-falsi evaluation; particle filter: once per step) through peer-mapped mailboxes (csrc/comm.cuh);
  * the ranks' quantised weight totals (one fp64 each)          -> `mb_comm_allgather`, gives every rank's CDF offset;
  * the ancestors' state                                        -> read directly from the owning rank's HBM over
    NVLink by the move kernel's fused gather (peer pointers), i.e. the redistribution all-to-all is not a
    separate pass.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only as plumbing: exchanging the 64-byte CUDA
IPC handles at start-up and the timing barriers of bench.py.
"""
import ctypes as C

import numpy as np

from . import _lib


def shard_range(n_total, rank, world):
    """contiguous equal shards; n_total must be divisible by world (the engines require equal shards)"""
    if n_total % world:
        raise _lib.MocatB200Error(f"n_total={n_total} must be divisible by the number of ranks {world}")
    n_local = n_total // world
    return rank * n_local, n_local


def owner_of(global_index, n_local):
    """(owner rank, local index) of a global particle index (host mirror of the kernels' arithmetic)"""
    g = np.asarray(global_index, dtype=np.int64)
    return g // n_local, g % n_local


def global_cdf_offsets(totals):
    """exclusive prefix of the ranks' exact weight totals = CDF offset of every rank (fp64, exact)"""
    t = np.asarray(totals, dtype=np.float64)
    return np.concatenate([[0.0], np.cumsum(t)])


def all_gather_bytes(payload: bytes, group=None):
    """gather one bytes object from every rank (works on nccl and gloo process groups)"""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [payload]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, payload, group=group)
    return out


class _RawCuda:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class ShardContext:
    """rank / world + the peer-mapped mailbox communicator + IPC-shared allocations"""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > _lib.MB_MAX_WORLD:
            raise _lib.MocatB200Error(f"at most {_lib.MB_MAX_WORLD} GPUs share one population")
        self.L = _lib.get()
        self.ctx = self.L.ctx()
        self.device = torch.device("cuda", torch.cuda.current_device())
        handle = (C.c_char * 64)()
        comm = C.c_void_p()
        self.L.call("mb_comm_create", self.ctx, self.rank, self.world, C.byref(comm), handle)
        handles = all_gather_bytes(bytes(handle), group)
        self.L.call("mb_comm_connect", comm, b"".join(handles))
        self.comm = comm
        self._allocs = []

    def alloc_shared(self, shape, dtype):
        """cudaMalloc'd tensor on this rank + the device pointers of the same tensor on every rank"""
        import torch
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
        nbytes = int(np.prod(shape)) * {"<f4": 4, "<f8": 8, "<i4": 4, "<i8": 8}[typestr]
        p = C.c_void_p()
        self.L.call("mb_alloc", self.ctx, nbytes, C.byref(p))
        handle = (C.c_char * 64)()
        self.L.call("mb_ipc_get_handle", self.ctx, p, handle)
        handles = all_gather_bytes(bytes(handle), self.group)
        peers = []
        for r, h in enumerate(handles):
            if r == self.rank:
                peers.append(p.value)
            else:
                q = C.c_void_p()
                self.L.call("mb_ipc_open", self.ctx, h, C.byref(q))
                peers.append(q.value)
        t = torch.as_tensor(_RawCuda(p.value, shape, typestr), device=self.device)
        self._allocs.append((p, peers))
        return t, peers

    def free_shared(self, tensors):
        """collective release of alloc_shared() buffers (every rank passes its tensors in the same order): the peer
        mappings are closed first, then -- after a barrier, so that no rank still maps it -- the local allocation"""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        mine = []
        for t in tensors:
            for k, (p, peers) in enumerate(self._allocs):
                if p.value == t.data_ptr():
                    mine.append(self._allocs.pop(k))
                    break
        for p, peers in mine:
            for r, q in enumerate(peers):
                if r != self.rank:
                    self.L.call("mb_ipc_close", self.ctx, C.c_void_p(q))
        if dist.is_initialized() and self.world > 1:
            dist.barrier(self.group)
        for p, peers in mine:
            self.L.call("mb_free", self.ctx, p)

    def allgather(self, src, nd, dst):
        self.L.call("mb_comm_allgather", self.comm, _lib.ptr(src), int(nd), _lib.ptr(dst), None, _lib.stream())


def _make_shard(sc, n_local, x_peers, cdf_peers, totals, anc_peers, lw_peers, ws_peers):
    sh = _lib.Shard()
    sh.rank, sh.world, sh.n_local, sh.n_total = sc.rank, sc.world, int(n_local), int(n_local) * sc.world
    for r in range(sc.world):
        sh.x_peers[r] = x_peers[r]
        sh.cdf_peers[r] = cdf_peers[r] if cdf_peers is not None else None
        sh.anc_peers[r] = anc_peers[r]
        sh.lw_peers[r] = lw_peers[r]
        sh.ws_peers[r] = ws_peers[r]
    sh.totals = totals.data_ptr()
    return sh


def _sharded_alloc(eng, sc):
    """IPC-shared buffers of a sharded engine: both particle buffers, the ancestor array (the fused systematic resampler
    writes an output's ancestor into the owning rank's array), the log-weights and the resampler workspace (every rank
    fills its own share of the outputs of heavy source tiles, reading their weights from the owner) and, for
    multinomial resampling, the rank-relative CDF and the strata histogram."""
    import torch
    shape = tuple(eng._x_shape()) if hasattr(eng, "_x_shape") else tuple(eng.xbuf[0].shape)
    xs = [sc.alloc_shared(shape, torch.float32) for _ in range(2)]
    eng.xbuf = [xs[0][0], xs[1][0]]
    lw_len = eng._lw_full.numel() if hasattr(eng, "_lw_full") else eng.n
    lw_full, lw_peers = sc.alloc_shared((lw_len,), torch.float32)
    lw_full.zero_()
    if hasattr(eng, "_lw_full"):
        eng._lw_full = lw_full
    eng.lw = lw_full[:eng.n]
    eng.rs_ws, ws_peers = sc.alloc_shared((eng.rs_ws.numel(),), torch.int64)
    eng.rs_ws.zero_()
    eng.anc, anc_peers = sc.alloc_shared((eng.n,), torch.int32)
    eng._shared = [eng.xbuf[0], eng.xbuf[1], lw_full, eng.rs_ws, eng.anc]
    cdf_peers = None
    if eng.resampling == _lib.RESAMPLE_MULTINOMIAL:
        eng.cdf, cdf_peers = sc.alloc_shared((eng.n,), torch.float64)
        eng.hist_local, hist_peers = sc.alloc_shared((eng.B,), torch.int32)
        eng._shared += [eng.cdf, eng.hist_local]
        eng.hist_peers = (C.c_void_p * sc.world)(*hist_peers)
    eng.totals = torch.zeros(sc.world, dtype=torch.float64, device=sc.device)      # fp64 totals / uint64 bit patterns
    eng._barrier_out = torch.zeros(sc.world, dtype=torch.float64, device=sc.device)
    eng.shards = [_make_shard(sc, eng.n, xs[k][1], cdf_peers, eng.totals, anc_peers, lw_peers, ws_peers) for k in range(2)]


def _sharded_resample_kernels(eng, sc, st):
    """Every launch is predicated on the replicated control block.
    systematic : integer tile sums -> exchange of the 8-byte shard totals -> every rank scans ITS OWN source tiles and
                 writes the ancestors into the arrays of the ranks that own the outputs; tiles with collapsed weight are
                 only recorded -> exchange (barrier) -> every rank
                 fills ITS OWN share of the heavy tiles' outputs, reading their weights from the owner -> exchange
                 (barrier) before the step kernel overwrites the weights a peer may still be reading.
    multinomial: scan (rank-relative, exact) + local strata histogram -> one exchange (weight totals; also the barrier
                 before the peers read each other's histograms) -> histogram sum over ranks -> global sorted-uniform
                 ancestor search over the peer-mapped CDFs."""
    L, ptr = eng.L, _lib.ptr
    ctl = ptr(eng.ctl.t)
    if eng.resampling == _lib.RESAMPLE_SYSTEMATIC:
        L.call("mb_rs_tile_sums", eng.ctx, ptr(eng.rs_ws), ptr(eng.lw), eng.n, eng.n_total, 1, ctl, 0, st)
        L.call("mb_comm_allgather", sc.comm, ptr(eng.rs_ws), 1, ptr(eng.totals), ctl, st)
        L.call("mb_rs_ancestors", eng.ctx, ptr(eng.rs_ws), ptr(eng.lw), eng.n, eng.n_total, 1, ctl, 0, -1, ptr(eng.totals),
               C.byref(eng.shards[eng.cur]), ptr(eng.anc), st)
        L.call("mb_comm_allgather", sc.comm, ptr(eng.totals), 1, ptr(eng._barrier_out), ctl, st)
        L.call("mb_rs_heavy", eng.ctx, ptr(eng.rs_ws), ptr(eng.lw), eng.n, eng.n_total, 1, ctl, 0, -1, ptr(eng.totals),
               C.byref(eng.shards[eng.cur]), ptr(eng.anc), st)
        L.call("mb_comm_allgather", sc.comm, ptr(eng.totals), 1, ptr(eng._barrier_out), ctl, st)
        return
    L.call("mb_cumsum_lw", eng.ctx, ptr(eng.lw), eng.n, ctl, 2, ptr(eng.cdf), st)
    L.call("mb_strata_hist", eng.ctx, eng.n, eng.gid0, eng.B, eng.seed, 0, ctl, ptr(eng.hist_local), 1, st)
    L.call("mb_comm_allgather", sc.comm, ptr(eng.cdf[eng.n - 1:]), 1, ptr(eng.totals), ctl, st)
    L.call("mb_strata_reduce", eng.ctx, sc.comm, eng.hist_peers, sc.world, eng.B, ptr(eng.hist), 0, ctl, st)
    L.call("mb_ancestors_sorted", eng.ctx, None, eng.n_total, C.byref(eng.shards[eng.cur]), eng.resampling,
           ptr(eng.hist), ptr(eng.offsets), eng.B, eng.seed, 0, eng.gid0, eng.n_total, ptr(eng.anc), eng.n, ctl, st)


def ShardedSMCEngine(sc, target, move, temper, n_local, seed, resampling=_lib.RESAMPLE_MULTINOMIAL, schedule=None):
    """engine.SMCEngine whose population is sharded over sc.world GPUs (n_total = world * n_local)"""
    import torch
    from . import engine

    class _Eng(engine.SMCEngine):
        def __init__(self):
            super().__init__(target, move, temper, n_local, seed, resampling=resampling, gid0=sc.rank * n_local,
                             n_total=sc.world * n_local, schedule=schedule)
            self.sc = sc
            self.comm = sc.comm
            _sharded_alloc(self, sc)

        def close(self):
            """collective: release the IPC-shared buffers"""
            shared, self._shared = self._shared, []
            self._graphs = [None, None]
            self.xbuf, self.lw, self.rs_ws, self.cdf, self.anc = [None, None], None, None, None, None
            sc.free_shared(shared)

        def _shard_ref(self):
            return C.byref(self.shards[self.cur])

        def _enqueue(self, events=None):
            st = _lib.stream()
            if events:
                events[0].record()
            L, ptr = self.L, _lib.ptr
            body = self._cond_begin(st)
            try:
                _sharded_resample_kernels(self, sc, body)
            finally:
                self._cond_end(st)
            if events:
                events[1].record()
            src, dst = self.xbuf[self.cur], self.xbuf[self.cur ^ 1]
            L.call("mb_smc_move", self.ctx, C.byref(self.target), C.byref(self.move), ptr(src), ptr(dst), self.ld, self.n,
                   ptr(self.anc), ptr(self.lw), ptr(self.up), ptr(self.lik), ptr(self.alpha), self.seed, self.gid0,
                   ptr(self.ctl.t), self._shard_ref(), st)
            if events:
                events[2].record()
            self._temper(advance=True)
            if events:
                events[3].record()

    return _Eng()


_SC = None
_SMC_POOL = {}


def shard_context():
    """process-wide ShardContext (created on first use; collective: every rank must call it)"""
    global _SC
    if _SC is None:
        _SC = ShardContext()
    return _SC


def acquire_sharded_smc(target, move, temper, n_local, seed, resampling, schedule=None):
    """pooled ShardedSMCEngine (IPC allocations and handle exchange are collective and cost milliseconds:
    repeated runs of one configuration reuse them).  Every rank takes the same decisions."""
    sc = shard_context()
    sched = None if schedule is None else np.asarray(schedule, np.float64)
    key = (bytes(target), bytes(move), temper.max_temperature, temper.ess_retain, temper.ess_resample, temper.tol,
           temper.max_search_iter, temper.max_iter, int(n_local), int(resampling),
           None if sched is None else sched.tobytes())
    eng = _SMC_POOL.get(key)
    if eng is None:
        if len(_SMC_POOL) >= 2:
            _SMC_POOL.pop(next(iter(_SMC_POOL))).close()           # collective: frees its IPC-shared HBM
        eng = ShardedSMCEngine(sc, target, move, temper, n_local, seed, resampling=resampling, schedule=schedule)
        _SMC_POOL[key] = eng
    eng.seed = int(seed)
    return eng


_PF_POOL = {}


def acquire_sharded_pf(ssm, n_local, seed, ess_threshold, resampling):
    """pooled ShardedPFEngine (one entry: config C3 holds 2 x 16 GB / world per rank).  Collective; every rank takes the
    same decisions.  `generation` is bumped per hand-out (see engine.PFEngine.acquire)."""
    sc = shard_context()
    key = (int(ssm.kind), int(ssm.dim), int(n_local), int(resampling))
    eng = _PF_POOL.get(key)
    if eng is None:
        for old in list(_PF_POOL.values()):
            old.close()
        _PF_POOL.clear()
        eng = ShardedPFEngine(sc, ssm, n_local, seed, ess_threshold=ess_threshold, resampling=resampling)
        _PF_POOL[key] = eng
    eng.ssm, eng.seed, eng.ess_threshold = ssm, int(seed), float(ess_threshold)
    eng.generation += 1
    return eng


def release_pools():
    """collective: close every pooled sharded engine (frees their IPC-shared HBM)"""
    for pool in (_PF_POOL, _SMC_POOL):
        for eng in list(pool.values()):
            if hasattr(eng, "close"):
                eng.close()
        pool.clear()


def ShardedPFEngine(sc, ssm, n_local, seed, ess_threshold=0.5, resampling=_lib.RESAMPLE_MULTINOMIAL):
    """engine.PFEngine whose population is sharded over sc.world GPUs"""
    from . import engine
    if n_local % 32:
        raise _lib.MocatB200Error("sharded particle filter: n_local must be a multiple of 32 (tile / pair alignment)")

    class _Eng(engine.PFEngine):
        def __init__(self):
            super().__init__(ssm, n_local, seed, ess_threshold=ess_threshold, resampling=resampling,
                             gid0=sc.rank * n_local, n_total=sc.world * n_local)
            self.sc = sc
            _sharded_alloc(self, sc)

        def _alloc_x(self, dev):
            return None                                     # the IPC-shared buffers of _sharded_alloc replace them

        def _comm(self):
            return sc.comm

        def close(self):
            """collective: release the IPC-shared buffers (they are not owned by torch)"""
            shared, self._shared = self._shared, []
            self.xbuf, self.lw, self._lw_full, self.rs_ws, self.anc = [None, None], None, None, None, None
            sc.free_shared(shared)

        def _shard_ref(self):
            return C.byref(self.shards[self.cur])

        def _resample_kernels(self, st):
            _sharded_resample_kernels(self, sc, st)

    return _Eng()


def ShardedABCEngine(sc, gk, n_local, seed, **kw):
    """engine.ABCEngine whose population is sharded over sc.world GPUs: global quantile threshold
    (radix histograms added over the ranks), global column variances and acceptance statistics (NCCL all-reduce of
    small records between the stages of mb_abc_adapt_stage), sharded resampling, ancestor state read over NVLink."""
    import torch
    import torch.distributed as dist
    from . import engine
    ptr = _lib.ptr

    class _Eng(engine.ABCEngine):
        def __init__(self):
            super().__init__(gk, n_local, seed, gid0=sc.rank * n_local, n_total=sc.world * n_local, **kw)
            self.sc = sc
            _sharded_alloc(self, sc)
            self._peer = {}
            for name in ("upbuf", "distbuf", "alphabuf"):
                bufs, peers = [], []
                for k in range(2):
                    t, p = sc.alloc_shared((self.n,), torch.float32)
                    t.zero_()
                    bufs.append(t); peers.append((C.c_void_p * sc.world)(*p))
                    self._shared.append(t)
                setattr(self, name, bufs)
                self._peer[name] = peers
            self._peer["xbuf"] = [(C.c_void_p * sc.world)(*[self.shards[k].x_peers[r] for r in range(sc.world)])
                                  for k in range(2)]
            self.ws = torch.zeros(_lib.MB_ABC_WS_BYTES // 8, dtype=torch.int64, device=sc.device)
            self._ws_f64, self._ws_i32 = self.ws.view(torch.float64), self.ws.view(torch.int32)

        def _resample_kernels(self, st):
            _sharded_resample_kernels(self, sc, st)

        def _stage(self, stage, advance):
            self.L.call("mb_abc_adapt_stage", self.ctx, stage, ptr(self.x), self.ld, self.n, self.n_total, self.d,
                        ptr(self.dist), ptr(self.lw), ptr(self.alpha), ptr(self.stepsize), self.ess_retain,
                        self.ess_resample, self.termination_alpha, self.max_iter, ptr(self._schedule),
                        1 if advance else 0, ptr(self.ws), ptr(self.ctl.t), ptr(self.ctl.hist), _lib.stream())

        def _adapt(self, advance):
            d = self.d
            self._stage(0, advance)
            dist.all_reduce(self._ws_f64[16:16 + 1 + 2 * d])               # column sums (ws + 128)
            if self._schedule is None:
                for p_ in (1, 2, 3):
                    self._stage(p_, advance)
                    dist.all_reduce(self._ws_i32[256:256 + 2048])          # radix histogram of this pass (ws + 1024)
                    self._stage(10 + p_, advance)
                self._stage(4, advance)
                dist.all_reduce(self.ws[80:81])                            # entries <= selected (ws + 640)
                dist.all_reduce(self.ws[81:82], op=dist.ReduceOp.MIN)      # smallest key above it (ws + 648)
            self._stage(5, advance)
            dist.all_reduce(self.ws[8:11])                                 # alive / previously alive / acceptance (ws + 64)
            self._stage(6, advance)

        def update(self):
            st = _lib.stream()
            self._resample_kernels(st)
            c, o = self.cur, self.cur ^ 1
            self.L.call("mb_abc_move_sharded", self.ctx, C.byref(self.gk), self.mcmc_steps, ptr(self.xbuf[o]), self.ld, self.n,
                        ptr(self.anc), ptr(self.upbuf[o]), ptr(self.distbuf[o]), ptr(self.lw), ptr(self.alphabuf[o]),
                        ptr(self.stepsize), self.seed, sc.rank, sc.world, self._peer["xbuf"][c], self._peer["upbuf"][c],
                        self._peer["distbuf"][c], self._peer["alphabuf"][c], ptr(self.ctl.t), st)
            self.cur = o
            # nobody may overwrite the buffers of parity c (next update's output) while a peer still reads them: the
            # first collective of the adaptation below orders this step's moves of all ranks before the next step's
            self._adapt(advance=True)
            self.enqueued += 1

        def close(self):
            shared, self._shared = self._shared, []
            sc.free_shared(shared)

    return _Eng()
