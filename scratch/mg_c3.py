"""per-rank timing of the sharded C3 step: torchrun --nproc-per-node N scratch/mg_c3.py [n_total] [ess_threshold]"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib, engine, models, parallel, ssm
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
sc = parallel.shard_context()
pf = parallel.ShardedPFEngine(sc, models.make_lorenz96(dim=40), n_total // world, 0, ess_threshold=thr, resampling=_lib.RESAMPLE_SYSTEMATIC)
scen = ssm.Lorenz96(dim=40); y = scen.simulate(0.05 * np.arange(40), 0).y.astype(np.float32)
yd = torch.as_tensor(y, device="cuda")
pf.init(yd[0])
for t in range(1, 6): pf.step(yd[t])
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
st = _lib.stream(); L = pf.L; ptr = _lib.ptr; ctl = ptr(pf.ctl.t)
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for t in range(6, 26):
    m = [ev() for _ in range(7)]
    pf.t += 1
    m[0].record()
    L.call("mb_rs_tile_sums", pf.ctx, ptr(pf.rs_ws), ptr(pf.lw), pf.n, pf.n_total, 1, ctl, 0, st); m[1].record()
    L.call("mb_comm_allgather", sc.comm, ptr(pf.rs_ws), 1, ptr(pf.totals), ctl, st); m[2].record()
    import ctypes as C
    L.call("mb_rs_ancestors", pf.ctx, ptr(pf.rs_ws), ptr(pf.lw), pf.n, pf.n_total, 1, ctl, 0, -1, ptr(pf.totals), C.byref(pf.shards[pf.cur]), ptr(pf.anc), st); m[3].record()
    L.call("mb_comm_allgather", sc.comm, ptr(pf.totals), 1, ptr(pf._barrier_out), ctl, st); m[4].record()
    L.call("mb_rs_heavy", pf.ctx, ptr(pf.rs_ws), ptr(pf.lw), pf.n, pf.n_total, 1, ctl, 0, -1, ptr(pf.totals), C.byref(pf.shards[pf.cur]), ptr(pf.anc), st)
    L.call("mb_comm_allgather", sc.comm, ptr(pf.totals), 1, ptr(pf._barrier_out), ctl, st); m[5].record()
    pf._step_kernel(yd[t], st); m[6].record()
    rows.append(m)
torch.cuda.synchronize()
names = ["tile_sums", "exch1", "passB", "exch2", "passC+exch3", "pf_step", "imported", "remote"]
med = [float(np.median([r[k].elapsed_time(r[k + 1]) for r in rows])) for k in range(6)]
per_step = [[r[k].elapsed_time(r[k + 1]) for k in range(6)] for r in rows[:6]]
out = [None] * world
nimp = int((pf.anc < 0).sum().item())
owner = pf.anc.to(torch.int64) // pf.n
nrem = int((owner != rank).sum().item())
med = med + [nimp, nrem]
dist.all_gather_object(out, (rank, med, float(pf.ctl.read()['ess']), per_step))
if rank == 0:
    print("n_total", n_total, "world", world, "thr", thr)
    for r, md, ess, ps in out:
        print("rank %d: " % r + "  ".join("%s %.3f" % (nm, v) for nm, v in zip(names, md)) + "  | sum %.3f  ess %.2f" % (sum(md[:6]), ess))
    for k in range(6):
        print("step", k, " passB per rank:", " ".join("%.3f" % o[3][k][2] for o in out), " passC+x3:", " ".join("%.3f" % o[3][k][4] for o in out), " pf:", " ".join("%.3f" % o[3][k][5] for o in out))
dist.barrier(); dist.destroy_process_group()
