import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib, engine, models, parallel
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
sc = parallel.ShardContext()
n_local, d, seed = 40_000, 5, 9
tgt = models.make_target(_lib.LIK_RASTRIGIN, d, prior_std=3.0, a=1.0)
mk = lambda: (models.make_move(_lib.MOVE_MALA, 0.1), models.make_temper(max_iter=25))
eng = parallel.ShardedSMCEngine(sc, tgt, *mk(), n_local, seed, resampling=_lib.RESAMPLE_SYSTEMATIC)
ref = None
if rank == 0:
    ref = engine.SMCEngine(tgt, *mk(), n_local * world, seed, resampling=_lib.RESAMPLE_SYSTEMATIC); ref.use_graphs = False
def compare(tag):
    xs = eng.values().contiguous(); g = [torch.empty_like(xs) for _ in range(world)]; dist.all_gather(g, xs)
    ls = eng.lw.contiguous(); gl = [torch.empty_like(ls) for _ in range(world)]; dist.all_gather(gl, ls)
    an = eng.anc.contiguous(); ga = [torch.empty_like(an) for _ in range(world)]; dist.all_gather(ga, an)
    c = eng.ctl.read()
    if rank == 0:
        cr = ref.ctl.read()
        same = float(np.mean(np.all(torch.cat(g).cpu().numpy() == ref.values().cpu().numpy(), axis=1)))
        samel = float(np.mean(torch.cat(gl).cpu().numpy() == ref.lw.cpu().numpy()))
        samea = float(np.mean(torch.cat(ga).cpu().numpy() == ref.anc.cpu().numpy()))
        print(f"{tag}: beta {c['beta']:.12g} vs {cr['beta']:.12g} | ess {c['ess']:.6f} vs {cr['ess']:.6f} | res {c['resampled']} {cr['resampled']} next {c['resample']} {cr['resample']} | x same {same:.4f} lw same {samel:.4f} anc same {samea:.4f} iters {c['search_iters']} {cr['search_iters']}", flush=True)
eng.startup()
if rank == 0: ref.startup()
compare("startup")
for it in range(8):
    eng.update()
    if rank == 0: ref.update()
    compare(f"iter {it+1}")
dist.barrier(); dist.destroy_process_group()
