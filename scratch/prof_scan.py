"""ncu target: scan + sorted ancestors at n = 1e8"""
import sys, ctypes as C
import torch
sys.path.insert(0, '.')
from mocat_b200 import _lib, engine, models
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
tgt = models.make_target(_lib.LIK_RASTRIGIN, 5, prior_std=3.0, a=1.0)
eng = engine.SMCEngine(tgt, models.make_move(_lib.MOVE_MALA, 0.1), models.make_temper(max_iter=1 << 30), n, 1,
                       resampling=_lib.RESAMPLE_SYSTEMATIC)
eng.use_graphs = False
eng.startup()
L, ptr = eng.L, _lib.ptr
for _ in range(3):
    L.call("mb_cumsum_lw", eng.ctx, ptr(eng.lw), eng.n, ptr(eng.ctl.t), 1, ptr(eng.cdf), _lib.stream())
    L.call("mb_ancestors_sorted", eng.ctx, ptr(eng.cdf), eng.n, None, eng.resampling, ptr(eng.hist), ptr(eng.offsets),
           eng.B, eng.seed, 1, 0, eng.n, ptr(eng.anc), eng.n, None, _lib.stream())
torch.cuda.synchronize()
