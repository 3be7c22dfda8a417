"""debug (library built with -DMB_TEMPER_TRACE): per-phase ns of the resident tempering kernel, block 0"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from mocat_b200 import _lib, engine, models
tgt = models.make_target(_lib.LIK_RASTRIGIN, 5, prior_std=3.0, a=1.0)
eng = engine.SMCEngine(tgt, models.make_move(_lib.MOVE_MALA, 0.1), models.make_temper(max_iter=1 << 30), 1_000_000, 1)
eng.startup()
for _ in range(25): eng.update()
torch.cuda.synchronize()
it = int(eng.ctl.read()['iter'])
h = eng.ctl.read_hist(it + 1)
for i in range(it - 8, it + 1):
    ev = h['search_iters'][i] + 1
    tot = h['beta'][i] + h['ess'][i] + h['log_z'][i] + h['alpha_mean'][i] + h['lse'][i]
    print(f"iter {i}: evals {ev}  compute {h['ess'][i]/ev:7.0f}  publish+gather {h['log_z'][i]/ev:7.0f}  merge {h['alpha_mean'][i]/ev:7.0f}  between {h['beta'][i]/ev:7.0f} ns/eval | tail {h['lse'][i]:7.0f}  total {tot:8.0f} ns")
