"""A few bootstrap-PF steps of config C3 for ncu: python scratch/prof_c3.py [n] [steps] [r_std]"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib, engine, models
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
r_std = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
d = 40
s = models.make_lorenz96(dim=d, r_std=r_std)
pf = engine.PFEngine(s, n, 3, ess_threshold=2.0, resampling=_lib.RESAMPLE_SYSTEMATIC)
rng = np.random.default_rng(0)
ys = torch.as_tensor((rng.standard_normal((steps + 1, d)) * 3.0 + 2.0).astype(np.float32), device="cuda")
pf.init(ys[0])
for t in range(1, steps + 1):
    pf.step(ys[t])
torch.cuda.synchronize()
print(pf.ctl.read()['ess'])
