"""A few SMC-ABC population steps (config C5, g-and-k m = 8) for ncu: python scratch/prof_abc.py [n] [steps]"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib, engine, models
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
z = np.random.default_rng(0).standard_normal(8)
data = np.sort(3.0 + 1.0 * (1 + 0.8 * np.tanh(2.0 * z / 2)) * z * (1 + z * z) ** 0.5)
eng = engine.ABCEngine(models.make_gk(data), n, 0, max_iter=1 << 30, resampling=_lib.RESAMPLE_MULTINOMIAL)
eng.startup()
for _ in range(steps):
    eng.update()
torch.cuda.synchronize()
print(eng.ctl.read()['iter'])
