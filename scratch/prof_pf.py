"""ncu / timing target: Lorenz-96 d=40 bootstrap-filter step"""
import sys, time, ctypes as C
import torch
sys.path.insert(0, '.')
from mocat_b200 import _lib, engine, models
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
s = models.make_lorenz96(dim=40)
pf = engine.PFEngine(s, n, 3, ess_threshold=0.0, resampling=_lib.RESAMPLE_SYSTEMATIC)   # threshold 0: never resample
y = torch.randn(40, device="cuda") + 2
pf.init(y)
for _ in range(2): pf.step(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): pf.step(y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"n={n} L96 pf step (no resampling; incl. predicated resample launches): {ms:.3f} ms  {n*328/ms/1e6:.0f} GB/s")
