"""C3 resampler kernels in the collapsed-weight state of the filter (n = 1e8, after `steps` filter steps): three rounds of
tile sums + ancestors (+ heavy), for ncu:  ncu -k regex:rf_ --metrics gpu__time_duration.sum python scratch/rs_profile.py"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from mocat_b200 import _lib, engine, models
from mocat_b200._lib import ptr, stream
from oracle import models as om
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = _lib.get(); ctx = L.ctx(); d = 40
pf = engine.PFEngine(models.make_lorenz96(dim=d), n, 3, ess_threshold=2.0, resampling=_lib.RESAMPLE_SYSTEMATIC)
_, ys = om.Lorenz96SSM(dim=d).simulate(steps + 2, np.random.default_rng(0), spinup=1000)
yd = torch.as_tensor(ys.astype(np.float32), device="cuda")
pf.init(yd[0])
for t in range(1, steps + 1): pf.step(yd[t])
torch.cuda.synchronize()
c = pf.ctl.read(); print("ess", c['ess'])
st = stream()
for _ in range(3):
    L.call("mb_rs_tile_sums", ctx, ptr(pf.rs_ws), ptr(pf.lw), n, n, 1, ptr(pf.ctl.t), 1, st)
    L.call("mb_rs_ancestors", ctx, ptr(pf.rs_ws), ptr(pf.lw), n, n, 1, ptr(pf.ctl.t), 1, -1, None, None, ptr(pf.anc), st)
torch.cuda.synchronize()
print("heavy jobs", int(pf.rs_ws[1].item() & 0xffffffff))
a = pf.anc.cpu().numpy()
u, cnt = np.unique(a, return_counts=True)
print("distinct ancestors", len(u), "max offspring", cnt.max(), "outputs from ancestors with <= 64 offspring", int(cnt[cnt <= 64].sum()))
