"""A few SVGD iterations of config C4 (median heuristic every iteration) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_svgd.csv python scratch/c4_launches.py"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import mocat_b200 as mocat
from mocat_b200 import kernels

n, d, N = 32768, 50, 1024
rng = np.random.default_rng(0)
A = rng.normal(size=(N, d)).astype(np.float32)
t = (rng.random(N) < 1 / (1 + np.exp(-A @ rng.normal(size=d)))).astype(np.float32)
sc = mocat.scenarios.LogisticRegression(A, t)
class SVGDMedian(mocat.SVGD):
    def adapt(self, st, extra):
        extra.parameters.kernel_params.bandwidth = kernels.median_bandwidth_update(st.value)
        return st, extra
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
out = mocat.run(sc, SVGDMedian(max_iter=iters, stepsize=0.05, keep_history=False), n=n, random_key=1)
torch.cuda.synchronize()
print("done", iters)
