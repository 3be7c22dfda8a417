"""is the SVGD iteration (config C4) host-bound?  enqueue time of 200 iterations vs their completion"""
import sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0, '/root/repo')
import mocat_b200 as mocat
from mocat_b200 import kernels
n, d, N = 32768, 50, 1024
rng = np.random.default_rng(0)
A = rng.standard_normal((N, d)).astype(np.float32)
lab = (rng.random(N) < 1.0 / (1.0 + np.exp(-(A @ rng.standard_normal(d))))).astype(np.float32)
sc = mocat.scenarios.LogisticRegression(A, lab)
class SVGDMedian(mocat.SVGD):
    def adapt(self, st, extra):
        extra.parameters.kernel_params.bandwidth = kernels.median_bandwidth_update(st.value)
        return st, extra
for rep in range(3):
    smp = SVGDMedian(max_iter=200, stepsize=0.05, keep_history=False)
    smp.n = n
    extra = mocat.cdict(); extra.random_key = rep
    st, extra = smp.startup(sc, n, None, extra)
    torch.cuda.synchronize()
    pr = cProfile.Profile() if rep == 2 else None
    t0 = time.perf_counter()
    if pr: pr.enable()
    for _ in range(200):
        st, extra = smp.update(sc, st, extra)
    if pr: pr.disable()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("rep %d: enqueue %.1f ms, complete %.1f ms  (%.3f / %.3f ms per iteration)" % (rep, (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t1 - t0) * 5, (t2 - t0) * 5))
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
