"""where the end-to-end call of config C3 spends its wall time: python scratch/e2e_c3_profile.py [n] [T]"""
import sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0, '/root/repo')
import mocat_b200 as mocat
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sc = mocat.ssm.Lorenz96(40)
sim = sc.simulate(np.arange(T) * 0.05, 0)
y, t = sim.y.astype(np.float32), sim.t
def run(key, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = mocat.ssm.run_particle_filter_for_marginals(sc, mocat.ssm.BootstrapFilter(), y, t, key, n=n, ess_threshold=2.0,
                                                      resampling='systematic', **kw)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out
print("first call %.4f s" % run(1)[0])
print("second call %.4f s" % run(2)[0])
print("moments=False %.4f s" % run(3, moments=False)[0])
pr = cProfile.Profile(); pr.enable(); dt, out = run(4); pr.disable()
print("profiled call %.4f s" % dt)
pstats.Stats(pr).sort_stats("cumtime").print_stats(18)
# device-only loop for reference
eng = out.engine
yd = torch.as_tensor(y, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(1, T): eng.step(yd[i])
torch.cuda.synchronize(); print("99 bare steps %.4f s" % (time.perf_counter() - t0))
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): eng.moments()
torch.cuda.synchronize(); print("moments kernel %.3f ms" % ((time.perf_counter() - t0) * 50))
