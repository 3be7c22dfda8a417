import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import mocat_b200 as mocat
n, D = 1_000_000, 5
x0 = torch.empty((n, D), dtype=torch.float32).pin_memory(); x0.copy_(torch.randn(n, D) * 3)
sc = mocat.scenarios.Rastrigin(dim=D, a=1.0, prior_std=3.0)
def run(iters=200):
    smp = mocat.MetropolisedSMCSampler(mocat.Underdamped(stepsize=0.1), max_iter=iters, keep_history=False, check_every=iters)
    t0 = time.perf_counter(); out = mocat.run(sc, smp, n, random_key=1, initial_state=mocat.cdict(value=x0.numpy())); torch.cuda.synchronize()
    return time.perf_counter() - t0, out
run(); run()
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); dt, out = run(); pr.disable()
print("wall", dt, "iters", len(out.temperature) - 1)
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
# raw copy speeds
a = torch.randn(n, D, device="cuda"); torch.cuda.synchronize()
t0 = time.perf_counter(); b = a.cpu().numpy(); print("pageable D2H 20MB: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
pin = torch.empty((n, D), dtype=torch.float32).pin_memory()
t0 = time.perf_counter(); pin.copy_(a, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter(); c = pin.numpy().copy(); t2 = time.perf_counter()
print("pinned D2H 20MB: %.2f ms, host copy %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter(); d = torch.as_tensor(x0.numpy(), device="cuda"); torch.cuda.synchronize(); print("H2D from pinned numpy view 20MB: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
