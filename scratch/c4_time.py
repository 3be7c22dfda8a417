"""C4 timing: SVGD iterations/s, n=32768 d=50 logistic regression (N=1024 data), median/mean bandwidth."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import mocat_b200 as mocat
from mocat_b200 import engine, kernels

n, d, N = 32768, 50, 1024
rng = np.random.default_rng(0)
A = rng.normal(size=(N, d)).astype(np.float32)
t = (rng.random(N) < 1 / (1 + np.exp(-A @ rng.normal(size=d)))).astype(np.float32)
sc = mocat.scenarios.LogisticRegression(A, t)
X = torch.randn(n, d, device='cuda')
def tm(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
Xs = X * 0.2
print("logistic pg ms (auto variant)", tm(lambda: sc._potential_grad_device(Xs, 1.0)))
U, G = sc._potential_grad_device(Xs, 1.0)
U0, G0 = engine.logistic_potential_grad(sc.features, sc.labels, 0.0, 1.0, 1.0, Xs, 0)
print("  tc vs fp32: max|dG|/max|G| = %.2e, rel rms = %.2e, max rel dU = %.2e" % (
    (G - G0).abs().max().item() / G0.abs().max().item(), ((G - G0).pow(2).mean().sqrt() / G0.pow(2).mean().sqrt()).item(),
    ((U - U0).abs() / U0.abs()).max().item()))
h = torch.tensor([7.0], device='cuda')
print("phi tc ms", tm(lambda: engine.svgd_phi(X, G, h, 1)))
print("mean bw tc ms", tm(lambda: kernels.mean_bandwidth_update(X, 1), 5), kernels.mean_bandwidth_update(X, 1).item(), kernels.mean_bandwidth_update(X, 0).item())
print("median bw tc ms", tm(lambda: kernels.median_bandwidth_update(X, 1), 5), kernels.median_bandwidth_update(X, 1).item(), kernels.median_bandwidth_update(X, 0).item())
class SVGDMedian(mocat.SVGD):
    def adapt(self, st, extra):
        extra.parameters.kernel_params.bandwidth = kernels.median_bandwidth_update(st.value)
        return st, extra
for cls, iters in ((mocat.SVGD, 200), (SVGDMedian, 200)):
    variant = cls.__name__
    s = cls(max_iter=iters, stepsize=0.05, keep_history=False)
    mocat.run(sc, s, n=n, random_key=1)
    torch.cuda.synchronize(); t0 = time.time()
    out = mocat.run(sc, s, n=n, random_key=1)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"SVGD {variant}: {iters/dt:.1f} iters/s ({dt/iters*1e3:.2f} ms/iter)")
