import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import engine
from oracle import svgd as osvgd
def run(n, d, spread=0.7, shift=1.0, check=True):
    rng = np.random.default_rng(n + d)
    X = rng.standard_normal((n, d)) * spread + shift
    G = rng.standard_normal((n, d))
    h = 0.9 * np.sqrt(d) * spread
    Xd, Gd = (torch.as_tensor(a.astype(np.float32), device="cuda") for a in (X, G))
    hd = torch.tensor([h], dtype=torch.float32, device="cuda")
    phi0 = engine.svgd_phi(Xd, Gd, hd, 0)
    phi1 = engine.svgd_phi(Xd, Gd, hd, 1)
    torch.cuda.synchronize()
    p0, p1 = phi0.cpu().numpy(), phi1.cpu().numpy()
    scale = np.abs(p0).max()
    msg = f"n={n} d={d}: tc vs simt max abs diff / max|phi| = {np.abs(p1 - p0).max() / scale:.3e}, nan={np.isnan(p1).sum()}"
    if check and n <= 4096:
        ref = osvgd.phi(X.astype(np.float32).astype(np.float64), G.astype(np.float32).astype(np.float64), np.float32(h))
        msg += f" | simt vs fp64 {np.abs(p0 - ref).max() / scale:.3e} | tc vs fp64 {np.abs(p1 - ref).max() / scale:.3e}"
    print(msg, flush=True)
    return Xd, Gd, hd
for n, d in [(128, 5), (256, 5), (300, 2), (1000, 50), (4096, 50)]:
    run(n, d)
Xd, Gd, hd = run(32768, 50, check=False)
for variant in (1, 0):
    for _ in range(2): engine.svgd_phi(Xd, Gd, hd, variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): engine.svgd_phi(Xd, Gd, hd, variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"variant {variant}: {ms:.3f} ms/phi  => {3.24e11 / (ms * 1e-3) / 1e12:.1f} algorithmic TFLOP/s", flush=True)
