import sys, os, shutil, numpy as np, torch, ctypes
sys.path.insert(0, '/root/repo')
lib = sys.argv[1] if len(sys.argv) > 1 else None
from mocat_b200 import _lib
if lib: _lib.LIB_PATH = lib; _lib._LIB = _lib.Library(lib)
from mocat_b200 import engine
n, d = 32768, 50
rng = np.random.default_rng(0)
X = torch.as_tensor((rng.standard_normal((n, d)) * 0.7 + 1).astype(np.float32), device="cuda")
G = torch.as_tensor(rng.standard_normal((n, d)).astype(np.float32), device="cuda")
h = torch.tensor([0.9 * np.sqrt(d) * 0.7], dtype=torch.float32, device="cuda")
ref = engine.svgd_phi(X, G, h, 0)
for _ in range(3): out = engine.svgd_phi(X, G, h, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = engine.svgd_phi(X, G, h, 1)
e1.record(); torch.cuda.synchronize()
err = ((out - ref).abs().max() / ref.abs().max()).item()
print(f"{lib}: {e0.elapsed_time(e1)/20:.4f} ms/phi  rel err {err:.3e}")
