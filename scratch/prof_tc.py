import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import engine
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 50
rng = np.random.default_rng(0)
X = torch.as_tensor((rng.standard_normal((n, d)) * 0.7 + 1).astype(np.float32), device="cuda")
G = torch.as_tensor(rng.standard_normal((n, d)).astype(np.float32), device="cuda")
h = torch.tensor([0.9 * np.sqrt(d) * 0.7], dtype=torch.float32, device="cuda")
for _ in range(3): engine.svgd_phi(X, G, h, 1)
torch.cuda.synchronize()
