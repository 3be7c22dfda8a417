import sys, torch
sys.path.insert(0, '.')
from mocat_b200 import engine
X = torch.randn(32768, 50, device='cuda')
for _ in range(2):
    engine.pairdist_bandwidth(X, "median", 1); engine.pairdist_bandwidth(X, "mean", 1)
torch.cuda.synchronize()
