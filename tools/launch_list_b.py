"""Workload behind profiles/r2_launches_level0_nmf_session3.csv (run under
`ncu --metrics gpu__time_duration.sum --clock-control none`): level 0 on the C3 graph twice, then
the NMF loop with sklearn's schedule (10 iterations + the checks around them) at n = 4 M, r = 8."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from graphrole_b200.graph import level0
from graphrole_b200.graph.generators import barabasi_albert_csr
from graphrole_b200.roles import factor

dev = torch.device('cuda', 0)
g = barabasi_albert_csr(10_000_000, 20, seed=0, device=dev)
for _ in range(2):
    level0.device_features(g)
torch.cuda.synchronize()
del g
torch.cuda.empty_cache()
gen = torch.Generator(device=dev).manual_seed(0)
n, f, r = 4_000_000, 512, 8
X = torch.rand(n, f, device=dev, generator=gen)
W = torch.rand(n, r, device=dev, generator=gen) + 0.1
H = torch.rand(r, f, device=dev, generator=gen) + 0.1
solver = factor.NmfSolver(n, f, r, dev)
solver.update(X, W, H, max_iter=10, tol=1e-30, check_every=10)
torch.cuda.synchronize()
print('done', solver.last_path)
