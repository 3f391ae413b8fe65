"""Scan of mesh sizes: exact-row-sum coarse solve against the flux form and sin(pi x) (uniform and jittered meshes)."""
import os, sys, math, numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
rng = np.random.default_rng(1)
sizes = sorted(set([2 ** p + d for p in range(11, 25) for d in (-1, 0, 1, 5)] + [int(x) for x in 10 ** rng.uniform(4, 7.4, 40)]))
worst = (0.0, 0)
for n in sizes:
    nodes = batch.mesh_linspace(-1.0, 1.0, n)
    uf = batch.fem_p1_solve(nodes, coarse_solver='flux')
    ue = batch.fem_p1_solve(nodes, coarse_solver='assembled_exact')
    d = (ue - uf).abs().max().item()
    s = (ue - torch.sin(math.pi * nodes)).abs().max().item()
    flag = ' <<<' if d > 2e-13 else ''
    if d > worst[0]: worst = (d, n)
    print('n=%9d  exact-flux %.2e  exact-sin %.2e%s' % (n, d, s, flag))
print('worst exact-flux %.2e at n=%d' % worst)
for n in (10 ** 5 + 3, 10 ** 6 + 7, 10 ** 7 + 1):
    w = 1.0 + 0.5 * torch.rand(n - 1, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(n)) * 2 - 0.5
    x = torch.cat([torch.zeros(1, dtype=torch.float64, device='cuda'), torch.cumsum(w, 0)])
    nodes = (-1.0 + 2.0 * x / x[-1]).contiguous()
    uf = batch.fem_p1_solve(nodes, coarse_solver='flux')
    ue = batch.fem_p1_solve(nodes, coarse_solver='assembled_exact')
    print('jittered n=%9d  exact-flux %.2e' % (n, (ue - uf).abs().max().item()))
