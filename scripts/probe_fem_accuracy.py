"""K1 accuracy probe: GPU coarse solves against SuperLU / LAPACK on identical data and against the analytic discrete solution."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
from oracle import fem_p1
for n in (1001, 10001, 100001, 1000001):
    nodes = np.linspace(-1, 1, n)
    d = torch.from_numpy(nodes).cuda()
    ua = batch.fem_p1_solve(d, coarse_solver='assembled').cpu().numpy()
    uf = batch.fem_p1_solve(d, coarse_solver='flux').cpu().numpy()
    sp = fem_p1.solve_fem_p1(nodes)
    bd = fem_p1.solve_fem_p1(nodes, solver='banded')
    ex = fem_p1.c_factor(2.0 / (n - 1)) * np.sin(np.pi * nodes)
    print('n=%8d  gpuA-superlu %.2e  gpuA-lapack %.2e  superlu-lapack %.2e | gpuA-analytic %.2e  gpuF-analytic %.2e  superlu-analytic %.2e'
          % (n, np.max(np.abs(ua - sp)), np.max(np.abs(ua - bd)), np.max(np.abs(sp - bd)), np.max(np.abs(ua - ex)),
             np.max(np.abs(uf - ex)), np.max(np.abs(sp - ex))))
