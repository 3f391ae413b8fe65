"""One launch of each hot kernel at its headline size, for `ncu --set full` (profiles/r02_*)."""
import os, sys, math
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E = 10 ** 7
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
fine = torch.empty((E, 32), dtype=torch.float64, device='cuda')
u = torch.empty(E + 1, dtype=torch.float64, device='cuda')
err = batch.new_error_accumulator()
for _ in range(2):      # the second round is the one to read (first-launch effects in the first)
    batch.fem_p1_solve(nodes, coarse_solver='assembled', out=u)
    batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine, err3=err)
    batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine)
Ed = 10 ** 6
nd = batch.mesh_linspace(-1.0, 1.0, Ed + 1)
ud = batch.fem_p1_solve(nd, coarse_solver='flux')
batch.lssvr_dual_batch(nd, ud, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine[:Ed])
E4, R = 10 ** 4, 64
ks = torch.arange(1, R + 1, dtype=torch.float64, device='cuda')
n4 = torch.from_numpy(0.3 + np.linspace(-1.0, 1.0, E4 + 1) * 1e-2).cuda()
u4 = torch.sin(math.pi * ks[:, None] * n4[None, :]).contiguous()
batch.lssvr_dual_multi(n4, u4, ks, 25, 1e4, N=128, F=32, want_coef=False, want_fine=True)
torch.cuda.synchronize()
print('profile_r02 done')
