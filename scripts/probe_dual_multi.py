"""BASELINE configs[4] dual solve (N = 128, R = 64) once per M, timed; used for A/B runs and as the ncu target."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E, R = int(os.environ.get('E', 10 ** 4)), 64
Ms = [int(m) for m in os.environ.get('MS', '5,25').split(',')]
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
ks = torch.arange(1, R + 1, dtype=torch.float64, device='cuda')
u = batch.fem_p1_solve_multi(nodes, ks, coarse_solver='flux')
for M in Ms:
    fn = lambda: batch.lssvr_dual_multi(nodes, u, ks, M, 1e4, N=128, F=32, want_coef=False, want_fine=True)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print('M=%d  best %.4f ms  median %.4f ms' % (M, ts[0], ts[len(ts) // 2]))
