"""BASELINE configs[1]: dual LSSVR at N = 12 (two 7 x 7 parity blocks in registers), 1e6 elements; timing and ncu target."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E = int(os.environ.get('E', 10 ** 6))
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
u = batch.fem_p1_solve(nodes, coarse_solver='flux')
fine = torch.empty((E, 32), dtype=torch.float64, device='cuda')
fn = lambda: batch.lssvr_dual_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine)
fp = lambda: batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine)
for name, f in (('dual', fn), ('primal', fp)):
    for _ in range(5): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print('%s E=%d best %.4f ms median %.4f ms -> %.3e element solves/s' % (name, E, ts[0], ts[10], E / (ts[10] * 1e-3)))
