"""Primal / dual element kernels with SAMPLED forcing ([N][E] array read from HBM), 1e7 elements, M = 9, N = 12, F = 32."""
import os, sys, math, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E, N, M, F = int(os.environ.get('E', 10 ** 7)), 12, 9, 32
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
u = torch.sin(math.pi * nodes)
x = nodes[:-1].unsqueeze(0) + (nodes[1:] - nodes[:-1]).unsqueeze(0) * torch.linspace(0, 1, N, dtype=torch.float64, device='cuda').unsqueeze(1)
f = (math.pi ** 2 * torch.sin(math.pi * x)).contiguous()
del x
fine = torch.empty((E, F), dtype=torch.float64, device='cuda')
for name, fn in (('primal samples', lambda: batch.lssvr_primal_batch(nodes, u, M, 1e4, N=N, F=F, forcing=f, want_coef=False, want_fine=True, fine_out=fine)),
                 ('primal sine   ', lambda: batch.lssvr_primal_batch(nodes, u, M, 1e4, N=N, F=F, want_coef=False, want_fine=True, fine_out=fine)),
                 ('dual samples  ', lambda: batch.lssvr_dual_batch(nodes, u, M, 1e4, N=N, F=F, forcing=f, want_coef=False, want_fine=True, fine_out=fine))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nbytes = (272 + (96 if 'samples' in name else 0)) * E
    print('%s E=%d %.4f ms -> %.0f GB/s of %d B/element' % (name, E, ms, nbytes / (ms * 1e-3) / 1e9, nbytes // E))
