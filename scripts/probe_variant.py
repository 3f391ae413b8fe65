"""A/B probe of the primal element kernel: run with HFL_LIB=<path of a variant build> (see build.py, HFL_VARIANT) and
without, and compare.  Prints the plain, fused-error and coefficient-writing kernel times at the headline size and a
checksum of the fine grid so that two builds can be compared for equality.  Run on the GPU box."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch, _lib
E = int(os.environ.get('E', 10 ** 7))
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1); u = torch.sin(3.141592653589793 * nodes)
fine = torch.empty((E, 32), dtype=torch.float64, device='cuda')
coef = torch.empty((E, 9), dtype=torch.float64, device='cuda')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[0], ts[len(ts) // 2]
err = batch.new_error_accumulator()
rows = [
    ('plain', lambda: batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine)),
    ('fused', lambda: batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine, err3=err)),
    ('coef ', lambda: batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=True, want_fine=True, fine_out=fine, coef_out=coef)),
]
print('lib', _lib.LIB_PATH)
for name, fn in rows:
    try:
        best, med = t(fn)
        print('%s best %.4f ms median %.4f ms' % (name, best, med))
    except TypeError as e:
        print(name, 'skipped:', e)
print('checksum fine %.17g coef %.17g' % (fine.sum().item(), coef.abs().sum().item()))
