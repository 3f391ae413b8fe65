"""Store-path probe: the element kernel with the solve switched off (primal_debug=2) for every store variant,
next to plain write streams.  Run on the GPU box."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E = 10 ** 7
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1); u = torch.sin(3.141592653589793 * nodes)
fine = torch.empty((E, 32), dtype=torch.float64, device='cuda')
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for store in (1, 2, 3, 4, 5):
    batch.set_option('primal_store', store)
    row = []
    for dbg in (0, 2):
        batch.set_option('primal_debug', dbg)
        row.append(t(lambda: batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine)))
    print('store %d: full %.4f ms, solve off %.4f ms' % (store, row[0], row[1]))
batch.set_option('primal_debug', 0); batch.set_option('primal_store', 0)
