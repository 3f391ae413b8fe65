"""General-operator element kernel (-(a u')' + c u = f), 1e7 elements, M = 9, N = 12, F = 32: timing and ncu target."""
import os, sys, math, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E, N, M, F = int(os.environ.get('E', 10 ** 7)), 12, 9, 32
nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
u = torch.sin(math.pi * nodes)
x = nodes[:-1].unsqueeze(0) + (nodes[1:] - nodes[:-1]).unsqueeze(0) * torch.linspace(0, 1, N, dtype=torch.float64, device='cuda').unsqueeze(1)
a = (1.0 + 0.5 * x).contiguous(); da = torch.full_like(x, 0.5); c = (2.0 + torch.cos(x)).contiguous()
up, d1, d2 = torch.sin(math.pi * x), math.pi * torch.cos(math.pi * x), -math.pi ** 2 * torch.sin(math.pi * x)
f = (-(da * d1 + a * d2) + c * up).contiguous()
del x, up, d1, d2
fn = lambda: batch.lssvr_general_batch(nodes, u, a, f, M, 1e4, N=N, F=F, da=da, c=c, want_coef=False, want_fine=True)
for _ in range(3): fn()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
    s.record(); out = fn(); t.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(t))
ts.sort()
fine = out[1]
xs = nodes[:-1].unsqueeze(1) + (nodes[1:] - nodes[:-1]).unsqueeze(1) * torch.linspace(0, 1, F, dtype=torch.float64, device='cuda')
print('general E=%d best %.4f ms median %.4f ms -> %.3e element solves/s, %.0f GB/s of 656 B; max|u - sin| %.2e'
      % (E, ts[0], ts[5], E / (ts[5] * 1e-3), 656 * E / (ts[5] * 1e-3) / 1e9, (fine - torch.sin(math.pi * xs)).abs().max().item()))
