"""hfl_evaluate_points (evaluate_solution, P:184-211) at scale: 1e7 elements, 1e8 query points, random and sorted."""
import os, sys, math, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch
E, P = 10 ** 7, int(os.environ.get('P', 10 ** 8))
for mesh in ('uniform', 'jittered'):
    nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
    if mesh == 'jittered':
        g = torch.Generator(device='cuda').manual_seed(0)
        h = (nodes[1:] - nodes[:-1]) * (1.0 + 0.5 * (2.0 * torch.rand(E, generator=g, device='cuda', dtype=torch.float64) - 1.0))
        nodes = torch.cat([torch.tensor([-1.0], device='cuda', dtype=torch.float64), -1.0 + torch.cumsum(h, 0) * (2.0 / h.sum())])
    u = torch.sin(math.pi * nodes)
    coef, _, _ = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=0)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = (2.2 * torch.rand(P, generator=g, device='cuda', dtype=torch.float64) - 1.1)
    for name, xs in (('random', x), ('sorted', torch.sort(x).values)):
        for _ in range(2): out = batch.evaluate_points(nodes, coef, xs)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = batch.evaluate_points(nodes, coef, xs); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        inside = (xs >= -1) & (xs <= 1)
        err = (out[inside] - torch.sin(math.pi * xs[inside])).abs().max().item()
        print('%s mesh, %s points: %.2f ms -> %.3e points/s, max|u - sin| inside %.2e, checksum %.17g' % (mesh, name, ms, P / (ms * 1e-3), err, out.sum().item()))
