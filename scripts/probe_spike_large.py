"""SPIKE-partitioned coarse solve at bench scale on ONE GPU: G ranges of E_local elements each, solved one after the other,
interface system, correction; error of the assembled result against sin(pi x) and of the interface values."""
import os, sys, math, numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch, dist as hdist
G = int(os.environ.get('G', 2)); El = int(os.environ.get('EL', 10 ** 7)); E = G * El
for solver in ('assembled', 'assembled_exact', 'flux'):
    gathered, parts = [], []
    for r in range(G):
        nl = hdist.local_nodes_linspace(-1.0, 1.0, E, G, r)
        y, react = batch.fem_p1_solve(nl, coarse_solver=solver, want_reaction=True)
        gathered += react.cpu().tolist(); parts.append((nl, y))
    iface = batch.spike_interface_solve(gathered)
    xs = [gathered[4 * r] for r in range(G)] + [gathered[4 * (G - 1) + 1]]
    print(solver, 'interface values', iface, 'expected', [math.sin(math.pi * x) for x in xs])
    print(solver, 'reactions', [gathered[4 * r + 2] for r in range(G)], [gathered[4 * r + 3] for r in range(G)],
          'expected -/+ pi cos(pi x) - flux of the linear part')
    worst = 0.0
    for r, (nl, y) in enumerate(parts):
        u = batch.fem_apply_bc(nl, y.clone(), iface[r], iface[r + 1])
        worst = max(worst, torch.max(torch.abs(u - torch.sin(math.pi * nl))).item())
    print(solver, 'max |u - sin| after correction: %.3e' % worst)
