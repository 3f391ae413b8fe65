"""Timing + parity summary of the BASELINE.json configs that are not the headline bench line.

configs[0]  shipped script (25 nodes, M=8), through the drop-in class, against the reference's golden outputs
configs[1]  dual LSSVR, 1e6 elements, degree 8
configs[4]  Legendre degree sweep 4-24 (M = 5..25), N = 128 collocation points, R = 64 forcing frequencies, dual form

Prints one JSON line per measurement (kept under profiles/).  Run on the GPU box:  python scripts/bench_configs.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import FEMLSSVRPrimalSolver, batch  # noqa: E402
from oracle import fem_p1, kkt  # noqa: E402


def t_ms(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def config0():
    g = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'config1.json')))
    t0 = time.perf_counter()
    s = FEMLSSVRPrimalSolver(25, lssvr_M=8, lssvr_gamma=1e4, global_domain=(-1, 1))
    s.solve()
    xs = np.linspace(-1, 1, 201)
    vals = s.evaluate_solution(xs)
    wall = time.perf_counter() - t0
    ref = np.array(g['values'])[:201]
    print(json.dumps({'config': 0, 'what': 'shipped script through the drop-in class (25 nodes, M=8, gamma=1e4, 201 points)',
                      'wall_s_incl_first_call': wall, 'max_abs_diff_vs_reference_output': float(np.max(np.abs(vals - ref))),
                      'max_err_vs_sin': float(np.max(np.abs(vals - np.sin(np.pi * xs))))}))


def config1():
    E, M = 10 ** 6, 9
    nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
    for solver in ('assembled', 'flux'):
        u = batch.fem_p1_solve(nodes, coarse_solver=solver)
        fine = torch.empty((E, 32), dtype=torch.float64, device='cuda')
        err = batch.new_error_accumulator()
        ms = t_ms(lambda: batch.lssvr_dual_batch(nodes, u, M, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine))
        batch.lssvr_dual_batch(nodes, u, M, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=fine, err3=err)
        l2, mx = batch.finish_error(err)
        _, fp, _ = batch.lssvr_primal_batch(nodes, u, M, 1e4, N=12, F=32, want_coef=False, want_fine=True)
        k1 = t_ms(lambda: batch.fem_p1_solve(nodes, coarse_solver=solver, out=u))
        print(json.dumps({'config': 1, 'what': 'dual LSSVR, 1e6 elements, M=9, N=12, F=32, coarse solver ' + solver,
                          'K4_dual_ms': ms, 'K1_ms': k1, 'element_solves_per_s': E / (ms * 1e-3),
                          'hbm_GBps_algorithmic': 272 * E / (ms * 1e-3) / 1e9,
                          'dual_vs_primal_max_abs': float(torch.max(torch.abs(fp - fine)).item()),
                          'fine_l2_vs_sin': l2, 'fine_max_vs_sin': mx}))


def config2_variants(E=10 ** 7):
    """the headline shape on a non-uniform mesh and with random data, where nothing could be hoisted
    across elements even in principle (the kernel never hoists: same launch, same code path)."""
    M, N, F = 9, 12, 32
    fine = torch.empty((E, F), dtype=torch.float64, device='cuda')
    gen = torch.Generator(device='cuda').manual_seed(0)
    # (a) uniform mesh, device forcing (the bench.py workload)
    nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
    u = torch.sin(np.pi * nodes)
    ms = t_ms(lambda: batch.lssvr_primal_batch(nodes, u, M, 1e4, N=N, F=F, want_coef=False, want_fine=True, fine_out=fine))
    print(json.dumps({'config': 2, 'variant': 'uniform mesh, sine forcing on device', 'K2K3_ms': ms,
                      'element_solves_per_s': E / (ms * 1e-3), 'GBps_algorithmic_272B': 272 * E / (ms * 1e-3) / 1e9}))
    # (b) jittered mesh: widths (2/E)(1 + 0.5 U(-1,1)), renormalised
    w = 1.0 + 0.5 * (2.0 * torch.rand(E, dtype=torch.float64, device='cuda', generator=gen) - 1.0)
    x = torch.cat([torch.zeros(1, dtype=torch.float64, device='cuda'), torch.cumsum(w, 0)])
    nodes_j = (-1.0 + 2.0 * x / x[-1]).contiguous()
    u_j = torch.sin(np.pi * nodes_j)
    ms = t_ms(lambda: batch.lssvr_primal_batch(nodes_j, u_j, M, 1e4, N=N, F=F, want_coef=False, want_fine=True, fine_out=fine))
    err = batch.new_error_accumulator()
    batch.lssvr_primal_batch(nodes_j, u_j, M, 1e4, N=N, F=F, want_coef=False, want_fine=True, fine_out=fine, err3=err)
    l2, mx = batch.finish_error(err)
    print(json.dumps({'config': 2, 'variant': 'non-uniform (jittered) mesh, sine forcing on device', 'K2K3_ms': ms,
                      'element_solves_per_s': E / (ms * 1e-3), 'GBps_algorithmic_272B': 272 * E / (ms * 1e-3) / 1e9,
                      'fine_max_vs_sin': mx}))
    # (c) random data: u ~ U(-1,1), f ~ N(0,1) streamed as samples [N][E] (+96 B/element)
    u_r = 2.0 * torch.rand(E + 1, dtype=torch.float64, device='cuda', generator=gen) - 1.0
    f_r = torch.randn((N, E), dtype=torch.float64, device='cuda', generator=gen)
    ms = t_ms(lambda: batch.lssvr_primal_batch(nodes_j, u_r, M, 1e4, N=N, F=F, forcing=f_r, want_coef=False, want_fine=True,
                                               fine_out=fine))
    idx = torch.arange(0, 200, device='cuda')
    ref = kkt.evaluate_fine(kkt.lssvr_primal_kkt_batch(nodes_j[:201].cpu().numpy(), u_r[:201].cpu().numpy(),
                                                        f_r[:, :200].cpu().numpy().T.copy(), M, 1e4), F)
    worst = float(np.max(np.abs(fine[idx].cpu().numpy() - ref)) / np.max(np.abs(ref)))
    print(json.dumps({'config': 2, 'variant': 'jittered mesh, random nodal data, random forcing samples', 'K2K3_ms': ms,
                      'element_solves_per_s': E / (ms * 1e-3), 'GBps_algorithmic_368B': 368 * E / (ms * 1e-3) / 1e9,
                      'max_rel_diff_vs_oracle_sample': worst}))


def config4(E=10 ** 4, R=64):
    nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
    ks = torch.arange(1, R + 1, dtype=torch.float64, device='cuda')
    u = torch.empty((R, E + 1), dtype=torch.float64, device='cuda')
    t_loop = t_ms(lambda: [batch.fem_p1_solve(nodes, k_freq=float(k), coarse_solver='flux', out=u[k - 1]) for k in range(1, R + 1)], n=2, warm=1)
    u_loop = u.clone()
    t0 = t_ms(lambda: batch.fem_p1_solve_multi(nodes, ks, coarse_solver='flux', out=u), n=10, warm=3)
    assert torch.equal(u, u_loop), 'multi-RHS coarse solve differs from the per-frequency solves'
    nh = nodes.cpu().numpy()
    for M in (5, 9, 13, 17, 21, 25):
        err = torch.zeros((R, 3), dtype=torch.float64, device='cuda')
        ms = t_ms(lambda: batch.lssvr_dual_multi(nodes, u, ks, M, 1e4, N=128, F=32, want_coef=False, want_fine=True), n=3, warm=1)
        _, fine, st = batch.lssvr_dual_multi(nodes, u, ks, M, 1e4, N=128, F=32, want_coef=False, want_fine=True,
                                             want_status=True, err3=err)
        torch.cuda.synchronize()
        # parity on a sample: primal oracle for 3 frequencies, first 50 elements
        worst = 0.0
        for r in (0, 7, 63):
            k = float(r + 1)
            f = fem_p1.forcing(np.linspace(nh[:50], nh[1:51], 128, axis=0), k).T.copy()
            ref = kkt.evaluate_fine(kkt.lssvr_primal_kkt_batch(nh[:51], u[r, :51].cpu().numpy(), f, M, 1e4), 32)
            worst = max(worst, float(np.max(np.abs(fine[r, :50].cpu().numpy() - ref)) / np.max(np.abs(ref))))
        e = err.cpu().numpy()
        print(json.dumps({'config': 4, 'what': 'dual, N=128 (130x130 systems), R=64 forcings sin(k pi x) k=1..64, E=%d, M=%d' % (E, M),
                          'K4_dual_multi_ms': ms, 'K1_64_rhs_one_call_ms': t0, 'K1_64_separate_calls_ms': t_loop, 'rhs_solves_per_s': E * R / (ms * 1e-3),
                          'element_factorisations_per_s': E / (ms * 1e-3), 'failed_elements': int(st.sum().item()),
                          'max_rel_diff_vs_primal_oracle_sample': worst, 'fine_max_vs_sin_worst_k': float(e[:, 1].max())}))


if __name__ == '__main__':
    which = sys.argv[1:] or ['0', '1', '2', '4']
    if '0' in which:
        config0()
    if '1' in which:
        config1()
    if '2' in which:
        config2_variants()
    if '4' in which:
        config4()
