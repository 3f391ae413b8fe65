"""General elliptic operator -(a u')' + c u = f: the per-element kernel against its oracle,
against the Poisson kernels when a = 1, a' = c = 0, and on a manufactured solution."""
import numpy as np
import pytest
import torch

from hybrid_fem_lssvr_b200 import batch
from oracle import fem_p1, general, kkt
from gpu_util import dev, jittered_mesh, rel

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pts(nodes, N):
    return np.linspace(nodes[:-1], nodes[1:], N, axis=0)          # [N, E]


@pytest.mark.parametrize('M', [3, 5, 8, 9, 12])
@pytest.mark.parametrize('E', [1, 33, 1000])
def test_against_oracle(M, E):
    N, F, gamma = 12, 32, 1e4
    nodes = jittered_mesh(E, seed=E + M)
    x = _pts(nodes, N)
    rng = np.random.default_rng(M)
    a = 1.0 + 0.5 * np.sin(2.0 * x) ** 2
    da = 2.0 * np.sin(2.0 * x) * np.cos(2.0 * x)
    c = 3.0 + x
    f = 10.0 * np.cos(3.0 * x) + rng.normal(size=x.shape) * 0.0
    u = rng.uniform(-1, 1, E + 1)
    coef, fine, status = batch.lssvr_general_batch(dev(nodes), dev(u), dev(a), dev(f), M, gamma, N=N, F=F, da=dev(da),
                                                   c=dev(c), want_fine=True, want_status=True)
    torch.cuda.synchronize()
    ref = general.lssvr_general_kkt_batch(nodes, u, a.T.copy(), da.T.copy(), c.T.copy(), f.T.copy(), M, gamma)
    assert not status.cpu().numpy().any()
    fr = kkt.evaluate_fine(ref, F)
    assert rel(fine.cpu().numpy(), fr) <= TOL
    assert np.max(np.abs(coef.cpu().numpy() - ref)) <= TOL * max(1.0, np.max(np.abs(ref)))


def test_reduces_to_the_poisson_kernel():
    E, M, N = 4097, 9, 12
    nodes = jittered_mesh(E, seed=3)
    x = _pts(nodes, N)
    f = (np.pi ** 2) * np.sin(np.pi * x)
    u = np.sin(np.pi * nodes)
    ones = np.ones_like(x)
    _, fg, _ = batch.lssvr_general_batch(dev(nodes), dev(u), dev(ones), dev(f), M, 1e4, N=N, F=32, want_fine=True)
    _, fp, _ = batch.lssvr_primal_batch(dev(nodes), dev(u), M, 1e4, N=N, F=32, forcing=dev(f), want_fine=True)
    _, fs, _ = batch.lssvr_primal_batch(dev(nodes), dev(u), M, 1e4, N=N, F=32, forcing='sine', want_fine=True)
    assert torch.max(torch.abs(fg - fp)).item() <= 1e-12
    assert torch.max(torch.abs(fg - fs)).item() <= 1e-12


def test_manufactured_solution():
    """u = sin(pi x), a = 1 + x^2/2, c = 2: with exact nodal values the reconstruction is within 1e-9 of u."""
    E, M, N, F = 200, 9, 12, 32
    nodes = np.linspace(-1, 1, E + 1)
    x = _pts(nodes, N)
    a, da, c = 1.0 + 0.5 * x * x, x, 2.0 + 0.0 * x
    uu, du, d2 = np.sin(np.pi * x), np.pi * np.cos(np.pi * x), -np.pi ** 2 * np.sin(np.pi * x)
    f = -(a * d2 + da * du) + c * uu
    _, fine, status = batch.lssvr_general_batch(dev(nodes), dev(np.sin(np.pi * nodes)), dev(a), dev(f), M, 1e4, N=N, F=F,
                                                da=dev(da), c=dev(c), want_fine=True, want_status=True)
    assert not status.cpu().numpy().any()
    assert np.max(np.abs(fine.cpu().numpy() - np.sin(np.pi * kkt.fine_points(nodes, F)))) <= 1e-9


def _coef_funcs():
    a = lambda x: 1.0 + 0.5 * x * x
    da = lambda x: x
    c = lambda x: 2.0 + 0.0 * x
    f = lambda x: -((1.0 + 0.5 * x * x) * (-np.pi ** 2 * np.sin(np.pi * x)) + x * np.pi * np.cos(np.pi * x)) + 2.0 * np.sin(np.pi * x)
    return a, da, c, f


def _gauss_samples(nodes, fn):
    h = np.diff(nodes)
    return np.stack([fn(nodes[:-1] + h * batch.GAUSS_X[q]) for q in range(2)])          # [2, E]


@pytest.mark.parametrize('n', [2, 3, 50, 2048, 2049, 5001])
def test_general_coarse_solve_vs_oracle(n):
    a, da, c, f = _coef_funcs()
    nodes = jittered_mesh(n - 1, seed=n)
    u = batch.fem_p1_solve_general(dev(nodes), dev(_gauss_samples(nodes, a)), dev(_gauss_samples(nodes, f)),
                                   cq=dev(_gauss_samples(nodes, c)), u_left=0.1, u_right=-0.2).cpu().numpy()
    ref = fem_p1.solve_fem_p1_general(nodes, a, c, f, 0.1, -0.2)
    assert u[0] == 0.1 and u[-1] == -0.2
    assert np.max(np.abs(u - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref)))


def test_general_coarse_solve_reduces_to_poisson():
    n = 4097
    nodes = np.linspace(-1, 1, n)
    one = lambda x: 1.0 + 0.0 * x
    f = lambda x: np.pi ** 2 * np.sin(np.pi * x)
    ug = batch.fem_p1_solve_general(dev(nodes), dev(_gauss_samples(nodes, one)), dev(_gauss_samples(nodes, f))).cpu().numpy()
    up = batch.fem_p1_solve(dev(nodes)).cpu().numpy()
    assert np.max(np.abs(ug - up)) <= 1e-11


def test_general_pipeline_end_to_end():
    """Hybrid method for -(a u')' + c u = f with u = sin(pi x): coarse P1 solve -> element LSSVR -> fine grid.
    The enhancement removes the interpolation error: what is left is the (O(h^2)) nodal error of the P1 solve."""
    a, da, c, f = _coef_funcs()
    E, M, N, F = 400, 9, 12, 32
    nodes = np.linspace(-1, 1, E + 1)
    d_nodes = dev(nodes)
    u = batch.fem_p1_solve_general(d_nodes, dev(_gauss_samples(nodes, a)), dev(_gauss_samples(nodes, f)),
                                   cq=dev(_gauss_samples(nodes, c)))
    nodal_err = np.max(np.abs(u.cpu().numpy() - np.sin(np.pi * nodes)))
    x = _pts(nodes, N)
    _, fine, status = batch.lssvr_general_batch(d_nodes, u, dev(a(x)), dev(f(x)), M, 1e4, N=N, F=F, da=dev(da(x)), c=dev(c(x)),
                                                want_fine=True, want_status=True)
    assert not status.cpu().numpy().any()
    fine_err = np.max(np.abs(fine.cpu().numpy() - np.sin(np.pi * kkt.fine_points(nodes, F))))
    assert 1e-7 < nodal_err < 1e-3
    assert fine_err <= 1.05 * nodal_err + 1e-9


@pytest.mark.parametrize('M', [13, 16, 24])
def test_generic_kernel_large_M(M):
    """M beyond the register-resident instantiations (3..12): the run-time-bound kernel, same oracle."""
    N, F, gamma, E = 32, 32, 1e4, 150
    nodes = jittered_mesh(E, seed=M)
    x = _pts(nodes, N)
    a = 1.0 + 0.5 * np.sin(2.0 * x) ** 2
    da = 2.0 * np.sin(2.0 * x) * np.cos(2.0 * x)
    c = 3.0 + x
    f = 10.0 * np.cos(3.0 * x)
    u = np.random.default_rng(M).uniform(-1, 1, E + 1)
    coef, fine, status = batch.lssvr_general_batch(dev(nodes), dev(u), dev(a), dev(f), M, gamma, N=N, F=F, da=dev(da),
                                                   c=dev(c), want_fine=True, want_status=True)
    ref = general.lssvr_general_kkt_batch(nodes, u, a.T.copy(), da.T.copy(), c.T.copy(), f.T.copy(), M, gamma)
    assert not status.cpu().numpy().any()
    assert rel(fine.cpu().numpy(), kkt.evaluate_fine(ref, F)) <= 1e-9      # cond of the degree-23 Gram matrix shows at 1e-10
    assert rel(kkt.evaluate_fine(coef.cpu().numpy(), F), fine.cpu().numpy()) <= 1e-12


@pytest.mark.parametrize('F', [32, 16, 33])
def test_fast_and_fallback_fine_paths_agree(F):
    """F = 32 takes the Horner + TMA kernel, other F the shared-memory transpose kernel: same coefficients, fine rows equal
    to the coefficients evaluated on the grid; ragged element counts around the CTA tile of 128."""
    for E in (1, 127, 128, 129, 1000):
        nodes = jittered_mesh(E, seed=E)
        x = _pts(nodes, 12)
        a, f = 1.0 + x * x, np.cos(2.0 * x)
        u = np.sin(nodes)
        coef, fine, _ = batch.lssvr_general_batch(dev(nodes), dev(u), dev(a), dev(f), 9, 1e4, N=12, F=F, da=dev(2.0 * x),
                                                  want_fine=True)
        assert rel(fine.cpu().numpy(), kkt.evaluate_fine(coef.cpu().numpy(), F)) <= 1e-13


@pytest.mark.parametrize('G', [2, 4])
def test_partitioned_general_solve_on_one_gpu(G):
    """dist.fem_p1_solve_general_distributed with every "rank" on one device: the three local solves, the six end
    residuals and the interface system reproduce the global general-operator solve."""
    from hybrid_fem_lssvr_b200 import dist as hdist
    E = 6000 + 7
    nodes = jittered_mesh(E, seed=G)
    af, cf, ff = (lambda x: 1.0 + 0.5 * np.sin(2.0 * x) ** 2), (lambda x: 2.0 + x), (lambda x: 10.0 * np.cos(3.0 * x))
    ref = fem_p1.solve_fem_p1_general(nodes, af, cf, ff, u_left=0.3, u_right=-0.2)

    def samples(x, fn):
        x0, h = x[:-1], np.diff(x)
        return dev(np.stack([fn(x0 + h * batch.GAUSS_X[0]), fn(x0 + h * batch.GAUSS_X[1])]))
    recs, parts = [], []
    for r in range(G):
        e0, e1 = hdist.partition(E, G, r)
        x = nodes[e0:e1 + 1]
        nl, aq, cq, fq = dev(x), samples(x, af), samples(x, cf), samples(x, ff)
        zero = torch.zeros_like(fq)
        y = batch.fem_p1_solve_general(nl, aq, fq, cq)
        v = batch.fem_p1_solve_general(nl, aq, zero, cq, u_left=1.0, u_right=0.0)
        w = batch.fem_p1_solve_general(nl, aq, zero, cq, u_left=0.0, u_right=1.0)
        ry = hdist._general_end_residuals(nl, y, aq, cq, fq, True)
        rv = hdist._general_end_residuals(nl, v, aq, cq, fq, False)
        rw = hdist._general_end_residuals(nl, w, aq, cq, fq, False)
        recs.append([ry[0], ry[1], rv[0], rv[1], rw[0], rw[1]])
        parts.append((e0, e1, y, v, w))
    U = hdist.general_interface_solve(np.array(recs), 0.3, -0.2)
    for r, (e0, e1, y, v, w) in enumerate(parts):
        u = (y + U[r] * v + U[r + 1] * w).cpu().numpy()
        assert np.max(np.abs(u - ref[e0:e1 + 1])) <= 1e-10
    # single "rank": the wrapper itself
    u1 = hdist.fem_p1_solve_general_distributed(dev(nodes), samples(nodes, af), samples(nodes, ff), samples(nodes, cf),
                                                u_left=0.3, u_right=-0.2).cpu().numpy()
    assert np.max(np.abs(u1 - ref)) <= 1e-10
