"""The reference's call surface (P:20-22, P:107-211) served by the GPU library: same entry points,
host numpy in and out, compared with what the reference itself produced (tests/golden)."""
import numpy as np
import pytest
import torch

from hybrid_fem_lssvr_b200 import FEMLSSVRPrimalSolver, batch, lssvr_primal, poisson_rhs
from oracle import fem_p1, kkt
from gpu_util import dev

pytestmark = pytest.mark.gpu


def test_shipped_main_program(golden_config1):
    """P:214-225: 25 nodes, M = 8, gamma = 1e4, 201 test points."""
    g = golden_config1
    solver = FEMLSSVRPrimalSolver(25, lssvr_M=8, lssvr_gamma=1e4, global_domain=(-1, 1))
    solver.solve()
    assert np.array_equal(solver.fem_nodes, np.array(g['nodes']))
    assert np.max(np.abs(solver.fem_values - np.array(g['fem_values']))) <= 1e-13
    assert len(solver.lssvr_functions) == 24
    for i in (0, 5, 23):
        fn = solver.lssvr_functions[i]
        assert np.max(np.abs(fn.coef - np.array(g['coef'][i]))) <= 1e-10
        assert list(fn.domain) == [g['nodes'][i], g['nodes'][i + 1]]
        assert abs(fn(g['nodes'][i]) - (0.0 if i == 0 else g['fem_values'][i])) <= 1e-12   # callable like Legendre
    xs = np.array(g['x_points'])
    vals = solver.evaluate_solution(xs)
    assert vals.shape == xs.shape
    assert np.max(np.abs(vals - np.array(g['values']))) <= 1e-10     # reference evaluate_solution output
    computed = solver.evaluate_solution(np.linspace(-1, 1, 201))
    assert abs(np.max(np.abs(computed - np.sin(np.pi * np.linspace(-1, 1, 201)))) - 3.274e-06) <= 2e-9


def test_evaluate_points_bit_exact_with_numpy_legval(golden_config1):
    """Same coefficients in -> the device legval reproduces Legendre.__call__ bit for bit."""
    g = golden_config1
    xs = np.array(g['x_points'])
    out = batch.evaluate_points(dev(g['nodes']), dev(np.array(g['coef'])), dev(xs)).cpu().numpy()
    assert np.array_equal(out, np.array(g['values']))
    rng = np.random.default_rng(0)
    nodes = np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, 500)]))
    coef = rng.normal(size=(501, 11))
    xq = np.concatenate([rng.uniform(-1.2, 1.2, 5000), nodes])
    out = batch.evaluate_points(dev(nodes), dev(coef), dev(xq)).cpu().numpy()
    assert np.array_equal(out, kkt.evaluate_solution(nodes, coef, xq))
    # uniform and nearly uniform meshes take the guess-and-gallop lookup: query points on the nodes (left element wins),
    # a hair to either side of them, outside the mesh, and meshes of 1..40 elements
    for E in (1, 2, 3, 31, 32, 33, 40, 2000):
        for jitter in (0.0, 0.3):
            nodes = np.linspace(-0.7, 1.9, E + 1)
            if jitter:
                nodes[1:-1] += jitter * (nodes[1] - nodes[0]) * rng.uniform(-1, 1, E - 1)
            coef = rng.normal(size=(E, 6))
            xq = np.concatenate([nodes, np.nextafter(nodes, -np.inf), np.nextafter(nodes, np.inf), rng.uniform(-1.0, 2.2, 3000)])
            out = batch.evaluate_points(dev(nodes), dev(coef), dev(xq)).cpu().numpy()
            assert np.array_equal(out, kkt.evaluate_solution(nodes, coef, xq)), (E, jitter)


def test_lssvr_primal_signature_and_golden(golden_elements):
    for c in golden_elements['cases'][:8]:
        k = c['k_freq']
        out = lssvr_primal(lambda x: (k * np.pi) ** 2 * np.sin(k * np.pi * x), [c['xmin'], c['xmax']], c['u_xmin'],
                           c['u_xmax'], c['M'], c['gamma'])
        assert isinstance(out, np.polynomial.Legendre) and list(out.domain) == [c['xmin'], c['xmax']]
        runs = np.array(c['coef_runs'])
        spread = max(np.max(np.abs(runs[i] - runs[j])) for i in range(3) for j in range(i))
        assert np.min(np.max(np.abs(runs - out.coef), axis=1)) <= max(1e-10 * max(1.0, np.max(np.abs(out.coef))), 3 * spread)
    # boundary flags (P:68-79): the FEM value is ignored at the global boundary
    a = lssvr_primal(poisson_rhs, [-1.0, -0.9], 0.123, -0.3, 8, 1e4, is_left_boundary=True)
    b = lssvr_primal(poisson_rhs, [-1.0, -0.9], 0.0, -0.3, 8, 1e4)
    assert np.array_equal(a.coef, b.coef)
    c_ = lssvr_primal(poisson_rhs, [-0.95, -0.9], 0.123, -0.3, 8, 1e4, is_left_boundary=True)   # not at the boundary
    assert abs(c_(-0.95) - 0.123) <= 1e-12


def test_custom_rhs_and_flux_solver():
    """A host callable as rhs_func (P:20) goes through sampled forcing; 'flux' coarse solver agrees."""
    s1 = FEMLSSVRPrimalSolver(41, lssvr_M=9, lssvr_gamma=1e4)
    s1.solve()
    s2 = FEMLSSVRPrimalSolver(41, lssvr_M=9, lssvr_gamma=1e4, rhs_func=lambda x: np.pi ** 2 * np.sin(np.pi * x),
                              coarse_solver='flux')
    s2.solve()
    xs = np.linspace(-1, 1, 333)
    assert np.max(np.abs(s1.evaluate_solution(xs) - s2.evaluate_solution(xs))) <= 1e-10
    ref = fem_p1.solve_fem_p1(np.linspace(-1, 1, 41))
    assert np.max(np.abs(s1.fem_values - ref)) <= 1e-12


def test_non_sine_rhs_func_reaches_both_stages():
    """a callable that is NOT the sine family must drive the coarse solve too (P:129-136 use the same
    rhs as P:45).  -u'' = 12 x^2 - 2 with u(+-1) = 0 has the solution u = x^2 - x^4; the coarse nodal values are
    checked against the general-operator oracle and the hybrid solution against the exact one."""
    f = lambda x: 12.0 * x ** 2 - 2.0
    s = FEMLSSVRPrimalSolver(65, lssvr_M=7, lssvr_gamma=1e6, rhs_func=f)
    s.solve()
    nodes = np.linspace(-1, 1, 65)
    ref = fem_p1.solve_fem_p1_general(nodes, lambda x: np.ones_like(x), lambda x: np.zeros_like(x), f)
    assert np.max(np.abs(s.fem_values - ref)) <= 1e-12
    xs = np.linspace(-1, 1, 501)
    exact = xs ** 2 - xs ** 4
    # the 2-point Gauss load is exact for this quadratic forcing, so the P1 nodal values are exact and the degree-6
    # element solves reproduce the quartic
    assert np.max(np.abs(s.fem_values - (nodes ** 2 - nodes ** 4))) <= 1e-12
    assert np.max(np.abs(s.evaluate_solution(xs) - exact)) <= 1e-9
    # and the sine family given as a callable still equals the device path
    k = FEMLSSVRPrimalSolver(33, lssvr_M=8, lssvr_gamma=1e4, rhs_func=poisson_rhs, k_freq=5.0)   # k_freq is ignored: rhs_func wins
    k.solve()
    assert np.max(np.abs(k.fem_values - fem_p1.solve_fem_p1(np.linspace(-1, 1, 33), 1.0))) <= 1e-12


def test_evaluate_solution_keeps_shape_for_any_layout_and_small_M_is_refused():
    s = FEMLSSVRPrimalSolver(25, lssvr_M=8, lssvr_gamma=1e4)
    s.solve()
    x = np.linspace(-1.2, 1.2, 24).reshape(4, 6)
    base = s.evaluate_solution(x)
    for view in (x.T, np.asfortranarray(x), x[:, ::2]):
        out = s.evaluate_solution(view)
        assert out.shape == view.shape
        assert np.array_equal(out, s.evaluate_solution(np.ascontiguousarray(view)))
    assert np.array_equal(s.evaluate_solution(x.T), base.T)
    with pytest.raises(ValueError):
        lssvr_primal(poisson_rhs, [-1.0, -0.9], 0.0, -0.3, 2, 1e4)


def test_streamed_forcing_samples_equal_a_plain_copy():
    """host_api.stream_samples (pinned double buffer + copy stream) delivers the same [N, E] array as evaluating the
    callable on all collocation points at once and copying it in one piece; chunk sizes that do and do not divide E."""
    from hybrid_fem_lssvr_b200 import host_api
    E, N = 100003, 12
    nodes = np.sort(np.random.default_rng(5).uniform(-1, 1, E + 1))
    f = lambda x: np.exp(x) * np.cos(7.0 * x)
    ref = f(np.linspace(nodes[:-1], nodes[1:], N, axis=0))
    for chunk in (E, 4096, 33333, 1):
        if chunk == 1:
            got = host_api.stream_samples(f, nodes[:50], nodes[1:51], np.linspace(0, 1, N), 'cuda', chunk_elements=1)
            assert np.array_equal(got.cpu().numpy(), ref[:, :50])
            continue
        got = host_api.stream_samples(f, nodes[:-1], nodes[1:], np.linspace(0, 1, N), 'cuda', chunk_elements=chunk)
        assert np.array_equal(got.cpu().numpy(), ref)
    gq = host_api.stream_samples(f, nodes[:-1], nodes[1:], batch.GAUSS_X, 'cuda', chunk_elements=7777)
    h = np.diff(nodes)
    assert np.array_equal(gq.cpu().numpy(), f(np.stack([nodes[:-1] + batch.GAUSS_X[0] * h, nodes[:-1] + batch.GAUSS_X[1] * h])))


def test_dual_form_through_the_class():
    """The 'Dual' script's entry points (D:100-203 are P:107-211): same class, form='dual'."""
    xs = np.linspace(-1, 1, 201)
    a = FEMLSSVRPrimalSolver(25, lssvr_M=8, lssvr_gamma=1e4)
    a.solve()
    b = FEMLSSVRPrimalSolver(25, lssvr_M=8, lssvr_gamma=1e4, form='dual')
    b.solve()
    assert np.max(np.abs(a.evaluate_solution(xs) - b.evaluate_solution(xs))) <= 1e-10
    assert not b.element_status.any()


def test_empty_and_single_element_batches():
    nodes = dev(np.array([0.25]))
    u = dev(np.array([0.5]))
    coef, fine, status = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_fine=True, want_status=True)
    assert coef.shape == (0, 9) and fine.shape == (0, 32) and status.shape == (0,)
    coef, fine, status = batch.lssvr_dual_batch(nodes, u, 9, 1e4, N=12, F=32, want_fine=True, want_status=True)
    assert coef.shape == (0, 9)
    s = FEMLSSVRPrimalSolver(2, lssvr_M=5, lssvr_gamma=1e2)      # one element, both ends Dirichlet
    s.solve()
    assert len(s.lssvr_functions) == 1 and abs(s.evaluate_solution(np.array([-1.0]))[0]) <= 1e-12


def test_host_pipeline_matches_device_path():
    """The host-buffer pipeline of bench.py's e2e leg (pinned host in, chunked D2H out) against one device-resident launch."""
    from hybrid_fem_lssvr_b200 import host_api
    E = 10007
    pipe = host_api.HostPipeline(E, 9, 1e4, N=12, F=32, chunks=3)
    nodes_h = pipe.pinned_nodes()
    nodes_h.copy_(torch.from_numpy(np.linspace(-1.0, 1.0, E + 1)))
    fine_h, u_h, (l2, mx) = pipe.run(nodes_h)
    nodes = nodes_h.cuda()
    u = batch.fem_p1_solve(nodes)
    err = batch.new_error_accumulator()
    _, fine, _ = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, err3=err)
    assert torch.equal(u.cpu(), u_h)
    assert torch.equal(fine.cpu(), fine_h)
    l2d, mxd = batch.finish_error(err)
    assert abs(l2 - l2d) <= 1e-9 * l2d and mx == mxd
    assert pipe.d2h_bytes == 8 * E * 32 + 8 * (E + 1) + 24 and pipe.h2d_bytes == 8 * (E + 1)
    fine_d, u_h2, (l2r, mxr) = pipe.run(nodes_h, fetch_fine=False)          # fine grid left on the device
    assert torch.equal(fine_d.cpu(), fine_h) and torch.equal(u_h2, u_h) and mxr == mx
