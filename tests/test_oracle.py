"""The oracle against the golden vectors produced by the reference itself, against an
80-digit solve of the same QP, and against analytic known answers."""
import numpy as np
import pytest

from oracle import dual, fem_p1, general, kkt, kkt_mp, ref_loader


def _rhs(k):
    return lambda x: (k * np.pi) ** 2 * np.sin(k * np.pi * x)


def test_config1_coefficients_match_reference(golden_config1):
    g = golden_config1
    nodes, u = np.array(g['nodes']), np.array(g['fem_values'])
    ref = np.array(g['coef'])
    E = len(nodes) - 1
    for i in range(E):
        w = kkt.lssvr_primal_kkt(_rhs(1), [nodes[i], nodes[i + 1]], u[i], u[i + 1], g['M'], g['gamma'],
                                 is_left_boundary=(i == 0), is_right_boundary=(i == E - 1))
        assert np.max(np.abs(w - ref[i])) <= 1e-10, i      # SLSQP lands within ~1e-11 of the minimiser


def test_config1_batch_matches_scalar(golden_config1):
    g = golden_config1
    nodes, u = np.array(g['nodes']), np.array(g['fem_values']).copy()
    u[0] = 0.0
    u[-1] = 0.0
    f = fem_p1.forcing(kkt.fine_points(nodes, g['N']))
    wb = kkt.lssvr_primal_kkt_batch(nodes, u, f, g['M'], g['gamma'])
    assert np.max(np.abs(wb - np.array(g['coef']))) <= 1e-10
    for i in (0, 5, 23):
        w = kkt.lssvr_primal_kkt(_rhs(1), [nodes[i], nodes[i + 1]], u[i], u[i + 1], g['M'], g['gamma'])
        assert np.max(np.abs(wb[i] - w)) <= 1e-13      # shared ideal abscissae vs per-element mapped ones


def test_config1_evaluate_solution_matches_reference(golden_config1):
    g = golden_config1
    vals = kkt.evaluate_solution(np.array(g['nodes']), np.array(g['coef']), np.array(g['x_points']))
    # same coefficients, same numpy legval -> identical up to the last bit
    assert np.max(np.abs(vals - np.array(g['values']))) <= 1e-15


def test_golden_elements(golden_elements):
    for c in golden_elements['cases']:
        w = kkt.lssvr_primal_kkt(_rhs(c['k_freq']), [c['xmin'], c['xmax']], c['u_xmin'], c['u_xmax'], c['M'], c['gamma'])
        runs = np.array(c['coef_runs'])
        spread = max(np.max(np.abs(runs[0] - runs[1])), np.max(np.abs(runs[0] - runs[2])), np.max(np.abs(runs[1] - runs[2])))
        # P:84 starts SLSQP from an unseeded random point: the reference reproduces itself only to `spread`
        tol = max(1e-10 * max(1.0, np.max(np.abs(w))), 3.0 * spread)
        assert np.min(np.max(np.abs(runs - w), axis=1)) <= tol, c


def test_known_answer_element5():
    """element 5 of the shipped configuration."""
    nodes = np.linspace(-1, 1, 25)
    u = fem_p1.solve_fem_p1(nodes)
    w, lam = kkt.lssvr_primal_kkt(_rhs(1), [nodes[5], nodes[6]], u[5], u[6], 8, 1e4, return_multipliers=True)
    expect = [-9.8861914775214310e-01, -1.7056636175178885e-02, 5.6557859379054764e-03, 1.9498843289255627e-05,
              -2.7701001112037983e-06, -5.3048445928204440e-09, 4.7940259609960324e-10, 6.3560617850959453e-13]
    assert np.max(np.abs(w - np.array(expect))) <= 1e-14
    assert abs(lam[0] - 0.48578125578848214) <= 1e-12 and abs(lam[1] - 0.502837891963661) <= 1e-12


@pytest.mark.parametrize('h,M,N,gamma', [(1 / 12, 8, 12, 1e4), (2e-4, 9, 12, 1e4), (2e-7, 9, 12, 1e4),
                                         (0.5, 12, 12, 1e2), (1e-2, 13, 128, 1e6)])
def test_kkt_matches_mpmath(h, M, N, gamma):
    xmin = 0.3
    xmax = xmin + h
    x = np.linspace(xmin, xmax, N)
    f = _rhs(3)(x)
    A, B, _ = kkt.element_matrices(xmin, xmax, M, N)
    w, _ = kkt.solve_kkt(A, B, f, np.array([0.4, -0.2]), gamma)
    wm, _ = kkt_mp.lssvr_primal_mp(f, xmin, xmax, 0.4, -0.2, M, gamma)
    xi = np.linspace(-1, 1, 32)
    um = np.array([float(v) for v in kkt_mp.evaluate_mp(wm, xi.tolist())])
    uf = np.polynomial.legendre.legval(xi, w)
    assert np.max(np.abs(uf - um)) <= 1e-12 * max(1.0, np.max(np.abs(um)))


def test_fem_analytic_factor():
    """On a uniform mesh u_i = c(h) sin(k pi x_i) exactly; c(1/12) = 1.0000032740710444."""
    assert abs(fem_p1.c_factor(1 / 12) - 1.0000032740710444) <= 2e-15
    for n, k in ((25, 1.0), (25, 8.0), (1001, 1.0), (1001, 3.0)):
        nodes = np.linspace(-1, 1, n)
        u = fem_p1.solve_fem_p1(nodes, k)
        expect = fem_p1.c_factor(2.0 / (n - 1), k) * np.sin(k * np.pi * nodes)
        assert np.max(np.abs(u - expect)) <= 5e-11 * (n / 25), (n, k)
    u = fem_p1.solve_fem_p1(np.linspace(-1, 1, 25))
    assert abs(u[1] - (-0.25881989249446236)) <= 1e-14
    assert abs(np.max(np.abs(u - np.sin(np.pi * np.linspace(-1, 1, 25)))) - 3.2740710449452592e-06) <= 1e-13


def test_fem_solver_spread_small_at_1e4():
    """two FP64 solvers on identical data agree to ~1e-11 at 1e4 elements."""
    nodes = np.linspace(-1, 1, 10001)
    a = fem_p1.solve_fem_p1(nodes, solver='spsolve')
    b = fem_p1.solve_fem_p1(nodes, solver='banded')
    assert np.max(np.abs(a - b)) <= 1e-10


@pytest.mark.skipif(not ref_loader.reference_available(), reason='reference tree not present (GPU box)')
def test_binary128_coarse_oracle_is_pinned():
    """oracle/c fem_p1_quad (binary128 Thomas on the reference's rounded system) against a 60-digit mpmath Thomas solve
    of the same double-precision entries, and against SuperLU where SuperLU is still accurate."""
    import mpmath as mp
    from oracle import c_port
    rng = np.random.default_rng(0)
    n = 3001
    w = 1 + 0.5 * rng.uniform(-1, 1, n - 1)
    x = np.concatenate([[0.0], np.cumsum(w)])
    nodes = -1 + 2 * x / x[-1]
    for exact in (False, True):
        off, diag, b = fem_p1.assemble_p1(nodes, 2.0)
        mp.mp.dps = 60
        c = [mp.mpf(0)] * n
        g = [mp.mpf(0)] * n
        g[0] = mp.mpf(0.25)
        for i in range(1, n - 1):
            l, r = mp.mpf(off[i - 1]), mp.mpf(off[i])
            d = mp.mpf(diag[i]) if not exact else -(l + r)
            den = d - l * c[i - 1]
            c[i] = r / den
            g[i] = (mp.mpf(b[i]) - l * g[i - 1]) / den
        sol = [mp.mpf(0)] * n
        sol[n - 1] = mp.mpf(-0.5)
        for i in range(n - 2, 0, -1):
            sol[i] = g[i] - c[i] * sol[i + 1]
        q = c_port.fem_p1_quad(nodes, 2.0, exact_rowsum=exact, u_left=0.25, u_right=-0.5)
        assert q[0] == 0.25 and q[-1] == -0.5
        assert float(max(abs(mp.mpf(q[i]) - sol[i]) for i in range(1, n - 1))) <= 3e-16
    for n in (25, 1001, 10001):
        nodes = np.linspace(-1, 1, n)
        assert np.max(np.abs(c_port.fem_p1_quad(nodes) - fem_p1.solve_fem_p1(nodes))) <= 1e-10


def test_reference_function_runs_and_agrees():
    ns = ref_loader.load_reference_functions()
    np.random.seed(3)
    out = ns['lssvr_primal'](ns['poisson_rhs'], [-0.25, -1 / 6], -0.70, -0.5, 8, 1e4)
    w = kkt.lssvr_primal_kkt(ns['poisson_rhs'], [-0.25, -1 / 6], -0.70, -0.5, 8, 1e4)
    assert np.max(np.abs(out.coef - w)) <= 1e-10
    assert list(out.domain) == [-0.25, -1 / 6]


@pytest.mark.parametrize('E,M,N,k', [(24, 8, 12, 1), (1000, 9, 12, 1), (10 ** 6, 9, 12, 1), (10 ** 4, 5, 128, 1),
                                      (10 ** 4, 13, 128, 64), (10 ** 4, 25, 128, 64)])
def test_dual_oracle_strong_duality(E, M, N, k):
    """The dual system gives the primal minimiser (the only oracle a dual implementation can have here)."""
    h = 2.0 / E
    x = np.linspace(0.3, 0.3 + h, N)
    f = _rhs(k)(x)
    g = np.sin(k * np.pi * np.array([0.3, 0.3 + h]))
    wd = dual.lssvr_dual(f, 0.3, 0.3 + h, g[0], g[1], M, 1e4)
    wp = kkt.lssvr_primal_kkt_batch(np.array([0.3, 0.3 + h]), g, f[None, :], M, 1e4)[0]
    V = np.polynomial.legendre.legvander(np.linspace(-1, 1, 32), M - 1)
    assert np.max(np.abs(V @ wd - V @ wp)) <= 1e-12 * np.max(np.abs(V @ wp))


def test_c_port_matches_numpy_oracle():
    """oracle/c/hfl_oracle.c (the CPU baseline of bench.py) against the numpy restatement."""
    from oracle import c_port
    try:
        c_port.load()
    except Exception as ex:          # no compiler here: the baseline falls back to the numpy port
        pytest.skip('C port not buildable: %r' % ex)
    nodes = np.sort(np.concatenate([[-1.0, 1.0], np.random.default_rng(0).uniform(-1, 1, 499)]))
    u = fem_p1.solve_fem_p1(nodes, 2.0, solver='banded')
    assert np.max(np.abs(c_port.fem_p1(nodes, 2.0) - u)) <= 1e-10
    f = fem_p1.forcing(np.linspace(nodes[:-1], nodes[1:], 12, axis=0), 2.0).T.copy()
    ref = kkt.lssvr_primal_kkt_batch(nodes, u, f, 9, 1e4)
    coef, fine, mx = c_port.primal_batch(nodes, u, 9, 1e4, N=12, k_freq=2.0, F=32)
    fr = kkt.evaluate_fine(ref, 32)
    assert np.max(np.abs(fine - fr)) <= 1e-10 * np.max(np.abs(fr))
    assert np.max(np.abs(coef - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_general_operator_oracle_reduces_to_poisson():
    """oracle/general.py with a = 1, a' = c = 0 is oracle/kkt.py (which is pinned on the reference)."""
    nodes = np.sort(np.concatenate([[-1.0, 1.0], np.random.default_rng(0).uniform(-1, 1, 49)]))
    u = np.sin(np.pi * nodes)
    pts = np.linspace(nodes[:-1], nodes[1:], 12, axis=0).T
    f = fem_p1.forcing(pts)
    wg = general.lssvr_general_kkt_batch(nodes, u, np.ones_like(f), np.zeros_like(f), np.zeros_like(f), f, 9, 1e4)
    assert np.max(np.abs(wg - kkt.lssvr_primal_kkt_batch(nodes, u, f, 9, 1e4))) <= 1e-15


@pytest.mark.parametrize('M,h', [(3, 2e-3), (9, 2e-3), (9, 0.25), (12, 1e-5)])
def test_general_operator_oracle_matches_mpmath(M, h):
    N = 12
    nodes = np.array([0.1, 0.1 + h])
    x = np.linspace(nodes[0], nodes[1], N)[None, :]
    a, da, c = 1.0 + 0.5 * np.sin(2 * x) ** 2, np.sin(4 * x), 3.0 + x
    f = 10.0 * np.cos(3.0 * x)
    u = np.array([0.3, -0.2])
    w = general.lssvr_general_kkt_batch(nodes, u, a, da, c, f, M, 1e4)[0]
    wm = general.lssvr_general_mp(nodes[0], nodes[1], u[0], u[1], a[0], da[0], c[0], f[0], M, 1e4)
    V = np.polynomial.legendre.legvander(np.linspace(-1, 1, 32), M - 1)
    assert np.max(np.abs(V @ w - V @ wm)) <= 1e-12 * np.max(np.abs(V @ wm))


@pytest.mark.parametrize('M,N', [(9, 12), (5, 12), (13, 128), (25, 128)])
def test_moment_form_of_the_dual_solution(M, N):
    """The identity the dual kernels' table paths rest on (csrc/hfl_dual_small.cu, csrc/hfl_dual_parity.cu), restated in
    numpy and checked against the primal KKT oracle.  Once tau is negligible the parity blocks C C^T of the dual system
    are element independent, the solution is the linear map G = C_P^T (C_P C_P^T)^-1 of the right-hand side over the
    pivots of a rank-revealing Cholesky, and for a resolved sine forcing the right-hand side is a Taylor polynomial in
    y = x_b^2, so  w_par = fac sum_m y^m mom[m] + g_par mom[5]  with six element-independent vectors per parity."""
    import math
    from numpy.polynomial import legendre as npleg
    gamma, kf = 1e4, 3.0
    _, D, _ = kkt.reference_tables(M, N)
    nhd = N // 2
    blocks = []
    for par in (0, 1):
        cols = np.arange(par, M, 2)
        C = np.vstack([-D[nhd:, cols], np.ones((1, len(cols)))]).astype(np.longdouble)      # rows xi+_j, then the constraint
        K = C @ C.T
        # diagonally pivoted Cholesky, stopped at 1e-13 of the largest diagonal entry
        perm, Kw, d0 = [], K.copy(), np.max(np.diag(K))
        rem = list(range(nhd + 1))
        while rem:
            j = max(rem, key=lambda i: Kw[i, i])
            if not Kw[j, j] > 1e-13 * d0:
                break
            perm.append(j); rem.remove(j)
            l = Kw[:, j] / np.sqrt(Kw[j, j])
            Kw = Kw - np.outer(l, l)
        P = np.array(perm)
        G = C[P].T @ np.linalg.inv(K[np.ix_(P, P)].astype(np.float64)).astype(np.longdouble)    # (ma, r)
        coll = P < nhd
        c = (2 * P[coll] + 1).astype(np.longdouble)
        tay = [1.0, -1.0 / 2, 1.0 / 24, -1.0 / 720, 1.0 / 40320] if par == 0 else [1.0, -1.0 / 6, 1.0 / 120, -1.0 / 5040, 1.0 / 362880]
        mom = [G[:, coll] @ (tay[m] * c ** (2 * m + par)) for m in range(5)]
        mom.append(G[:, ~coll].sum(axis=1))
        blocks.append((cols, np.array(mom, dtype=np.float64)))
    # a few fine elements around x = 0.3
    nodes = 0.3 + np.array([0.0, 1.1e-4, 1.9e-4, 3.2e-4])
    u = np.sin(kf * np.pi * nodes) + 1e-3 * np.array([0.3, -0.2, 0.5, 0.1])
    xs = np.linspace(nodes[:-1], nodes[1:], N, axis=0).T
    ref = kkt.lssvr_primal_kkt_batch(nodes, u, (kf * np.pi) ** 2 * np.sin(kf * np.pi * xs), M, gamma)
    for e in range(3):
        xl, xr = nodes[e], nodes[e + 1]
        h = xr - xl
        xb = np.pi * kf * h * 0.5 / (N - 1)
        y, ak = xb * xb, 0.25 * h * h * (kf * np.pi) ** 2
        S, Cc = math.sin(kf * np.pi * 0.5 * (xl + xr)), math.cos(kf * np.pi * 0.5 * (xl + xr))
        w = np.zeros(M)
        for par, (cols, mom) in enumerate(blocks):
            fac = ak * S if par == 0 else ak * Cc * xb
            gpar = 0.5 * (u[e] + u[e + 1]) if par == 0 else 0.5 * (u[e + 1] - u[e])
            w[cols] = fac * sum(y ** m * mom[m] for m in range(5)) + gpar * mom[5]
        fine_w = npleg.legval(np.linspace(-1, 1, 33), w)
        fine_r = npleg.legval(np.linspace(-1, 1, 33), ref[e])
        assert np.max(np.abs(fine_w - fine_r)) <= 1e-10 * np.max(np.abs(fine_r)), (M, N, e)
