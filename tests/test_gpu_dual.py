"""Parity of the batched dual LSSVR kernel (K4) with the dual oracle and, by strong duality, with the
primal oracle and the primal kernel.  The reference ships no dual code."""
import numpy as np
import pytest
import torch

from hybrid_fem_lssvr_b200 import batch
from oracle import dual, fem_p1, kkt
from gpu_util import dev, jittered_mesh, oracle_coef, rel, sine_samples

pytestmark = pytest.mark.gpu
TOL = 1e-10


# Stated tolerances (relative to max |u| on the fine grid, against the primal KKT oracle, which tests/test_oracle.py pins
# to the 80-digit mpmath solve at <= 1e-14).  The bar is 1e-10 (BASELINE.json north_star) except where the DUAL FORMULA
# itself cannot deliver it in FP64: w = A^T alpha + B^T beta cancels when the collocation residual e = alpha / gamma is
# large (under-resolved forcing, random samples, one element of width 2).  Those cases carry the fixed number below -
# measured on B200, with the loss of an FP64 LU solve of the same dual system printed beside it for context - instead of a
# tolerance derived from the oracle at run time.
STATED = {
    # N = 128, M = 5 (degree 4 against 128 collocation points, frequencies up to k = 64): measured 2.2e-9 at k = 64, 4.9e-10 at
    # k = 48, <= 6.3e-11 below; an FP64 LU solve of the same 130 x 130 dual system loses 1.3e-9 / 2.2e-9 there
    'large_system/M=5': 5e-9,
    # N(0, 1) forcing samples at 12 points, M = 9: measured 3.0e-6; FP64 LU of the dual system: 1.6e-5
    'samples/random': 1e-5,
}


def check_dual(case, achieved, lu_loss=None):
    tol = STATED.get(case, TOL)
    print('DUALTOL %-40s achieved %.3e  stated %.1e  fp64-LU-of-the-dual-system %s' %
          (case, achieved, tol, 'n/a' if lu_loss is None else '%.3e' % lu_loss))
    assert achieved <= tol, (case, achieved, tol)


def _run_dual(nodes, u, M, gamma, N=12, F=32, k=1.0, samples=None, **kw):
    forcing = 'sine' if samples is None else dev(samples)
    coef, fine, status = batch.lssvr_dual_batch(dev(nodes), dev(u), M, gamma, N=N, F=F, forcing=forcing, k_freq=k,
                                                want_fine=F > 0, want_status=True, **kw)
    torch.cuda.synchronize()
    return coef.cpu().numpy(), (fine.cpu().numpy() if fine is not None else None), status.cpu().numpy()


def test_config1_dual_equals_reference_primal(golden_config1):
    """The 'Dual' script is the primal script: its outputs are the golden primal outputs."""
    g = golden_config1
    nodes, u = np.array(g['nodes']), np.array(g['fem_values']).copy()
    u[0] = u[-1] = 0.0
    coef, fine, status = _run_dual(nodes, u, g['M'], g['gamma'], N=g['N'])
    assert not status.any()
    assert rel(fine, kkt.evaluate_fine(np.array(g['coef']), 32)) <= TOL


@pytest.mark.parametrize('E', [1, 3, 4, 5, 1000, 10 ** 5])
def test_small_system_vs_oracles(E):
    """BASELINE configs[1] shape: M = 9, N = 12 (14 x 14 systems), one warp per element."""
    M, N, gamma, k = 9, 12, 1e4, 1.0
    nodes = jittered_mesh(E, seed=E)
    u = fem_p1.c_factor(2.0 / E) * np.sin(np.pi * nodes) if E > 10 else np.cos(nodes)
    coef, fine, status = _run_dual(nodes, u, M, gamma, N=N, k=k)
    assert not status.any()
    sl = slice(0, min(E, 300))
    f = sine_samples(nodes[:sl.stop + 1], N, k)
    ref_p = kkt.lssvr_primal_kkt_batch(nodes[:sl.stop + 1], u[:sl.stop + 1], f.T.copy(), M, gamma)
    ref_d = dual.lssvr_dual_batch(nodes[:sl.stop + 1], u[:sl.stop + 1], f.T.copy(), M, gamma)
    fd, fp = kkt.evaluate_fine(ref_d, 32), kkt.evaluate_fine(ref_p, 32)
    if E >= 3:
        assert rel(fd, fp) <= 1e-12     # strong duality (oracle check); E = 1 is one element of width 2
    check_dual('small_system/E=%d' % E, rel(fine[sl], fp), rel(fd, fp))


def test_dual_matches_primal_kernel_full_size():
    """BASELINE configs[1]: 1e6 elements, degree 8.  Property: dual and primal kernels agree everywhere."""
    E = 10 ** 6
    nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
    u = batch.fem_p1_solve(nodes, coarse_solver='flux')
    _, fp, _ = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True)
    err3 = batch.new_error_accumulator()
    _, fd, st = batch.lssvr_dual_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, want_status=True,
                                       err3=err3)
    torch.cuda.synchronize()
    assert int(st.sum().item()) == 0
    assert torch.max(torch.abs(fp - fd)).item() <= TOL
    l2, mx = batch.finish_error(err3)
    assert mx <= 1e-10


@pytest.mark.parametrize('M,team', [(5, 0), (9, 0), (13, 0), (17, 0), (21, 0), (25, 0), (9, 2), (25, 2), (9, 3), (25, 3)])
def test_large_system_multi_rhs(M, team):
    """BASELINE configs[4]: N = 128, R forcing frequencies sharing one factorisation.  team = 0: left-looking parity
    kernel (two 65 x 65 blocks); 3: parity split in shared memory; 2: full 130 x 130 system kernel."""
    E, N, F, gamma, R = 64, 128, 32, 1e4, 8
    batch.set_option('dual_team', team)
    nodes = np.linspace(-1, 1, E + 1) * 0.01 + 0.3          # h = 3.1e-4: k h <= 0.02, resolved for every k
    ks = np.array([1.0, 2.0, 5.0, 8.0, 16.0, 32.0, 48.0, 64.0])
    u = np.stack([np.sin(k * np.pi * nodes) for k in ks])
    err3 = torch.zeros((R, 3), dtype=torch.float64, device='cuda')
    try:
        coef, fine, status = batch.lssvr_dual_multi(dev(nodes), dev(u), dev(ks), M, gamma, N=N, F=F, want_fine=True,
                                                    want_status=True, err3=err3)
        torch.cuda.synchronize()
    finally:
        batch.set_option('dual_team', 0)
    assert rel(kkt.evaluate_fine(coef[3].cpu().numpy(), F), fine[3].cpu().numpy()) <= 1e-13     # coef and fine agree
    assert not status.cpu().numpy().any()
    for r, k in enumerate(ks):
        ref = oracle_coef(nodes, u[r], M, gamma, N, k=k)
        fp = kkt.evaluate_fine(ref, F)
        sl = slice(0, 8)
        ref_d = dual.lssvr_dual_batch(nodes[:9], u[r][:9], sine_samples(nodes[:9], N, k).T.copy(), M, gamma)
        check_dual('large_system/M=%d' % M if M == 5 else 'large_system/M=%d/k=%g/team=%d' % (M, k, team),
                   rel(fine[r].cpu().numpy(), fp), rel(kkt.evaluate_fine(ref_d, F), fp[sl]))
    e = err3.cpu().numpy()
    assert np.all(e[:, 1] < 1e-6) and np.all(e[:, 2] == 0)


@pytest.mark.parametrize('N,M', [(32, 7), (64, 9), (70, 6), (128, 12), (136, 9), (160, 10)])
@pytest.mark.parametrize('scale', [1.0, 1e-3])
def test_parity_left_looking_sizes(N, M, scale):
    """The left-looking parity kernel (default for even N with N + 2 > 32) on a coarse mesh (tau not negligible: every pivot is taken, rank = N/2 + 1) and on a fine one (early stop at the
    numerical rank), sine forcing and sampled forcing with the boundary correction: same answers as the
    shared-memory parity kernel and within the dual tolerance of the primal oracle."""
    E, F, gamma = 37, 32, 1e4
    rng = np.random.default_rng(N + M)
    nodes = jittered_mesh(E, seed=N) * scale + 0.1
    u = np.sin(np.pi * nodes) + 0.1 * rng.uniform(-1, 1, E + 1) * scale ** 2
    fs = np.exp(np.linspace(nodes[:-1], nodes[1:], N, axis=0))
    bc2 = torch.tensor([0.25, -0.5], dtype=torch.float64, device='cuda')
    ub = u + (0.25 * (nodes[-1] - nodes) - 0.5 * (nodes - nodes[0])) / (nodes[-1] - nodes[0])
    out = {}
    for team in (0, 3):
        batch.set_option('dual_team', team)
        try:
            out[team] = (_run_dual(nodes, u, M, gamma, N=N, F=F, k=1.0), _run_dual(nodes, u, M, gamma, N=N, F=F, samples=fs, bc2=bc2))
        finally:
            batch.set_option('dual_team', 0)
    for which, ref in ((0, oracle_coef(nodes, u, M, gamma, N, k=1.0)), (1, oracle_coef(nodes, ub, M, gamma, N, f_samples=fs))):
        fp = kkt.evaluate_fine(ref, F)
        f_samp = sine_samples(nodes, N, 1.0).T.copy() if which == 0 else fs.T.copy()
        ref_d = dual.lssvr_dual_batch(nodes, u if which == 0 else ub, f_samp, M, gamma)
        case = 'parity_left/%s/N=%d/M=%d/%s' % ('coarse' if scale == 1.0 else 'fine', N, M, 'sine' if which == 0 else 'samples')
        coef, fine, status = out[0][which]
        assert not status.any()
        check_dual(case, rel(fine, fp), rel(kkt.evaluate_fine(ref_d, F), fp))
        assert rel(kkt.evaluate_fine(coef, F), fine) <= 1e-13
        check_dual(case, rel(out[3][which][1], fp))


def test_samples_forcing_and_boundary_correction():
    E, M, N = 500, 9, 12
    rng = np.random.default_rng(2)
    nodes = jittered_mesh(E, seed=2)
    y = rng.uniform(-1, 1, E + 1)
    f = rng.normal(size=(N, E))
    bl, br = 0.2, -0.4
    u = y + (bl * (nodes[-1] - nodes) + br * (nodes - nodes[0])) / (nodes[-1] - nodes[0])
    bc2 = torch.tensor([bl, br], dtype=torch.float64, device='cuda')
    coef, fine, status = _run_dual(nodes, y, M, 1e4, samples=f, bc2=bc2)
    ref = oracle_coef(nodes, u, M, 1e4, N, f_samples=f)
    ref_d = dual.lssvr_dual_batch(nodes, u, f.T.copy(), M, 1e4)
    assert not status.any()
    fp = kkt.evaluate_fine(ref, 32)
    check_dual('samples/random', rel(fine, fp), rel(kkt.evaluate_fine(ref_d, 32), fp))      # random samples: large residual
    # smooth samples (a resolved forcing): the plain 1e-10 bar
    fs = np.exp(np.linspace(nodes[:-1], nodes[1:], N, axis=0))
    coef, fine, status = _run_dual(nodes, y, M, 1e4, samples=fs, bc2=bc2)
    ref = oracle_coef(nodes, u, M, 1e4, N, f_samples=fs)
    assert rel(fine, kkt.evaluate_fine(ref, 32)) <= TOL


def test_team_kernel_small_system():
    """The generic team kernel (warp per element) on the 14 x 14 system, forced through the option switch."""
    E, M, N = 300, 9, 12
    nodes = jittered_mesh(E, seed=11)
    u = np.sin(np.pi * nodes)
    batch.set_option('dual_team', 1)
    try:
        coef, fine, status = _run_dual(nodes, u, M, 1e4, N=N, k=1.0)
    finally:
        batch.set_option('dual_team', 0)
    coef2, fine2, _ = _run_dual(nodes, u, M, 1e4, N=N, k=1.0)
    ref = oracle_coef(nodes, u, M, 1e4, N, k=1.0)
    assert not status.any()
    assert rel(fine, kkt.evaluate_fine(ref, 32)) <= TOL and rel(fine2, kkt.evaluate_fine(ref, 32)) <= TOL


@pytest.mark.parametrize('mesh', ['fine', 'mixed', 'fine_high_k', 'fine_many_rhs'])
def test_factor_reuse_is_bitwise(mesh):
    """Left-looking kernel: while tau stays below half an ulp of every diagonal entry the element matrix K + tau J is
    the same floating-point matrix, and the kernel keeps the factor of the previous element.  Same bits as
    factorising every element ('dual_reuse_factor' = 0), on a fine mesh (every element after a CTA's first reuses)
    and on a mesh that alternates coarse and fine elements (the cache must be dropped and rebuilt).  'fine_high_k': one
    frequency the fine elements do not resolve to the Taylor bound (k h / 2 >= 2^-7), which keeps the elements out of the
    barrier-free pass: factor kept per team, that right-hand side pivot by pivot, the others through the moment tables.
    'fine_many_rhs': 64 frequencies, so that every CTA runs several iterations of the barrier-free loop."""
    E, N, F, M, gamma = 3001, 128, 32, 13, 1e4           # more elements than resident CTAs: several per CTA
    rng = np.random.default_rng(5)
    w = rng.uniform(0.5e-4, 1.5e-4, E)
    if mesh == 'mixed':
        w[rng.uniform(size=E) < 0.3] = 2e-2
    nodes = 0.1 + np.concatenate([[0.0], np.cumsum(w)])
    ks = np.array([1.0, 3.0, 200.0 if mesh == 'fine_high_k' else 7.0])
    if mesh == 'fine_many_rhs':          # BASELINE configs[4] count: several passes of the barrier-free loop per CTA
        ks = np.arange(1.0, 65.0)
    u = np.stack([np.sin(k * np.pi * nodes) for k in ks])
    out = {}
    for reuse in (1, 0):
        batch.set_option('dual_reuse_factor', reuse)
        try:
            out[reuse] = batch.lssvr_dual_multi(dev(nodes), dev(u), dev(ks), M, gamma, N=N, F=F, want_fine=True, want_status=True)
            torch.cuda.synchronize()
        finally:
            batch.set_option('dual_reuse_factor', 1)
    for x, y in zip(out[1], out[0]):
        assert torch.equal(x, y)
    assert not out[1][2].cpu().numpy().any()
    # and the answer is the oracle's (sample of elements, fine ones)
    fine_el = np.nonzero(w < 1e-3)[0][:6]
    for e in fine_el:
        ref = oracle_coef(nodes[e:e + 2], u[1][e:e + 2], M, gamma, N, k=ks[1])
        fp = kkt.evaluate_fine(ref, F)
        assert rel(out[1][1][1, e:e + 1].cpu().numpy(), fp) <= TOL


@pytest.mark.parametrize('M', [5, 9, 12])
@pytest.mark.parametrize('fused', [False, True])
def test_register_kernel_moment_form_matches_factorisation(M, fused):
    """N = 12 register kernel: on elements whose tau is below half an ulp of the diagonal and that resolve the forcing, the
    solution is the plan's one linear map of the right-hand side (moment tables built on the host in long double); every
    other element of the same launch takes the factorisation, out of line.  A mesh that mixes both kinds: the two forms
    agree to rounding, with and without the fused error norms (the elements that take the factorisation bit for bit), and the
    moment form sits on the primal oracle."""
    E, N, F, gamma, k = 5003, 12, 32, 1e4, 3.0
    rng = np.random.default_rng(M)
    w = rng.uniform(0.5e-4, 1.5e-4, E)
    coarse = rng.uniform(size=E) < 0.25
    w[coarse] = 3e-2
    nodes = 0.05 + np.concatenate([[0.0], np.cumsum(w)])
    u = np.sin(k * np.pi * nodes)
    out = {}
    for reuse in (1, 0):
        batch.set_option('dual_reuse_factor', reuse)
        try:
            err3 = torch.zeros(3, dtype=torch.float64, device='cuda') if fused else None
            coef, fine, status = batch.lssvr_dual_batch(dev(nodes), dev(u), M, gamma, N=N, F=F, k_freq=k, want_fine=True,
                                                        want_status=True, err3=err3)
            torch.cuda.synchronize()
            out[reuse] = (coef.cpu().numpy(), fine.cpu().numpy(), status.cpu().numpy())
        finally:
            batch.set_option('dual_reuse_factor', 1)
    assert not out[1][2].any() and not out[0][2].any()
    assert rel(out[1][1], out[0][1]) <= 1e-12 and rel(out[1][0], out[0][0]) <= 1e-12
    assert not np.array_equal(out[1][1][~coarse], out[0][1][~coarse])          # the moment form did run on the fine elements
    assert np.array_equal(out[1][1][coarse], out[0][1][coarse])                # and the coarse ones took the same factorisation
    for e in np.nonzero(~coarse)[0][:6]:
        ref = oracle_coef(nodes[e:e + 2], u[e:e + 2], M, gamma, N, k=k)
        check_dual('moment_form/M=%d' % M, rel(out[1][1][e:e + 1], kkt.evaluate_fine(ref, F)))


@pytest.mark.parametrize('F', [20, 33, 64])
@pytest.mark.parametrize('scale', [1.0, 1e-3])
def test_left_looking_kernel_many_rhs_and_fine_grids(F, scale):
    """Left-looking parity kernel with more right-hand sides than a team has threads (R = 100 > 96: two passes of the TEAM
    loop) and fine grids that are odd, not a multiple of 16, or wider than one 16-half-point block, on a coarse mesh (TEAM
    pass, factorisation per element) and a fine one (barrier-free STREAM pass): every right-hand side equals the
    single-right-hand-side launch, the fine grid is the evaluation of the returned coefficients, status is clean."""
    E, N, M, gamma, R = 700, 32, 7, 1e4, 100
    nodes = jittered_mesh(E, seed=F) * scale + 0.2
    ks = np.linspace(0.5, 3.0, R)
    u = np.stack([np.sin(k * np.pi * nodes) for k in ks])
    coef, fine, status = batch.lssvr_dual_multi(dev(nodes), dev(u), dev(ks), M, gamma, N=N, F=F, want_fine=True, want_status=True)
    torch.cuda.synchronize()
    coef, fine = coef.cpu().numpy(), fine.cpu().numpy()
    assert not status.cpu().numpy().any()
    assert not np.isnan(fine).any() and not np.isnan(coef).any()
    for r in (0, 57, 95, 96, 99):
        assert rel(fine[r], kkt.evaluate_fine(coef[r], F)) <= 1e-13
        c1, f1, _ = _run_dual(nodes, u[r], M, gamma, N=N, F=F, k=ks[r])
        assert rel(coef[r], c1) <= 1e-12 and rel(fine[r], f1) <= 1e-12
