"""Parity of the batched primal LSSVR kernels (K2 + K3 + fused K5) with the oracle, through the C ABI.

Tolerance: 1e-10 relative (BASELINE.json north_star) on coefficients-as-functions, i.e. on the fine
grid relative to max|u|, and on the coefficient vector relative to its largest entry."""
import numpy as np
import pytest
import torch

from hybrid_fem_lssvr_b200 import batch
from oracle import fem_p1, kkt
from gpu_util import dev, jittered_mesh, oracle_coef, rel, sine_samples

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _run(nodes, u, M, gamma, N=12, F=32, k=1.0, samples=None, **kw):
    forcing = 'sine' if samples is None else dev(samples)
    coef, fine, status = batch.lssvr_primal_batch(dev(nodes), dev(u), M, gamma, N=N, F=F, forcing=forcing, k_freq=k,
                                                  want_fine=F > 0, want_status=True, **kw)
    torch.cuda.synchronize()
    return coef.cpu().numpy(), (fine.cpu().numpy() if fine is not None else None), status.cpu().numpy()


def test_config1_against_reference_golden(golden_config1):
    g = golden_config1
    nodes, u = np.array(g['nodes']), np.array(g['fem_values']).copy()
    u[0] = u[-1] = 0.0                                  # boundary-flag branches P:68-79
    coef, fine, status = _run(nodes, u, g['M'], g['gamma'], N=g['N'])
    assert not status.any()
    assert np.max(np.abs(coef - np.array(g['coef']))) <= TOL          # reference lssvr_primal output
    assert rel(fine, kkt.evaluate_fine(np.array(g['coef']), 32)) <= TOL


def test_golden_elements_against_reference(golden_elements):
    for c in golden_elements['cases']:
        nodes = np.array([c['xmin'], c['xmax']])
        u = np.array([c['u_xmin'], c['u_xmax']])
        for samples in (None, sine_samples(nodes, 12, c['k_freq'])):
            coef, _, status = _run(nodes, u, c['M'], c['gamma'], k=float(c['k_freq']), samples=samples, F=0)
            runs = np.array(c['coef_runs'])
            spread = max(np.max(np.abs(runs[i] - runs[j])) for i in range(3) for j in range(i))
            tol = max(TOL * max(1.0, np.max(np.abs(coef))), 3.0 * spread)   # see tests/golden/make_golden.py
            assert status[0] == 0
            assert np.min(np.max(np.abs(runs - coef[0]), axis=1)) <= tol, c


@pytest.mark.parametrize('M', [3, 4, 5, 8, 9, 12, 14])
@pytest.mark.parametrize('store', [1, 2, 3, 4, 5])
def test_specialised_kernel_vs_oracle(M, store):
    E, N, gamma, k = 1000 + 37, 12, 1e4, 3.0
    nodes = jittered_mesh(E, seed=M)
    u = np.random.default_rng(M).uniform(-1, 1, E + 1)
    batch.set_option('primal_store', store)
    try:
        coef, fine, status = _run(nodes, u, M, gamma, N=N, k=k)
    finally:
        batch.set_option('primal_store', 0)
    ref = oracle_coef(nodes, u, M, gamma, N, k=k)
    assert not status.any()
    assert rel(fine, kkt.evaluate_fine(ref, 32)) <= TOL
    assert np.max(np.abs(coef - ref)) <= TOL * max(1.0, np.max(np.abs(ref)))


def test_store_variants_bitwise_identical():
    E = 4097
    nodes = jittered_mesh(E, seed=1)
    u = np.sin(np.pi * nodes)
    outs = []
    for store in (1, 2, 3, 4, 5):
        batch.set_option('primal_store', store)
        outs.append(_run(nodes, u, 9, 1e4)[1])
    batch.set_option('primal_store', 0)
    assert all(np.array_equal(outs[0], o) for o in outs[1:])


@pytest.mark.parametrize('E', [1, 2, 31, 32, 33, 127, 128, 129, 4095])
def test_ragged_sizes(E):
    nodes = jittered_mesh(E, seed=E)
    u = np.cos(nodes)
    coef, fine, status = _run(nodes, u, 9, 1e4, k=2.0)
    ref = oracle_coef(nodes, u, 9, 1e4, 12, k=2.0)
    assert fine.shape == (E, 32) and not status.any()
    assert rel(fine, kkt.evaluate_fine(ref, 32)) <= TOL


@pytest.mark.parametrize('M,N,F', [(9, 12, 16), (9, 12, 5), (9, 12, 33), (9, 7, 32), (9, 13, 32), (9, 128, 32),
                                   (16, 128, 32), (20, 128, 32), (25, 128, 64), (5, 128, 32), (13, 128, 8), (32, 256, 256)])
def test_generic_shapes_vs_oracle(M, N, F):
    E, gamma, k = 300, 1e4, 2.0
    nodes = np.linspace(-1, 1, E + 1)        # h = 1/150: the bubble is resolved, cond(G22) grows with M
    u = np.sin(2 * np.pi * nodes) + 0.1 * nodes
    coef, fine, status = _run(nodes, u, M, gamma, N=N, F=F, k=k)
    ref = oracle_coef(nodes, u, M, gamma, N, k=k)
    assert not status.any()
    tol = TOL if M <= 20 else 1e-8           # cond(G22) ~ 1e7 at M = 25, N = 128
    assert rel(fine, kkt.evaluate_fine(ref, F)) <= tol


def test_samples_forcing_random_data():
    """SURVEY.md section 8d random-data variant: u ~ U(-1,1), f ~ N(0,1); nothing can be hoisted."""
    E, M, N = 2000, 9, 12
    rng = np.random.default_rng(3)
    nodes = jittered_mesh(E, seed=3)
    u = rng.uniform(-1, 1, E + 1)
    f = rng.normal(size=(N, E))
    coef, fine, status = _run(nodes, u, M, 1e4, samples=f)
    ref = oracle_coef(nodes, u, M, 1e4, N, f_samples=f)
    assert not status.any()
    assert rel(fine, kkt.evaluate_fine(ref, 32)) <= TOL
    # a resolved bubble: wide elements, strong forcing -> per-coefficient check (hard part 6)
    nodes = np.linspace(-1, 1, 9)
    u = rng.uniform(-1, 1, 9)
    f = 50.0 * rng.normal(size=(N, 8))
    coef, fine, _ = _run(nodes, u, M, 1e4, samples=f)
    ref = oracle_coef(nodes, u, M, 1e4, N, f_samples=f)
    assert np.max(np.abs(ref[:, 2:])) > 1e-2
    assert np.max(np.abs(coef - ref)) <= TOL * np.max(np.abs(ref))


@pytest.mark.parametrize('gamma', [1e-2, 1.0, 1e4, 1e8])
@pytest.mark.parametrize('E', [4, 24, 10 ** 4])
def test_gamma_and_width_range(gamma, E):
    nodes = np.linspace(-1, 1, E + 1)
    u = fem_p1.c_factor(2.0 / E) * np.sin(np.pi * nodes)
    sl = slice(0, min(E, 500))
    coef, fine, status = _run(nodes, u, 8, gamma)
    ref = oracle_coef(nodes[:sl.stop + 1], u[:sl.stop + 1], 8, gamma, 12, k=1.0)
    assert not status.any()
    assert rel(fine[sl], kkt.evaluate_fine(ref, 32)) <= TOL


def test_boundary_correction_on_the_fly():
    E = 777
    nodes = jittered_mesh(E, seed=9)
    y = np.random.default_rng(9).uniform(-1, 1, E + 1)
    bl, br = 0.37, -1.2
    L = nodes[-1] - nodes[0]
    u = y + (bl * (nodes[-1] - nodes) + br * (nodes - nodes[0])) / L
    bc2 = torch.tensor([bl, br], dtype=torch.float64, device='cuda')
    coef, fine, _ = _run(nodes, y, 9, 1e4, bc2=bc2)
    ref = oracle_coef(nodes, u, 9, 1e4, 12, k=1.0)
    assert rel(fine, kkt.evaluate_fine(ref, 32)) <= TOL
    u_dev = batch.fem_apply_bc(dev(nodes), dev(y), bl, br).cpu().numpy()
    assert np.max(np.abs(u_dev - u)) <= 1e-15 * 4


def test_breakdown_falls_back_to_linear_interpolant():
    """P:171-176: an element whose solve fails is replaced by the linear interpolant of its nodal values.  A NaN node
    makes the two elements that touch it break down deterministically (pivot test fails): status = 1 for exactly those
    two, coefficients {(u_L + u_R)/2, (u_R - u_L)/2, 0, ...}, a linear fine row, err3[2] = 2; every other element is
    bit-identical to the clean run."""
    E, M = 64, 9
    nodes = np.linspace(-1, 1, E + 1)
    u = np.sin(np.pi * nodes) + 0.05 * nodes
    clean_coef, clean_fine, clean_status = _run(nodes, u, M, 1e4, err3=batch.new_error_accumulator())   # same kernel instantiation
    assert not clean_status.any()
    bad_nodes = nodes.copy()
    bad_nodes[20] = np.nan
    for form in (batch.lssvr_primal_batch, batch.lssvr_dual_batch):
        err3 = batch.new_error_accumulator()
        coef, fine, status = form(dev(bad_nodes), dev(u), M, 1e4, N=12, F=32, want_fine=True, want_status=True, err3=err3)
        torch.cuda.synchronize()
        coef, fine, status = coef.cpu().numpy(), fine.cpu().numpy(), status.cpu().numpy()
        bad = np.array([19, 20])
        assert np.array_equal(np.nonzero(status)[0], bad)
        assert int(err3[2].item()) == 2
        assert np.array_equal(coef[bad, 2:], np.zeros((2, M - 2)))
        assert np.array_equal(coef[bad, 0], 0.5 * (u[bad] + u[bad + 1]))
        assert np.array_equal(coef[bad, 1], 0.5 * (u[bad + 1] - u[bad]))
        xi = np.linspace(-1, 1, 32)
        lin = coef[bad, 0:1] + coef[bad, 1:2] * xi[None, :]
        assert np.max(np.abs(fine[bad] - lin)) <= 1e-15
        good = np.setdiff1d(np.arange(E), bad)
        if form is batch.lssvr_primal_batch:
            assert np.array_equal(coef[good], clean_coef[good]) and np.array_equal(fine[good], clean_fine[good])
        else:
            assert np.max(np.abs(fine[good] - clean_fine[good])) <= 1e-10


def test_fused_error_matches_standalone_and_numpy():
    E, k = 5000, 1.0
    nodes = jittered_mesh(E, seed=4)
    u = np.sin(np.pi * nodes) + 1e-6 * np.cos(3 * nodes)
    err3 = batch.new_error_accumulator()
    coef, fine, _ = _run(nodes, u, 9, 1e4, k=k, err3=err3)
    l2_f, mx_f = batch.finish_error(err3)
    l2_s, mx_s = batch.finish_error(batch.error_fine(dev(nodes), dev(fine), k))
    x = kkt.fine_points(nodes, 32)
    d = fine - np.sin(np.pi * x)
    w = np.ones(32); w[0] = w[-1] = 0.5
    l2_n = np.sqrt(np.sum((np.diff(nodes) / 31.0)[:, None] * w[None, :] * d * d))
    mx_n = np.max(np.abs(d))
    assert abs(l2_f - l2_n) <= 1e-9 * l2_n and abs(l2_s - l2_n) <= 1e-9 * l2_n
    assert abs(mx_f - mx_n) <= 1e-9 * mx_n + 1e-15 and abs(mx_s - mx_n) <= 1e-9 * mx_n + 1e-15


def test_full_size_properties():
    """BASELINE configs[2]: 1e7 elements, M = 9, N = 12, F = 32.  Size-independent properties:
    the boundary rows hold (end values = nodal values), neighbouring elements agree at shared nodes,
    the reconstruction is as close to sin(pi x) as the nodal data allow, and a random sample of
    elements matches the oracle."""
    E = 10 ** 7
    nodes = batch.mesh_linspace(-1.0, 1.0, E + 1)
    u = torch.sin(np.pi * nodes)
    err3 = batch.new_error_accumulator()
    _, fine, status = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True,
                                               want_status=True, err3=err3)
    torch.cuda.synchronize()
    assert int(status.sum().item()) == 0
    assert torch.max(torch.abs(fine[:, 0] - u[:-1])).item() <= 1e-12
    assert torch.max(torch.abs(fine[:, -1] - u[1:])).item() <= 1e-12
    assert torch.max(torch.abs(fine[1:, 0] - fine[:-1, -1])).item() <= 1e-12
    l2, mx = batch.finish_error(err3)
    assert mx <= 1e-12 and l2 <= 1e-12
    idx = np.sort(np.random.default_rng(0).choice(E - 1, 2000, replace=False))
    nh = nodes.cpu().numpy()
    uh = u.cpu().numpy()
    fh = fine[torch.from_numpy(idx).cuda()].cpu().numpy()
    for j, e in enumerate(idx[:200]):
        ref = oracle_coef(nh[e:e + 2], uh[e:e + 2], 9, 1e4, 12, k=1.0)
        assert rel(fh[j:j + 1], kkt.evaluate_fine(ref, 32)) <= TOL


@pytest.mark.parametrize('M,F', [(9, 16), (9, 64), (3, 64), (14, 16), (14, 64)])
def test_specialised_other_fine_sizes(M, F):
    """F = 16 and F = 64 have their own instantiations (TMA stores, fused error norms, coefficient output)."""
    E, N, gamma, k = 2000 + 5, 12, 1e4, 2.0
    nodes = jittered_mesh(E, seed=F + M)
    u = np.sin(2 * np.pi * nodes) + 1e-6 * np.cos(3 * nodes)      # a nodal perturbation, so that the error norms are not round-off
    err3 = batch.new_error_accumulator()
    coef, fine, status = _run(nodes, u, M, gamma, N=N, F=F, k=k, err3=err3)
    ref = oracle_coef(nodes, u, M, gamma, N, k=k)
    fr = kkt.evaluate_fine(ref, F)
    assert not status.any()
    assert rel(fine, fr) <= TOL
    assert np.max(np.abs(coef - ref)) <= TOL * max(1.0, np.max(np.abs(ref)))
    _, fine2, _ = batch.lssvr_primal_batch(dev(nodes), dev(u), M, gamma, N=N, F=F, k_freq=k, want_coef=False, want_fine=True)
    assert np.max(np.abs(fine - fine2.cpu().numpy())) <= 1e-14            # plain and fused-error instantiations (different FMA contraction)
    x = kkt.fine_points(nodes, F)
    d = fine - np.sin(k * np.pi * x)
    w = np.ones(F); w[0] = w[-1] = 0.5
    l2 = np.sqrt(np.sum((np.diff(nodes) / (F - 1.0))[:, None] * w[None, :] * d * d))
    l2g, mxg = batch.finish_error(err3)
    assert abs(l2g - l2) <= 1e-9 * l2 and abs(mxg - np.max(np.abs(d))) <= 1e-9 * np.max(np.abs(d)) + 1e-15
