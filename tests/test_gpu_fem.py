"""Parity of the coarse P1 FEM solve (K1), the mesh generator and the SPIKE pieces with the oracle."""
import numpy as np
import pytest
import torch

from hybrid_fem_lssvr_b200 import batch
from oracle import fem_p1
from gpu_util import dev, jittered_mesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('a,b,n', [(-1.0, 1.0, 25), (-1.0, 1.0, 10 ** 6 + 1), (0.1, 0.7, 1000), (-3.0, 2.0, 2)])
def test_linspace_bit_exact(a, b, n):
    out = batch.mesh_linspace(a, b, n).cpu().numpy()
    assert np.array_equal(out, np.linspace(a, b, n))
    if n > 10:
        part = batch.mesh_linspace(a, b, n, 7, n - 9).cpu().numpy()
        assert np.array_equal(part, np.linspace(a, b, n)[7:n - 2])


@pytest.mark.parametrize('solver', ['assembled', 'flux', 'assembled_exact'])
@pytest.mark.parametrize('n,k', [(2, 1.0), (3, 1.0), (9, 1.0), (25, 1.0), (25, 8.0), (2048, 1.0), (2049, 1.0), (2050, 2.0),
                                 (4097, 1.0), (10001, 1.0), (10001, 3.0)])
def test_uniform_mesh_vs_oracle(solver, n, k):
    """1e-10 relative nodal parity with the restated reference solve is well-posed up to ~1e4 nodes
    (SURVEY.md section 0 fact 9)."""
    nodes = np.linspace(-1, 1, n)
    u = batch.fem_p1_solve(dev(nodes), k_freq=k, coarse_solver=solver).cpu().numpy()
    ref = fem_p1.solve_fem_p1(nodes, k)
    # nodal parity bar of BASELINE.json; well-posed up to ~1e4 nodes (SuperLU itself is 2.9e-12 off the exact
    # solution of its own system at 1e4 nodes, 1.2e-10 at 1e5 - scripts/probe_fem_accuracy.py)
    tol = 1e-10
    assert np.max(np.abs(u - ref)) <= tol * max(1.0, np.max(np.abs(ref)))
    if n >= 9:
        expect = fem_p1.c_factor(2.0 / (n - 1), k) * np.sin(k * np.pi * nodes)   # analytic discrete solution
        # the assembled system itself is a few 1e-11 away from the analytic solution at ~2e3 nodes (rounded diagonal)
        assert np.max(np.abs(u - expect)) <= (tol if solver == 'assembled' else 1e-12)


@pytest.mark.parametrize('solver', ['assembled', 'flux', 'assembled_exact'])
@pytest.mark.parametrize('n', [100, 5000, 2048 * 3 + 1])
def test_jittered_mesh_and_dirichlet_data(solver, n):
    nodes = jittered_mesh(n - 1, seed=n)
    ul, ur = 0.3, -0.8
    u = batch.fem_p1_solve(dev(nodes), k_freq=2.0, u_left=ul, u_right=ur, coarse_solver=solver).cpu().numpy()
    ref0 = fem_p1.solve_fem_p1(nodes, 2.0)
    ref = ref0 + (ul * (nodes[-1] - nodes) + ur * (nodes - nodes[0])) / (nodes[-1] - nodes[0])
    assert u[0] == ul and u[-1] == ur
    assert np.max(np.abs(u - ref)) <= 1e-10


@pytest.mark.parametrize('solver', ['assembled', 'flux'])
@pytest.mark.parametrize('n,R', [(2, 3), (25, 1), (2049, 5), (10001, 64)])
def test_multi_rhs_rows_equal_single_solves(solver, n, R):
    """hfl_fem_p1_solve_multi: row r is bit-identical to the single solve with k_freqs[r] (BASELINE configs[4])."""
    nodes = dev(jittered_mesh(n - 1, seed=n) if n > 2 else np.array([-1.0, 1.0]))
    ks = torch.tensor([0.5 + 1.25 * r for r in range(R)], dtype=torch.float64, device='cuda')
    u = batch.fem_p1_solve_multi(nodes, ks, u_left=0.25, u_right=-0.5, coarse_solver=solver)
    assert u.shape == (R, n)
    for r in range(R):
        one = batch.fem_p1_solve(nodes, k_freq=float(ks[r]), u_left=0.25, u_right=-0.5, coarse_solver=solver)
        assert torch.equal(u[r], one), (solver, n, r)
    with pytest.raises(Exception):
        batch.fem_p1_solve_multi(nodes, ks, out=torch.empty(R * n + 1, dtype=torch.float64, device='cuda'))


@pytest.mark.parametrize('solver,exact', [('assembled', False), ('assembled_exact', True)])
@pytest.mark.parametrize('mesh', ['uniform-10001', 'uniform-100001', 'jittered-60000'])
def test_against_binary128_solution_of_the_same_system(solver, exact, mesh):
    """The reference's rounded system (and its exact-row-sum variant) solved in IEEE binary128 (oracle/c, pinned against a
    60-digit solve in tests/test_oracle.py) is the exact solution to double precision.  The GPU elimination in row-sum
    form stays within 1e-12 of it where SuperLU / LAPACK on the same matrix are 1e-10 away (1e5 nodes)."""
    from oracle import c_port
    kind, n = mesh.split('-')
    n = int(n)
    nodes = np.linspace(-1, 1, n) if kind == 'uniform' else jittered_mesh(n - 1, seed=7)
    ref = c_port.fem_p1_quad(nodes, 2.0, exact_rowsum=exact, u_left=0.25, u_right=-0.5)
    u = batch.fem_p1_solve(dev(nodes), k_freq=2.0, u_left=0.25, u_right=-0.5, coarse_solver=solver).cpu().numpy()
    err = np.max(np.abs(u - ref))
    print('%s %s: |gpu - binary128| = %.2e' % (solver, mesh, err))
    assert err <= 1e-12


def test_headline_size_against_binary128():
    """BASELINE configs[2] size, the mode bench.py runs at N = 1 ('assembled'): 1e7 + 1 nodes against the binary128
    solution of the same rounded system.  Double precision direct solvers are 1e-8 apart from each other here
    (SURVEY.md fact 9); the system itself is 4.5e-5 away from sin(pi x) (rounded diagonal), which the flux form removes."""
    from oracle import c_port
    n = 10 ** 7 + 1
    nodes = np.linspace(-1, 1, n)
    ref = c_port.fem_p1_quad(nodes)
    d_nodes = dev(nodes)
    u = batch.fem_p1_solve(d_nodes, coarse_solver='assembled').cpu().numpy()
    err = np.max(np.abs(u - ref))
    exact = np.sin(np.pi * nodes)
    print('n = 1e7+1: |gpu assembled - binary128| = %.2e; |binary128 - sin| = %.2e' % (err, np.max(np.abs(ref - exact))))
    assert err <= 2e-11
    assert u[0] == 0.0 and u[-1] == 0.0
    ue = batch.fem_p1_solve(d_nodes, coarse_solver='assembled_exact').cpu().numpy()
    refe = c_port.fem_p1_quad(nodes, exact_rowsum=True)
    erre = np.max(np.abs(ue - refe))
    print('n = 1e7+1: |gpu exact row sums - binary128| = %.2e; |binary128 - sin| = %.2e' % (erre, np.max(np.abs(refe - exact))))
    assert erre <= 1e-12 and np.max(np.abs(refe - exact)) <= 1e-12


def test_peer_exchange_timeout_is_loud():
    """A peer that never delivers: the receive spin expires, the status word is set, the missing doubles arrive as NaN,
    the fused interface solve writes NaN, and the host-side check raises (no silent stale payload)."""
    from hybrid_fem_lssvr_b200 import _lib, dist as hdist
    lib = _lib.load()
    nbytes = int(lib.hfl_peer_buffer_bytes())
    bufs = [torch.zeros(nbytes // 8, dtype=torch.int64, device='cuda') for _ in range(2)]
    me = hdist.PeerExchange(rank=0, buffers=[b.data_ptr() for b in bufs])
    batch.set_option('peer_spin_log2', 10)
    try:
        out = me.all_gather(dev(np.array([1.0, 2.0, 3.0])), hdist.PeerExchange.CHANNEL_ERROR)
        bc2 = me.spike_exchange(dev(np.array([-1.0, 0.0, 0.1, 0.2])), 0.0, 0.0)
        torch.cuda.synchronize()
    finally:
        batch.set_option('peer_spin_log2', 24)
    o = out.cpu().numpy()
    assert np.array_equal(o[0], [1.0, 2.0, 3.0]) and np.isnan(o[1]).all()
    assert np.isnan(bc2.cpu().numpy()).all()
    assert me.timed_out()
    with pytest.raises(_lib.HflError):
        me.check()
    with pytest.raises(_lib.HflError):
        hdist.finish_gathered_error(out, exchange=me)


def test_large_mesh_reported_spread():
    """Beyond ~1e4 nodes FP64 solvers disagree with each other on identical data.  The row-sum (GTH) elimination
    of the assembled solve stays ~1e-14 from the exact solution of the reference's rounded system, so it must be
    at least as close to SuperLU as LAPACK is; and both must show the SAME deviation from the analytic discrete
    solution (2.9e-6 at 1e6 nodes: that deviation belongs to the rounded matrix, not to a solver), which the
    flux form does not have."""
    n = 10 ** 6 + 1
    nodes = np.linspace(-1, 1, n)
    ref = fem_p1.solve_fem_p1(nodes)
    alt = fem_p1.solve_fem_p1(nodes, solver='banded')
    spread = np.max(np.abs(ref - alt))
    d_nodes = dev(nodes)
    u = batch.fem_p1_solve(d_nodes, coarse_solver='assembled').cpu().numpy()
    assert np.max(np.abs(u - ref)) <= 1.5 * spread
    uf = batch.fem_p1_solve(d_nodes, coarse_solver='flux').cpu().numpy()
    exact = fem_p1.c_factor(2.0 / (n - 1)) * np.sin(np.pi * nodes)
    assert np.max(np.abs(uf - exact)) <= 1e-11
    da, ds = np.max(np.abs(u - exact)), np.max(np.abs(ref - exact))
    assert abs(da - ds) <= 0.05 * ds        # same inherent deviation of the assembled system
    print('spread spsolve-vs-banded %.3e, gpu-assembled-vs-spsolve %.3e, gpu-flux-vs-analytic %.3e'
          % (spread, np.max(np.abs(u - ref)), np.max(np.abs(uf - exact))))


@pytest.mark.parametrize('n', [2048 * 1024 + 5, 10 ** 7 + 1])
def test_top_level_chunks_and_full_size(n):
    """Top-level chunk length > 1 (more than 1024 tiles) and the BASELINE size; property: the flux form
    reproduces the analytic discrete solution, the assembled form satisfies its own equations."""
    nodes = batch.mesh_linspace(-1.0, 1.0, n)
    uf = batch.fem_p1_solve(nodes, coarse_solver='flux')
    h = 2.0 / (n - 1)
    exact = fem_p1.c_factor(h) * torch.sin(np.pi * nodes)
    assert torch.max(torch.abs(uf - exact)).item() <= 1e-10
    ua = batch.fem_p1_solve(nodes, coarse_solver='assembled')
    # the assembled system has cond ~ n^2: round-off, not discretisation, sets its accuracy here (fact 9);
    # it must still be a sane solution with exact boundary rows
    diff = torch.max(torch.abs(ua - uf)).item()
    print('n=%d assembled-vs-flux %.3e' % (n, diff))
    assert diff <= 1e-2
    assert ua[0].item() == 0.0 and ua[-1].item() == 0.0


@pytest.mark.parametrize('solver', ['assembled_exact', 'flux'])
def test_spike_partitioned_solve_at_bench_scale(solver):
    """Two ranges of 2e6 elements each (h = 5e-7), solved one after the other on one device: interface values and the
    corrected nodal values against sin(pi x) (discretisation error ~h^2 = 3e-13).  The reference's rounded diagonal
    ('assembled') is 1e-5 off here and 1e-3 off at 1e7 elements per range: the partitioned solve uses the exact
    row sums (HFL_COARSE_ASSEMBLED_EXACT) or the flux form."""
    import math
    from hybrid_fem_lssvr_b200 import dist as hdist
    G, El = 2, 2 * 10 ** 6
    gathered, parts = [], []
    for r in range(G):
        nl = hdist.local_nodes_linspace(-1.0, 1.0, G * El, G, r)
        y, react = batch.fem_p1_solve(nl, coarse_solver=solver, want_reaction=True)
        gathered += react.cpu().tolist()
        parts.append((nl, y))
    iface = batch.spike_interface_solve(gathered)
    assert abs(iface[1]) <= 1e-11
    for r, (nl, y) in enumerate(parts):
        u = batch.fem_apply_bc(nl, y.clone(), iface[r], iface[r + 1])
        assert torch.max(torch.abs(u - torch.sin(math.pi * nl))).item() <= 1e-11


@pytest.mark.parametrize('solver', ['assembled', 'flux', 'assembled_exact'])
@pytest.mark.parametrize('G', [2, 4, 8])
def test_spike_partitioned_solve_on_one_gpu(solver, G):
    """All ranks' data on one device: local zero-Dirichlet solves + interface system + linear correction
    reproduce the global solve (host and device interface solvers)."""
    from hybrid_fem_lssvr_b200 import _lib, dist as hdist
    E = 8000 + 3          # nodal parity with the restated reference is well-posed up to ~1e4 nodes
    nodes = jittered_mesh(E, seed=G)
    ref = fem_p1.solve_fem_p1(nodes, 1.0)
    gathered, ys = [], []
    for r in range(G):
        e0, e1 = hdist.partition(E, G, r)
        nl = dev(nodes[e0:e1 + 1])
        y, react = batch.fem_p1_solve(nl, coarse_solver=solver, want_reaction=True)
        rec = react.cpu().tolist()
        assert rec[0] == nodes[e0] and rec[1] == nodes[e1]
        gathered += rec
        ys.append((nl, y))
    iface = batch.spike_interface_solve(gathered)
    dg = dev(np.array(gathered))
    for r in range(G):
        e0, e1 = hdist.partition(E, G, r)
        bc2 = torch.empty(2, dtype=torch.float64, device='cuda')
        _lib.check(_lib.load().hfl_spike_interface_solve_device(G, batch._ptr(dg), 0.0, 0.0, r, batch._ptr(bc2),
                                                                batch._stream()), 'device interface solve')
        assert np.max(np.abs(np.array(bc2.cpu().tolist()) - np.array(iface[r:r + 2]))) <= 1e-14
        nl, y = ys[r]
        u = batch.fem_apply_bc(nl, y.clone(), iface[r], iface[r + 1]).cpu().numpy()
        assert np.max(np.abs(u - ref[e0:e1 + 1])) <= 1e-10


def test_nodal_error_norms():
    n, k = 101, 4.0
    nodes = np.linspace(-1, 1, n)
    u = batch.fem_p1_solve(dev(nodes), k_freq=k)
    l2, mx = batch.finish_error(batch.error_nodal(dev(nodes), u, k))
    d = u.cpu().numpy() - np.sin(k * np.pi * nodes)
    w = np.zeros(n); w[1:-1] = 0.5 * (nodes[2:] - nodes[:-2]); w[0] = 0.5 * (nodes[1] - nodes[0]); w[-1] = w[0]
    assert mx > 1e-7 and abs(mx - np.max(np.abs(d))) <= 1e-9 * mx
    assert abs(l2 - np.sqrt(np.sum(w * d * d))) <= 1e-9 * l2


def test_peer_allgather_protocol_on_one_gpu():
    """hfl_peer_allgather with three "ranks" in one process: three buffers on one device, one stream per rank (the
    kernels wait for each other, so they must be able to run concurrently).  Several epochs on two channels, both
    widths used by the partitioned path; every rank must see every rank's doubles bit for bit."""
    import ctypes as C
    from hybrid_fem_lssvr_b200 import _lib, dist as hdist
    lib = _lib.load()
    G = 3
    nbytes = int(lib.hfl_peer_buffer_bytes())
    bufs = [torch.zeros(nbytes // 8, dtype=torch.int64, device='cuda') for _ in range(G)]
    ranks = [hdist.PeerExchange(rank=r, buffers=[b.data_ptr() for b in bufs]) for r in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    rng = np.random.default_rng(0)
    torch.cuda.synchronize()
    for epoch in range(5):
        for channel, W in ((hdist.PeerExchange.CHANNEL_INTERFACE, 4), (hdist.PeerExchange.CHANNEL_ERROR, 3)):
            data = rng.normal(size=(G, W)) * 10.0 ** rng.integers(-300, 300, size=(G, W))
            outs = []
            for r in range(G):
                with torch.cuda.stream(streams[r]):
                    outs.append(ranks[r].all_gather(dev(data[r]), channel))
            torch.cuda.synchronize()
            for r in range(G):
                assert np.array_equal(outs[r].cpu().numpy(), data), (epoch, channel, r)
                assert not ranks[r].timed_out()
    # the fused exchange + interface solve against the host solve of the same records
    xs = np.array([-1.0, -0.2, 0.3, 1.0])
    recs = np.array([[xs[r], xs[r + 1], rng.normal(), rng.normal()] for r in range(G)])
    expect = batch.spike_interface_solve(recs.reshape(-1).tolist(), 0.25, -0.5)
    got = []
    for r in range(G):
        with torch.cuda.stream(streams[r]):
            got.append(ranks[r].spike_exchange(dev(recs[r]), 0.25, -0.5))
    torch.cuda.synchronize()
    for r in range(G):
        assert np.max(np.abs(got[r].cpu().numpy() - np.array(expect[r:r + 2]))) <= 1e-14
        assert not ranks[r].timed_out()


@pytest.mark.parametrize('n', [2048 * 300 + 5, 2048 * 6144, 2048 * 6144 + 1])
def test_large_meshes_three_solvers(n):
    """Meshes of 0.6e6 and 1.26e7 nodes (300 / 6144 / 6145 tiles: the top-level CTA holds the tile-head rows in shared
    memory for the first, in the workspace for the others): the exact-row-sum mode against the flux form (both
    well-conditioned, so they must agree to rounding) and against sin(pi x); the reference's rounded mode for sanity."""
    nodes = batch.mesh_linspace(-1.0, 1.0, n)
    uf = batch.fem_p1_solve(nodes, coarse_solver='flux')
    ue = batch.fem_p1_solve(nodes, coarse_solver='assembled_exact')
    assert torch.max(torch.abs(ue - uf)).item() <= 1e-12
    assert torch.max(torch.abs(ue - torch.sin(np.pi * nodes))).item() <= 1e-12
    ua = batch.fem_p1_solve(nodes, coarse_solver='assembled')
    assert torch.max(torch.abs(ua - uf)).item() <= 1e-2 and ua[0].item() == 0.0 and ua[-1].item() == 0.0


@pytest.mark.parametrize('n', [4 * 10 ** 6 + 1, 16384 * 1100 + 3])
def test_top_level_with_rows_in_the_workspace(n):
    """The top-level CTA keeps the tile-head rows in shared memory while they fit; beyond ~8.5e7 nodes they live in the
    workspace.  fem_top_smem_kb = 64 forces that path at testable sizes: 245 tiles (one head per top-level thread) and
    1101 tiles (chunks of two heads per thread); same answers as with the rows in shared memory, bit for bit."""
    nodes = batch.mesh_linspace(-1.0, 1.0, n)
    ref = batch.fem_p1_solve(nodes, coarse_solver='assembled_exact')
    refa = batch.fem_p1_solve(nodes, coarse_solver='assembled')
    batch.set_option('fem_top_smem_kb', 64)
    try:
        ue = batch.fem_p1_solve(nodes, coarse_solver='assembled_exact')
        ua = batch.fem_p1_solve(nodes, coarse_solver='assembled')
    finally:
        batch.set_option('fem_top_smem_kb', 227)
    assert torch.equal(ue, ref) and torch.equal(ua, refa)
    assert torch.max(torch.abs(ue - torch.sin(np.pi * nodes))).item() <= 1e-12
