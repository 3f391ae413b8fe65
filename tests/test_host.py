"""Host-side logic that needs no GPU: sampling of rhs callables, partitioning, and the SPIKE
exchange over a 2-process gloo group (local solves stood in by the oracle, which only tests may do)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybrid_fem_lssvr_b200 import api
from hybrid_fem_lssvr_b200 import dist as hdist
from oracle import fem_p1


def test_collocation_points_match_reference_linspace():
    nodes = np.array([-1.0, -0.3, 0.2, 1.0])
    pts = api._collocation_points(nodes, 12)
    assert pts.shape == (12, 3)
    for e in range(3):
        assert np.array_equal(pts[:, e], np.linspace(nodes[e], nodes[e + 1], 12))   # P:40, bit for bit


def test_sample_rhs_vectorised_scalar_and_constant():
    pts = np.linspace(0, 1, 6).reshape(2, 3)
    assert np.array_equal(api._sample_rhs(api.poisson_rhs, pts), np.pi ** 2 * np.sin(np.pi * pts))
    assert np.array_equal(api._sample_rhs(lambda x: 2.5, pts), np.full((2, 3), 2.5))

    def scalar_only(x):
        if isinstance(x, np.ndarray):
            raise TypeError('scalar only')
        return 3.0 * x
    assert np.allclose(api._sample_rhs(scalar_only, pts), 3.0 * pts)


@pytest.mark.parametrize('E,world', [(10, 1), (10, 3), (7, 8), (10 ** 8, 8), (24, 5)])
def test_partition_is_contiguous_and_balanced(E, world):
    ranges = [hdist.partition(E, world, r) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == E
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        hdist.partition(E, world, world)


def _jittered_mesh(E, seed=5):
    """Non-uniform mesh of [-1, 1]: widths (2/E)(1 + 0.5 U(-1, 1)), renormalised."""
    w = 1.0 + 0.5 * np.random.default_rng(seed).uniform(-1, 1, E)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return -1.0 + 2.0 * x / x[-1]


def _oracle_local_solve(nodes, k_freq, coarse_solver, out=None):
    """Stand-in for the CUDA local solve: zero-Dirichlet solve on the range + end-node residuals."""
    x = nodes.numpy()
    y = fem_p1.solve_fem_p1(x, k_freq, solver='banded')
    off, _, _ = fem_p1.assemble_p1(x, k_freq)
    kloc = -off
    h = np.diff(x)
    gx = fem_p1._GX
    kp2 = (k_freq * np.pi) ** 2

    def contrib(e, right):
        tot = 0.0
        for q in range(2):
            fq = kp2 * np.sin(k_freq * np.pi * (h[e] * gx[q] + x[e]))
            tot += fq * (gx[q] if right else 1.0 - gx[q]) * h[e] * 0.5
        return tot
    rl = contrib(0, False) + kloc[0] * (y[1] - y[0])
    rr = contrib(len(h) - 1, True) + kloc[-1] * (y[-2] - y[-1])
    return torch.from_numpy(y), torch.tensor([x[0], x[-1], rl, rr], dtype=torch.float64)


def _spike_worker(rank, world, port, E, k_freq, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        nodes_g = _jittered_mesh(E)
        e0, e1 = hdist.partition(E, world, rank)
        nodes = torch.from_numpy(nodes_g[e0:e1 + 1].copy())
        y, bc2 = hdist.fem_p1_solve_distributed(nodes, k_freq=k_freq, local_solve=_oracle_local_solve)
        x = nodes.numpy()
        L = x[-1] - x[0]
        u = y.numpy() + (bc2[0].item() * (x[-1] - x) + bc2[1].item() * (x - x[0])) / L
        err = torch.tensor([float(rank + 1), float(rank), 1.0 if rank == 1 else 0.0], dtype=torch.float64)
        l2, mx, failed = hdist.reduce_error(err)
        assert hdist.finish_gathered_error(hdist.gather_error(err)) == (l2, mx, failed)
        np.save(os.path.join(out_dir, 'u%d.npy' % rank), u)
        np.save(os.path.join(out_dir, 'r%d.npy' % rank), np.array([l2, mx, failed]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [2, 3])
def test_spike_exchange_over_gloo(tmp_path, world):
    E, k = 301, 2.0
    mp.spawn(_spike_worker, args=(world, _free_port(), E, k, str(tmp_path)), nprocs=world, join=True)
    nodes_g = _jittered_mesh(E)
    u_ref = fem_p1.solve_fem_p1(nodes_g, k, solver='banded')
    for r in range(world):
        e0, e1 = hdist.partition(E, world, r)
        u = np.load(tmp_path / ('u%d.npy' % r))
        assert np.max(np.abs(u - u_ref[e0:e1 + 1])) <= 1e-11, r
        l2, mx, failed = np.load(tmp_path / ('r%d.npy' % r))
        assert abs(l2 - np.sqrt(world * (world + 1) / 2)) <= 1e-14 and mx == world - 1 and failed == 1


# ---- general operator -(a u')' + c u = f across ranks (dist.fem_p1_solve_general_distributed), CPU stand-in for the
# local solves: the exchange, the interface system and the superposition are what is tested here
_A = lambda x: 1.0 + 0.5 * np.sin(2.0 * x) ** 2      # noqa: E731
_C = lambda x: 2.0 + x                               # noqa: E731
_F = lambda x: 10.0 * np.cos(3.0 * x)                # noqa: E731


def _gauss_samples(nodes, fn):
    x0, h = nodes[:-1], np.diff(nodes)
    return np.stack([fn(x0 + h * fem_p1._GX[0]), fn(x0 + h * fem_p1._GX[1])])


def _oracle_general_local(nodes, aq, fq, cq, ul, ur):
    """Local P1 solve from the SAMPLES (as the kernels do): assemble with the oracle's rule from aq / cq / fq."""
    x = nodes.numpy()
    n = x.size
    h = np.diff(x)
    gx = fem_p1._GX
    kd = np.zeros(n); ko = np.zeros(n - 1); b = np.zeros(n)
    a_, f_ = aq.numpy(), fq.numpy()
    c_ = cq.numpy() if cq is not None else np.zeros_like(a_)
    for q in range(2):
        w = h * 0.5
        pl, pr = 1.0 - gx[q], gx[q]
        ka = a_[q] / h / h * w
        kd[:-1] += ka + c_[q] * pl * pl * w
        kd[1:] += ka + c_[q] * pr * pr * w
        ko += -ka + c_[q] * pl * pr * w
        b[:-1] += f_[q] * pl * w
        b[1:] += f_[q] * pr * w
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    lo, up, dg = ko.copy(), ko.copy(), kd.copy()
    dg[0] = dg[-1] = 1.0
    up[0] = 0.0
    lo[-1] = 0.0
    b[0], b[-1] = ul, ur
    return torch.from_numpy(spla.spsolve(sp.diags([lo, dg, up], [-1, 0, 1], format='csr'), b))


def _general_worker(rank, world, port, E, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        nodes_g = _jittered_mesh(E)
        e0, e1 = hdist.partition(E, world, rank)
        x = nodes_g[e0:e1 + 1].copy()
        aq, cq, fq = (torch.from_numpy(_gauss_samples(x, fn)) for fn in (_A, _C, _F))
        u = hdist.fem_p1_solve_general_distributed(torch.from_numpy(x), aq, fq, cq, u_left=0.3, u_right=-0.2,
                                                   local_solve=_oracle_general_local)
        np.save(os.path.join(out_dir, 'g%d.npy' % rank), u.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_general_operator_interface_system_over_gloo(tmp_path, world):
    E = 257
    mp.spawn(_general_worker, args=(world, _free_port(), E, str(tmp_path)), nprocs=world, join=True)
    nodes_g = _jittered_mesh(E)
    ref = fem_p1.solve_fem_p1_general(nodes_g, _A, _C, _F, u_left=0.3, u_right=-0.2)
    for r in range(world):
        e0, e1 = hdist.partition(E, world, r)
        u = np.load(tmp_path / ('g%d.npy' % r))
        assert np.max(np.abs(u - ref[e0:e1 + 1])) <= 1e-11, r
