"""Out-of-bounds and race checks of our own (compute-sanitizer is closed on this pool: `profiles/r02_sanitizer.txt`).

Every output buffer a kernel writes through the C ABI sits between two guard bands of NaN canaries (4096 doubles
each): after the launch the bands must be untouched and the interior fully written.  Every launch runs twice on the same
inputs and must be bit-identical (a shared-memory or inter-CTA race shows up as run-to-run differences on one of the
sizes: tile / chunk / warp boundaries inside the range, ragged tails, single elements)."""
import numpy as np
import pytest
import torch

from hybrid_fem_lssvr_b200 import batch
from gpu_util import dev, jittered_mesh

pytestmark = pytest.mark.gpu
PAD = 4096


class Guarded:
    def __init__(self, *shape):
        n = int(np.prod(shape))
        self.raw = torch.full((n + 2 * PAD,), float('nan'), dtype=torch.float64, device='cuda')
        self.view = self.raw[PAD:PAD + n].view(*shape)

    def check(self, written=True):
        torch.cuda.synchronize()
        assert torch.isnan(self.raw[:PAD]).all() and torch.isnan(self.raw[-PAD:]).all(), 'guard band overwritten'
        if written:
            assert not torch.isnan(self.view).any(), 'output not fully written'
        return self.view.clone()


@pytest.mark.parametrize('n', [2, 3, 8, 9, 2047, 2048, 2049, 2048 * 8 - 1, 2048 * 8, 2048 * 8 + 1, 2048 * 17 + 5])
@pytest.mark.parametrize('mode', ['assembled', 'assembled_exact', 'flux'])
def test_coarse_solve_stays_inside_its_output(n, mode):
    nodes = dev(jittered_mesh(n - 1, seed=n) if n > 2 else np.array([-1.0, 1.0]))
    outs = []
    for _ in range(2):
        g = Guarded(n)
        batch.fem_p1_solve(nodes, k_freq=3.0, u_left=0.2, u_right=-0.1, coarse_solver=mode, out=g.view)
        outs.append(g.check())
    assert torch.equal(outs[0], outs[1])
    R = 3
    ks = dev(np.array([1.0, 2.0, 5.5]))
    g = Guarded(R, n)
    batch.fem_p1_solve_multi(nodes, ks, coarse_solver=mode, out=g.view)
    g.check()


@pytest.mark.parametrize('E', [1, 31, 32, 33, 127, 128, 129, 1000, 4099])
@pytest.mark.parametrize('store', [1, 2, 3, 4, 5])
def test_element_kernels_stay_inside_their_outputs(E, store):
    nodes = dev(jittered_mesh(E, seed=E))
    u = dev(np.random.default_rng(E).uniform(-1, 1, E + 1))
    batch.set_option('primal_store', store)
    try:
        for kw in (dict(), dict(err3=True), dict(coef=True)):
            outs = []
            for _ in range(2):
                gf, gc = Guarded(E, 32), Guarded(E, 9)
                batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef='coef' in kw, want_fine=True, fine_out=gf.view,
                                         coef_out=gc.view if 'coef' in kw else None,
                                         err3=batch.new_error_accumulator() if 'err3' in kw else None)
                outs.append((gf.check(), gc.check(written='coef' in kw)))
            assert torch.equal(outs[0][0], outs[1][0])
            if 'coef' in kw:
                assert torch.equal(outs[0][1], outs[1][1])
    finally:
        batch.set_option('primal_store', 0)
    outs = []
    for _ in range(2):
        gf = Guarded(E, 32)
        batch.lssvr_dual_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, fine_out=gf.view)
        outs.append(gf.check())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize('F', [16, 64, 33])
def test_other_fine_grid_sizes_stay_inside(F):
    E = 777
    nodes = dev(jittered_mesh(E, seed=F))
    u = dev(np.cos(np.linspace(0, 3, E + 1)))
    for M in (9, 20):
        gf = Guarded(E, F)
        batch.lssvr_primal_batch(nodes, u, M, 1e4, N=12, F=F, want_coef=False, want_fine=True, fine_out=gf.view,
                                 err3=batch.new_error_accumulator())
        gf.check()


@pytest.mark.parametrize('N,M', [(128, 25), (128, 5), (64, 9)])
def test_dual_multi_is_deterministic(N, M):
    """The left-looking parity kernel synchronises its two teams with named barriers: two runs must agree bit for bit,
    on a fine mesh (early stop at the numerical rank) and on a coarse one (full rank, spill columns in global memory)."""
    ks = dev(np.array([1.0, 3.0, 8.0, 16.0]))
    for nodes in (dev(0.3 + np.linspace(-1, 1, 38) * 1e-3), dev(np.linspace(-1, 1, 38))):
        us = torch.sin(np.pi * ks[:, None] * nodes[None, :]).contiguous()
        a = batch.lssvr_dual_multi(nodes, us, ks, M, 1e4, N=N, F=32, want_coef=True, want_fine=True)
        b = batch.lssvr_dual_multi(nodes, us, ks, M, 1e4, N=N, F=32, want_coef=True, want_fine=True)
        torch.cuda.synchronize()
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert not torch.isnan(a[1]).any()
