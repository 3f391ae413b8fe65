"""Every kernel of libhfl.so once, at small sizes, for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck  python scripts/sanitize_r02.py
    compute-sanitizer --tool racecheck python scripts/sanitize_r02.py

Covers K1 in its three modes plus the general operator and the multi-right-hand-side launch (several chunk-CTA / tile
counts, a tile boundary inside the mesh), the element kernel's store paths (direct, shared-memory transpose, TMA boxes,
TMA rows, warp-cooperative) with and without the fused norms and the coefficient output, F = 16 / 64 and the generic
kernel, both dual kernels (registers, left-looking parity, right-looking parity, full system), the general-operator
kernel, point evaluation, the error kernels and the peer-exchange kernel (including its expiry path)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch, dist as hdist, _lib

def dev(x): return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
rng = np.random.default_rng(0)
done = []

# ---- K1
for n in (2, 9, 2047, 2048 * 2 + 5, 2048 * 9 + 1, 2048 * 17):
    w = 1 + 0.5 * rng.uniform(-1, 1, n - 1)
    nodes = dev(-1 + 2 * np.concatenate([[0], np.cumsum(w)]) / w.sum())
    for mode in ('assembled', 'assembled_exact', 'flux'):
        batch.fem_p1_solve(nodes, k_freq=2.0, u_left=0.1, u_right=-0.2, coarse_solver=mode, want_reaction=True)
    ks = dev(np.array([1.0, 2.5, 7.0]))
    batch.fem_p1_solve_multi(nodes, ks, coarse_solver='assembled')
    batch.fem_p1_solve_multi(nodes, ks, coarse_solver='flux')
    if n > 2:
        aq = dev(1 + rng.random((2, n - 1))); fq = dev(rng.normal(size=(2, n - 1))); cq = dev(rng.random((2, n - 1)))
        batch.fem_p1_solve_general(nodes, aq, fq, cq)
        batch.fem_p1_solve_general(nodes, aq, fq)
done.append('K1 x6 sizes x (3 modes + multi + general)')

# ---- element kernel
E = 3000 + 13
w = 1 + 0.5 * rng.uniform(-1, 1, E)
nodes = dev(-1 + 2 * np.concatenate([[0], np.cumsum(w)]) / w.sum())
u = batch.fem_p1_solve(nodes)
fs = dev(rng.normal(size=(12, E)))
bc2 = dev(np.array([0.1, -0.3]))
for store in (1, 2, 3, 4, 5):
    batch.set_option('primal_store', store)
    for M in (5, 9, 12):
        batch.lssvr_primal_batch(nodes, u, M, 1e4, N=12, F=32, want_coef=False, want_fine=True)
        batch.lssvr_primal_batch(nodes, u, M, 1e4, N=12, F=32, want_coef=True, want_fine=True, want_status=True,
                                 err3=batch.new_error_accumulator(), bc2=bc2)
        batch.lssvr_primal_batch(nodes, u, M, 1e4, N=12, F=32, forcing=fs, want_coef=True, want_fine=True)
batch.set_option('primal_store', 0)
for F in (0, 16, 64, 33):
    batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=F, want_coef=True, want_fine=F > 0,
                             err3=batch.new_error_accumulator() if F else None)
batch.lssvr_primal_batch(nodes, u, 20, 1e4, N=15, F=32, want_coef=True, want_fine=True, err3=batch.new_error_accumulator())
done.append('element kernel: 5 store paths x 3 M x 3 variants, F = 0/16/64/33, generic')

# ---- dual kernels
batch.lssvr_dual_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=True, want_fine=True, want_status=True, err3=batch.new_error_accumulator())
batch.lssvr_dual_batch(nodes, u, 9, 1e4, N=12, F=32, forcing=fs, want_coef=False, want_fine=True, bc2=bc2)
Es = 24
ns = dev(0.3 + np.linspace(-1, 1, Es + 1) * 1e-3)
ks = dev(np.array([1.0, 3.0, 8.0, 16.0]))
us = torch.sin(np.pi * ks[:, None] * ns[None, :]).contiguous()
for team in (0, 3, 2, 1):
    batch.set_option('dual_team', team)
    for (N, M) in ((128, 25), (128, 5), (64, 9), (13, 7)):
        if team == 1 and N > 32:
            continue
        batch.lssvr_dual_multi(ns, us, ks, M, 1e4, N=N, F=32, want_coef=True, want_fine=True, want_status=True,
                               err3=torch.zeros((4, 3), dtype=torch.float64, device='cuda'))
batch.set_option('dual_team', 0)
nc = dev(np.linspace(-1, 1, 40))            # coarse mesh: full rank, spill columns of the left-looking kernel
uc = torch.sin(np.pi * ks[:, None] * nc[None, :]).contiguous()
batch.lssvr_dual_multi(nc, uc, ks, 25, 1e4, N=128, F=32, want_coef=True, want_fine=True)
done.append('dual kernels: registers, left-looking / right-looking parity, full system, spill columns')

# ---- general operator, point evaluation, error kernels
a = dev(1 + rng.random((12, E))); da = dev(rng.normal(size=(12, E))); c = dev(rng.random((12, E)))
batch.lssvr_general_batch(nodes, u, a, fs, 9, 1e4, N=12, F=32, da=da, c=c, bc2=bc2, want_coef=True, want_fine=True, want_status=True)
coef, fine, _ = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=True, want_fine=True)
batch.evaluate_points(nodes, coef, dev(rng.uniform(-1.2, 1.2, 5000)))
batch.evaluate_points(dev(np.linspace(-1, 1, E + 1)), coef, dev(np.sort(rng.uniform(-1, 1, 5000))))
batch.error_fine(nodes, fine, 1.0)
batch.error_nodal(nodes, u, 1.0)
batch.mesh_linspace(-1.0, 1.0, 12345, 7, 1000)
done.append('general operator, point evaluation, error kernels, linspace')

# ---- peer exchange: one rank alone (expiry path) and a two-rank exchange on two streams
lib = _lib.load()
nb = int(lib.hfl_peer_buffer_bytes())
bufs = [torch.zeros(nb // 8, dtype=torch.int64, device='cuda') for _ in range(2)]
batch.set_option('peer_spin_log2', 8)
me = hdist.PeerExchange(rank=0, buffers=[b.data_ptr() for b in bufs])
me.all_gather(dev(np.array([1.0, 2.0, 3.0])), hdist.PeerExchange.CHANNEL_ERROR)
me.spike_exchange(dev(np.array([-1.0, 0.0, 0.1, 0.2])), 0.0, 0.0)
torch.cuda.synchronize()
assert me.timed_out()
done.append('peer exchange kernel (expiry path)')
torch.cuda.synchronize()
print('sanitize_r02: ok -', '; '.join(done), '- launches', _lib.launch_count())
