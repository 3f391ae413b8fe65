"""Multi-GPU host logic: contiguous element ranges, one process per GPU.

Elements are independent given the nodal values, so K2-K5 shard with no exchange.  The coarse solve
needs ONE exchange (SPIKE): every rank solves its own range with zero Dirichlet data at both ends,
the ranks all-gather four doubles each ({x_first, x_last, r_left, r_right}), every rank solves the
(G-1)-unknown interface system redundantly, and the correction - a linear function on the range - is
applied on the fly by the element kernel (d_bc2) or by hfl_fem_apply_bc.  Error norms end with one
all-reduce(sum) and one all-reduce(max).  All messages are a few doubles: latency-bound.
"""
import math

import torch
import torch.distributed as dist

from . import _lib, batch


class PeerExchange:
    """All-gather of a few doubles per rank through NVLink peer memory (csrc/hfl_peer.cu) instead of NCCL: same
    result as dist.all_gather_into_tensor, a few microseconds per call instead of 15-25.  One object per process
    group; construction is collective (swaps the IPC handles through torch.distributed).  Raises when peer memory is
    not available (ranks on different nodes, IPC disabled): callers keep NCCL as the fallback.

    `buffers` (tests): a list of G raw device pointers created in ONE process, one "rank" per stream."""

    CHANNEL_INTERFACE, CHANNEL_ERROR = 0, 1

    def __init__(self, group=None, device=None, rank=None, buffers=None):
        import ctypes as C
        self._lib = _lib.load()
        self._peers, self._own = [], None
        if buffers is not None:
            self.world, self.rank = len(buffers), int(rank)
            self.device = torch.device(device or 'cuda')
            ptrs = list(buffers)
        else:
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
            self.device = torch.device(device or ('cuda:%d' % torch.cuda.current_device()))
            # every step that can fail on one rank only is followed by a collective agreement, so that either all ranks
            # end up with a working exchange or all of them raise (and fall back to NCCL together)
            own = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            err = None
            try:
                with torch.cuda.device(self.device):
                    _lib.check(self._lib.hfl_peer_buffer_create(C.byref(own), handle), 'hfl_peer_buffer_create')
                self._own = own
            except Exception as exc:
                err = exc
            mine = torch.tensor([0 if err else 1] + list(handle), dtype=torch.uint8, device=self.device)
            allh = torch.empty(65 * self.world, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().reshape(self.world, 65)
            if int(allh[:, 0].min()) == 0:
                self.close()
                raise _lib.HflError('peer-memory buffer creation failed on a rank: %s' % (err or 'another rank'))
            ptrs = []
            try:
                for r in range(self.world):
                    if r == self.rank:
                        ptrs.append(own.value)
                        continue
                    h = (C.c_ubyte * 64)(*allh[r, 1:].tolist())
                    peer = C.c_void_p()
                    with torch.cuda.device(self.device):
                        _lib.check(self._lib.hfl_peer_buffer_open(h, C.byref(peer)), 'hfl_peer_buffer_open')
                    self._peers.append(peer)
                    ptrs.append(peer.value)
            except Exception as exc:
                err = exc
            good = torch.tensor([0.0 if err else 1.0], dtype=torch.float64, device=self.device)
            dist.all_reduce(good, op=dist.ReduceOp.MIN, group=group)     # also: every buffer is open before the first store
            if good.item() == 0.0:
                self.close()
                raise _lib.HflError('peer-memory buffer could not be opened on a rank: %s' % (err or 'another rank'))
        self.bufs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        # epochs live on the device (HFL_PEER_EPOCH_DEVICE): every launch is identical from step to step, so a step that
        # contains an exchange can be captured once and replayed as a CUDA graph
        self.EPOCH_DEVICE = 0xFFFFFFFF

    def all_gather(self, src, channel, out=None):
        """out [G, W] <- every rank's src [W] (W <= 4 float64), stream-ordered on the current stream."""
        W = src.numel()
        out = torch.empty((self.world, W), dtype=torch.float64, device=src.device) if out is None else out
        _lib.check(self._lib.hfl_peer_allgather(self.world, self.rank, W, batch._ptr(src), batch._ptr(self.bufs), self.EPOCH_DEVICE, channel,
                                                batch._ptr(out), batch._ptr(self.status), batch._stream()),
                   'hfl_peer_allgather')
        return out

    def spike_exchange(self, iface4, u_left, u_right):
        """The interface exchange in one launch: all-gather of the 4-double records + interface solve -> bc2 [2]."""
        gathered = torch.empty(4 * self.world, dtype=torch.float64, device=iface4.device)
        bc2 = torch.empty(2, dtype=torch.float64, device=iface4.device)
        ch = self.CHANNEL_INTERFACE
        _lib.check(self._lib.hfl_peer_spike_exchange(self.world, self.rank, batch._ptr(iface4), batch._ptr(self.bufs), self.EPOCH_DEVICE, ch,
                                                     float(u_left), float(u_right), batch._ptr(gathered), batch._ptr(bc2),
                                                     batch._ptr(self.status), batch._stream()), 'hfl_peer_spike_exchange')
        return bc2

    def timed_out(self):
        """True when a receive spin expired since construction (host sync)."""
        return bool(self.status.item())

    def check(self):
        """Raise when a receive spin expired since construction (host sync).  The kernels have already poisoned what they
        delivered with NaN; this turns it into an exception at the host's next synchronisation point."""
        if self.timed_out():
            raise _lib.HflError('peer-memory exchange timed out on rank %d: a peer did not deliver its record within the '
                                'receive spin (hfl_set_option peer_spin_log2); results of this step are NaN' % self.rank)

    def close(self):
        for peer in self._peers:
            self._lib.hfl_peer_buffer_close(peer)
        self._peers = []
        if self._own is not None:
            self._lib.hfl_peer_buffer_destroy(self._own)
            self._own = None


def partition(E_global, world, rank):
    """Contiguous element range [e0, e1) of `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError('rank %d outside [0, %d)' % (rank, world))
    base, rem = divmod(int(E_global), int(world))
    e0 = rank * base + min(rank, rem)
    return e0, e0 + base + (1 if rank < rem else 0)


def local_nodes_linspace(a, b, E_global, world, rank, device='cuda'):
    """Nodes of this rank's elements from the global numpy.linspace(a, b, E_global + 1)."""
    e0, e1 = partition(E_global, world, rank)
    return batch.mesh_linspace(a, b, E_global + 1, e0, e1 - e0 + 1, device=device)


def _device_local_solve(nodes, k_freq, coarse_solver, out=None):
    return batch.fem_p1_solve(nodes, k_freq=k_freq, u_left=0.0, u_right=0.0, coarse_solver=coarse_solver,
                              out=out, want_reaction=True)


def fem_p1_solve_distributed(nodes_local, k_freq=1.0, u_left=0.0, u_right=0.0, coarse_solver='assembled_exact',
                             group=None, local_solve=None, device_interface=True, out=None, exchange=None):
    """SPIKE coarse solve.  Returns (y_local, bc2): the local zero-Dirichlet solve and the two
    interface values {U_rank, U_rank+1} as a 2-vector on the same device as `nodes_local`;
    u_local = y_local + linear correction (see hfl_fem_apply_bc).

    coarse_solver: 'assembled_exact' (default) or 'flux'.  The interface system takes the correction to be linear on
    each range, which holds for these two; the reference's rounded diagonal ('assembled') adds a spurious reaction term
    eps k_i u_i whose discrete harmonic functions are not linear - harmless up to ~1e5 nodes per range (1e-10 parity,
    tests), 1e-3 at 1e7 nodes per range.

    exchange: a PeerExchange (NVLink peer memory) to use instead of the NCCL all-gather.

    `local_solve(nodes, k_freq, coarse_solver, out) -> (y, iface4)` defaults to the CUDA kernels; the
    CPU tests inject a stand-in to exercise the exchange logic over gloo.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    solve = local_solve or _device_local_solve
    y, mine = solve(nodes_local, k_freq, coarse_solver, out)
    if world > 1 and exchange is not None and mine.is_cuda and device_interface:
        return y, exchange.spike_exchange(mine, u_left, u_right)
    if world > 1 and exchange is not None:
        gathered = exchange.all_gather(mine, PeerExchange.CHANNEL_INTERFACE).reshape(-1)
    elif world > 1:
        gathered = torch.empty(4 * world, dtype=torch.float64, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    if mine.is_cuda and device_interface:
        bc2 = torch.empty(2, dtype=torch.float64, device=mine.device)
        _lib.check(_lib.load().hfl_spike_interface_solve_device(
            world, batch._ptr(gathered), float(u_left), float(u_right), rank, batch._ptr(bc2), batch._stream()),
            'hfl_spike_interface_solve_device')
    else:
        iface = batch.spike_interface_solve(gathered.cpu().tolist(), u_left, u_right)
        bc2 = torch.tensor([iface[rank], iface[rank + 1]], dtype=torch.float64, device=mine.device)
    return y, bc2


def gather_error(err3, group=None, out=None, exchange=None):
    """Stream-ordered half of the error reduction: all-gather of the per-rank accumulators into [G, 3]
    (no host synchronisation; call finish_gathered_error when the numbers are needed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return err3.reshape(1, 3)
    out = torch.empty((world, 3), dtype=torch.float64, device=err3.device) if out is None else out
    if exchange is not None:
        return exchange.all_gather(err3, PeerExchange.CHANNEL_ERROR, out=out)
    dist.all_gather_into_tensor(out.reshape(-1), err3, group=group)
    return out


def finish_gathered_error(gathered, exchange=None):
    """(L2, max, failed) from the [G, 3] accumulators: sum, max, sum.  This is the host synchronisation point of a
    step: with `exchange` given, a receive spin that expired anywhere in the step raises here."""
    g = gathered.cpu()
    if exchange is not None:
        exchange.check()
    if bool(torch.isnan(g).any()):
        raise _lib.HflError('error accumulators are NaN: an interface exchange delivered NaN (timeout) or the solve diverged')
    return math.sqrt(float(g[:, 0].sum())), float(g[:, 1].max()), int(g[:, 2].sum())


def reduce_error(err3, group=None):
    """Global (L2, max, failed) from per-rank accumulators: all-reduce(sum) on [0] and [2], all-reduce(max) on [1]."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        s = err3[[0, 2]].clone()
        m = err3[1:2].clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        vals = [s[0].item(), m[0].item(), s[1].item()]
    else:
        vals = err3.tolist()
    return math.sqrt(vals[0]), vals[1], int(vals[2])


# ---------------------------------------------------------------------------------------------------------------------
# General operator -(a u')' + c u = f on a mesh split into contiguous ranges.  Discrete
# homogeneous solutions are no longer linear, so every rank solves its range three times with the same matrix: y (the
# load, zero Dirichlet data), v (no load, u = 1 at its left end), w (no load, u = 1 at its right end); u = y + U_r v +
# U_{r+1} w for the interface values U.  At every interface the residuals of the two neighbouring end nodes must cancel:
# a tridiagonal system in U_1 .. U_{G-1} built from six end-node residuals per rank (one all-gather of 6 doubles).

def _general_end_residuals(nodes, z, aq, cq, fq, with_load):
    """(r_left, r_right): residual of the un-enforced P1 rows of the two end nodes for the nodal vector z; host scalars."""
    x = torch.stack([nodes[0], nodes[1], nodes[-2], nodes[-1]]).tolist()
    zz = torch.stack([z[0], z[1], z[-2], z[-1]]).tolist()
    E = nodes.numel() - 1
    g0, g1 = batch.GAUSS_X
    out = []
    for side, e in ((0, 0), (1, E - 1)):
        h = x[1] - x[0] if side == 0 else x[3] - x[2]
        hw = 0.5 * h
        a0, a1 = aq[0, e].item(), aq[1, e].item()
        c0, c1 = (cq[0, e].item(), cq[1, e].item()) if cq is not None else (0.0, 0.0)
        f0, f1 = (fq[0, e].item(), fq[1, e].item()) if with_load else (0.0, 0.0)
        pl0, pl1, pr0, pr1 = 1.0 - g0, 1.0 - g1, g0, g1
        ka = (a0 + a1) / (h * h) * hw
        mll = (c0 * pl0 * pl0 + c1 * pl1 * pl1) * hw
        mlr = (c0 * pl0 * pr0 + c1 * pl1 * pr1) * hw
        mrr = (c0 * pr0 * pr0 + c1 * pr1 * pr1) * hw
        if side == 0:
            out.append((ka + mll) * zz[0] + (-ka + mlr) * zz[1] - (f0 * pl0 + f1 * pl1) * hw)
        else:
            out.append((-ka + mlr) * zz[2] + (ka + mrr) * zz[3] - (f0 * pr0 + f1 * pr1) * hw)
    return out


def general_interface_solve(records, u_left=0.0, u_right=0.0):
    """Interface values U_0 .. U_G from the gathered records [G][6] = {ry_l, ry_r, rv_l, rv_r, rw_l, rw_r} per rank."""
    import numpy as np
    rec = np.asarray(records, dtype=np.float64).reshape(-1, 6)
    G = rec.shape[0]
    U = np.zeros(G + 1)
    U[0], U[G] = u_left, u_right
    if G > 1:
        m = G - 1
        A = np.zeros((m, m))
        b = np.zeros(m)
        for r in range(1, G):          # interface between rank r-1 (right end) and rank r (left end)
            ryr, rvr, rwr = rec[r - 1, 1], rec[r - 1, 3], rec[r - 1, 5]
            ryl, rvl, rwl = rec[r, 0], rec[r, 2], rec[r, 4]
            i = r - 1
            A[i, i] = rwr + rvl
            b[i] = -(ryr + ryl)
            if r - 1 >= 1:
                A[i, i - 1] = rvr
            else:
                b[i] -= rvr * u_left
            if r + 1 <= G - 1:
                A[i, i + 1] = rwl
            else:
                b[i] -= rwl * u_right
        U[1:G] = np.linalg.solve(A, b)
    return U


def fem_p1_solve_general_distributed(nodes_local, aq, fq, cq=None, u_left=0.0, u_right=0.0, group=None, local_solve=None):
    """Partitioned coarse solve of -(a u')' + c u = f: returns the nodal values of this rank's range (interface values
    included).  aq, fq, cq: [2, E_local] samples at the Gauss points (batch.fem_p1_solve_general).  One all-gather of six
    doubles per rank; the interface system is solved on the host (this wrapper synchronises, unlike the Poisson path).

    `local_solve(nodes, aq, fq, cq, u_left, u_right) -> u` defaults to the CUDA kernels (CPU tests inject a stand-in)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    solve = local_solve or (lambda n, a, f, c, ul, ur: batch.fem_p1_solve_general(n, a, f, c, u_left=ul, u_right=ur))
    zero = torch.zeros_like(fq)
    y = solve(nodes_local, aq, fq, cq, 0.0, 0.0)
    v = solve(nodes_local, aq, zero, cq, 1.0, 0.0)
    w = solve(nodes_local, aq, zero, cq, 0.0, 1.0)
    ry = _general_end_residuals(nodes_local, y, aq, cq, fq, True)
    rv = _general_end_residuals(nodes_local, v, aq, cq, fq, False)
    rw = _general_end_residuals(nodes_local, w, aq, cq, fq, False)
    mine = torch.tensor([ry[0], ry[1], rv[0], rv[1], rw[0], rw[1]], dtype=torch.float64, device=nodes_local.device)
    if world > 1:
        gathered = torch.empty(6 * world, dtype=torch.float64, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    U = general_interface_solve(gathered.cpu().numpy(), u_left, u_right)
    return y + float(U[rank]) * v + float(U[rank + 1]) * w
