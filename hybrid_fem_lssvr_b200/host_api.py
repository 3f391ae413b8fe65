"""Host-buffer pipeline: the hot path as a caller with HOST memory sees it (what the reference's
`solver.solve(); solver.evaluate_solution(...)` sequence delivers: host arrays in, host arrays out).

Pinned host nodes -> H2D -> K1 coarse solve -> K2/K3/K5 over element chunks -> D2H of each chunk's
fine rows while the next chunk computes (two device staging buffers, compute + copy streams).
"""
import numpy as np
import torch

from . import batch


def stream_samples(func, lo, hi, rel_points, device, chunk_elements=1 << 20):
    """Samples of a host callable on every element, streamed to the device through pinned memory.

    Row j of the result holds func(lo + rel_points[j] * (hi - lo)) for every element (lo, hi: host arrays [E] of the
    element end points): rel_points = linspace(0, 1, N) gives the collocation samples d_f[j * E + e] of
    HFL_FORCING_SAMPLES (P:40, P:45), the two Gauss abscissae the samples of hfl_fem_p1_solve_general.  The callable is
    evaluated chunk by chunk into one of two pinned staging buffers while the previous chunk's cudaMemcpyAsync runs on
    a copy stream, so at 1e7 elements the 960 MB of samples never sit in pageable memory and the copy hides behind the
    host evaluation.  Returns a [len(rel_points), E] float64 CUDA tensor, ready for the current stream."""
    from .api import _sample_rhs          # vectorised / scalar / constant callables
    device = torch.device(device)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    rel = np.asarray(rel_points, dtype=np.float64).reshape(-1, 1)
    E, N = lo.size, rel.shape[0]
    out = torch.empty((N, E), dtype=torch.float64, device=device)
    if E == 0:
        return out
    chunk = max(1, min(int(chunk_elements), E))
    stage = [torch.empty((N, chunk), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    done = [None, None]
    copy_stream = torch.cuda.Stream(device=device)
    copy_stream.wait_stream(torch.cuda.current_stream(device))      # `out` exists before the first copy lands in it
    for c, e0 in enumerate(range(0, E, chunk)):
        e1 = min(E, e0 + chunk)
        buf = stage[c & 1]
        if done[c & 1] is not None:
            done[c & 1].synchronize()                                # the copy that last used this staging buffer
        a, b = lo[e0:e1], hi[e0:e1]
        if N > 1 and rel[0, 0] == 0.0 and rel[-1, 0] == 1.0 and np.allclose(rel[:, 0], np.linspace(0.0, 1.0, N), rtol=0, atol=0):
            pts = np.linspace(a, b, N, axis=0)                       # the reference's own points (P:40), bit for bit
        else:
            pts = a[None, :] + rel * (b - a)[None, :]
        buf.numpy()[:, :e1 - e0] = _sample_rhs(func, pts)
        with torch.cuda.stream(copy_stream):
            out[:, e0:e1].copy_(buf[:, :e1 - e0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        done[c & 1] = ev
    torch.cuda.current_stream(device).wait_stream(copy_stream)
    for ev in done:                      # the staging buffers die with this frame: their copies must have been issued AND done
        if ev is not None:
            ev.synchronize()
    return out


class HostPipeline:
    def __init__(self, E, M, gamma, N=12, F=32, k_freq=1.0, coarse_solver='assembled', device='cuda', chunks=8):
        self.E, self.M, self.gamma, self.N, self.F = int(E), int(M), float(gamma), int(N), int(F)
        self.k_freq, self.coarse_solver = float(k_freq), coarse_solver
        self.device = torch.device(device)
        self.chunks = max(1, min(int(chunks), self.E))
        per = (self.E + self.chunks - 1) // self.chunks
        self.ranges = [(s, min(self.E, s + per)) for s in range(0, self.E, per)]
        self.nodes_d = torch.empty(self.E + 1, dtype=torch.float64, device=self.device)
        self.u_d = torch.empty(self.E + 1, dtype=torch.float64, device=self.device)
        self.stage = [torch.empty((per, self.F), dtype=torch.float64, device=self.device) for _ in range(2)]
        self.err3 = batch.new_error_accumulator(self.device)
        self.fine_h = torch.empty((self.E, self.F), dtype=torch.float64, pin_memory=True)
        self.u_h = torch.empty(self.E + 1, dtype=torch.float64, pin_memory=True)
        self.err_h = torch.empty(3, dtype=torch.float64, pin_memory=True)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 8 * (self.E + 1)
        self.d2h_bytes = 8 * self.E * self.F + 8 * (self.E + 1) + 24

    def _run_resident(self, nodes_h):
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, 'fine_d', None) is None:
            self.fine_d = torch.empty((self.E, self.F), dtype=torch.float64, device=self.device)
        self.nodes_d.copy_(nodes_h, non_blocking=True)
        self.err3.zero_()
        batch.fem_p1_solve(self.nodes_d, k_freq=self.k_freq, coarse_solver=self.coarse_solver, out=self.u_d)
        batch.lssvr_primal_batch(self.nodes_d, self.u_d, self.M, self.gamma, N=self.N, F=self.F, forcing='sine',
                                 k_freq=self.k_freq, want_coef=False, want_fine=True, fine_out=self.fine_d, err3=self.err3)
        self.u_h.copy_(self.u_d, non_blocking=True)
        self.err_h.copy_(self.err3, non_blocking=True)
        cur.synchronize()
        return self.fine_d, self.u_h, (float(self.err_h[0]) ** 0.5, float(self.err_h[1]))

    def pinned_nodes(self):
        return torch.empty(self.E + 1, dtype=torch.float64, pin_memory=True)

    def run(self, nodes_h, fetch_fine=True):
        """nodes_h: pinned host tensor [E + 1].  Returns (fine_h [E, F], u_h [E + 1], (l2, max)) on the host.
        fetch_fine=False: the fine grid stays on the device (one launch over all elements into `self.fine_d`); only the
        nodal values and the error norms come back."""
        if not fetch_fine:
            return self._run_resident(nodes_h)
        cur = torch.cuda.current_stream(self.device)
        self.nodes_d.copy_(nodes_h, non_blocking=True)
        self.err3.zero_()
        batch.fem_p1_solve(self.nodes_d, k_freq=self.k_freq, coarse_solver=self.coarse_solver, out=self.u_d)
        k1_done = torch.cuda.Event()
        k1_done.record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(k1_done)
            self.u_h.copy_(self.u_d, non_blocking=True)
        freed = [None, None]
        for c, (e0, e1) in enumerate(self.ranges):
            buf = self.stage[c & 1]
            if freed[c & 1] is not None:
                cur.wait_event(freed[c & 1])
            view = buf[:e1 - e0]
            batch.lssvr_primal_batch(self.nodes_d[e0:e1 + 1], self.u_d[e0:e1 + 1], self.M, self.gamma, N=self.N,
                                     F=self.F, forcing='sine', k_freq=self.k_freq, want_coef=False, want_fine=True,
                                     fine_out=view, err3=self.err3)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                self.fine_h[e0:e1].copy_(view, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.copy_stream)
            freed[c & 1] = done
        self.err_h.copy_(self.err3, non_blocking=True)
        cur.synchronize()
        self.copy_stream.synchronize()
        return self.fine_h, self.u_h, (float(self.err_h[0]) ** 0.5, float(self.err_h[1]))
