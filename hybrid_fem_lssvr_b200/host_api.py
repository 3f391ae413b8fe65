"""Host-buffer pipeline: the hot path as a caller with HOST memory sees it (what the reference's
`solver.solve(); solver.evaluate_solution(...)` sequence delivers: host arrays in, host arrays out).

Pinned host nodes -> H2D -> K1 coarse solve -> K2/K3/K5 over element chunks -> D2H of each chunk's
fine rows while the next chunk computes (two device staging buffers, compute + copy streams).
"""
import torch

from . import batch


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

    def pinned_nodes(self):
        return torch.empty(self.E + 1, dtype=torch.float64, pin_memory=True)

    def run(self, nodes_h):
        """nodes_h: pinned host tensor [E + 1].  Returns (fine_h [E, F], u_h [E + 1], (l2, max)) on the host."""
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
