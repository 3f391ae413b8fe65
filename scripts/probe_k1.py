"""K1 A/B probe straight through the C ABI (ctypes), usable with any build of libhfl.so: python scripts/probe_k1.py LIB [LIB ...]
Times hfl_fem_p1_solve (assembled and flux) on 1e7 + 1 nodes, interleaving the libraries round-robin."""
import ctypes as C, os, sys, torch
n = int(os.environ.get('N_NODES', 10 ** 7 + 1))
libs = []
for path in sys.argv[1:]:
    L = C.CDLL(os.path.abspath(path))
    L.hfl_fem_p1_workspace_bytes.restype = C.c_size_t; L.hfl_fem_p1_workspace_bytes.argtypes = [C.c_int64]
    L.hfl_fem_p1_solve.restype = C.c_int
    L.hfl_fem_p1_solve.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]
    libs.append((path, L))
nodes = torch.linspace(-1, 1, n, dtype=torch.float64, device='cuda')
u = torch.empty_like(nodes)
ws = torch.empty(int(max(L.hfl_fem_p1_workspace_bytes(n) for _, L in libs)), dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def run(L, solver):
    rc = L.hfl_fem_p1_solve(n, nodes.data_ptr(), 1.0, 0.0, 0.0, solver, u.data_ptr(), None, ws.data_ptr(), ws.numel(), st)
    assert rc == 0
res = {}
for rnd in range(4):
    for path, L in libs:
        for solver in (0, 1):
            for _ in range(5): run(L, solver)
            torch.cuda.synchronize()
            ts = []
            for _ in range(20):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); run(L, solver); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            ts.sort(); res.setdefault((path, solver), []).append(ts[len(ts) // 2])
for (path, solver), v in res.items():
    print('%-50s %-9s median per round (ms): %s' % (path, 'assembled' if solver == 0 else 'flux', ' '.join('%.4f' % x for x in v)))
