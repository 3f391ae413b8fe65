"""K1 accuracy A/B through the C ABI: python scripts/probe_k1_accuracy.py LIB [LIB ...]
For several mesh sizes: exact-row-sum and assembled modes against the flux form of the FIRST library and sin(pi x)."""
import ctypes as C, os, sys, math, torch
libs = []
for path in sys.argv[1:]:
    L = C.CDLL(os.path.abspath(path))
    L.hfl_fem_p1_workspace_bytes.restype = C.c_size_t; L.hfl_fem_p1_workspace_bytes.argtypes = [C.c_int64]
    L.hfl_fem_p1_solve.restype = C.c_int
    L.hfl_fem_p1_solve.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]
    libs.append((os.path.basename(path), L))
st = torch.cuda.current_stream().cuda_stream
def run(L, n, nodes, solver):
    u = torch.empty_like(nodes)
    ws = torch.empty(int(L.hfl_fem_p1_workspace_bytes(n)), dtype=torch.uint8, device='cuda')
    rc = L.hfl_fem_p1_solve(n, nodes.data_ptr(), 1.0, 0.0, 0.0, solver, u.data_ptr(), None, ws.data_ptr(), ws.numel(), st)
    assert rc == 0
    torch.cuda.synchronize()
    return u
sizes = [int(x) for x in os.environ.get('SIZES', '').split(',') if x] or [
    2048 * 300 + 5, 2048 * 1024, 2048 * 1024 + 5, 2048 * 2048, 2048 * 3000 + 7, 10 ** 7 + 1, 2048 * 6144, 2048 * 6144 + 1]
for n in sizes:
    nodes = torch.linspace(-1, 1, n, dtype=torch.float64, device='cuda')
    ex = torch.sin(math.pi * nodes)
    uf = run(libs[0][1], n, nodes, 1)
    line = 'n=%9d flux-sin %.1e |' % (n, (uf - ex).abs().max().item())
    for name, L in libs:
        ue = run(L, n, nodes, 2); ua = run(L, n, nodes, 0)
        d = (ue - uf).abs()
        line += ' %s: exact-flux %.2e (at %d) assembled-flux %.2e |' % (name, d.max().item(), int(d.argmax().item()), (ua - uf).abs().max().item())
    print(line)
if len(libs) > 1:
    n = 10 ** 7 + 1
    nodes = torch.linspace(-1, 1, n, dtype=torch.float64, device='cuda')
    a0 = run(libs[0][1], n, nodes, 0); a1 = run(libs[1][1], n, nodes, 0)
    print('assembled, n=1e7+1: max |%s - %s| = %.3e' % (libs[0][0], libs[1][0], (a0 - a1).abs().max().item()))
