"""Debug aid: run the assembled / exact coarse solve, read the tile records and the tile-head values back from the
workspace and re-solve the tile-head system on the host (row-sum Thomas in numpy) to see which level loses accuracy."""
import ctypes as C, os, sys, math, numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import k1_cr_prototype as P
L = C.CDLL(os.path.abspath(sys.argv[1]))
L.hfl_fem_p1_workspace_bytes.restype = C.c_size_t; L.hfl_fem_p1_workspace_bytes.argtypes = [C.c_int64]
L.hfl_fem_p1_solve.restype = C.c_int
L.hfl_fem_p1_solve.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]
st = torch.cuda.current_stream().cuda_stream

def safe_thomas(l, sg, r, b):
    n = l.size
    c = np.zeros(n); g = np.zeros(n)
    q = 1.0; bp = 0.0
    for i in range(n):
        s = sg[i] - l[i] * q
        inv = 1.0 / (s - r[i])
        bp = (b[i] - l[i] * bp) * inv
        c[i] = r[i] * inv
        q = s * inv
        g[i] = bp
    x = np.zeros(n); x[-1] = g[-1]
    for i in range(n - 2, -1, -1):
        x[i] = g[i] - c[i] * x[i + 1]
    return x

for n in [int(x) for x in os.environ.get('SIZES', '614405,10000001').split(',')]:
    nodes = torch.linspace(-1, 1, n, dtype=torch.float64, device='cuda')
    u = torch.empty_like(nodes)
    ws = torch.zeros(int(L.hfl_fem_p1_workspace_bytes(n)), dtype=torch.uint8, device='cuda')
    assert L.hfl_fem_p1_solve(n, nodes.data_ptr(), 1.0, 0.0, 0.0, 2, u.data_ptr(), None, ws.data_ptr(), ws.numel(), st) == 0
    torch.cuda.synchronize()
    nt = (n + 2047) // 2048
    w = ws.view(torch.float64).cpu().numpy()
    rec = w[:12 * nt].reshape(12, nt).T.copy()
    utop = w[12 * nt:13 * nt].copy()
    rows = P.top_rows(rec)
    ref = safe_thomas(*rows)
    ex = np.sin(math.pi * nodes.cpu().numpy())
    heads_exact = ex[::2048][:nt]
    print('n=%d nt=%d: |utop_gpu - host_solve(rec)| = %.2e   |host_solve(rec) - sin| = %.2e   |utop_gpu - sin| = %.2e   |u - sin| = %.2e'
          % (n, nt, np.max(np.abs(utop - ref)), np.max(np.abs(ref - heads_exact)), np.max(np.abs(utop - heads_exact)),
             np.max(np.abs(u.cpu().numpy() - ex))))

# field-by-field comparison of the GPU tile records with the numpy model (exact row sums), n = 614405
n = 614405
nodes_h = np.linspace(-1, 1, n)
nodes = torch.from_numpy(nodes_h).cuda()
u = torch.empty_like(nodes)
ws = torch.zeros(int(L.hfl_fem_p1_workspace_bytes(n)), dtype=torch.uint8, device='cuda')
assert L.hfl_fem_p1_solve(n, nodes.data_ptr(), 1.0, 0.0, 0.0, 2, u.data_ptr(), None, ws.data_ptr(), ws.numel(), st) == 0
torch.cuda.synchronize()
nt = (n + 2047) // 2048
w = ws.view(torch.float64).cpu().numpy()
rec = w[:12 * nt].reshape(12, nt).T.copy()
l, sg, r, b = P.rows_assembled(nodes_h, 1.0, True)
TS = 2048
pad = nt * TS - n
l, r = np.concatenate([l, np.zeros(pad)]), np.concatenate([r, np.zeros(pad)])
sg, b = np.concatenate([sg, np.ones(pad)]), np.concatenate([b, np.zeros(pad)])
worst = np.zeros(12)
for c in range(nt):
    sl = slice(c * TS, (c + 1) * TS)
    ref, _ = P.tile_reduce(l[sl], sg[sl], r[sl], b[sl], 256, 8)
    rel = np.abs(rec[c] - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rec[c])[ref == 0]
    worst = np.maximum(worst, rel)
    if c in (1, 150):
        print('tile', c, 'gpu', rec[c]); print('        ref', ref)
print('worst relative deviation per record field:', ' '.join('%.1e' % x for x in worst))
refs = []
for c in range(nt):
    sl = slice(c * TS, (c + 1) * TS)
    refs.append(P.tile_reduce(l[sl], sg[sl], r[sl], b[sl], 256, 8)[0])
refs = np.array(refs)
for f in (5, 6, 9, 10):
    tgt = -1.0 / 2048 if f in (6, 9) else -(1.0 - 1.0 / 2048)
    dg = rec[1:-1, f] / tgt - 1.0; dr = refs[1:-1, f] / tgt - 1.0
    print('field %2d: gpu/ideal - 1: mean %.2e std %.2e min %.2e max %.2e | numpy model: mean %.2e std %.2e' %
          (f, dg.mean(), dg.std(), dg.min(), dg.max(), dr.mean(), dr.std()))
heads_exact = np.sin(math.pi * nodes_h)[::2048][:nt]
def err(rc):
    return np.max(np.abs(safe_thomas(*P.top_rows(rc)) - heads_exact))
print('host solve: gpu records %.2e, numpy-model records %.2e' % (err(rec), err(refs)))
for grp, name in (((0, 1, 2), 'head row l, sigma, r'), ((3,), 'head b'), ((4, 8), 'y first/last'), ((5, 10), 'big couplings'),
                  ((6, 9), 'small couplings'), ((7, 11), 'leaks')):
    mix = rec.copy(); mix[:, list(grp)] = refs[:, list(grp)]
    print('  gpu records with %-22s from the numpy model: %.2e' % (name, err(mix)))
for f in (6, 9):
    d = rec[1:-1, f] / refs[1:-1, f] - 1.0
    print('field %d gpu/numpy - 1: mean %.3e std %.3e min %.3e max %.3e; first tiles:' % (f, d.mean(), d.std(), d.min(), d.max()), ' '.join('%.1e' % x for x in d[:12]))
d = (rec[1:-1, 6] / refs[1:-1, 6]) / (rec[1:-1, 9] / refs[1:-1, 9]) - 1.0
print('asymmetry f6 vs f9: mean %.3e std %.3e' % (d.mean(), d.std()))
heads = w[19 * nt:19 * nt + nt * 768].reshape(nt, 3, 256)
for c in (1, 100):
    sl = slice(c * TS, (c + 1) * TS)
    _, (Ld, Rd, Bd) = P.tile_reduce(l[sl], sg[sl], r[sl], b[sl], 256, 8)
    for lvname, idx in (('d=1', np.arange(1, 256, 2)), ('d=2', np.arange(2, 256, 4)), ('d=4', np.arange(4, 256, 8)), ('d=8', np.arange(8, 256, 16)),
                        ('d=16', np.arange(16, 256, 32)), ('d=32', np.arange(32, 256, 64)), ('d=64', np.array([64, 192])), ('d=128', np.array([128]))):
        dl = heads[c, 0, idx] / Ld[idx] - 1.0; dr = heads[c, 1, idx] / Rd[idx] - 1.0
        print('tile %d level %-5s Ld gpu/numpy-1: mean %9.2e max|.| %.2e   Rd: mean %9.2e max|.| %.2e   (Ld+Rd+1 gpu: %.1e)' %
              (c, lvname, dl.mean(), np.abs(dl).max(), dr.mean(), np.abs(dr).max(), np.abs(heads[c, 0, idx] + heads[c, 1, idx] + 1).max()))
