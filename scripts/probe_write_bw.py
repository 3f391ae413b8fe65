"""Write-bandwidth probes on the GPU box: what a store-only stream of 2.56 GB reaches (context for the
roofline fraction of the store-dominated element kernel)."""
import sys, os, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from hybrid_fem_lssvr_b200 import batch

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n

n = 32 * 10 ** 7
buf = torch.empty(n, dtype=torch.float64, device='cuda')
src = torch.randn(n, dtype=torch.float64, device='cuda')
lib = batch._lib.load()
ms = t(lambda: batch._lib.check(lib.hfl_mesh_linspace(-1.0, 1.0, n, 0, n, batch._ptr(buf), batch._stream()), 'linspace'))
print('linspace write (distinct values, STG.64) 2.56 GB: %.4f ms  %.0f GB/s' % (ms, 8 * n / ms / 1e6))
ms = t(lambda: buf.fill_(1.0))
print('torch fill_(1.0)                          2.56 GB: %.4f ms  %.0f GB/s' % (ms, 8 * n / ms / 1e6))
ms = t(lambda: buf.copy_(src))
print('torch copy_ (read 2.56 + write 2.56 GB)           : %.4f ms  %.0f GB/s' % (ms, 16 * n / ms / 1e6))
ms = t(lambda: torch.add(src, 1.0, out=buf))
print('torch add scalar (read + write)                   : %.4f ms  %.0f GB/s' % (ms, 16 * n / ms / 1e6))

for blocks, threads, pattern, tile in ((148 * 16, 256, 0, 2), (148 * 8, 256, 0, 2), (148 * 4, 128, 0, 2), (148 * 4, 128, 1, 4096),
                                       (148 * 3, 128, 1, 4096), (148 * 8, 128, 1, 4096), (148 * 16, 128, 1, 4096),
                                       (148 * 4, 256, 1, 4096), (148 * 4, 128, 1, 1024), (148 * 4, 128, 1, 32768)):
    ms = t(lambda: batch._lib.check(lib.hfl_store_probe(blocks, threads, n, pattern, tile, batch._ptr(buf), batch._stream()), 'probe'))
    print('store probe blocks=%5d threads=%3d pattern=%d tile=%6d doubles: %.4f ms  %.0f GB/s' % (blocks, threads, pattern, tile, ms, 8 * n / ms / 1e6))
