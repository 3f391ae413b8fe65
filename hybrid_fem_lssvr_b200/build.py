"""Build libhfl.so (hand-written CUDA for sm_100a + the C ABI of include/hfl.h) in-tree.

    python -m hybrid_fem_lssvr_b200.build          # incremental
    python -m hybrid_fem_lssvr_b200.build --force

nvcc cross-compiles without a GPU.  The shared object lands next to this file so it travels with
the repository snapshot to the GPU box; it is git-ignored.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# HFL_VARIANT=name HFL_EXTRA_NVCC_FLAGS="-D..." builds an A/B copy (libhfl_name.so, objects in build_name/)
VARIANT = os.environ.get('HFL_VARIANT', '')
OBJ = os.path.join(HERE, 'build' + ('_' + VARIANT if VARIANT else ''))
LIB = os.path.join(HERE, 'libhfl%s.so' % ('_' + VARIANT if VARIANT else ''))
SOURCES = ['hfl_abi.cu', 'hfl_primal.cu', 'hfl_fem.cu', 'hfl_flux.cu', 'hfl_eval.cu', 'hfl_dual.cu',
           'hfl_dual_small.cu', 'hfl_dual_parity.cu', 'hfl_peer.cu', 'hfl_general.cu', 'hfl_primal_f16.cu', 'hfl_primal_f64.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr'] + os.environ.get('HFL_EXTRA_NVCC_FLAGS', '').split()


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force, hdr_m, verbose):
    obj = os.path.join(OBJ, src.replace('.cu', '.o'))
    spath = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(spath), hdr_m)):
        return obj, False
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', spath, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj, True


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libhfl.so.  Returns the library path."""
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_m, verbose), SOURCES))
    objs = [o for o, _ in res]
    if (force or any(ch for _, ch in res) or not os.path.exists(LIB)
            or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs)):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-cudart', 'static', '-Xlinker', '--no-undefined', '-ldl', '-lrt', '-lpthread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
