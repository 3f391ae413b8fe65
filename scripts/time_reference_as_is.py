"""Reference-as-is timing, BUILD CONTAINER ONLY (needs /root/reference): the reference's own
lssvr_primal (P:20-105, SLSQP), executed unchanged through oracle/ref_loader.py, on all 24 elements of the shipped
configuration (P:216-220: 25 nodes, M = 8, gamma = 1e4) - one process, then a multiprocessing pool over the host's
cores.  Writes profiles/r02_reference_as_is.json, which bench.py attaches to its cpu_baseline object as the stated
reference-as-is baseline (the GPU box has no /root/reference).

    python scripts/time_reference_as_is.py
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from oracle import fem_p1, ref_loader  # noqa: E402

NODES = np.linspace(-1.0, 1.0, 25)
U = fem_p1.solve_fem_p1(NODES)
_ns = None


def _solve(i):
    global _ns
    if _ns is None:
        _ns = ref_loader.load_reference_functions()
    np.random.seed(1234 + i)
    f = _ns['lssvr_primal'](_ns['poisson_rhs'], [NODES[i], NODES[i + 1]], U[i], U[i + 1], 8, 1e4,
                            is_left_boundary=(i == 0), is_right_boundary=(i == 23), global_domain_range=(-1, 1))
    return f.coef.tolist()


def main():
    import scipy
    _solve(0)                                    # warm-up (imports, first SLSQP call)
    t0 = time.perf_counter()
    for i in range(24):
        _solve(i)
    serial = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    with mp.get_context('fork').Pool(cores) as pool:
        pool.map(_solve, range(cores))           # warm the workers
        t0 = time.perf_counter()
        reps = 4
        pool.map(_solve, list(range(24)) * reps)
        pooled = time.perf_counter() - t0
    out = {'what': 'reference lssvr_primal (P:20-105, scipy SLSQP) executed unchanged on the 24 elements of the shipped '
                   'configuration (25 nodes, M=8, gamma=1e4, N=12)',
           'where': 'build container (no GPU)', 'cpu': platform.processor() or platform.machine(), 'cores': cores,
           'numpy': np.__version__, 'scipy': scipy.__version__,
           'serial_24_elements_s': serial, 'solves_per_s_one_process': 24 / serial,
           'pool_processes': cores, 'pool_solves': 24 * reps, 'pool_wall_s': pooled,
           'solves_per_s_whole_host': 24 * reps / pooled}
    path = os.path.join(os.path.dirname(__file__), '..', 'profiles', 'r02_reference_as_is.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
