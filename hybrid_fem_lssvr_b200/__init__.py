"""hybrid_fem_lssvr_b200: the hybrid FEM + LSSVR hot path on NVIDIA B200 (sm_100a, FP64).

Drop-in surface of the reference script (same names as 1D-Possion/Hybrid-FEM-LSSVR.py) in ``api``,
device-resident batched entry points in ``batch``, multi-GPU host logic in ``dist``.  All arithmetic
is hand-written CUDA in ``libhfl.so`` behind the C ABI of ``include/hfl.h``; there is no CPU path.
"""
from . import _lib, batch, dist                                               # noqa: F401
from .api import (FEMLSSVRPrimalSolver, lssvr_primal, main_boundary_condition_left,   # noqa: F401
                  main_boundary_condition_right, poisson_rhs, true_solution)

__all__ = ['FEMLSSVRPrimalSolver', 'lssvr_primal', 'true_solution', 'poisson_rhs',
           'main_boundary_condition_left', 'main_boundary_condition_right', 'batch', 'dist']
