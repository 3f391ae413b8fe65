"""Load the reference's own functions without importing the reference module.

TEST INFRASTRUCTURE, build container only.  /root/reference does not exist on
the GPU box, so nothing that runs there may call this; it is used by
tests/golden/make_golden.py to generate the committed golden vectors and by
CPU tests that skip when the reference tree is absent.

``import`` of P: fails at P:3 (matplotlib) and P:5 (skfem), neither installed.
The module-level ``def``s (P:8-105: true_solution, poisson_rhs, the two boundary
functions and lssvr_primal) need only numpy and scipy, so they are taken out of
the parsed AST and executed, unchanged, in a namespace that provides those.
"""
import ast
import os

REFERENCE_PRIMAL = '/root/reference/1D-Possion/Hybrid-FEM-LSSVR.py'
REFERENCE_DUAL = '/root/reference/1D-Possion/Hybrid-FEM-LSSVR-Dual.py'


def reference_available(path=REFERENCE_PRIMAL):
    return os.path.isfile(path)


def load_reference_functions(path=REFERENCE_PRIMAL, with_class=False):
    """Return a dict with the reference's module-level functions (and optionally the class)."""
    import numpy as np
    from scipy.optimize import minimize
    from numpy.polynomial.legendre import Legendre, leggauss
    with open(path, 'r') as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            or (with_class and isinstance(n, ast.ClassDef))]
    ns = {'np': np, 'minimize': minimize, 'Legendre': Legendre, 'leggauss': leggauss}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
    return ns
