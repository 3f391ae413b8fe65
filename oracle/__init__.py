"""CPU oracle for the hybrid FEM + LSSVR hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / scipy / mpmath, plus a small C port
under ``oracle/c``) of the algorithm in the reference scripts

    P: /root/reference/1D-Possion/Hybrid-FEM-LSSVR.py
    D: /root/reference/1D-Possion/Hybrid-FEM-LSSVR-Dual.py

It exists to check the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it; nothing under ``hybrid_fem_lssvr_b200/`` does, and the product path raises
when the CUDA library is missing rather than falling back to anything here.

How the oracle is pinned
------------------------
* ``lssvr_primal`` (P:20-105): PINNED.  The reference function itself was
  executed in the build container (AST-extracted from P:, see
  ``oracle/ref_loader.py``) on the shipped configuration and on seeded random
  elements; inputs and outputs are committed under ``tests/golden/`` together
  with the generating script ``tests/golden/make_golden.py``.  The closed-form
  KKT restatement in ``oracle/kkt.py`` reproduces them to the accuracy SLSQP
  reaches (~1e-11) and is itself checked against an 80-digit mpmath solve of
  the same QP (``oracle/kkt_mp.py``).
* ``evaluate_solution`` (P:184-211): PINNED the same way (numpy only).
* ``solve_fem`` (P:117-145): PARITY UNPINNED.  It needs scikit-fem 11.0.0,
  which is not installed and cannot be fetched; ``oracle/fem_p1.py`` restates
  the assembly from scikit-fem's documented behaviour.  The restatement is
  anchored on the analytic discrete solution c(h)*sin(pi x).
* dual LSSVR: the reference contains no dual implementation (D: is a copy of
  P:); ``oracle/dual.py`` derives it and its oracle is "same w as the primal".
"""
