"""CPU oracle for femo-b200 -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's arithmetic lives in un-vendored third-party
packages (fenics-dolfinx 0.5.1, basix 0.5.x, UFL 2022.2, FFCx 0.5.x, PETSc +
MUMPS; /root/reference/README.md:21) that cannot be installed here, and the
reference ships no tests or golden vectors.  This package restates the
algorithm of the hot path from the reference's call sites and is pinned only
by exact symbolic integration of the reference's weak forms typed into sympy
(tests/test_oracle_symbolic.py: every family, 1e-13), analytic known answers,
manufactured solutions and finite differences (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under femo_b200/ does.
"""
