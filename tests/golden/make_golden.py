"""Generates tests/golden/*.npz from the oracle (seeded inputs, tiny meshes).

The reference ships no golden vectors and cannot run offline,
so these fixtures freeze the ORACLE's outputs: they guard the oracle against
regressions and give the GPU tests a second, file-based comparison target.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import mesh as om, families as fam, assembly as asm, solvers  # noqa: E402
from _cases import square_boundary_lists  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(famid, n):
    m = om.unit_square_tri(n)
    rng = np.random.default_rng(100 * famid + n)
    if famid == 1:
        F = fam.PoissonP1(m)
        x = m.coords
        F.u_ex = 1.0 / (2 * np.pi ** 2) * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])
        bc = asm.DirichletBC(F.N, square_boundary_lists(m.coords), 0.0)
    else:
        F = fam.NonlinearPoissonP1(m)
        bc = None
    u = rng.standard_normal(F.N)
    f = rng.standard_normal(F.M)
    sp = solvers.StatePath(F, bc)
    A = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), None)
    Abc = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), bc)
    D = asm.assemble_matrix(F.dRdm(0, u, f), (F.N, F.M), None)
    fs = 0.1 * np.ones(F.M) if famid == 2 else f
    if famid == 2:
        us, info = sp.solve_snes(np.zeros(F.N), [fs])
    else:
        us, info = sp.solve_newton(np.zeros(F.N), [fs])
    (g,), lam = sp.total_derivative(0, us, [fs])
    np.savez_compressed(
        os.path.join(HERE, 'family%d_n%d.npz' % (famid, n)),
        u=u, f=f, R=asm.assemble_vector(F.residual(u, f), F.N),
        rowptr=A.indptr.astype(np.int32), col=A.indices.astype(np.int32), J=A.data, Jbc=Abc.data,
        d_rowptr=D.indptr.astype(np.int32), d_col=D.indices.astype(np.int32), D=D.data,
        out=asm.assemble_scalar(F.output(0, u, f)),
        out_du=asm.assemble_vector(F.output_du(0, u, f), F.N),
        out_dm=asm.assemble_vector(F.output_dm(0, 0, u, f), F.M),
        newton_F=sp.newton_F(u, [f]),
        f_solve=fs, u_solved=us, newton_its=info['iterations'], lam=lam, total=g)


if __name__ == '__main__':
    for famid in (1, 2):
        for n in (2, 4, 16):
            dump(famid, n)
    print('wrote', sorted(p for p in os.listdir(HERE) if p.endswith('.npz')))
