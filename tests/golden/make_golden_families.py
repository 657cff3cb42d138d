"""Generates tests/golden/case_*.npz for the families beyond 1 and 2 (beam, SIMP Q1, SIMP hexahedra, P2, motor
magnetostatics, motor mesh motion) from the oracle, on the seeded cases of tests/_cases*.py.

Same role as make_golden.py: the reference ships no golden vectors and cannot run offline, so the fixtures freeze
the ORACLE; tests/test_golden_families.py checks the oracle (CPU) and the CUDA path (GPU) against them.
Run from the repo root:  python tests/golden/make_golden_families.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import assembly as asm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    """name -> builder of a seeded case (engine problem not uploaded: no GPU needed)."""
    from _cases import Case
    from _cases34 import BeamCase, SimpCase, HexCase
    from _cases_motor import MotorCase, MotorMMCase
    from femo_b200 import engine as E
    return {
        'beam50': lambda **k: BeamCase(50, seed=21, **k),
        'simp_q1_8x4': lambda **k: SimpCase(8, 4, seed=22, **k),
        'simp_hex_4x3x2': lambda **k: HexCase(4, 3, 2, seed=23, **k),
        'nlpoisson_p2_n4': lambda **k: Case(E.FAMILY_NLPOISSON_P2, 4, seed=24, **k),
        'motor_em_4x12': lambda **k: MotorCase(4, 12, seed=25, **k),
        'motor_mm_4x12': lambda **k: MotorMMCase(4, 12, seed=26, **k),
    }


def fields(c):
    m = c.m if hasattr(c, 'm') else c.f
    return c.u, m


def oracle_arrays(c):
    F = c.F
    u, m = fields(c)
    A = asm.assemble_matrix(F.jacobian(u, m), (F.N, F.N), None)
    D = asm.assemble_matrix(F.dRdm(0, u, m), (F.N, F.M), None)
    out = dict(u=u, m=m, R=asm.assemble_vector(F.residual(u, m), F.N), rowptr=A.indptr.astype(np.int32),
               col=A.indices.astype(np.int32), J=A.data, d_rowptr=D.indptr.astype(np.int32), d_col=D.indices.astype(np.int32),
               D=D.data, n_outputs=F.n_outputs)
    for k in range(F.n_outputs):
        out['out%d' % k] = asm.assemble_scalar(F.output(k, u, m))
        out['out%d_du' % k] = asm.assemble_vector(F.output_du(k, u, m), F.N)
        out['out%d_dm' % k] = asm.assemble_vector(F.output_dm(k, 0, u, m), F.M)
    return out


if __name__ == '__main__':
    for name, make in cases().items():
        np.savez_compressed(os.path.join(HERE, 'case_%s.npz' % name), **oracle_arrays(make(upload=False)))
    print('wrote', sorted(p for p in os.listdir(HERE) if p.startswith('case_')))
