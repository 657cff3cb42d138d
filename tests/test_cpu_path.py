"""The C++/OpenMP restatement (oracle/cpu_path.cpp: the CPU baseline of bench.py and the fast second oracle of the
GPU tests at n >= 1024) pinned against the numpy oracle: element-level quantities to round-off, the solved state,
adjoint and total derivative against SuperLU."""
import numpy as np
import pytest

from oracle import mesh as om, families as fam, assembly as asm, solvers, cpu_path
from _cases import relerr


@pytest.mark.parametrize('n', [1, 2, 5, 16])
def test_assembly_matches_numpy_oracle(n):
    F = fam.NonlinearPoissonP1(om.unit_square_tri(n))
    rng = np.random.default_rng(n)
    u, f = rng.standard_normal(F.N), rng.standard_normal(F.M)
    assert relerr(cpu_path.residual(n, u, f), asm.assemble_vector(F.residual(u, f), F.N)) < 1e-12
    A = cpu_path.jacobian(n, u)
    Ao = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), None)
    assert abs(A - Ao).max() < 1e-12 * abs(Ao).max()
    J, dJdu, dJdf = cpu_path.output(n, u, f)
    Jo = asm.assemble_scalar(F.output(0, u, f))
    assert abs(J - Jo) <= 1e-12 * abs(Jo)
    assert relerr(dJdu, asm.assemble_vector(F.output_du(0, u, f), F.N)) < 1e-12
    assert relerr(dJdf, asm.assemble_vector(F.output_dm(0, 0, u, f), F.M)) < 1e-12


@pytest.mark.parametrize('n', [16, 48, 96])
def test_state_adjoint_gradient_match_direct_solve(n):
    F = fam.NonlinearPoissonP1(om.unit_square_tri(n))
    sp = solvers.StatePath(F, None)
    f = 0.1 * np.ones(F.M)
    r = cpu_path.step(n, f, krylov_rtol=1e-12)
    assert r['converged']
    uo, oinfo = sp.solve_snes(np.zeros(F.N), [f])
    assert r['newton_its'] == oinfo['iterations']
    assert relerr(r['u'], uo) < 1e-9
    (go,), lamo = sp.total_derivative(0, uo, [f])
    assert relerr(r['lam'], lamo) < 1e-8
    assert relerr(r['grad'], go) < 1e-8
    assert abs(r['J'] - asm.assemble_scalar(F.output(0, uo, f))) < 1e-10 * abs(r['J'])


def test_multigrid_iterations_mesh_independent():
    its = [cpu_path.step(n, 0.1)['krylov_its'] for n in (64, 256)]
    assert max(its) <= 14 and abs(its[0] - its[1]) <= 3, its
