"""CUDA path vs oracle on identical meshes and seeded inputs, through the C ABI.

Tolerances are BASELINE.json's: assembled residual / Jacobian values within
1e-12 relative, states within solver tolerance, dJ/dm within 1e-8 relative."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import assembly as asm
from _cases import Case, relerr, relerr_entry

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize('famid', [1, 2])
@pytest.mark.parametrize('n,ny', [(1, 1), (4, 4), (16, 16), (33, 7)])
def test_assembly_matches_oracle(cuda_device, famid, n, ny):
    c = Case(famid, n, ny, seed=n)
    F, p, m = c.F, c.p, [c.f]
    R = p.assemble_residual().cpu().numpy()
    assert relerr(R, asm.assemble_vector(F.residual(c.u, *m), F.N)) < TOL
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    A = asm.assemble_matrix(F.jacobian(c.u, *m), (F.N, F.N), None)
    Abc = asm.assemble_matrix(F.jacobian(c.u, *m), (F.N, F.N), c.bc)
    assert relerr(vals.cpu().numpy(), A.data) < TOL
    assert relerr(vals_bc.cpu().numpy(), Abc.data) < TOL
    D = asm.assemble_matrix(F.dRdm(0, c.u, *m), (F.N, F.M), None)
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), D.data) < TOL
    # entrywise (not max-norm) relative agreement of every assembled value
    assert relerr_entry(R, asm.assemble_vector(F.residual(c.u, *m), F.N)) < 1e-11
    assert relerr_entry(vals.cpu().numpy(), A.data) < 1e-11
    assert relerr_entry(vals_bc.cpu().numpy(), Abc.data) < 1e-11
    assert relerr_entry(p.assemble_dRdm(0).cpu().numpy(), D.data) < 1e-12
    J = p.assemble_output(0)
    Jo = asm.assemble_scalar(F.output(0, c.u, *m))
    assert abs(J - Jo) <= TOL * abs(Jo)
    gu = p.assemble_output_grad(0, 0).cpu().numpy()
    assert relerr(gu, asm.assemble_vector(F.output_du(0, c.u, *m), F.N)) < TOL
    gm = p.assemble_output_grad(0, 1).cpu().numpy()
    assert relerr(gm, asm.assemble_vector(F.output_dm(0, 0, c.u, *m), F.M)) < TOL


@pytest.mark.parametrize('famid,bc', [(1, True), (2, True), (2, False)])
def test_newton_rhs_lifting(cuda_device, famid, bc):
    rng = np.random.default_rng(5)
    g = rng.standard_normal((17 + 1) ** 2)
    c = Case(famid, 17, seed=3, bc=bc, g=g if bc else None)
    vals, _ = c.p.assemble_jacobian(plain=True, bc=False)
    b = c.p.newton_rhs(vals).cpu().numpy()
    assert relerr(b, c.sp.newton_F(c.u, [c.f])) < TOL


@pytest.mark.parametrize('famid', [1, 2])
def test_spmv_forward_and_transposed(cuda_device, famid):
    c = Case(famid, 21, 13, seed=2)
    rng = np.random.default_rng(1)
    vals, _ = c.p.assemble_jacobian()
    A = c.csr(0, vals)
    x = rng.standard_normal(c.F.N)
    assert relerr(c.p.spmv(0, vals, c.p.to_device(x)).cpu().numpy(), A @ x) < 1e-13
    assert relerr(c.p.spmv(0, vals, c.p.to_device(x), transpose=True).cpu().numpy(), A.T @ x) < 1e-13
    dv = c.p.assemble_dRdm(0)
    D = c.csr(1, dv)
    xm = rng.standard_normal(c.F.M)
    assert relerr(c.p.spmv(1, dv, c.p.to_device(xm)).cpu().numpy(), D @ xm) < 1e-13
    assert relerr(c.p.spmv(1, dv, c.p.to_device(x), transpose=True).cpu().numpy(), D.T @ x) < 1e-13


def test_transposed_solve_nonsymmetric_values(cuda_device):
    """The adjoint solve must use the true transpose (quirk B3): perturb the
    values so A != A^T on the symmetric pattern and compare with SuperLU."""
    c = Case(1, 12, seed=7)
    _, vals_bc = c.p.assemble_jacobian(plain=False, bc=True)
    rng = np.random.default_rng(0)
    v = vals_bc.cpu().numpy()
    rp, col = c.p.pattern(0)
    rows = np.repeat(np.arange(c.F.N), np.diff(rp))
    v = v * (1.0 + 0.05 * rng.random(v.size) * (rows != col))
    d_v = c.p.to_device(v)
    A = c.csr(0, d_v)
    b = rng.standard_normal(c.F.N)
    # CG needs SPD: solve with the symmetric part (explicit zeros kept so data stays aligned)
    AT = A.T.tocsr()
    AT.sort_indices()
    assert np.array_equal(AT.indices, A.indices)
    vs = 0.5 * (A.data + AT.data)
    As = c.csr(0, c.p.to_device(vs))
    x, info = c.p.linear_solve(c.p.to_device(vs), c.p.to_device(b), transpose=True, rtol=1e-12)
    assert info['converged'], info
    assert relerr(x.cpu().numpy(), spla.spsolve(As.tocsc(), b)) < 1e-9
    # and the transposed SpMV on the genuinely nonsymmetric values
    y = c.p.spmv(0, d_v, c.p.to_device(b), transpose=True).cpu().numpy()
    assert relerr(y, A.T @ b) < 1e-13


@pytest.mark.parametrize('famid,kind', [(1, 'Newton'), (2, 'SNES'), (2, 'Newton')])
def test_state_solve_matches_oracle(cuda_device, famid, kind):
    c = Case(famid, 16, seed=4)
    f = 0.1 * np.ones(c.F.M) if famid == 2 else c.f
    c.set_input(f)
    c.set_state(np.zeros(c.F.N))
    info = c.p.newton_solve(kind=kind, krylov_rtol=1e-13)
    u = c.d_u.cpu().numpy()
    if kind == 'SNES':
        uo, oinfo = c.sp.solve_snes(np.zeros(c.F.N), [f])
    else:
        uo, oinfo = c.sp.solve_newton(np.zeros(c.F.N), [f])
        assert info['iterations'] == 3          # quirk B1
    assert relerr(u, uo) < 1e-9
    assert info['iterations'] == oinfo['iterations']


@pytest.mark.parametrize('famid', [1, 2])
def test_total_derivative_matches_oracle(cuda_device, famid):
    """dJ/dm through the reference's callback chain."""
    c = Case(famid, 16, seed=9)
    f = 0.1 * np.ones(c.F.M) if famid == 2 else c.f
    c.set_input(f)
    c.set_state(np.zeros(c.F.N))
    c.p.newton_solve(kind='SNES' if famid == 2 else 'Newton', krylov_rtol=1e-13)
    u = c.d_u.cpu().numpy()
    p = c.p
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    dv = p.assemble_dRdm(0)
    dJdu = p.assemble_output_grad(0, 0)
    lam, info = p.linear_solve(vals_bc if c.bc is not None else vals, dJdu, transpose=True, rtol=1e-13)
    assert info['converged']
    g = p.assemble_output_grad(0, 1).cpu().numpy() - p.spmv(1, dv, lam, transpose=True).cpu().numpy()
    (go,), lamo = c.sp.total_derivative(0, u, [f])
    assert relerr(lam.cpu().numpy(), lamo) < 1e-8
    assert relerr(g, go) < 1e-8


@pytest.mark.parametrize('famid,n,ny', [(1, 64, 64), (1, 37, 21), (2, 64, 64), (2, 50, 23), (2, 9, 9)])
def test_multigrid_pcg_matches_direct_solve(cuda_device, famid, n, ny):
    """GMG-preconditioned CG (nested and non-nested level sizes, with and without
    Dirichlet rows) against SuperLU, and mesh-independent iteration counts."""
    c = Case(famid, n, ny, seed=1, mg=True)
    u0 = 0.3 * np.sin(3 * c.omesh.coords[:, 0]) * np.cos(2 * c.omesh.coords[:, 1])
    c.set_state(u0)
    vals, vals_bc = c.p.assemble_jacobian(plain=True, bc=True)
    v = vals_bc if c.bc is not None else vals
    A = c.csr(0, v)
    rng = np.random.default_rng(2)
    b = rng.standard_normal(c.F.N)
    x, info = c.p.linear_solve(v, c.p.to_device(b), rtol=1e-12, precond=2)
    assert info['converged'], info
    assert relerr(x.cpu().numpy(), spla.spsolve(A.tocsc(), b)) < 1e-9
    assert info['iterations'] <= 30, info
    xj, ij = c.p.linear_solve(v, c.p.to_device(b), rtol=1e-12, precond=0)
    assert ij['iterations'] > info['iterations']


def test_newton_with_multigrid(cuda_device):
    c = Case(2, 48, seed=4, mg=True)
    f = 0.1 * np.ones(c.F.M)
    c.set_input(f)
    c.set_state(np.zeros(c.F.N))
    info = c.p.newton_solve(kind='SNES', krylov_rtol=1e-12, precond=2)
    uo, oinfo = c.sp.solve_snes(np.zeros(c.F.N), [f])
    assert relerr(c.d_u.cpu().numpy(), uo) < 1e-9
    assert info['iterations'] == oinfo['iterations']
    assert info['krylov_iterations'] <= 40


import glob as _glob
import os as _os
_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('path', sorted(_glob.glob(_os.path.join(_GOLD, 'family*_n*.npz'))))
def test_against_golden_fixtures(cuda_device, path):
    """CUDA path vs the committed oracle dumps (tests/golden/make_golden.py)."""
    z = np.load(path)
    famid = int(_os.path.basename(path)[6])
    n = int(_os.path.basename(path).split('_n')[1].split('.')[0])
    c = Case(famid, n, mg=True)
    c.set_state(z['u'])
    c.set_input(z['f'])
    p = c.p
    rp, col = p.pattern(0)
    assert np.array_equal(rp, z['rowptr']) and np.array_equal(col, z['col'])
    rp, col = p.pattern(1)
    assert np.array_equal(rp, z['d_rowptr']) and np.array_equal(col, z['d_col'])
    assert relerr(p.assemble_residual().cpu().numpy(), z['R']) < TOL
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    assert relerr(vals.cpu().numpy(), z['J']) < TOL
    assert relerr(vals_bc.cpu().numpy(), z['Jbc']) < TOL
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), z['D']) < TOL
    assert abs(p.assemble_output(0) - float(z['out'])) <= TOL * abs(float(z['out']))
    assert relerr(p.assemble_output_grad(0, 0).cpu().numpy(), z['out_du']) < TOL
    assert relerr(p.assemble_output_grad(0, 1).cpu().numpy(), z['out_dm']) < TOL
    assert relerr(p.newton_rhs(vals).cpu().numpy(), z['newton_F']) < TOL
    # state solve + adjoint total derivative
    c.set_input(z['f_solve'])
    c.set_state(np.zeros(c.F.N))
    info = p.newton_solve(kind='SNES' if famid == 2 else 'Newton', krylov_rtol=1e-13, precond=2)
    assert info['iterations'] == int(z['newton_its'])
    assert relerr(c.d_u.cpu().numpy(), z['u_solved']) < 1e-9
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    lam, li = p.linear_solve(vals_bc if c.bc is not None else vals, p.assemble_output_grad(0, 0), transpose=True,
                             rtol=1e-13, precond=2)
    g = p.assemble_output_grad(0, 1).cpu().numpy() - p.spmv(1, p.assemble_dRdm(0), lam, transpose=True).cpu().numpy()
    assert relerr(lam.cpu().numpy(), z['lam']) < 1e-8
    assert relerr(g, z['total']) < 1e-8


def test_bitwise_reproducible_run_to_run(cuda_device):
    """Determinism (sorted segmented reduction, fixed reduction trees): two runs give identical bits."""
    outs = []
    for _ in range(2):
        c = Case(2, 40, 24, seed=11, mg=True)
        c.set_input(0.1 * np.ones(c.F.M))
        c.set_state(np.zeros(c.F.N))
        c.p.newton_solve(kind='SNES', krylov_rtol=1e-12, precond=2)
        vals, _ = c.p.assemble_jacobian()
        outs.append((c.d_u.cpu().numpy().copy(), vals.cpu().numpy().copy(), c.p.assemble_output(0)))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]


def test_edge_cases(cuda_device):
    """Smallest meshes, empty Dirichlet set, BC change after upload, zero inputs."""
    c = Case(1, 1, 1, seed=0, bc=False)                      # one square = two cells, no BC
    R = c.p.assemble_residual().cpu().numpy()
    assert relerr(R, asm.assemble_vector(c.F.residual(c.u, c.f), c.F.N)) < TOL
    _, vbc = c.p.assemble_jacobian(plain=False, bc=True)     # no BC set: BC'd copy equals the plain matrix
    assert relerr(vbc.cpu().numpy(), asm.assemble_matrix(c.F.jacobian(c.u, c.f), (c.F.N, c.F.N)).data) < TOL
    # install Dirichlet data after upload, then clear it again
    from _cases import square_boundary_lists
    lists = square_boundary_lists(c.omesh.coords)
    c.p.set_bc(lists)
    bc = asm.DirichletBC(c.F.N, lists, 0.0)
    _, vbc = c.p.assemble_jacobian(plain=False, bc=True)
    assert relerr(vbc.cpu().numpy(), asm.assemble_matrix(c.F.jacobian(c.u, c.f), (c.F.N, c.F.N), bc).data) < TOL
    c.p.set_bc([])
    _, vbc = c.p.assemble_jacobian(plain=False, bc=True)
    assert relerr(vbc.cpu().numpy(), asm.assemble_matrix(c.F.jacobian(c.u, c.f), (c.F.N, c.F.N)).data) < TOL
    # all-zero state and input
    c2 = Case(2, 5, 3, seed=1)
    c2.set_state(np.zeros(c2.F.N))
    c2.set_input(np.zeros(c2.F.M))
    assert relerr(c2.p.assemble_residual().cpu().numpy(),
                  asm.assemble_vector(c2.F.residual(np.zeros(c2.F.N), np.zeros(c2.F.M)), c2.F.N)) < TOL


def test_hypothesis_random_fields(cuda_device):
    """Seeded random coefficient fields of varying magnitude (SURVEY.md section 8c golden-vector plan)."""
    from hypothesis import given, settings, strategies as st
    c = Case(2, 9, 6, seed=0)

    @settings(max_examples=15, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.floats(1e-3, 1e3))
    def run(seed, scale):
        rng = np.random.default_rng(seed)
        u, f = scale * rng.standard_normal(c.F.N), rng.standard_normal(c.F.M) / scale
        c.set_state(u)
        c.set_input(f)
        assert relerr(c.p.assemble_residual().cpu().numpy(), asm.assemble_vector(c.F.residual(u, f), c.F.N)) < 1e-11
        vals, _ = c.p.assemble_jacobian()
        assert relerr(vals.cpu().numpy(), asm.assemble_matrix(c.F.jacobian(u, f), (c.F.N, c.F.N)).data) < 1e-11
    run()
