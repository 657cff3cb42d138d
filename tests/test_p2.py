"""Quadratic Lagrange nonlinear Poisson (family 9, config 2 "P2"): layouts vs the oracle on the CPU,
CUDA assembly / solves vs the oracle on the GPU, p-multigrid (P2 -> P1 -> lattice hierarchy)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from femo_b200 import engine as E
from oracle import assembly as asm, mesh as om
from _cases import Case, relerr

P2 = E.FAMILY_NLPOISSON_P2


@pytest.mark.parametrize('n,ny', [(1, 1), (4, 3), (16, 16)])
def test_p2_layout_bit_exact(n, ny):
    c = Case(P2, n, ny, upload=False)
    F, p = c.F, c.p
    ev, ce = om.triangle_edges(c.omesh)
    pev, pce = p.edges()
    assert np.array_equal(pev, ev) and np.array_equal(pce, ce)
    assert p.N == F.N == c.omesh.nverts + len(ev) and p.M == [F.M]
    u, f = np.zeros(F.N), np.ones(F.M)
    for which, blocks, shape in ((0, F.jacobian(u, f), (F.N, F.N)), (1, F.dRdm(0, u, f), (F.N, F.M))):
        rp, col = p.pattern(which)
        orp, ocol = asm.pattern(blocks, shape)
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)


def test_p2_sizes_of_survey_8d():
    """n = 2000 gives the same 16 008 001 dofs as P1 at n = 4000."""
    n = 2000
    assert (n + 1) ** 2 + (2 * n * (n + 1) + n * n) == 16008001
    c = Case(P2, 8, upload=False)
    assert c.p.N == 17 * 17


def test_p2_oracle_reproduces_quadratics():
    """P2 contains x^2 + y^2: with f = u^3 - lap(u) the interior residual rows vanish to round-off."""
    from oracle import families as fam
    m = om.unit_square_tri(6)
    F = fam.NonlinearPoissonP2(m)
    ev, _ = om.triangle_edges(m)
    X = np.concatenate([m.coords, 0.5 * (m.coords[ev[:, 0]] + m.coords[ev[:, 1]])])
    u = X[:, 0] ** 2 + X[:, 1] ** 2
    # cellwise constant f cannot represent u^3 + 4, so test the Laplacian part: residual of the linearised
    # operator applied to u equals -int 4 v on interior rows
    A = asm.assemble_matrix(F.jacobian(np.zeros(F.N), np.zeros(F.M)), (F.N, F.N), None)   # stiffness + Nitsche
    rhs = -asm.assemble_matrix(F.dRdm(0, u, np.zeros(F.M)), (F.N, F.M), None) @ np.full(F.M, -4.0)   # int (-lap u) v
    interior = np.ones(F.N, dtype=bool)
    fc, fl = m.exterior_facets()
    interior[np.unique(F.cell_dofs[fc])] = False
    r = A @ u - rhs
    assert np.abs(r[interior]).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('n,ny', [(1, 1), (4, 4), (16, 16), (33, 7)])
def test_gpu_p2_assembly_matches_oracle(cuda_device, n, ny):
    c = Case(P2, n, ny, seed=n)
    F, p, m = c.F, c.p, [c.f]
    TOL = 1e-12
    assert relerr(p.assemble_residual().cpu().numpy(), asm.assemble_vector(F.residual(c.u, *m), F.N)) < TOL
    vals, _ = p.assemble_jacobian(plain=True, bc=False)
    assert relerr(vals.cpu().numpy(), asm.assemble_matrix(F.jacobian(c.u, *m), (F.N, F.N), None).data) < TOL
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), asm.assemble_matrix(F.dRdm(0, c.u, *m), (F.N, F.M), None).data) < TOL
    Jo = asm.assemble_scalar(F.output(0, c.u, *m))
    assert abs(p.assemble_output(0) - Jo) <= TOL * abs(Jo)
    assert relerr(p.assemble_output_grad(0, 0).cpu().numpy(), asm.assemble_vector(F.output_du(0, c.u, *m), F.N)) < TOL
    assert relerr(p.assemble_output_grad(0, 1).cpu().numpy(), asm.assemble_vector(F.output_dm(0, 0, c.u, *m), F.M)) < TOL
    assert relerr(p.newton_rhs(vals).cpu().numpy(), c.sp.newton_F(c.u, m)) < TOL


@pytest.mark.gpu
def test_gpu_p2_state_adjoint_and_pmultigrid(cuda_device):
    """SNES state and adjoint gradient vs the oracle (SuperLU); the p-multigrid hierarchy (P2 -> P1 on the same
    mesh -> coarsened lattices) must beat Jacobi-CG by a wide margin and be mesh independent."""
    c = Case(P2, 24, seed=3, upload=False, mg=True)
    assert c.p.mg_levels >= 3
    c.upload()
    c.set_input(0.1 + 0.05 * np.random.default_rng(1).standard_normal(c.F.M))
    c.set_state(np.zeros(c.F.N))
    info = c.p.newton_solve(kind='SNES', precond=2, krylov_rtol=1e-12)
    assert info['converged'] > 0, info
    uo, _ = c.sp.solve_snes(np.zeros(c.F.N), [c.f])
    u = c.d_u.cpu().numpy()
    assert relerr(u, uo) < 1e-8
    vals, _ = c.p.assemble_jacobian()
    lam, li = c.p.linear_solve(vals, c.p.assemble_output_grad(0, 0), transpose=True, rtol=1e-13, precond=2, max_it=400)
    assert li['converged'], li
    g = c.p.assemble_output_grad(0, 1).cpu().numpy() - c.p.spmv(1, c.p.assemble_dRdm(0), lam, transpose=True).cpu().numpy()
    (go,), lamo = c.sp.total_derivative(0, uo, [c.f])
    assert relerr(lam.cpu().numpy(), lamo) < 1e-7
    assert relerr(g, go) < 1e-7
    its = {}
    for n in (32, 128):
        d = Case(P2, n, seed=4, upload=False, mg=True)
        d.upload()
        d.set_state(np.zeros(d.F.N))
        v, _ = d.p.assemble_jacobian()
        b = d.p.to_device(np.random.default_rng(2).standard_normal(d.F.N))
        x, i2 = d.p.linear_solve(v, b, rtol=1e-10, precond=2, max_it=400)
        assert i2['converged'], i2
        its[n] = i2['iterations']
        if n == 32:
            A = d.csr(0, v)
            assert relerr(x.cpu().numpy(), spla.spsolve(A.tocsc(), b.cpu().numpy())) < 1e-7
            _, i0 = d.p.linear_solve(v, b, rtol=1e-10, precond=0, max_it=100000, check_every=20)
            assert i2['iterations'] * 4 < i0['iterations'], (i2, i0)
    assert its[128] <= its[32] + 6, its
