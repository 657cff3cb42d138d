"""CUDA path vs oracle for the Euler-Bernoulli beam (config 3) and SIMP Q1
elasticity (config 4): assembly to 1e-12, states / adjoint gradients within solver
tolerance, analytic known answers of SURVEY.md section 4."""
import numpy as np
import pytest

from oracle import assembly as asm
from _cases import relerr
from _cases34 import BeamCase, SimpCase, HexCase, set_state, set_input

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize('make', [lambda: BeamCase(50, seed=1), lambda: BeamCase(7, seed=2),
                                  lambda: SimpCase(8, 4, seed=3), lambda: SimpCase(80, 40, seed=4, rho_lo=1e-4),
                                  lambda: HexCase(4, 3, 2, seed=11), lambda: HexCase(9, 5, 7, seed=12, rho_lo=1e-3)])
def test_assembly_matches_oracle(cuda_device, make):
    c = make()
    F, p, m = c.F, c.p, [c.m]
    assert relerr(p.assemble_residual().cpu().numpy(), asm.assemble_vector(F.residual(c.u, *m), F.N)) < TOL
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    assert relerr(vals.cpu().numpy(), asm.assemble_matrix(F.jacobian(c.u, *m), (F.N, F.N), None).data) < TOL
    assert relerr(vals_bc.cpu().numpy(), asm.assemble_matrix(F.jacobian(c.u, *m), (F.N, F.N), c.bc).data) < TOL
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), asm.assemble_matrix(F.dRdm(0, c.u, *m), (F.N, F.M), None).data) < TOL
    assert relerr(p.newton_rhs(vals).cpu().numpy(), c.sp.newton_F(c.u, m)) < TOL
    for k in range(2):
        Jo = asm.assemble_scalar(F.output(k, c.u, *m))
        assert abs(p.assemble_output(k) - Jo) <= TOL * max(abs(Jo), 1e-300)
        assert relerr(p.assemble_output_grad(k, 0).cpu().numpy(), asm.assemble_vector(F.output_du(k, c.u, *m), F.N)) < TOL
        assert relerr(p.assemble_output_grad(k, 1).cpu().numpy(), asm.assemble_vector(F.output_dm(k, 0, c.u, *m), F.M)) < TOL


def test_beam_known_answer(cuda_device):
    """Tip load -1, E=1, L=1, b=h=0.1: tip deflection P L^3/(3 E I) = -4.0e4,
    compliance 4.0e4, volume 0.01 (Hermite cubics are exact here)."""
    c = BeamCase(50)
    set_input(c, np.full(50, 0.1))
    set_state(c, np.zeros(c.F.N))
    info = c.p.newton_solve(kind='Newton', precond=3, krylov_rtol=1e-12)
    assert info['iterations'] == 3
    u = c.d_u.cpu().numpy()
    assert abs(u[100] + 4.0e4) < 1e-5 * 4.0e4
    assert abs(c.p.assemble_output(0) - 4.0e4) < 1e-5 * 4.0e4
    assert abs(c.p.assemble_output(1) - 0.01) < 1e-14
    uo, _ = c.sp.solve_newton(np.zeros(c.F.N), [c.m])
    assert relerr(u, uo) < 1e-8


@pytest.mark.parametrize('make,precond', [(lambda: BeamCase(50, seed=5), 3), (lambda: SimpCase(8, 4, seed=6), 3),
                                          (lambda: SimpCase(24, 12, seed=7), 0), (lambda: HexCase(6, 4, 3, seed=13), 0)])
def test_state_and_adjoint_match_oracle(cuda_device, make, precond):
    c = make()
    p, F = c.p, c.F
    set_state(c, np.zeros(F.N))
    info = p.newton_solve(kind='Newton', precond=precond, krylov_rtol=1e-13, krylov_max_it=200000, check_every=50)
    assert info['iterations'] == 3
    uo, _ = c.sp.solve_newton(np.zeros(F.N), [c.m])
    u = c.d_u.cpu().numpy()
    assert relerr(u, uo) < 1e-8
    k = 0 if isinstance(c, BeamCase) else 1                       # compliance
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    lam, li = p.linear_solve(vals_bc, p.assemble_output_grad(k, 0), transpose=True, rtol=1e-13, precond=precond,
                             max_it=200000, check_every=50)
    assert li['converged']
    g = p.assemble_output_grad(k, 1).cpu().numpy() - p.spmv(1, p.assemble_dRdm(0), lam, transpose=True).cpu().numpy()
    (go,), lamo = c.sp.total_derivative(k, uo, [c.m])
    assert relerr(lam.cpu().numpy(), lamo) < 1e-7
    assert relerr(g, go) < 1e-7


def test_beam_thickness_optimisation_reaches_reference_optimum(cuda_device):
    """The reference embeds the OpenMDAO optimum `thick_ref`
    (run_thickness_opt_cantilever_beam.py:252-261).  Minimise compliance s.t. volume = b h L with
    SLSQP through the femo_b200 API (FEA + FEAModel + Simulator) and compare."""
    import scipy.optimize as so
    from femo_b200.fea.fea_b200 import (FEA, createIntervalMesh, FunctionSpace, Function, TestFunction, Constant,
                                         locate_entities_boundary, locate_dofs_topological, meshtags, Measure)
    from femo_b200.forms.beam import pdeRes, compliance, volume
    from femo_b200.csdl_opt import FEAModel, Simulator
    E, L, b, h, nel = 1.0, 1.0, 0.1, 0.1, 50
    mesh = createIntervalMesh(nel, 0.0, L)
    fea = FEA(mesh)
    t = Function(FunctionSpace(mesh, ('DG', 0)))
    V = FunctionSpace(mesh, ('Hermite', 3))
    u = Function(V)
    f = Constant(mesh, -1.0)
    endpoint = locate_entities_boundary(mesh, 0, lambda x: np.isclose(x[0], L))
    assert endpoint.tolist() == [nel]
    facet_tag = meshtags(mesh, 0, endpoint, np.full(len(endpoint), 100, dtype=np.int32))
    ds_ = Measure('ds', domain=mesh, subdomain_data=facet_tag, metadata={"quadrature_degree": 4})
    residual_form = pdeRes(u, TestFunction(V), t, f, ds_(100), E, b)
    fea.add_input('thickness', t)
    fea.add_state(name='displacements', function=u, residual_form=residual_form, arguments=['thickness'])
    fea.add_output(name='compliance', type='scalar', form=compliance(u, f, ds_(100)), arguments=['thickness', 'displacements'])
    fea.add_output(name='volume', type='scalar', form=volume(t, b, L), arguments=['thickness'])
    ubc = Function(V)
    start = locate_entities_boundary(mesh, 0, lambda x: np.isclose(x[0], 0))
    loc = locate_dofs_topological(V, 0, start)
    fea.add_strong_bc(ubc, [loc[0:1], loc[1:2]])
    fea.REPORT = False
    model = FEAModel(fea=[fea], debug_mode=False)
    model.create_input('thickness', shape=nel, val=h)
    sim = Simulator(model)
    sim.run()
    assert abs(sim['compliance'][0] - 4.0e4) < 1e-5 * 4.0e4 and abs(sim['volume'][0] - 0.01) < 1e-14

    def fun(x):
        sim['thickness'] = x
        sim.run()
        g = sim.compute_totals('compliance', 'thickness')[('compliance', 'thickness')]
        return float(sim['compliance'][0]), g

    def vol(x):
        sim['thickness'] = x
        sim.run()
        return float(sim['volume'][0]) - b * h * L

    def dvol(x):
        return sim.compute_totals('volume', 'thickness')[('volume', 'thickness')]
    res = so.minimize(fun, np.full(nel, h), jac=True, method='SLSQP', bounds=[(1e-2, 10.0)] * nel,
                      constraints=[dict(type='eq', fun=vol, jac=dvol)], options=dict(maxiter=300, ftol=1e-12))
    thick_ref = np.array([
        0.14915754, 0.14764328, 0.14611321, 0.14456715, 0.14300421, 0.14142417, 0.13982611, 0.13820976, 0.13657406,
        0.13491866, 0.13324268, 0.13154528, 0.12982575, 0.12808305, 0.12631658, 0.12452477, 0.12270701, 0.12086183,
        0.11898809, 0.11708424, 0.11514904, 0.11318072, 0.11117762, 0.10913764, 0.10705891, 0.10493903, 0.10277539,
        0.10056526, 0.09830546, 0.09599246, 0.09362243, 0.09119084, 0.08869265, 0.08612198, 0.08347229, 0.08073573,
        0.07790323, 0.07496382, 0.07190453, 0.06870925, 0.0653583, 0.06182632, 0.05808044, 0.05407658, 0.04975295,
        0.0450185, 0.03972912, 0.03363155, 0.02620192, 0.01610863])
    assert np.abs(res.x - thick_ref).max() < 2e-4, np.abs(res.x - thick_ref).max()


@pytest.mark.parametrize('nx,ny,rho_lo', [(64, 32, 0.3), (80, 40, 1e-3), (50, 21, 0.2)])
def test_simp_multigrid_pcg(cuda_device, nx, ny, rho_lo):
    """GMG-preconditioned CG for the vector Q1 elasticity operator (rediscretised coarse levels with
    power-mean densities) against SuperLU, nested and non-nested level sizes, moderate and high contrast."""
    import scipy.sparse.linalg as spla
    from _cases34 import csr
    c = SimpCase(nx, ny, seed=8, upload=False, rho_lo=rho_lo)
    levels = c.p.enable_multigrid()
    assert levels >= 3
    from _cases34 import _upload
    _upload(c)
    _, vals_bc = c.p.assemble_jacobian(plain=False, bc=True)
    A = csr(c, 0, vals_bc)
    b = np.random.default_rng(1).standard_normal(c.F.N)
    b[c.bc.dofs] = 0.0
    x, info = c.p.linear_solve(vals_bc, c.p.to_device(b), rtol=1e-11, precond=2, max_it=500)
    assert info['converged'], info
    xo = spla.spsolve(A.tocsc(), b)
    assert relerr(x.cpu().numpy(), xo) < 1e-7
    xj, ij = c.p.linear_solve(vals_bc, c.p.to_device(b), rtol=1e-11, precond=0, max_it=200000, check_every=100)
    assert info['iterations'] < ij['iterations'] / 4, (info, ij)


@pytest.mark.parametrize('nx,ny,nz,rho_lo', [(16, 8, 8, 0.3), (20, 9, 6, 0.05)])
def test_hex_multigrid_pcg(cuda_device, nx, ny, nz, rho_lo):
    """GMG-preconditioned CG on the 3-D hexahedral SIMP operator (trilinear transfers, rediscretised coarse
    levels) against SuperLU; nested and non-nested level sizes."""
    import scipy.sparse.linalg as spla
    from _cases34 import csr, _upload
    c = HexCase(nx, ny, nz, seed=14, upload=False, rho_lo=rho_lo)
    levels = c.p.enable_multigrid()
    assert levels >= 3
    _upload(c)
    _, vals_bc = c.p.assemble_jacobian(plain=False, bc=True)
    A = csr(c, 0, vals_bc)
    b = np.random.default_rng(1).standard_normal(c.F.N)
    b[c.bc.dofs] = 0.0
    x, info = c.p.linear_solve(vals_bc, c.p.to_device(b), rtol=1e-11, precond=2, max_it=500)
    assert info['converged'], info
    xo = spla.spsolve(A.tocsc(), b)
    assert relerr(x.cpu().numpy(), xo) < 1e-7
    xj, ij = c.p.linear_solve(vals_bc, c.p.to_device(b), rtol=1e-11, precond=0, max_it=200000, check_every=100)
    assert info['iterations'] < ij['iterations'] / 3, (info, ij)
    # SpMV with 81-entry rows (multi-lane row sums) against scipy
    xr = np.random.default_rng(2).standard_normal(c.F.N)
    y = c.p.spmv(0, vals_bc, c.p.to_device(xr)).cpu().numpy()
    assert relerr(y, A @ xr) < 1e-13


def test_hex_general_element_kernels(cuda_device, monkeypatch):
    """Uniform boxes take the K0 fast paths (Jacobian reduced from rho^p x one unit block, matrix-free V-cycle);
    FEMO_NO_MATFREE forces the general per-cell quadrature kernels: both must match the oracle and each other."""
    from _cases34 import csr
    fast = HexCase(5, 4, 3, seed=31)
    vf, vfb = fast.p.assemble_jacobian(plain=True, bc=True)
    monkeypatch.setenv('FEMO_NO_MATFREE', '1')
    gen = HexCase(5, 4, 3, seed=31)
    vg, vgb = gen.p.assemble_jacobian(plain=True, bc=True)
    A = asm.assemble_matrix(gen.F.jacobian(gen.u, gen.m), (gen.F.N, gen.F.N), None).data
    assert relerr(vg.cpu().numpy(), A) < TOL and relerr(vf.cpu().numpy(), A) < TOL
    assert relerr(vfb.cpu().numpy(), vgb.cpu().numpy()) < 1e-13
    # V-cycle with the assembled fp32 copies instead of the matrix-free operator: same solution
    c = HexCase(16, 8, 8, seed=32, upload=False)
    c.p.enable_multigrid()
    from _cases34 import _upload
    _upload(c)
    _, vb = c.p.assemble_jacobian(plain=False, bc=True)
    b = np.random.default_rng(3).standard_normal(c.F.N)
    b[c.bc.dofs] = 0.0
    x1, i1 = c.p.linear_solve(vb, c.p.to_device(b), rtol=1e-11, precond=2, max_it=300)
    x2, i2 = c.p.linear_solve(vb, c.p.to_device(b), rtol=1e-11, precond=2, max_it=300, mg_precision=1)
    assert i1['converged'] and i2['converged'] and abs(i1['iterations'] - i2['iterations']) <= 2
    assert relerr(x1.cpu().numpy(), x2.cpu().numpy()) < 1e-8


def test_hex_bsr3_spmv_and_solve(cuda_device, monkeypatch):
    """BSR-3 (K7): the 3x3-block copy of the hexahedral dR/du values gives the CSR product, and the CG recurrence that
    streams it converges to the same solution in the same number of iterations as the scalar-CSR recurrence."""
    from _cases34 import csr, _upload
    c = HexCase(12, 7, 5, seed=5, upload=False)
    c.p.enable_multigrid()
    _upload(c)
    _, vb = c.p.assemble_jacobian(plain=False, bc=True)
    A = csr(c, 0, vb)
    x = np.random.default_rng(3).standard_normal(c.F.N)
    y = c.p.spmv_bsr3(vb, c.p.to_device(x)).cpu().numpy()
    assert relerr(y, A @ x) < 1e-13
    b = np.random.default_rng(4).standard_normal(c.F.N)
    b[c.bc.dofs] = 0.0
    monkeypatch.setenv('FEMO_BSR', '1')                     # opt-in: the recurrence streams the 3x3-block copy
    x1, i1 = c.p.linear_solve(vb, c.p.to_device(b), rtol=1e-11, precond=2, max_it=300)
    monkeypatch.delenv('FEMO_BSR')
    x0, i0 = c.p.linear_solve(vb, c.p.to_device(b), rtol=1e-11, precond=2, max_it=300)
    assert i1['converged'] and i0['converged'] and i1['iterations'] == i0['iterations']
    assert relerr(x1.cpu().numpy(), x0.cpu().numpy()) < 1e-9
