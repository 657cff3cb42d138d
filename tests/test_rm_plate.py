"""Reissner-Mindlin plate family (FEMO_FAMILY_RM_PLATE; the flat case of the reference's shell examples,
examples/test_shell_m3l/shell_pde.py:219-311).  The reference's forms live in the un-vendored shell_analysis_fenicsx, so
parity is UNPINNED; the oracle (oracle/rm_plate.py) is pinned by its own energy (finite differences, symmetry, rigid
motions) and by the Kirchhoff thin-plate limit, the engine's integer layout is compared with `==`, and the CUDA kernels
are compared with the oracle at 1e-12."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from femo_b200 import engine as E
from oracle import mesh as om, assembly as asm
from oracle.rm_plate import RMPlate
from _cases import relerr


def _fields(F, seed=0):
    rng = np.random.default_rng(seed)
    u = 1e-2 * rng.standard_normal(F.N)
    t = 0.05 + 0.02 * rng.random(F.M)
    f = rng.standard_normal(F.M)
    return u, t, f


def test_oracle_energy_consistency():
    F = RMPlate(om.unit_square_tri(5, 4), E=2.0e5, nu=0.3, pen=1.0e6)
    u, t, f = _fields(F)
    A = asm.assemble_matrix(F.jacobian(u, t, f), (F.N, F.N))
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    # the residual is the gradient of U(u) - int f w: R(u) = A u - load, energy output = 1/2 u K_int u
    R = asm.assemble_vector(F.residual(u, t, f), F.N)
    R0 = asm.assemble_vector(F.residual(np.zeros(F.N), t, f), F.N)
    assert relerr(R - R0, A @ u) < 1e-12
    Fint = RMPlate(F.mesh, clamped=[], E=2.0e5, nu=0.3)             # no penalty facets: pure elastic operator
    K = asm.assemble_matrix(Fint.jacobian(u, t, f), (F.N, F.N))
    en = asm.assemble_scalar(Fint.output(2, u, t, f))
    assert abs(en - 0.5 * u @ (K @ u)) < 1e-12 * abs(en)
    # rigid motions w = a + b x + c y, theta = grad w carry no elastic energy
    nv, ne = F.mesh.nverts, F.nedges
    X = F.mesh.coords
    from oracle.mesh import triangle_edges
    ev, _ = triangle_edges(F.mesh)
    Xe = 0.5 * (X[ev[:, 0]] + X[ev[:, 1]])
    rb = np.zeros(F.N)
    rb[:nv] = 0.3 + 0.7 * X[:, 0] - 0.2 * X[:, 1]
    rb[nv:nv + ne] = 0.3 + 0.7 * Xe[:, 0] - 0.2 * Xe[:, 1]
    rb[nv + ne::2], rb[nv + ne + 1::2] = 0.7, -0.2
    assert np.max(np.abs(K @ rb)) < 1e-9 * abs(K).max()
    # thickness derivatives by central differences
    h = 1e-6
    d = np.random.default_rng(1).standard_normal(F.M)
    D = asm.assemble_matrix(F.dRdm(0, u, t, f), (F.N, F.M))
    fd = (asm.assemble_vector(F.residual(u, t + h * d, f), F.N) - asm.assemble_vector(F.residual(u, t - h * d, f), F.N)) / (2 * h)
    assert relerr(D @ d, fd) < 1e-7
    Df = asm.assemble_matrix(F.dRdm(1, u, t, f), (F.N, F.M))
    fd = (asm.assemble_vector(F.residual(u, t, f + d), F.N) - asm.assemble_vector(F.residual(u, t, f), F.N))
    assert relerr(Df @ d, fd) < 1e-9          # difference of two O(K u) residuals: cancellation
    for k in range(3):
        g = asm.assemble_vector(F.output_dm(k, 0, u, t, f), F.M)
        fd = (asm.assemble_scalar(F.output(k, u, t + h * d, f)) - asm.assemble_scalar(F.output(k, u, t - h * d, f))) / (2 * h)
        assert abs(g @ d - fd) <= 1e-6 * max(abs(fd), 1e-12)
        du = np.random.default_rng(2).standard_normal(F.N)
        gu = asm.assemble_vector(F.output_du(k, u, t, f), F.N)
        fd = (asm.assemble_scalar(F.output(k, u + h * du, t, f)) - asm.assemble_scalar(F.output(k, u - h * du, t, f))) / (2 * h)
        assert abs(gu @ du - fd) <= 1e-6 * max(abs(fd), 1e-12)


def test_oracle_thin_clamped_plate_reaches_kirchhoff_limit():
    """Clamped unit square, uniform load, t/a = 0.01: centre deflection -> 0.00126 q a^4 / D (Timoshenko); no shear
    locking thanks to the reduced shear rule."""
    n, tv, Em, nu = 24, 0.01, 1.0e6, 0.3
    F = RMPlate(om.unit_square_tri(n), E=Em, nu=nu, pen=1.0e10)
    t, f, u0 = np.full(F.M, tv), np.ones(F.M), np.zeros(F.N)
    A = asm.assemble_matrix(F.jacobian(u0, t, f), (F.N, F.N))
    x = -spla.spsolve(A.tocsc(), asm.assemble_vector(F.residual(u0, t, f), F.N))
    D = Em * tv ** 3 / (12 * (1 - nu ** 2))
    assert abs(x[(n // 2) * (n + 1) + n // 2] / (0.00126 / D) - 1.0) < 0.02


def test_engine_layout_matches_oracle():
    """Mixed-space dof numbering, dR/du pattern and both dR/dm patterns: integer arrays compared with ==."""
    F = RMPlate(om.unit_square_tri(4, 3))
    p = E.EngineProblem(E.EngineMesh.unit_square(4, 3), E.FAMILY_RM_PLATE)
    assert p.N == F.N and p.M == [F.M, F.M]
    u, t, f = _fields(F)
    rp, col = p.pattern(0)
    orp, ocol = asm.pattern(F.jacobian(u, t, f), (F.N, F.N))
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    for s in (0, 1):
        rp, col = p.pattern(1 + s)
        orp, ocol = asm.pattern(F.dRdm(s, u, t, f), (F.N, F.M))
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)


@pytest.mark.gpu
@pytest.mark.parametrize('n,ny,clamp', [(1, 1, None), (5, 3, None), (12, 9, 'left')])
def test_cuda_kernels_match_oracle(cuda_device, n, ny, clamp):
    em = E.EngineMesh.unit_square(n, ny)
    m = om.unit_square_tri(n, ny)
    tagged = None
    if clamp == 'left':
        fc, fl = m.exterior_facets()
        lf = m.local_facets[fl]
        xm = 0.5 * (m.coords[m.cells[fc, lf[:, 0]]] + m.coords[m.cells[fc, lf[:, 1]]])
        tagged = np.nonzero(xm[:, 0] < 1e-12)[0].astype(np.int32)
    prm = [3.0e5, 0.25, 2.0e7, 2.7]
    F = RMPlate(m, clamped=tagged, E=prm[0], nu=prm[1], pen=prm[2], rho=prm[3])
    p = E.EngineProblem(em, E.FAMILY_RM_PLATE, prm, tagged=tagged)
    p.upload(0)
    u, t, f = _fields(F, seed=n)
    du, dt, df = p.to_device(u), p.to_device(t), p.to_device(f)
    p.set_coefficient(0, du); p.set_coefficient(1, dt); p.set_coefficient(2, df)
    TOL = 1e-12
    assert relerr(p.assemble_residual().cpu().numpy(), asm.assemble_vector(F.residual(u, t, f), F.N)) < TOL
    vals, _ = p.assemble_jacobian()
    assert relerr(vals.cpu().numpy(), asm.assemble_matrix(F.jacobian(u, t, f), (F.N, F.N)).data) < TOL
    for s in (0, 1):
        D = asm.assemble_matrix(F.dRdm(s, u, t, f), (F.N, F.M))
        assert relerr(p.assemble_dRdm(s).cpu().numpy(), D.data) < TOL
    for k in range(3):
        Jo = asm.assemble_scalar(F.output(k, u, t, f))
        assert abs(p.assemble_output(k) - Jo) <= TOL * abs(Jo)
        assert relerr(p.assemble_output_grad(k, 0).cpu().numpy(), asm.assemble_vector(F.output_du(k, u, t, f), F.N)) < TOL
        assert relerr(p.assemble_output_grad(k, 1).cpu().numpy(), asm.assemble_vector(F.output_dm(k, 0, u, t, f), F.M)) < TOL
        assert float(p.assemble_output_grad(k, 2).abs().max()) == 0.0


@pytest.mark.gpu
def test_state_solve_and_thickness_gradient(cuda_device):
    """Linear state solve through the reference's NewtonSolver path (linear_problem = True) and the adjoint gradient of the
    elastic energy wrt the nodal thickness against the oracle's direct solve."""
    n = 10
    m = om.unit_square_tri(n)
    prm = [1.0e5, 0.3, 1.0e7, 1.0]
    F = RMPlate(m, E=prm[0], nu=prm[1], pen=prm[2], rho=prm[3])
    p = E.EngineProblem(E.EngineMesh.unit_square(n), E.FAMILY_RM_PLATE, prm)
    p.upload(0)
    t = 0.05 + 0.02 * np.random.default_rng(0).random(F.M)
    f = np.ones(F.M)
    du, dt, df = p.new_vector(p.N, 0.0), p.to_device(t), p.to_device(f)
    p.set_coefficient(0, du); p.set_coefficient(1, dt); p.set_coefficient(2, df)
    info = p.newton_solve(kind='Newton', max_it=1, krylov_rtol=1e-13, krylov_max_it=200000, precond=0)
    A = asm.assemble_matrix(F.jacobian(np.zeros(F.N), t, f), (F.N, F.N))
    uo = -spla.spsolve(A.tocsc(), asm.assemble_vector(F.residual(np.zeros(F.N), t, f), F.N))
    assert relerr(du.cpu().numpy(), uo) < 1e-7
    vals, _ = p.assemble_jacobian()
    lam, li = p.linear_solve(vals, p.assemble_output_grad(2, 0), transpose=True, rtol=1e-13, max_it=200000)
    assert li['converged']
    g = p.assemble_output_grad(2, 1).cpu().numpy() - p.spmv(1, p.assemble_dRdm(0), lam, transpose=True).cpu().numpy()
    lamo = spla.spsolve(A.T.tocsc(), asm.assemble_vector(F.output_du(2, uo, t, f), F.N))
    go = asm.assemble_vector(F.output_dm(2, 0, uo, t, f), F.M) - asm.assemble_matrix(F.dRdm(0, uo, t, f), (F.N, F.M)).T @ lamo
    assert relerr(g, go) < 1e-6


@pytest.mark.gpu
def test_shell_example_api_check_totals(cuda_device):
    """The reference's shell example structure (shell_pde.py:219-311) through FEA + FEAModel + Simulator: thickness and
    load inputs, linear state solve, compliance / mass / elastic-energy outputs, adjoint totals vs finite differences."""
    from femo_b200.fea.fea_b200 import FEA, createUnitSquareMesh, Function
    from femo_b200.forms.shell import ShellPDE
    from femo_b200.csdl_opt import FEAModel, Simulator
    mesh = createUnitSquareMesh(6)
    pde = ShellPDE(mesh)
    fea = FEA(mesh)
    fea.PDE_SOLVER = 'Newton'
    fea.REPORT = False
    fea.linear_problem = True
    h, F_solid, w = Function(pde.VT), Function(pde.VF), Function(pde.W)
    res = pde.pdeRes(h, w, F_solid, 2.0e4, 0.3, penalty=True, pen=1.0e6)
    fea.add_input('thicknesses', h)
    fea.add_input('F_solid', F_solid)
    fea.add_state(name='disp_solid', function=w, residual_form=res, arguments=['thicknesses', 'F_solid'])
    fea.add_output(name='compliance', type='scalar', form=pde.compliance(), arguments=['disp_solid', 'thicknesses'])
    fea.add_output(name='mass', type='scalar', form=pde.mass(h, 2.7), arguments=['thicknesses'])
    fea.add_output(name='elastic_energy', type='scalar', form=pde.elastic_energy(), arguments=['thicknesses', 'disp_solid'])
    model = FEAModel(fea=[fea], debug_mode=False)
    nT = fea.inputs_dict['thicknesses']['shape']
    model.create_input('thicknesses', shape=nT, val=0.08 + 0.02 * np.random.default_rng(0).random(nT))
    model.create_input('F_solid', shape=nT, val=1.0)
    sim = Simulator(model)
    sim.run()
    # oracle
    m = om.unit_square_tri(6)
    Fo = RMPlate(m, E=2.0e4, nu=0.3, pen=1.0e6, rho=2.7)
    t, f = np.asarray(sim['thicknesses']), np.asarray(sim['F_solid'])
    A = asm.assemble_matrix(Fo.jacobian(np.zeros(Fo.N), t, f), (Fo.N, Fo.N))
    uo = -spla.spsolve(A.tocsc(), asm.assemble_vector(Fo.residual(np.zeros(Fo.N), t, f), Fo.N))
    assert relerr(sim['disp_solid'], uo) < 1e-7
    for k, name in enumerate(('compliance', 'mass', 'elastic_energy')):
        Jo = asm.assemble_scalar(Fo.output(k, uo, t, f))
        assert abs(float(np.ravel(sim[name])[0]) - Jo) < 1e-6 * abs(Jo)
    rep = sim.check_totals(['compliance', 'elastic_energy', 'mass'], ['thicknesses'], step=1e-6, compact_print=False)
    assert max(rep.values()) < 1e-4, rep
