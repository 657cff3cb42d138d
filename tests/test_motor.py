"""Motor magnetostatics family (config 5b): oracle pins (no GPU), layout equality, and the CUDA path
(forward-mode dual numbers) against the oracle (complex-step derivatives)."""
import numpy as np
import pytest

from oracle import motor, assembly as asm, solvers
from _cases import relerr
from _cases_motor import MotorCase, em_params


def test_mesh_tags_and_layout_bit_exact():
    c = MotorCase(6, 24, upload=False)
    assert np.array_equal(c.emesh.coords(), c.omesh.coords)
    assert np.array_equal(c.emesh.cells(), c.omesh.cells)
    fc, fl = c.emesh.exterior_facets()
    oc, ol = c.omesh.exterior_facets()
    assert np.array_equal(fc, oc) and np.array_equal(fl, ol)
    ids = np.unique(c.tags)
    assert set(range(1, 52)) <= set(ids.tolist())              # steel, 12 magnets, 36 windings, shaft all present
    u, uh = np.zeros(c.F.N), np.zeros(c.F.M)
    for which, blocks, shape in ((0, c.F.jacobian(u, uh), (c.F.N, c.F.N)), (1, c.F.dRdm(0, u, uh), (c.F.N, c.F.M))):
        rp, col = c.p.pattern(which)
        orp, ocol = asm.pattern(blocks, shape)
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)


def test_oracle_derivatives_vs_finite_differences():
    c = MotorCase(8, 36, upload=False)
    F, u, uh = c.F, c.u, c.m
    rng = np.random.default_rng(1)
    du, dh = rng.standard_normal(F.N), rng.standard_normal(F.M)
    R = lambda a, b: asm.assemble_vector(F.residual(a, b), F.N)
    A = asm.assemble_matrix(F.jacobian(u, uh), (F.N, F.N))
    h = 1e-7
    fd = (R(u + h * du, uh) - R(u - h * du, uh)) / (2 * h)
    assert np.abs(A @ du - fd).max() < 1e-6 * np.abs(fd).max()
    D = asm.assemble_matrix(F.dRdm(0, u, uh), (F.N, F.M))
    h = 1e-9
    fd = (R(u, uh + h * dh) - R(u, uh - h * dh)) / (2 * h)
    assert np.abs(D @ dh - fd).max() < 1e-6 * np.abs(fd).max()
    assert abs(A - A.T).max() > 1e-3 * abs(A).max()             # the nonlinear Nitsche coefficient breaks symmetry


def test_oracle_load_ramp_reaches_saturation():
    """run_motor_opt.py:231-250: five source increments, SNES each; the steel reaches the cubic/exponential branch."""
    c = MotorCase(12, 48, upload=False)
    x = np.zeros(c.F.N)
    for st in range(1, 6):
        c.F.js_scale = st / 5
        x, info = c.sp.solve_snes(x, [np.zeros(c.F.M)])
        assert info['reason'] in ('ABS', 'REL', 'STOL')
    gx, _, _, _ = c.F._kin(c.F.G, x[c.F.cell_dofs], np.zeros((c.omesh.ncells, 3, 2)))
    Bn = np.sqrt((gx ** 2).sum(axis=1))
    assert Bn[(c.tags == 1) | (c.tags == 2)].max() > 0.8


@pytest.mark.gpu
@pytest.mark.parametrize('nr,nth,uscale', [(6, 24, 1e-2), (12, 48, 1e-2), (8, 36, 3e-2)])
def test_gpu_assembly_matches_oracle(cuda_device, nr, nth, uscale):
    c = MotorCase(nr, nth, seed=nr, uscale=uscale)
    F, p = c.F, c.p
    TOL = 1e-11
    assert relerr(p.assemble_residual().cpu().numpy(), asm.assemble_vector(F.residual(c.u, c.m), F.N)) < TOL
    vals, _ = p.assemble_jacobian()
    assert relerr(vals.cpu().numpy(), asm.assemble_matrix(F.jacobian(c.u, c.m), (F.N, F.N)).data) < TOL
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), asm.assemble_matrix(F.dRdm(0, c.u, c.m), (F.N, F.M)).data) < TOL
    for k in range(2):
        Jo = asm.assemble_scalar(F.output(k, c.u, c.m))
        assert abs(p.assemble_output(k) - Jo) <= TOL * abs(Jo)
        assert relerr(p.assemble_output_grad(k, 0).cpu().numpy(), asm.assemble_vector(F.output_du(k, c.u, c.m), F.N)) < TOL
        assert relerr(p.assemble_output_grad(k, 1).cpu().numpy(), asm.assemble_vector(F.output_dm(k, 0, c.u, c.m), F.M)) < TOL


@pytest.mark.gpu
def test_gpu_gmres_and_state_solve(cuda_device):
    """Incremental nonlinear B-H solve (five source steps, SNES + GMRES) and the adjoint shape gradient
    d(int |B|^2 J dx)/d(uhat) against the oracle's direct solves."""
    import scipy.sparse.linalg as spla
    c = MotorCase(12, 48, seed=2)
    p, F = c.p, c.F
    uh = c.m
    # GMRES on the non-symmetric Jacobian (plain and transposed)
    vals, _ = p.assemble_jacobian()
    A = c.csr(0, vals)
    b = np.random.default_rng(0).standard_normal(F.N)
    for tr in (False, True):
        x, info = p.linear_solve(vals, p.to_device(b), transpose=tr, rtol=1e-12, method=1, max_it=20000, precond=1,
                                 cheb_degree=24, cheb_ratio=600.0)
        assert info['converged'], info
        assert relerr(x.cpu().numpy(), spla.spsolve((A.T if tr else A).tocsc(), b)) < 1e-7
    # load ramp
    c.d_u.zero_()
    x = np.zeros(F.N)
    for st in range(1, 6):
        p.set_param(6, st / 5)
        F.js_scale = st / 5
        info = p.newton_solve(kind='SNES', krylov_rtol=1e-12, krylov_max_it=20000, method=1, precond=1, cheb_degree=24,
                              cheb_ratio=600.0)
        x, oinfo = c.sp.solve_snes(x, [uh])
        assert info['converged'] in (1, 2, 3)
    u = c.d_u.cpu().numpy()
    assert relerr(u, x) < 1e-7
    # adjoint: dJ/duhat = pJ/puhat - (dR/duhat)^T A^-T pJ/pu
    vals, _ = p.assemble_jacobian()
    lam, li = p.linear_solve(vals, p.assemble_output_grad(0, 0), transpose=True, rtol=1e-12, method=1, max_it=20000,
                             precond=1, cheb_degree=24, cheb_ratio=600.0)
    assert li['converged']
    g = p.assemble_output_grad(0, 1).cpu().numpy() - p.spmv(1, p.assemble_dRdm(0), lam, transpose=True).cpu().numpy()
    (go,), lamo = c.sp.total_derivative(0, x, [uh])
    assert relerr(lam.cpu().numpy(), lamo) < 1e-6
    assert relerr(g, go) < 1e-6


@pytest.mark.gpu
def test_gpu_motor_api_check_totals(cuda_device):
    """The EM half of examples/em_motor_opt/run_motor_opt.py through FEA + FEAModel + Simulator:
    custom incremental solve, B-influence outputs, adjoint totals w.r.t. the mesh displacement vs FD."""
    from femo_b200.fea.fea_b200 import FEA, Mesh, FunctionSpace, VectorFunctionSpace, Function, TestFunction, meshtags, Measure
    from femo_b200.fea.utils_b200 import solveNonlinear
    from femo_b200.forms import motor as pde
    from femo_b200.csdl_opt import FEAModel, Simulator
    from femo_b200 import engine as E
    nr, nth = 8, 36
    mesh = Mesh(E.EngineMesh.annulus(nr, nth), 'triangle')
    om_ = motor.annulus_tri(nr, nth)
    tags = motor.motor_tags(om_)
    dx = Measure('dx', domain=mesh, subdomain_data=meshtags(mesh, 2, np.arange(mesh.num_cells), tags))
    Hc, p, s, vacuum_perm, angle, iq = 838.e3, 12, 36, 4e-7 * np.pi, 0., 282.2 / 0.00016231
    fea_em = FEA(mesh)
    fea_em.PDE_SOLVER = 'SNES'
    fea_em.REPORT = False
    uhat = Function(VectorFunctionSpace(mesh, ('CG', 1)))
    V = FunctionSpace(mesh, ('CG', 1))
    A_z = Function(V)
    res = pde.pdeResEM(A_z, TestFunction(V), uhat, iq, dx, p, s, Hc, vacuum_perm, angle, g=Function(V), nitsche=True,
                       sym=True)
    js = pde.JS(res)

    def solveIncrementalEM(res_, func, bc, report=False):
        func.vector.set(0.0)
        for i in range(5):
            js.set_scale((i + 1) / 5)
            solveNonlinear(res_, func, bc, 'SNES', False, False)
    fea_em.custom_solve = solveIncrementalEM
    fea_em.add_input('uhat', uhat, init_val=0.0)
    fea_em.add_state(name='A_z', function=A_z, residual_form=res, arguments=['uhat'])
    fea_em.add_output(name='B_influence_eddy_current', type='scalar',
                      form=pde.B_power_form(A_z, uhat, 2, dx, [1, 2]), arguments=['A_z', 'uhat'])
    fea_em.add_output(name='B_influence_hysteresis', type='scalar',
                      form=pde.B_power_form(A_z, uhat, 1.76835, dx, [1, 2]), arguments=['A_z', 'uhat'])
    model = FEAModel(fea=[fea_em], debug_mode=False)
    rng = np.random.default_rng(0)
    model.create_input('uhat', shape=fea_em.inputs_dict['uhat']['shape'], val=1e-4 * rng.standard_normal(2 * om_.nverts))
    sim = Simulator(model)
    sim.run()
    # oracle forward values
    F = motor.MotorEM(om_, tags)
    sp_ = solvers.StatePath(F, None)
    x = np.zeros(F.N)
    for st in range(1, 6):
        F.js_scale = st / 5
        x, _ = sp_.solve_snes(x, [sim['uhat']])
    assert relerr(sim['A_z'], x) < 1e-7
    for k, name in enumerate(('B_influence_eddy_current', 'B_influence_hysteresis')):
        assert abs(sim[name][0] - asm.assemble_scalar(F.output(k, x, sim['uhat']))) < 1e-7 * abs(sim[name][0])
    rep = sim.check_totals('B_influence_eddy_current', 'uhat', step=1e-7, directions=2, compact_print=False)
    assert max(rep.values()) < 1e-4, rep


# ---------------------------------------------------------------------------------------------
# config 5a: hyperelastic mesh motion
# ---------------------------------------------------------------------------------------------
def test_mm_layout_and_oracle_derivatives():
    from _cases_motor import MotorMMCase
    c = MotorMMCase(6, 18, upload=False)
    F, uh, g = c.F, c.u, c.m
    for which, blocks, shape in ((0, F.jacobian(uh, g), (F.N, F.N)), (1, F.dRdm(0, uh, g), (F.N, F.M))):
        rp, col = c.p.pattern(which)
        orp, ocol = asm.pattern(blocks, shape)
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    rng = np.random.default_rng(1)
    du = rng.standard_normal(F.N)
    R = lambda a, b: asm.assemble_vector(F.residual(a, b), F.N)
    h = 1e-9
    A = asm.assemble_matrix(F.jacobian(uh, g), (F.N, F.N))
    fd = (R(uh + h * du, g) - R(uh - h * du, g)) / (2 * h)
    assert np.abs(A @ du - fd).max() < 1e-7 * np.abs(fd).max()
    D = asm.assemble_matrix(F.dRdm(0, uh, g), (F.N, F.M))
    fd = (R(uh, g + h * du) - R(uh, g - h * du)) / (2 * h)
    assert np.abs(D @ du - fd).max() < 1e-7 * np.abs(fd).max()


@pytest.mark.gpu
@pytest.mark.parametrize('nr,nth', [(4, 12), (8, 24)])
def test_gpu_mm_assembly_matches_oracle(cuda_device, nr, nth):
    from _cases_motor import MotorMMCase
    c = MotorMMCase(nr, nth, seed=nr)
    F, p = c.F, c.p
    TOL = 1e-11
    assert relerr(p.assemble_residual().cpu().numpy(), asm.assemble_vector(F.residual(c.u, c.m), F.N)) < TOL
    vals, _ = p.assemble_jacobian()
    assert relerr(vals.cpu().numpy(), asm.assemble_matrix(F.jacobian(c.u, c.m), (F.N, F.N)).data) < TOL
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), asm.assemble_matrix(F.dRdm(0, c.u, c.m), (F.N, F.M)).data) < TOL
    for k in range(3):
        Jo = asm.assemble_scalar(F.output(k, c.u, c.m))
        assert abs(p.assemble_output(k) - Jo) <= TOL * abs(Jo)
        assert relerr(p.assemble_output_grad(k, 0).cpu().numpy(), asm.assemble_vector(F.output_du(k, c.u, c.m), F.N)) < TOL
        assert np.all(p.assemble_output_grad(k, 1).cpu().numpy() == 0.0)


@pytest.mark.gpu
def test_gpu_mm_state_and_adjoint(cuda_device):
    """SNES + GMRES mesh-motion solve for a prescribed radial edge displacement (weak Nitsche BC), and the
    adjoint derivative of the steel area w.r.t. the prescribed displacement, against the oracle."""
    from _cases_motor import MotorMMCase
    c = MotorMMCase(8, 24, seed=3)
    p, F = c.p, c.F
    g = c.radial_bc(0.02)
    c.d_m.copy_(p.to_device(g))
    c.d_u.zero_()
    kw = dict(method=1, precond=1, cheb_degree=24, cheb_ratio=600.0)
    info = p.newton_solve(kind='SNES', krylov_rtol=1e-12, krylov_max_it=40000, **kw)
    xo, oinfo = c.sp.solve_snes(np.zeros(F.N), [g])
    x = c.d_u.cpu().numpy()
    assert info['converged'] in (1, 2, 3)
    assert relerr(x, xo) < 1e-7
    nodes = c.mid * c.nth + np.arange(c.nth)
    assert np.abs(x[2 * nodes] - g[2 * nodes]).max() < 1e-3 * np.abs(g).max()      # weakly enforced
    vals, _ = p.assemble_jacobian()
    lam, li = p.linear_solve(vals, p.assemble_output_grad(2, 0), transpose=True, rtol=1e-12, max_it=40000, **kw)
    assert li['converged']
    gr = p.assemble_output_grad(2, 1).cpu().numpy() - p.spmv(1, p.assemble_dRdm(0), lam, transpose=True).cpu().numpy()
    (go,), lamo = c.sp.total_derivative(2, xo, [g])
    assert relerr(lam.cpu().numpy(), lamo) < 1e-6
    assert relerr(gr, go) < 1e-6


@pytest.mark.gpu
def test_gpu_motor_coupled_chain_totals(cuda_device):
    """examples/em_motor_opt/run_motor_opt.py end to end on the synthetic annulus: edge displacement uhat_bc ->
    mesh motion uhat (incremental SNES) -> magnetostatics A_z (incremental SNES) -> B influence; the total
    derivative w.r.t. uhat_bc chains the EM adjoint and the mesh-motion adjoint and is checked by FD."""
    from femo_b200.fea.fea_b200 import FEA, Mesh, FunctionSpace, VectorFunctionSpace, Function, TestFunction, meshtags, Measure
    from femo_b200.fea.utils_b200 import solveNonlinear, getFuncArray
    from femo_b200.forms import motor as pde
    from femo_b200.csdl_opt import FEAModel, Simulator
    from femo_b200 import engine as E
    nr, nth = 8, 24
    mesh = Mesh(E.EngineMesh.annulus(nr, nth), 'triangle')
    om_ = motor.annulus_tri(nr, nth)
    tags = motor.motor_tags(om_)
    dx = Measure('dx', domain=mesh, subdomain_data=meshtags(mesh, 2, np.arange(mesh.num_cells), tags))
    mid = nr // 2
    dS = pde.SideMeasure(*pde.annulus_circle_sides(mesh, mid))
    c0, c1 = pde.annulus_circle_sides(mesh, 0), pde.annulus_circle_sides(mesh, nr)
    ds = pde.SideMeasure(np.concatenate([c0[0], c1[0]]), np.concatenate([c0[1], c1[1]]))
    Hc, p, s, vacuum_perm, angle, iq = 838.e3, 12, 36, 4e-7 * np.pi, 0., 282.2 / 0.00016231
    VV = VectorFunctionSpace(mesh, ('CG', 1))
    # ---- mesh motion subproblem (run_motor_opt.py:94-208)
    fea_mm = FEA(mesh)
    fea_mm.PDE_SOLVER, fea_mm.REPORT = 'SNES', False
    uhat_bc, uhat = Function(VV), Function(VV)
    res_mm = pde.pdeResMM(uhat, TestFunction(VV), g=uhat_bc, nitsche=True, sym=True, dS_=dS(1000), ds_=ds(1000),
                          cell_tags=tags)

    def solveIncremental(res, func, bc, report=False):
        vec = np.copy(getFuncArray(uhat_bc))
        STEPS = 2
        func.vector.set(0.0)
        for i in range(STEPS):
            uhat_bc.vector.setArray(vec * (i + 1) / STEPS)
            solveNonlinear(res, func, bc, 'SNES', False, False)
        uhat_bc.vector.setArray(vec)
    fea_mm.custom_solve = solveIncremental
    fea_mm.add_input('uhat_bc', uhat_bc, init_val=0.0)
    fea_mm.add_state(name='uhat', function=uhat, residual_form=res_mm, arguments=['uhat_bc'])
    fea_mm.add_output(name='steel_area', type='scalar', form=pde.area_form(uhat, dx, [1, 2]), arguments=['uhat'])
    # ---- electromagnetic subproblem (:212-317)
    fea_em = FEA(mesh)
    fea_em.PDE_SOLVER, fea_em.REPORT = 'SNES', False
    V = FunctionSpace(mesh, ('CG', 1))
    A_z = Function(V)
    res_em = pde.pdeResEM(A_z, TestFunction(V), uhat, iq, dx, p, s, Hc, vacuum_perm, angle, g=Function(V), nitsche=True, sym=True)
    js = pde.JS(res_em)

    def solveIncrementalEM(res, func, bc, report=False):
        func.vector.set(0.0)
        for i in range(5):
            js.set_scale((i + 1) / 5)
            solveNonlinear(res, func, bc, 'SNES', False, False)
    fea_em.custom_solve = solveIncrementalEM
    fea_em.add_input('uhat', uhat, init_val=0.0)
    fea_em.add_state(name='A_z', function=A_z, residual_form=res_em, arguments=['uhat'])
    fea_em.add_output(name='B_influence_eddy_current', type='scalar', form=pde.B_power_form(A_z, uhat, 2, dx, [1, 2]),
                      arguments=['A_z', 'uhat'])
    model = FEAModel(fea=[fea_mm, fea_em], debug_mode=False)
    g0 = np.zeros(2 * om_.nverts)
    nodes = mid * nth + np.arange(nth)
    xy = om_.coords[nodes]
    g0[2 * nodes], g0[2 * nodes + 1] = 0.01 * xy[:, 0], 0.01 * xy[:, 1]
    model.create_input('uhat_bc', shape=g0.size, val=g0)
    sim = Simulator(model)
    sim.run()
    # forward values against the oracle chain
    from oracle import motor_mm
    parts = [motor_mm.circle_facets(om_, k) for k in (0, mid, nr)]
    fc = np.concatenate([q[0] for q in parts]); fl = np.concatenate([q[1] for q in parts]); o = np.lexsort((fl, fc))
    Fm = motor_mm.MotorMM(om_, (fc[o], fl[o]), tags)
    spm = solvers.StatePath(Fm, None)
    xm = np.zeros(Fm.N)
    for st in (0.5, 1.0):
        xm, _ = spm.solve_snes(xm, [st * g0])
    assert relerr(sim['uhat'], xm) < 1e-7
    Fe = motor.MotorEM(om_, tags)
    spe = solvers.StatePath(Fe, None)
    xe = np.zeros(Fe.N)
    for st in range(1, 6):
        Fe.js_scale = st / 5
        xe, _ = spe.solve_snes(xe, [xm])
    assert relerr(sim['A_z'], xe) < 1e-6
    Bo = asm.assemble_scalar(Fe.output(0, xe, xm))
    assert abs(sim['B_influence_eddy_current'][0] - Bo) < 1e-6 * abs(Bo)
    # chained adjoint totals vs finite differences
    rep = sim.check_totals('B_influence_eddy_current', 'uhat_bc', step=1e-7, directions=2, compact_print=False)
    assert max(rep.values()) < 1e-3, rep
    rep = sim.check_totals('steel_area', 'uhat_bc', step=1e-7, directions=2, compact_print=False)
    assert max(rep.values()) < 1e-4, rep
