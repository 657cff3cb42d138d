"""The reference-facing API (FEA + CSDL operations) on the GPU vs the oracle:
the tests read like the reference's example scripts."""
import numpy as np
import pytest

from oracle import mesh as om, families as fam, assembly as asm, solvers
from _cases import relerr, square_boundary_lists

pytestmark = pytest.mark.gpu


def _poisson_setup(n):
    """examples/poisson_opt/run_poisson_opt.py:22-153 against the femo_b200 API."""
    from femo_b200.fea.fea_b200 import (FEA, createUnitSquareMesh, FunctionSpace, Function, TestFunction,
                                         locate_dofs_geometrical, getFuncArray)
    from femo_b200.forms.poisson import pdeRes, outputForm
    PI = np.pi

    class Expression_f:
        def eval(self, x):
            return 1 / (1 + 1e-6 * 4 * np.power(PI, 4)) * np.sin(PI * x[0]) * np.sin(PI * x[1])

    class Expression_u:
        def eval(self, x):
            return 1 / (2 * np.power(PI, 2)) * np.sin(PI * x[0]) * np.sin(PI * x[1])

    mesh = createUnitSquareMesh(n)
    fea = FEA(mesh)
    Vf = FunctionSpace(mesh, ('DG', 0))
    f = Function(Vf)
    Vu = FunctionSpace(mesh, ('CG', 1))
    u = Function(Vu)
    v = TestFunction(Vu)
    u_ex = fea.add_exact_solution(Expression_u, Vu)
    f_ex = fea.add_exact_solution(Expression_f, Vf)
    output_form = outputForm(u, f, u_ex)
    ubc = Function(Vu)
    ubc.vector.set(0.0)
    locs = [locate_dofs_geometrical((Vu, Vu), lambda x, a=a, b=b: np.isclose(x[a], b, atol=1e-6))
            for a, b in ((0, 0.), (0, 1.), (1, 0.), (1, 1.))]
    fea.add_strong_bc(ubc, locs, Vu)
    residual_form = pdeRes(u, v, f)
    fea.add_input('f', f)
    fea.add_state(name='u', function=u, residual_form=residual_form, arguments=['f'])
    fea.add_output(name='l2_functional', type='scalar', form=output_form, arguments=['f', 'u'])
    fea.PDE_SOLVER = 'Newton'
    fea.REPORT = False
    return fea, getFuncArray(f_ex).copy(), getFuncArray(u_ex).copy()


def test_poisson_opt_forward_and_totals(cuda_device):
    from femo_b200.csdl_opt import FEAModel, Simulator
    n = 16
    fea, f_ex, u_ex = _poisson_setup(n)
    model = FEAModel(fea=[fea], debug_mode=False)
    model.create_input('f', shape=fea.inputs_dict['f']['shape'], val=0.1 * np.ones(fea.inputs_dict['f']['shape']) * 0.86)
    model.add_design_variable('f')
    model.add_objective('l2_functional', scaler=1e5)
    sim = Simulator(model)
    sim['f'] = f_ex
    sim.run()
    # oracle
    m = om.unit_square_tri(n)
    F = fam.PoissonP1(m, u_ex=u_ex)
    bc = asm.DirichletBC(F.N, square_boundary_lists(m.coords), 0.0)
    sp = solvers.StatePath(F, bc)
    uo, _ = sp.solve_newton(np.zeros(F.N), [f_ex])
    assert relerr(sim['u'], uo) < 1e-9
    Jo = asm.assemble_scalar(F.output(0, uo, f_ex))
    assert abs(sim['l2_functional_output_model.l2_functional'][0] - Jo) < 1e-9 * abs(Jo)
    tot = sim.compute_totals('l2_functional', 'f')[('l2_functional', 'f')]
    (go,), _ = sp.total_derivative(0, uo, [f_ex])          # reference-faithful (quirk B2)
    assert relerr(tot, go) < 1e-8
    # the totals differ from finite differences only through quirk B2
    (gc,), _ = sp.total_derivative(0, uo, [f_ex], consistent_bc=True)
    assert relerr(gc, go) > 1e-6


@pytest.mark.parametrize('degree', [1, 2])
def test_nonlinear_poisson_check_totals(cuda_device, degree):
    """examples/nonlinear_poisson_opt: SNES state solve, then adjoint totals vs
    central finite differences through the full callback chain (no strong BC, so
    reference-faithful and consistent totals coincide).  degree 2 = the P2 variant of the same script."""
    from femo_b200.fea.fea_b200 import FEA, createUnitSquareMesh, FunctionSpace, Function, TestFunction
    from femo_b200.forms.nonlinear_poisson import pdeRes, outputForm
    from femo_b200.csdl_opt import FEAModel, Simulator
    mesh = createUnitSquareMesh(12)
    fea = FEA(mesh)
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    Vu = FunctionSpace(mesh, ('CG', degree))
    u = Function(Vu)
    residual_form = pdeRes(u, TestFunction(Vu), f)
    fea.add_input('f', f)
    fea.add_state(name='u', function=u, residual_form=residual_form, arguments=['f'])
    fea.add_output(name='l2_functional', type='scalar', form=outputForm(u, f), arguments=['f', 'u'])
    fea.PDE_SOLVER = 'SNES'
    fea.REPORT = False
    model = FEAModel(fea=[fea], debug_mode=False)
    model.create_input('f', shape=fea.inputs_dict['f']['shape'], val=0.1)
    sim = Simulator(model)
    rep = sim.check_totals('l2_functional', 'f', step=1e-4, compact_print=False)
    assert max(rep.values()) < 1e-6
    # and against the oracle's reference-ordered chain
    m = om.unit_square_tri(12)
    F = fam.NonlinearPoissonP1(m) if degree == 1 else fam.NonlinearPoissonP2(m)
    sp = solvers.StatePath(F, None)
    f0 = 0.1 * np.ones(F.M)
    uo, _ = sp.solve_snes(np.zeros(F.N), [f0])
    assert relerr(sim['u'], uo) < 1e-9
    (go,), _ = sp.total_derivative(0, uo, [f0])
    assert relerr(sim.compute_totals('l2_functional', 'f')[('l2_functional', 'f')], go) < 1e-8


def test_api_error_behaviour(cuda_device):
    """Error conventions of the reference."""
    from femo_b200.fea.fea_b200 import FEA, createUnitSquareMesh, FunctionSpace, Function, assemble
    from femo_b200.forms.poisson import pdeRes
    mesh = createUnitSquareMesh(4)
    fea = FEA(mesh)
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    u = Function(FunctionSpace(mesh, ('CG', 1)))
    fea.add_input('f', f)
    with pytest.raises(ValueError):
        fea.add_input('f', f)                              # fea_dolfinx.py:101-102
    assert np.all(f.x.array == 1.0)                        # quirk B5: init_val overwrite
    r = assemble(pdeRes(u, None, f), dim=7)
    assert isinstance(r, TypeError)                        # returned, not raised (quirk B7)


def test_project_matches_oracle(cuda_device):
    """`project` (utils_dolfinx.py:549-583) as used at run_nonlinear_poisson_opt.py:165-169 and
    run_topo_opt_cantilever_beam.py:264-268, against the oracle's exact mass solve."""
    from femo_b200.fea.fea_b200 import createUnitSquareMesh, FunctionSpace, Function, project, getFuncArray
    from femo_b200.forms.nonlinear_poisson import u_ex_ufl, f_ex_ufl
    from oracle.projection import MassProjection
    n = 12
    mesh = createUnitSquareMesh(n)
    m = om.unit_square_tri(n)
    u_ex = Function(FunctionSpace(mesh, ('CG', 1)))
    project(u_ex_ufl, u_ex)
    assert relerr(getFuncArray(u_ex), MassProjection(m, 'CG', 'u_ex').project()) < 1e-10
    f_ex = Function(FunctionSpace(mesh, ('DG', 0)))
    project(f_ex_ufl, f_ex)
    assert relerr(getFuncArray(f_ex), MassProjection(m, 'DG', 'f_ex').project()) < 1e-10
    rho = Function(FunctionSpace(mesh, ('DG', 0)))
    r = np.random.default_rng(0).random(rho.function_space.dim)
    rho.vector.setArray(r)
    pen = Function(FunctionSpace(mesh, ('DG', 0)))
    project(rho ** 3, pen)
    assert relerr(getFuncArray(pen), r ** 3) < 1e-12
    g = Function(FunctionSpace(mesh, ('CG', 1)))
    project(rho, g)
    assert relerr(getFuncArray(g), MassProjection(m, 'CG', 'dg_pow', 1.0).project(r)) < 1e-10
    gl = Function(FunctionSpace(mesh, ('CG', 1)))
    project(u_ex_ufl, gl, lump_mass=True)
    assert relerr(getFuncArray(gl), getFuncArray(u_ex)) < 0.2


def test_topology_example_with_density_filter(cuda_device):
    """examples/beam_topo_opt with its GeneralFilterModel pre-processor: filtered density vs the oracle's
    KD-tree weights, forward solve vs oracle, adjoint totals through filter + state vs finite differences."""
    from femo_b200.fea.fea_b200 import (FEA, createRectangleMesh, FunctionSpace, VectorFunctionSpace, Function,
                                         TestFunction, Constant, locate_dofs_geometrical, locate_entities_boundary,
                                         meshtags, Measure, meshSize, DOLFIN_EPS)
    from femo_b200.forms.topo import pdeRes, averageFunc, compliance
    from femo_b200.csdl_opt import FEAModel, Simulator
    from femo_b200.csdl_opt.pre_processor.general_filter_model import GeneralFilterModel
    from oracle.filter import weight_matrix
    nx, ny, LX, LY = 24, 12, 160., 80.
    mesh = createRectangleMesh(np.array([0.0, 0.0]), np.array([LX, LY]), nx, ny)
    tb = locate_entities_boundary(mesh, 1, lambda x: np.logical_and(abs(x[1] - LY / 2) < LY / ny + DOLFIN_EPS * 1e10,
                                                                     abs(x[0] - LX) < DOLFIN_EPS * 1e10))
    ds_ = Measure('ds', domain=mesh, subdomain_data=meshtags(mesh, 1, tb, np.full(len(tb), 100, dtype=np.int32)),
                  metadata={"quadrature_degree": 4})
    fea = FEA(mesh)
    Vr = FunctionSpace(mesh, ('DG', 0))
    rho = Function(Vr)
    Vu = VectorFunctionSpace(mesh, ('CG', 1))
    u = Function(Vu)
    f = Constant(mesh, (0, -1 / 4))
    res = pdeRes(u, TestFunction(Vu), rho, f, dss=ds_(100), method='SIMP')
    fea.add_input('density', rho)
    fea.add_state(name='displacements', function=u, residual_form=res, arguments=['density'])
    fea.add_output(name='avg_density', type='scalar', form=averageFunc(rho), arguments=['density'])
    fea.add_output(name='compliance', type='scalar', form=compliance(u, f, dss=ds_(100)), arguments=['displacements'])
    fea.add_strong_bc(Function(Vu), [locate_dofs_geometrical((Vu, Vu), lambda x: np.isclose(x[0], 0., atol=1e-6))], Vu)
    fea.REPORT = False
    model = FEAModel(fea=[fea], debug_mode=False)
    coords = Vr.tabulate_dof_coordinates()
    h = meshSize(mesh)
    h_avg = (h.max() + h.min()) / 2
    nel = mesh.num_cells
    model.add(GeneralFilterModel(nel=nel, coordinates=coords, h_avg=h_avg), name='general_filter_model')
    np.random.seed(0)
    x0 = 0.3 + 0.6 * np.random.random(nel)
    model.create_input('density_unfiltered', shape=nel, val=x0)
    sim = Simulator(model)
    sim.run()
    W = weight_matrix(coords[:, :2], h_avg)
    assert relerr(sim['density'], W @ x0) < 1e-13
    assert abs(sim['avg_density'][0] - (W @ x0).mean()) < 1e-12
    rep = sim.check_totals('compliance', 'density_unfiltered', step=1e-5, compact_print=False)
    assert max(rep.values()) < 1e-5, rep
    rep = sim.check_totals('avg_density', 'density_unfiltered', step=1e-3, compact_print=False)
    assert max(rep.values()) < 1e-8, rep


def test_topology_example_3d_hex(cuda_device):
    """The same script on a hexahedral box (SURVEY.md section 8d, C4-3D scaled down): forward solve vs the oracle
    (GMG-PCG vs SuperLU), compliance / average-density totals w.r.t. the density vs finite differences."""
    from femo_b200.fea.fea_b200 import (FEA, createBoxMesh, FunctionSpace, VectorFunctionSpace, Function, TestFunction,
                                         Constant, locate_dofs_geometrical, locate_entities_boundary, meshtags, Measure,
                                         DOLFIN_EPS)
    from femo_b200.forms.topo import pdeRes, averageFunc, compliance
    from femo_b200.csdl_opt import FEAModel, Simulator
    from oracle import mesh as om, families as fam, assembly as asm, solvers
    nx, ny, nz, L = 12, 6, 4, (24., 12., 8.)
    mesh = createBoxMesh(np.zeros(3), np.array(L), nx, ny, nz)
    tb = locate_entities_boundary(mesh, 2, lambda x: np.logical_and(abs(x[1] - L[1] / 2) < L[1] / ny + DOLFIN_EPS * 1e10,
                                                                     abs(x[0] - L[0]) < DOLFIN_EPS * 1e10))
    assert len(tb) == 2 * nz
    ds_ = Measure('ds', domain=mesh, subdomain_data=meshtags(mesh, 2, tb, np.full(len(tb), 100, dtype=np.int32)))
    fea = FEA(mesh)
    Vr = FunctionSpace(mesh, ('DG', 0))
    rho = Function(Vr)
    Vu = VectorFunctionSpace(mesh, ('CG', 1))
    assert Vu.block == 3
    u = Function(Vu)
    f = Constant(mesh, (0, -1 / 4, 0))
    res = pdeRes(u, TestFunction(Vu), rho, f, dss=ds_(100), method='SIMP')
    fea.add_input('density', rho)
    fea.add_state(name='displacements', function=u, residual_form=res, arguments=['density'])
    fea.add_output(name='avg_density', type='scalar', form=averageFunc(rho), arguments=['density'])
    fea.add_output(name='compliance', type='scalar', form=compliance(u, f, dss=ds_(100)), arguments=['displacements'])
    fea.add_strong_bc(Function(Vu), [locate_dofs_geometrical((Vu, Vu), lambda x: np.isclose(x[0], 0., atol=1e-6))], Vu)
    fea.REPORT = False
    model = FEAModel(fea=[fea], debug_mode=False)
    np.random.seed(0)
    x0 = 0.3 + 0.6 * np.random.random(mesh.num_cells)
    model.create_input('density', shape=mesh.num_cells, val=x0)
    sim = Simulator(model)
    sim.run()
    assert res.fam.problem.mg_levels >= 2 and res.fam.precond == 2
    m = om.box_hex((0, 0, 0), L, nx, ny, nz)
    fc, fl = m.exterior_facets()
    F = fam.SimpHex8(m, tb)
    nodes = np.nonzero(np.isclose(m.coords[:, 0], 0.0, atol=1e-6))[0]
    bc = asm.DirichletBC(F.N, [np.stack([3 * nodes, 3 * nodes + 1, 3 * nodes + 2], axis=1).ravel()], 0.0)
    uo, _ = solvers.StatePath(F, bc).solve_newton(np.zeros(F.N), [x0])
    assert relerr(sim['displacements'], uo) < 1e-7
    Co = asm.assemble_scalar(F.output(1, uo, x0))
    assert abs(sim['compliance'][0] - Co) < 1e-7 * abs(Co)
    rep = sim.check_totals('compliance', 'density', step=1e-5, compact_print=False)
    assert max(rep.values()) < 1e-5, rep
    rep = sim.check_totals('avg_density', 'density', step=1e-3, compact_print=False)
    assert max(rep.values()) < 1e-8, rep


def test_density_filter_3d_matches_kdtree_oracle(cuda_device):
    """femo_filter_apply3 (the 3-D pre-processor of the hexahedral cantilever) against the oracle's KD-tree weights:
    W x and the reverse-mode action W^T y."""
    from femo_b200.csdl_opt.pre_processor.general_filter_model import GeneralFilterOperation
    from oracle.filter import weight_matrix
    nx, ny, nz, h = 9, 6, 5, (2.0, 2.0, 1.5)
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
    coords = np.stack([(I.ravel() + 0.5) * h[0], (J.ravel() + 0.5) * h[1], (K.ravel() + 0.5) * h[2]], axis=1)
    op = GeneralFilterOperation(nel=coords.shape[0], beta=2.0, coordinates=coords, h_avg=1.9)
    op.define()
    W = weight_matrix(coords, 1.9, 2.0)
    rng = np.random.default_rng(0)
    x, y = rng.random(coords.shape[0]), rng.standard_normal(coords.shape[0])
    out = {}
    op.compute({'density_unfiltered': x}, out)
    assert np.max(np.abs(out['density'] - W @ x)) < 1e-13
    assert np.max(np.abs(op.vjp('density', 'density_unfiltered', y) - W.T @ y)) < 1e-13
