"""Host-side logic of the API mirror that needs no GPU."""
import numpy as np
import pytest

from femo_b200.fea.fea_b200 import (FEA, createUnitSquareMesh, createRectangleMesh, createIntervalMesh, FunctionSpace,
                                     VectorFunctionSpace, Function, update, getFuncArray, locate_dofs_geometrical,
                                     locate_entities_boundary, locate_dofs_topological)


def test_function_update_semantics():
    mesh = createUnitSquareMesh(3)
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    assert len(getFuncArray(f)) == 18
    update(f, np.array([2.5]))                              # quirk B6: length-1 broadcasts
    assert np.all(getFuncArray(f) == 2.5)
    update(f, np.arange(18.0))
    assert np.array_equal(getFuncArray(f), np.arange(18.0))
    with pytest.raises(ValueError):
        update(f, np.arange(5.0))


def test_registry_and_shapes():
    mesh = createUnitSquareMesh(16)
    fea = FEA(mesh)
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    fea.add_input('f', f, init_val=0.3)
    assert fea.inputs_dict['f']['shape'] == 512 and np.all(getFuncArray(f) == 0.3)
    with pytest.raises(ValueError):
        fea.add_input('f', f)
    assert fea.PDE_SOLVER == 'Newton' and fea.linear_problem is False and fea.opt_iter == 0


def test_boundary_dof_location():
    mesh = createUnitSquareMesh(4)
    V = FunctionSpace(mesh, ('CG', 1))
    d = locate_dofs_geometrical((V, V), lambda x: np.isclose(x[0], 0.0, atol=1e-6))
    assert isinstance(d, list) and np.array_equal(d[0], np.arange(0, 25, 5))
    qm = createRectangleMesh(np.array([0.0, 0.0]), np.array([160.0, 80.0]), 8, 4)
    W = VectorFunctionSpace(qm, ('CG', 1))
    assert W.dim == 2 * 9 * 5
    dv = locate_dofs_geometrical((W, W), lambda x: np.isclose(x[0], 0.0, atol=1e-6))[0]
    assert np.array_equal(dv, np.stack([2 * np.arange(0, 45, 9), 2 * np.arange(0, 45, 9) + 1], 1).ravel())
    im = createIntervalMesh(50, 0.0, 1.0)
    start = locate_entities_boundary(im, 0, lambda x: np.isclose(x[0], 0))
    H = FunctionSpace(im, ('Hermite', 3))
    assert np.array_equal(locate_dofs_topological(H, 0, start), [0, 1])
    fac = locate_entities_boundary(im, 0, lambda x: np.isclose(x[0], 1.0))
    assert np.array_equal(fac, [50])


def test_compute_needs_gpu_and_fails_loudly():
    from femo_b200 import engine
    from femo_b200._lib import FemoError
    from femo_b200.fea.fea_b200 import assembleVector
    from femo_b200.forms.poisson import pdeRes
    if engine.device_count() > 0:
        pytest.skip('CUDA device present')
    mesh = createUnitSquareMesh(2)
    u = Function(FunctionSpace(mesh, ('CG', 1)))
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    with pytest.raises(FemoError):
        assembleVector(pdeRes(u, None, f))


def test_tracked_sources_skip_only_unchanged_uploads():
    """update() may skip a source it already holds only when the source is version-tracked storage at the same
    version; plain arrays, bumped versions and any other write to the Function invalidate the shortcut."""
    from femo_b200 import _hostops as H
    mesh = createUnitSquareMesh(3)
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    src = H.tracked(np.arange(18.0))
    update(f, src)
    v0 = f._host_ver
    assert np.array_equal(getFuncArray(f), np.arange(18.0)) and f._src is not None
    update(f, src)                                          # same storage, same version: nothing to do
    assert f._host_ver == v0
    np.asarray(src)[:] = 7.0                                # the owner writes ...
    update(f, src)                                          # ... but forgot to bump: by contract still skipped
    assert f._host_ver == v0
    src.bump()
    update(f, src)
    assert f._host_ver > v0 and np.all(getFuncArray(f) == 7.0)
    plain = np.full(18, 3.0)
    update(f, plain)
    v1 = f._host_ver
    update(f, plain)                                        # untracked arrays are always copied
    assert f._host_ver > v1 and f._src is None
    update(f, src)                                          # tracked again after something else wrote the Function
    assert np.all(getFuncArray(f) == 7.0)
    f.vector.set(1.5)                                       # constant fill invalidates the shortcut as well
    update(f, src)
    assert np.all(getFuncArray(f) == 7.0)
    assert H.tracked(np.zeros(4))[1:].version is None       # views are never tracked


def test_simulator_storage_versions(monkeypatch):
    """The Simulator stand-in bumps a variable's version on every write it performs and whenever it hands out a
    writable reference, so operations never see stale device copies."""
    from femo_b200.csdl_opt._csdl_compat import Simulator, Model
    from femo_b200 import _hostops as H

    class M(Model):
        def define(self):
            self.create_input('x', shape=5, val=1.0)
    sim = Simulator(M())
    x = sim.vars['x']
    assert isinstance(x, H.TrackedArray) and x.version == 1
    sim['x'] = np.arange(5.0)
    assert sim.vars['x'].version == 2
    ref = sim['x']                                          # writable reference leaves the simulator
    assert sim.vars['x'].version == 3 and ref is sim.vars['x']


def test_hostops_bulk_helpers_match_numpy():
    from femo_b200 import _hostops as H
    rng = np.random.default_rng(0)
    for n in (7, H.BIG + 3):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        d = a.copy()
        H.iadd(d, b, -2.5)
        assert np.allclose(d, a - 2.5 * b, rtol=1e-15, atol=1e-15)
        H.scaled_copy(d, b, 3.0)
        assert np.array_equal(d, 3.0 * b)
        H.copy(d, a)
        assert np.array_equal(d, a)
        H.fill(d, 4.0)
        assert np.all(d == 4.0)
        H.copy(d, np.array([2.0]))
        assert np.all(d == 2.0)


def test_density_filter_lattice_detection_and_weights_3d():
    """GeneralFilterOperation on a 3-D lattice of hexahedron centres: lattice detection and the sparse Jacobian
    (weight_triplets) against the oracle's KD-tree restatement of general_filter_model.py:67-90."""
    import scipy.sparse as sp
    from femo_b200.csdl_opt.pre_processor.general_filter_model import GeneralFilterOperation, _lattice
    from oracle.filter import weight_matrix
    nx, ny, nz, h = 7, 5, 4, (2.0, 1.5, 1.0)
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
    coords = np.stack([(I.ravel() + 0.5) * h[0], (J.ravel() + 0.5) * h[1], (K.ravel() + 0.5) * h[2]], axis=1)
    assert _lattice(coords) == (nx, ny, nz, h[0], h[1], h[2])
    assert _lattice(coords[:nx * ny, :2]) == (nx, ny, 1, h[0], h[1], 1.0)
    op = GeneralFilterOperation(nel=nx * ny * nz, beta=2.0, coordinates=coords, h_avg=1.6)
    op.define()
    r, c, v = op.weight_triplets()
    W = sp.csr_matrix((v, (r, c)), shape=(nx * ny * nz,) * 2)
    Wo = weight_matrix(coords, 1.6, 2.0)
    assert abs(W - Wo).max() < 1e-14


def test_project_cellwise_expressions_onto_dg0_on_quadrilaterals():
    """examples/beam_topo_opt/run_topo_opt_cantilever_beam.py:255-260: project(rho**3, penalized) and the RAMP interpolation
    project(rho / (1 + 8 (1 - rho)), penalized) on the quadrilateral design mesh; onto DG0 the L2 projection is cell-wise
    exact (diagonal mass matrix), no device needed."""
    from femo_b200.fea.fea_b200 import FunctionSpace, Function
    from femo_b200.fea.fem import Mesh
    from femo_b200.fea.utils_b200 import project, getFuncArray, setFuncArray
    from femo_b200 import engine as E
    mesh = Mesh(E.EngineMesh.rectangle_quad((0.0, 0.0), (2.0, 1.0), 8, 4), 'quadrilateral')
    V = FunctionSpace(mesh, ('DG', 0))
    rho, out = Function(V), Function(V)
    r = np.linspace(0.1, 1.0, 32)
    setFuncArray(rho, r)
    project(rho ** 3, out)
    assert np.allclose(getFuncArray(out), r ** 3, rtol=1e-15)
    project(rho / (1 + 8. * (1. - rho)), out)
    assert np.allclose(getFuncArray(out), r / (1 + 8 * (1 - r)), rtol=1e-15)
