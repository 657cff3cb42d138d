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
