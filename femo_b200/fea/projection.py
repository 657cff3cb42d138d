"""`project` of the reference (femo/fea/utils_dolfinx.py:549-583): L2 projection of
an expression onto a CG1 / DG0 function, on the device.

Expressions (the UFL objects the reference passes) are one of
  Expr('u_ex') / Expr('f_ex')   the analytic fields of examples/nonlinear_poisson_opt
  func ** p                     a DG0 function to a power (examples/beam_topo_opt:264-268)
  cell-wise arithmetic          of DG0 functions and numbers, e.g. rho / (1 + 8 * (1 - rho)) (RAMP, :259): onto a DG0 target on
                                ANY cell type (exact: the DG0 mass matrix is diagonal)
  func                          a CG1 or DG0 function
The mass matrix and right-hand side are assembled by the projection family's kernels
and solved with Jacobi-CG (the reference uses PETSc's default KSP at rtol 1e-5);
lump_mass=True divides the load vector by the row sums of the mass matrix.
"""
import numpy as np

from .fem import Function, Expr
from .family import FormFamily
from .. import engine as _E


def _source(v, target):
    mesh = target.function_space.mesh
    if isinstance(v, Expr) and v.kind in ('u_ex', 'f_ex'):
        return (0 if v.kind == 'u_ex' else 1), 1.0, None
    if isinstance(v, Expr) and v.kind == 'pow':
        f, p = v.args
        if f.function_space.family != 'DG':
            raise NotImplementedError('project: powers are supported for DG0 functions')
        return 2, float(p), f
    if isinstance(v, Function):
        return (2 if v.function_space.family == 'DG' else 3), 1.0, v
    raise TypeError('project: unsupported expression %r' % (v,))


def project(v, target_func, bcs=[], lump_mass=False):
    if bcs:
        raise NotImplementedError('project: Dirichlet conditions are not supported')
    V = target_func.function_space
    if V.family == 'DG' and V.block == 1 and isinstance(v, Expr):
        # piecewise-constant expression onto DG0, any cell type (quadrilaterals / hexahedra of the topology example): the
        # mass matrix is diag(cell volume), so the L2 projection IS the cell-wise value of the expression
        vals = v.cellwise(V)
        if vals is not None:
            target_func._assign(np.ascontiguousarray(np.broadcast_to(vals, (V.local_dim,)), dtype=np.float64))
            return target_func
    if V.mesh.cell_type != 'triangle' or V.block != 1:
        raise NotImplementedError('project: scalar CG1 / DG0 targets on triangle meshes')
    source, power, func = _source(v, target_func)
    target = 1 if V.family == 'DG' else 0
    if func is None:                                # analytic source: the input slot is a placeholder
        func = target_func.__dict__.setdefault('_femo_dummy_src', Function(type(V)(V.mesh, ('DG', 0))))
    fam = FormFamily.get(_E.FAMILY_MASS_P1, V.mesh, target_func, [func], params=[target, source, power])
    fam.precond = 0
    p = fam.sync()
    fam.apply_bcs([])
    ut = target_func.device_tensor(p)
    ut.zero_()
    vals, _ = p.assemble_jacobian(plain=True, bc=False)
    b = p.assemble_residual()                       # R(0) = -int g w dx
    if lump_mass:
        ones = p.new_vector(p.N, 1.0)
        rowsum = p.spmv(0, vals, ones)
        import ctypes as C
        from .._lib import lib, check
        check(lib.femo_pointwise_divide(p._h, -1.0, C.c_void_p(b.data_ptr()), C.c_void_p(rowsum.data_ptr()),
                                        C.c_void_p(ut.data_ptr()), p.N))
        target_func.mark_device_written()
        return target_func
    x = p.new_vector(p.N, 0.0)
    _, info = p.linear_solve(vals, b, x, rtol=1e-13, max_it=10000, check_every=5, precond=0)
    if not info['converged']:
        raise RuntimeError('project: mass solve did not converge: %r' % (info,))
    p.axpy(-1.0, x, ut)                             # u = -M^-1 R(0)
    target_func.mark_device_written()
    return target_func
