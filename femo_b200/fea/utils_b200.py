"""Free functions of the femo FEA layer, backed by the CUDA engine.

Same names, argument meaning and error behaviour as the reference's
femo/fea/utils_dolfinx.py (cited per function); the dolfinx/PETSc one-liners
they delegate to are replaced by libfemo_b200 calls through `engine.py`.
"""
from timeit import default_timer

import numpy as np

from .fem import (Mesh, Function, FunctionSpace, VectorFunctionSpace, Constant, TestFunction, TrialFunction,
                  Form, derivative, dirichletbc, locate_dofs_geometrical, locate_dofs_topological,
                  locate_entities_boundary, meshtags, Measure, ds, dx)
from .. import engine as _E
from .._lib import FemoError

DOLFIN_EPS = 3E-16

# Krylov settings that replace the reference's direct LU (KSP preonly + MUMPS,
# utils_dolfinx.py:405-408,476-512).  rtol is tight enough for the 1e-8
# derivative target of BASELINE.json.
KRYLOV = dict(rtol=1e-12, max_it=200000, check_every=1, precond=2, cheb_degree=2)


# ---- meshes (utils_dolfinx.py:136-153) -------------------------------------
def createUnitSquareMesh(n):
    """One process: the n x n unit square.  Under torchrun with femo_b200.dist.init() done (the reference's nominal
    MPI.COMM_WORLD, utils_dolfinx.py:32,140-153): this rank's y-slab of the same lattice -- owned rows plus a one-cell
    ghost layer -- and every Function on it exposes the rank's OWNED dofs."""
    from .. import dist as _D
    R, rank = _D._state.get('nranks', 1), _D._state.get('rank', 0)
    if not _D._state.get('initialised') or R == 1:
        return Mesh(_E.EngineMesh.unit_square(n), 'triangle')
    if n % R:
        raise ValueError('createUnitSquareMesh: n must be divisible by the number of ranks')
    rows = n // R
    a, b = rank * rows, rank * rows + rows
    crow0 = a - 1 if a > 0 else 0
    ncrows = b - crow0
    own0 = a - crow0
    slab = dict(nx=n, gny=n, rank=rank, nranks=R, crow0=crow0, ncrows=ncrows, own0=own0,
                own1=own0 + rows + (1 if rank == R - 1 else 0), cown0=own0, cown1=own0 + rows)
    em = _E.EngineMesh.unit_square(n, ncrows, lo=(0.0, crow0 / n), hi=(1.0, (crow0 + ncrows) / n))
    return Mesh(em, 'triangle', slab=slab)


def createIntervalMesh(n, x0, x1):
    return Mesh(_E.EngineMesh.interval(n, x0, x1), 'interval')


def createRectangleMesh(pt1, pt2, nx, ny):
    return Mesh(_E.EngineMesh.rectangle_quad(tuple(pt1), tuple(pt2), nx, ny), 'quadrilateral')


def createBoxMesh(pt1, pt2, nx, ny, nz):
    """dolfinx.mesh.create_box(..., CellType.hexahedron): the 3-D cantilever of SURVEY.md section 8d (C4-3D)."""
    return Mesh(_E.EngineMesh.box_hex(tuple(pt1), tuple(pt2), nx, ny, nz), 'hexahedron')


def meshSize(mesh):
    """utils_dolfinx.py:526-530"""
    return mesh.h()


# ---- array I/O (utils_dolfinx.py:155-167,300-311) ---------------------------
def getFuncArray(v):
    return v.vector.getArray()


def setFuncArray(v, v_array):
    v.vector[:] = v_array
    v.vector.assemble()
    v.vector.ghostUpdate()


def update(v, v_values):
    """Length-1 arrays broadcast to every dof (quirk B6, utils_dolfinx.py:308-309)."""
    if len(v_values) == 1:
        v.vector.set(v_values)
    else:
        setFuncArray(v, v_values)


def computePartials(form, function):
    return derivative(form, function)


def createFunction(function):
    return Function(function.function_space)


# ---- matrices ----------------------------------------------------------------
class Mat:
    """Device CSR matrix: values tensor + the family's shared pattern."""

    def __init__(self, fam, which, vals, transposed=False):
        self.fam, self.which, self.vals, self.transposed = fam, which, vals, transposed

    def _info(self):
        return self.fam.problem.pattern_info(self.which)

    def getSizes(self):
        i = self._info()
        return (i['cols'], i['rows']) if self.transposed else (i['rows'], i['cols'])

    getSize = getSizes
    size = property(getSizes)

    def assemble(self):
        return self

    def copy(self):
        return Mat(self.fam, self.which, self.vals.clone(), self.transposed)

    def transpose(self, *a):
        return Mat(self.fam, self.which, self.vals, not self.transposed)

    def getValuesCSR(self):
        import scipy.sparse as sp
        rp, col = self.fam.problem.pattern(self.which)
        data = self.vals.cpu().numpy()
        self.fam.problem.d2h_bytes += data.nbytes
        if not self.transposed:
            return rp, col, data
        i = self._info()
        T = sp.csr_matrix((data, col, rp), shape=(i['rows'], i['cols'])).T.tocsr()
        T.sort_indices()
        return T.indptr, T.indices, T.data

    def mult(self, x_tensor, out=None):
        return self.fam.problem.spmv(self.which, self.vals, x_tensor, transpose=self.transposed, out=out)

    def multTranspose(self, x_tensor, out=None):
        return self.fam.problem.spmv(self.which, self.vals, x_tensor, transpose=not self.transposed, out=out)

    def convert(self, kind='dense'):
        import scipy.sparse as sp
        ip, ix, d = self.getValuesCSR()
        return sp.csr_matrix((d, ix, ip), shape=self.getSizes()).toarray()


def _download(prob, tensor):
    """Device -> host through a pinned staging buffer (a ring of three per size,
    so a result stays valid while the next two are produced; callers that keep
    it longer copy it, as CSDL does on assignment)."""
    import torch
    a = prob.pinned_buffer(tensor.numel())
    torch.from_numpy(a).copy_(tensor)
    prob.d2h_bytes += a.nbytes
    return a


def _own(fam, slot, a):
    """Owned block of a local vector living in the space of `slot` (0 state, 1+s input s); identity on one rank."""
    own = fam.function_of(slot).function_space.own
    return a if own is None else a[own]


# ---- assembly (utils_dolfinx.py:169-222) -----------------------------------------
def assembleScalar(c):
    p = c.fam.sync()
    return p.assemble_output(c.out_id)


def assembleVector(v):
    """Residual / gradient vector; Dirichlet values are NOT applied (quirk B11)."""
    p = v.fam.sync()
    if v.kind == 'residual':
        return _own(v.fam, 0, _download(p, p.assemble_residual()))
    if v.kind == 'output_grad':
        return _own(v.fam, v.slot, _download(p, p.assemble_output_grad(v.out_id, v.slot)))
    raise TypeError('form of kind %r does not assemble to a vector' % v.kind)


def assembleMatrix(M, bcs=[]):
    fam = M.fam
    p = fam.sync()
    if M.kind == 'dRdu':
        if bcs:
            fam.apply_bcs(bcs)
            _, vals = p.assemble_jacobian(plain=False, bc=True)
        else:
            vals, _ = p.assemble_jacobian(plain=True, bc=False)
        return Mat(fam, 0, vals)
    if M.kind == 'dRdm':
        return Mat(fam, 1 + M.slot, p.assemble_dRdm(M.slot))
    raise TypeError('form of kind %r does not assemble to a matrix' % M.kind)


def assembleSystem(J, F, bcs=[], rhs=True):
    """BC'd Jacobian and lifted right-hand side (utils_dolfinx.py:189-202).
    The RHS uses x0=None, scale=1: b = R - A(:,bc) g, b[bc] = g.  One element
    pass produces both the BC'd matrix and the un-BC'd one (kept as `A.plain`,
    quirk B8); rhs=False skips the vector the reference's caller discards."""
    fam = J.fam
    p = fam.sync()
    fam.apply_bcs(bcs)
    if bcs:
        vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
        A = Mat(fam, 0, vals_bc)
        A.plain = Mat(fam, 0, vals)
    else:
        vals, _ = p.assemble_jacobian(plain=True, bc=False)
        A = Mat(fam, 0, vals)
        A.plain = Mat(fam, 0, vals)       # same values, no reference cycle (frees promptly)
    if not rhs:
        return A, None
    b = p.system_rhs(vals) if bcs else p.assemble_residual()
    return A, _own(fam, 0, _download(p, b))


def assemble(f, dim=0, bcs=[]):
    if dim == 0:
        return assembleScalar(f)
    elif dim == 1:
        return assembleVector(f)
    elif dim == 2:
        M = assembleMatrix(f, bcs=bcs)
        return convertToDense(M.copy())
    else:
        return TypeError("Invalid type for assembly.")      # returned, not raised (quirk B7)


def assemble_partials(of=None, wrt=None, dim=1):
    """utils_dolfinx.py:216-222"""
    return assemble(derivative(of, wrt), dim=dim)


# ---- linear algebra (utils_dolfinx.py:241-297) ------------------------------------
def transpose(A):
    return A.transpose()


def convertToCOO(A):
    import scipy.sparse as sp
    ip, ix, d = A.getValuesCSR()
    return sp.csr_matrix((d, ix, ip), shape=A.getSizes()).tocoo()


def convertToDense(A_petsc):
    return A_petsc.convert("dense")


def computeMatVecProductFwd(A, x):
    """y = A x, x a Function (utils_dolfinx.py:256-264)."""
    p = A.fam.problem
    return _own(A.fam, A.which if A.transposed else 0, _download(p, A.mult(x.device_tensor(p))))


def computeMatVecProductBwd(A, R):
    """y = A^T R (utils_dolfinx.py:275-287)."""
    p = A.fam.problem
    return _own(A.fam, 0 if A.transposed else A.which, _download(p, A.multTranspose(R.device_tensor(p))))


# ---- nonlinear solves (utils_dolfinx.py:319-449) ----------------------------------
def solveNonlinear(res, func, bc, solver, report, initialize):
    start = default_timer()
    fam = res.fam
    if func is not fam.state:
        raise ValueError('solveNonlinear: `func` is not the state of the residual form')
    if solver == 'Newton' and initialize is True:
        func.vector.set(0.1)                                   # utils_dolfinx.py:433-435
    p = fam.sync()
    fam.apply_bcs(bc)
    fam.ensure_amg()
    try:
        info = p.newton_solve(kind=solver, krylov_rtol=KRYLOV['rtol'], krylov_max_it=KRYLOV['max_it'],
                              check_every=KRYLOV['check_every'],
                              precond=KRYLOV['precond'] if fam.precond is None else fam.precond,
                              method=fam.method, **dict(dict(cheb_degree=KRYLOV['cheb_degree']), **fam.krylov_extra),
                              **getattr(fam, 'newton_extra', {}))
    finally:
        func.mark_device_written()
    if solver == 'SNES':
        print("Converged reason:", info['converged'])
    stop = default_timer()
    if report is True:
        print("Solve nonlinear finished in ", stop - start, "seconds")
    fam.last_solve_info = info
    return info


class KSP:
    """setUpKSP_MUMPS stand-in: remembers the operator; solve() runs the Krylov
    method.  Signature ksp.solve(b, x) as petsc4py."""

    def __init__(self, A):
        self.A = A

    def solve(self, b, x):
        _solve_into(self.A, b, x)


def _as_function(v):
    return v._f if hasattr(v, '_f') else v


def _solve_into(A, b, x):
    """Solve A x = b for Function/Vector arguments; result stays on the device."""
    fam = A.fam
    p = fam.problem
    bf, xf = _as_function(b), _as_function(x)
    xt = xf.device_tensor(p)
    xt.zero_()
    fam.ensure_amg(A.vals)
    _, info = p.linear_solve(A.vals, bf.device_tensor(p), xt, transpose=A.transposed, rtol=KRYLOV['rtol'],
                             max_it=KRYLOV['max_it'], check_every=KRYLOV['check_every'],
                             precond=KRYLOV['precond'] if fam.precond is None else fam.precond,
                             method=fam.method, **dict(dict(cheb_degree=KRYLOV['cheb_degree']), **fam.krylov_extra))
    xf.mark_device_written()
    fam.last_linear_info = info
    if not info['converged']:
        raise FemoError(-5, 'Krylov solve did not converge: %r' % (info,))
    return info


def solveKSP_mumps(A, b, x):
    """utils_dolfinx.py:476-493 (direct LU replaced by preconditioned Krylov)."""
    _solve_into(A, b, x)


solveKSP = solveKSP_mumps


def setUpKSP_MUMPS(A):
    return KSP(A)


from .projection import project  # noqa: E402,F401  (utils_dolfinx.py:549-583)
from .mesh_io import import_mesh, read_msh  # noqa: E402,F401  (utils_dolfinx.py:69-123)
from .fem import Expr  # noqa: E402,F401


# ---- errors and projection (utils_dolfinx.py:225-237,549-583) ------------------------
def errorNorm(v, v_ex, norm='L2'):
    """L2 norm of v - v_ex for Functions on the same P1/DG0 space (degree-2 rule)."""
    a = getFuncArray(v) if isinstance(v, Function) else v
    b = getFuncArray(v_ex) if isinstance(v_ex, Function) else v_ex
    f = v if isinstance(v, Function) else v_ex
    V = f.function_space
    mesh = V.mesh
    if mesh.cell_type != 'triangle' or norm != 'L2':
        raise NotImplementedError('errorNorm: only L2 on triangles')
    X = mesh.geometry.x[mesh.cells][:, :, :2]
    det = np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0]))
    e = (a - b)
    if V.family == 'DG':
        return float(np.sqrt(np.sum(0.5 * det * e * e)))
    ee = e[mesh.cells]
    s = (ee ** 2).sum(axis=1) + ee[:, 0] * ee[:, 1] + ee[:, 0] * ee[:, 2] + ee[:, 1] * ee[:, 2]
    return float(np.sqrt(np.sum(det / 12.0 * s)))


# ---- motor helpers (utils_dolfinx.py:126-134,514-546,587-641) ------------------------
def findNodeIndices(node_coordinates, coordinates):
    """Indices of the mesh vertices closest to `node_coordinates` (scipy KDTree, as the reference)."""
    from scipy.spatial import KDTree
    dist, node_indices = KDTree(coordinates).query(node_coordinates)
    return node_indices


def locateDOFs(coords, V, input='polar'):
    """Dofs (both components) of the vertices nearest to the given edge points of the mesh-motion problem."""
    coords = np.reshape(np.array(coords, dtype=np.float64), (-1, 2))
    if input == 'polar':
        theta, r = coords[:, 0].copy(), coords[:, 1].copy()
        coords = np.stack([r * np.cos(theta), r * np.sin(theta)], axis=1)
    node_indices = findNodeIndices(coords, V.tabulate_dof_coordinates()[:, :-1])
    edge_indices = np.empty(2 * len(node_indices))
    edge_indices[0::2] = 2 * node_indices
    edge_indices[1::2] = 2 * node_indices + 1
    return edge_indices.astype('int')


def move(mesh, u):
    """Add the displacement function u (vector CG1) to the mesh coordinates."""
    a = getFuncArray(u).reshape(-1, mesh.geometry.dim)
    mesh.geometry.x[:, :mesh.geometry.dim] += a


def moveBackward(mesh, u):
    a = getFuncArray(u).reshape(-1, mesh.geometry.dim)
    mesh.geometry.x[:, :mesh.geometry.dim] -= a


def createCustomMeasure(mesh, dim, SubdomainFunc, measure: str, tag: int):
    """Tagged measure from a geometric marker (quadrature degree 4 in the reference, :537)."""
    metadata = {"quadrature_degree": 4}
    subdomain = locate_entities_boundary(mesh, dim, SubdomainFunc)
    subdomain_tag = meshtags(mesh, dim, subdomain, np.full(len(subdomain), tag, dtype=np.int32))
    return Measure(measure, domain=mesh, subdomain_data=subdomain_tag, metadata=metadata)
