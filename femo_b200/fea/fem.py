"""Minimal finite-element objects the femo API is written against.

The reference builds its forms from dolfinx / UFL objects (Function,
FunctionSpace, TestFunction, dirichletbc, derivative, ...; imports at
femo/fea/fea_dolfinx.py:5-16 and femo/fea/utils_dolfinx.py:5-21).  UFL's
symbolic algebra is replaced here by *form families*: closed sets of forms whose
quadrature kernels exist in the CUDA engine.  A `Form` names (family, kind);
`derivative` moves between kinds exactly where the reference calls
ufl.derivative (utils_dolfinx.py:313-314).

Host arrays are authoritative at the API boundary (numpy in / numpy out, as
with PETSc Vec.getArray in the reference); device copies are torch tensors
used purely as buffers and are refreshed lazily.
"""
import numpy as np

from .. import _hostops as _H

from .. import engine as _E


# --------------------------------------------------------------------------
# mesh
# --------------------------------------------------------------------------
class _Topology:
    def __init__(self, dim, ncells):
        self.dim = dim
        self._ncells = ncells

    def index_map(self, dim):
        class _IM:
            pass
        im = _IM()
        im.size_local = im.size_global = self._ncells
        return im

    def create_connectivity(self, a, b):
        return None


class _Geometry:
    def __init__(self, x3, gdim):
        self.x = x3
        self.dim = gdim


class Mesh:
    """Structured mesh in canonical lattice numbering ("Numbering")."""

    def __init__(self, emesh, cell_type, slab=None):
        """slab: dict(nx, gny, rank, nranks, own0, own1, cown0, cown1) when this is one rank's y-slab of a partitioned
        lattice (local mesh = owned rows + one ghost cell row below + one ghost node row above, csrc/dist.cuh): the
        API-level arrays of Functions then hold the OWNED dofs only, as dolfinx's Vec.getArray does under MPI."""
        self._e = emesh
        self.cell_type = cell_type
        self.slab = slab
        xy = emesh.coords()
        x3 = np.zeros((emesh.nverts, 3))
        x3[:, :emesh.gdim] = xy
        self.geometry = _Geometry(x3, emesh.gdim)
        self.topology = _Topology({'interval': 1, 'hexahedron': 3}.get(cell_type, 2), emesh.ncells)
        self.cells = emesh.cells()
        self.num_cells = emesh.ncells
        self.num_vertices = emesh.nverts

    def edges(self):
        """Triangle meshes: (edge -> [min vertex, max vertex], cell -> edges) in the engine's canonical order
        (edges sorted lexicographically by that pair; local edge i is opposite local vertex i)."""
        if getattr(self, '_edges', None) is None:
            c = self.cells.astype(np.int64)
            a = np.stack([c[:, 1], c[:, 0], c[:, 0]], axis=1)
            b = np.stack([c[:, 2], c[:, 2], c[:, 1]], axis=1)
            base = self.num_vertices + 1
            uniq, inv = np.unique((np.minimum(a, b) * base + np.maximum(a, b)).ravel(), return_inverse=True)
            self._edges = (np.stack([uniq // base, uniq % base], axis=1).astype(np.int32), inv.reshape(-1, 3).astype(np.int32))
        return self._edges

    def h(self):
        """Cell diameters (dolfinx.cpp.mesh.h, utils_dolfinx.py:526-530)."""
        x = self.geometry.x[self.cells]
        d = np.zeros(self.num_cells)
        nv = x.shape[1]
        for a in range(nv):
            for b in range(a + 1, nv):
                d = np.maximum(d, np.linalg.norm(x[:, a] - x[:, b], axis=1))
        return d


# --------------------------------------------------------------------------
# spaces and functions
# --------------------------------------------------------------------------
class _IndexMap:
    def __init__(self, n):
        self.size_local = self.size_global = n


class _DofMap:
    def __init__(self, nnodes, bs):
        self.index_map = _IndexMap(nnodes)
        self.index_map_bs = bs


class FunctionSpace:
    """FunctionSpace(mesh, ('CG', 1) | ('DG', 0) | ('Hermite', 3)) (+ block size)."""

    def __init__(self, mesh, element, block=1):
        family, degree = element
        family = {'Lagrange': 'CG', 'P': 'CG', 'Q': 'CG', 'Discontinuous Lagrange': 'DG'}.get(family, family)
        if (family, degree) == ('DG', 0):
            nnodes = mesh.num_cells
        elif (family, degree) == ('CG', 1):
            nnodes = mesh.num_vertices
        elif (family, degree) == ('CG', 2):
            if mesh.cell_type != 'triangle' or block != 1:
                raise ValueError('femo_b200: CG2 kernels exist for scalar spaces on triangles')
            nnodes = mesh.num_vertices + mesh.edges()[0].shape[0]
        elif family == 'RMPlate':
            # mixed Reissner-Mindlin plate space: CG2 deflection, then CG1^2 rotations (forms/shell.py)
            if mesh.cell_type != 'triangle':
                raise ValueError('femo_b200: the RM plate space needs a triangle mesh')
            nnodes, block = 3 * mesh.num_vertices + mesh.edges()[0].shape[0], 1
        elif (family, degree) == ('Hermite', 3):
            nnodes, block = mesh.num_vertices, 2
        else:
            raise ValueError('femo_b200: no kernels for element %s%d' % (family, degree))
        self.mesh = mesh
        self.family, self.degree, self.block = family, degree, block
        self.num_nodes = nnodes
        self.dim = nnodes * block
        self.dofmap = _DofMap(nnodes, block)
        self.num_sub_spaces = block if block > 1 and family != 'Hermite' else 0
        self.own = None                       # slice of the owned dofs inside the local (owned + ghost) vector
        self.halo_kind = 0
        if mesh.slab is not None:
            sl = mesh.slab
            if (family, degree) == ('CG', 1):
                w = (sl['nx'] + 1) * block
                self.own = slice(sl['own0'] * w, sl['own1'] * w)
            elif (family, degree) == ('DG', 0):
                w = 2 * sl['nx'] * block
                self.own = slice(sl['cown0'] * w, sl['cown1'] * w)
                self.halo_kind = 1
            else:
                raise ValueError('femo_b200: partitioned meshes carry CG1 and DG0 spaces')
        self.local_dim = self.dim
        if self.own is not None:
            self.dim = self.own.stop - self.own.start

    def node_coordinates(self):
        if self.family == 'DG':
            return self.mesh.geometry.x[self.mesh.cells].mean(axis=1)
        if self.family == 'RMPlate':                # w at vertices, w at edge midpoints, rotations at the vertices (x2)
            ev = self.mesh.edges()[0]
            X = self.mesh.geometry.x
            return np.concatenate([X, 0.5 * (X[ev[:, 0]] + X[ev[:, 1]]), np.repeat(X, 2, axis=0)])
        if self.degree == 2:                       # vertices, then edge midpoints (engine's P2 numbering)
            ev = self.mesh.edges()[0]
            X = self.mesh.geometry.x
            return np.concatenate([X, 0.5 * (X[ev[:, 0]] + X[ev[:, 1]])])
        return self.mesh.geometry.x

    def tabulate_dof_coordinates(self):
        return self.node_coordinates()

    def __eq__(self, other):
        return (isinstance(other, FunctionSpace) and other.mesh is self.mesh and
                (other.family, other.degree, other.block) == (self.family, self.degree, self.block))

    def __hash__(self):
        return hash((id(self.mesh), self.family, self.degree, self.block))


def VectorFunctionSpace(mesh, element, dim=None):
    return FunctionSpace(mesh, element, block=mesh.geometry.dim if dim is None else dim)


class Vector:
    """The slice of petsc4py.Vec the reference uses on Function.vector."""

    def __init__(self, func):
        self._f = func

    def getArray(self):
        """Owned dofs (the whole vector on one rank), as Vec.getArray; a view into the host mirror."""
        a = self._f._host_array()
        own = self._f.function_space.own
        return a if own is None else a[own]

    array = property(getArray)

    def set(self, value):
        self._f._fill(float(np.ravel(value)[0]))

    def setArray(self, values):
        self._f._assign(values)

    def __setitem__(self, key, values):
        if isinstance(key, slice) and key == slice(None) and isinstance(values, np.ndarray):
            self._f._assign(values)              # whole-vector assignment: no D2H refresh needed first
            return
        a = self.getArray()
        a[key] = values
        self._f._host_changed()

    def __getitem__(self, key):
        return self.getArray()[key]

    def assemble(self):
        pass

    def ghostUpdate(self, *a, **k):
        pass

    def norm(self):
        return float(np.linalg.norm(self.getArray()))

    def __len__(self):
        return self._f.function_space.dim


class _XView:
    def __init__(self, func):
        self._f = func

    @property
    def array(self):
        # handing out a writable view: assume the caller modifies it
        a = self._f._host_array()
        self._f._host_changed()
        return a


class Function:
    def __init__(self, V, name=None):
        self.function_space = V
        self.name = name
        self._host = np.zeros(V.local_dim)      # local vector: owned dofs + ghosts on a partitioned mesh
        self._dev = None            # torch tensor on the problem's device (a buffer, nothing more)
        self._prob = None
        self._host_ver = 1          # bumped on every host write
        self._src = None            # (address, version) of the tracked array these values came from
        self._dev_ver = 0           # host version the device copy mirrors (-1: device is newer)
        self.vector = Vector(self)
        self.x = _XView(self)

    # -- host/device coherence ------------------------------------------------
    def _host_array(self):
        if self._dev is not None and self._dev_ver == -1:
            import torch
            torch.from_numpy(self._host).copy_(self._dev)          # D2H into the (pinned) host mirror
            self._dev_ver = self._host_ver
            if self._prob is not None:
                self._prob.d2h_bytes += self._host.nbytes
        return self._host

    def _host_changed(self):
        self._host_ver += 1
        self._src = None

    def _fill(self, value):
        self._src = None
        if self._dev is not None:
            self._dev.fill_(value)              # constant fields never cross PCIe
            self._dev_ver = -1
        else:
            self._host.fill(value)
            self._host_changed()

    def _assign(self, values):
        ver = values.version if isinstance(values, _H.TrackedArray) else None
        v = np.asarray(values, dtype=np.float64).ravel()
        own = self.function_space.own
        if v.size == 1:
            return self._fill(float(v[0]))
        elif own is not None and v.size == own.stop - own.start:
            # one rank's OWNED values: into the owned block of the local vector; the ghost rows are refreshed by a halo
            # exchange after the next upload (device_tensor).  Also taken when the rank has no ghost entries in this space
            # (rank 0 holds no ghost cell row): the exchange is collective, every rank must reach it (its neighbours
            # need this rank's rows)
            self._host_array()
            _H.copy(self._host[own], v)
            self._src = None
            ver = None
        else:
            src = (v.__array_interface__['data'][0], ver)
            if ver is not None and src == self._src:
                return                           # already holds exactly this version of that storage
            if v.size != self._host.size:
                raise ValueError('size mismatch: function has %d dofs, got %d values' % (self._host.size, v.size))
            if v.__array_interface__['data'][0] != self._host.__array_interface__['data'][0]:
                if self._dev is not None and _H.is_pinned(v):
                    # page-locked source (the backend's variable storage): DMA straight into the device
                    # buffer; the host mirror is refreshed lazily if somebody reads it
                    import torch
                    self._dev.copy_(torch.from_numpy(v))
                    self._dev_ver = -1
                    if self._prob is not None:
                        self._prob.h2d_bytes += v.nbytes
                    self._src = src if ver is not None else None
                    return
                _H.copy(self._host, v)
        self._dev_ver = 0 if self._dev_ver == -1 else self._dev_ver
        self._host_changed()
        if v.size != 1 and ver is not None:
            self._src = src

    def device_tensor(self, prob):
        """Device buffer holding the current values (H2D copy if the host is newer)."""
        import torch
        if self._dev is None:
            self._dev = torch.empty(self._host.size, dtype=torch.float64, device=prob.device)
            self._dev_ver = 0
            self._prob = prob
            pinned = torch.empty(self._host.size, dtype=torch.float64).pin_memory().numpy()
            pinned[:] = self._host                                 # re-home the host mirror in pinned memory
            self._host = pinned
        if self._dev_ver != -1 and self._dev_ver != self._host_ver:
            self._dev.copy_(torch.from_numpy(self._host), non_blocking=False)
            self._dev_ver = self._host_ver
            prob.h2d_bytes += self._host.nbytes
            if self.function_space.own is not None:       # partitioned: neighbours' owned rows into our ghost rows
                prob.halo(self._dev, self.function_space.halo_kind)
        return self._dev

    def mark_device_written(self):
        self._dev_ver = -1
        self._src = None

    # -- dolfinx-like surface ----------------------------------------------
    def interpolate(self, fn):
        X = self.function_space.node_coordinates().T        # (3, nnodes) as dolfinx passes it
        vals = np.asarray(fn(X), dtype=np.float64)
        bs = self.function_space.block
        if bs > 1 and self.function_space.family != 'Hermite':
            vals = np.asarray(vals).reshape(bs, -1).T.ravel()
        self._assign(vals)

    def rename(self, name, label=None):
        self.name = name

    def __pow__(self, p):
        return Expr('pow', self, p)

    # cell-wise algebra of DG0 functions (the UFL expressions the examples hand to `project`, e.g. the RAMP interpolation
    # rho / (1 + 8 (1 - rho)) of examples/beam_topo_opt/run_topo_opt_cantilever_beam.py:259): evaluated pointwise per cell
    def __add__(self, o): return Expr('op', np.add, self, o)
    def __radd__(self, o): return Expr('op', np.add, o, self)
    def __sub__(self, o): return Expr('op', np.subtract, self, o)
    def __rsub__(self, o): return Expr('op', np.subtract, o, self)
    def __mul__(self, o): return Expr('op', np.multiply, self, o)
    def __rmul__(self, o): return Expr('op', np.multiply, o, self)
    def __truediv__(self, o): return Expr('op', np.divide, self, o)
    def __rtruediv__(self, o): return Expr('op', np.divide, o, self)

    def copy(self):
        g = Function(self.function_space, self.name)
        g._assign(self._host_array())
        return g


class Expr:
    """Stand-in for the UFL expressions handed to `project`: ('u_ex'|'f_ex') analytic fields,
    ('pow', function, exponent), ('op', numpy ufunc, a, b) cell-wise arithmetic of DG0 functions and numbers."""

    def __init__(self, kind, *args):
        self.kind, self.args = kind, args

    def __add__(self, o): return Expr('op', np.add, self, o)
    def __radd__(self, o): return Expr('op', np.add, o, self)
    def __sub__(self, o): return Expr('op', np.subtract, self, o)
    def __rsub__(self, o): return Expr('op', np.subtract, o, self)
    def __mul__(self, o): return Expr('op', np.multiply, self, o)
    def __rmul__(self, o): return Expr('op', np.multiply, o, self)
    def __truediv__(self, o): return Expr('op', np.divide, self, o)
    def __rtruediv__(self, o): return Expr('op', np.divide, o, self)
    def __pow__(self, p): return Expr('op', np.power, self, p)

    def cellwise(self, V):
        """Value per cell when every leaf is a DG0 function on V's mesh or a number (None otherwise)."""
        def ev(a):
            if isinstance(a, Function):
                if a.function_space.family != 'DG' or a.function_space.mesh is not V.mesh:
                    return None
                return a._host_array()
            if isinstance(a, Expr):
                return a.cellwise(V)
            return float(a) if np.ndim(a) == 0 else None
        if self.kind == 'pow':
            b = ev(self.args[0])
            return None if b is None else np.power(b, float(self.args[1]))
        if self.kind == 'op':
            x, y = ev(self.args[1]), ev(self.args[2])
            return None if x is None or y is None else self.args[0](x, y)
        return None


class Constant:
    def __init__(self, mesh, value):
        self.mesh = mesh
        self.value = np.asarray(value, dtype=np.float64)

    def __float__(self):
        return float(self.value)


class TestFunction:
    def __init__(self, V):
        self.function_space = V


class TrialFunction(TestFunction):
    pass


# --------------------------------------------------------------------------
# boundary conditions
# --------------------------------------------------------------------------
class DirichletBC:
    def __init__(self, value, dofs, V=None):
        self.value = value
        d = dofs[0] if isinstance(dofs, (list, tuple)) else dofs
        self.dofs = np.asarray(d, dtype=np.int32).ravel()
        self.function_space = V if V is not None else getattr(value, 'function_space', None)

    def values(self, n):
        if isinstance(self.value, Function):
            return self.value._host_array()
        return np.full(n, float(np.ravel(getattr(self.value, 'value', self.value))[0]))

    def dof_values(self):
        """Current prescribed values at this condition's dofs (read live, as dolfinx does on every solve)."""
        if isinstance(self.value, Function):
            return np.ascontiguousarray(self.value._host_array()[self.dofs], dtype=np.float64)
        return np.full(self.dofs.size, float(np.ravel(getattr(self.value, 'value', self.value))[0]))


def dirichletbc(value, dofs, V=None):
    """dolfinx.fem.dirichletbc as used at fea_dolfinx.py:169-176."""
    return DirichletBC(value, dofs, V)


def locate_dofs_geometrical(V, marker):
    """dolfinx.fem.locate_dofs_geometrical for a (V, V) pair or a single space;
    returns dof indices (all components of blocked spaces)."""
    pair = isinstance(V, (tuple, list))
    space = V[0] if pair else V
    X = space.node_coordinates().T
    nodes = np.nonzero(np.asarray(marker(X)))[0].astype(np.int32)
    bs = space.block
    if bs > 1:
        nodes = (bs * nodes[:, None] + np.arange(bs, dtype=np.int32)[None, :]).ravel()
    return [nodes, nodes.copy()] if pair else nodes


def locate_entities_boundary(mesh, dim, marker):
    """Vertices (dim 0) or facets (dim tdim-1) on the boundary satisfying `marker`.
    Facets are returned as indices into the mesh's exterior-facet list."""
    X = mesh.geometry.x
    tdim = mesh.topology.dim
    if dim == 0:
        on = np.zeros(mesh.num_vertices, dtype=bool)
        fc, fl = mesh._e.exterior_facets()
        lf = _local_facets(mesh)
        on[np.unique(mesh.cells[fc][np.arange(fc.size)[:, None], lf[fl]])] = True
        ok = np.asarray(marker(X.T)) & on
        return np.nonzero(ok)[0].astype(np.int32)
    if dim != tdim - 1:
        raise ValueError('locate_entities_boundary: only vertices and facets are supported')
    fc, fl = mesh._e.exterior_facets()
    lf = _local_facets(mesh)
    fv = mesh.cells[fc][np.arange(fc.size)[:, None], lf[fl]]          # (nf, verts per facet)
    ok = np.ones(fc.size, dtype=bool)
    for k in range(fv.shape[1]):
        ok &= np.asarray(marker(X[fv[:, k]].T))
    return np.nonzero(ok)[0].astype(np.int32)


def _local_facets(mesh):
    if mesh.cell_type == 'triangle':
        return np.array([[1, 2], [0, 2], [0, 1]])
    if mesh.cell_type == 'quadrilateral':
        return np.array([[0, 1], [0, 2], [1, 3], [2, 3]])
    if mesh.cell_type == 'hexahedron':
        return np.array([[0, 1, 2, 3], [0, 1, 4, 5], [0, 2, 4, 6], [1, 3, 5, 7], [2, 3, 6, 7], [4, 5, 6, 7]])
    return np.array([[0], [1]])


def locate_dofs_topological(V, dim, entities):
    if dim != 0:
        raise ValueError('locate_dofs_topological: only vertex entities are supported')
    bs = V.block
    ent = np.asarray(entities, dtype=np.int32)
    if bs == 1:
        return ent
    return (bs * ent[:, None] + np.arange(bs, dtype=np.int32)[None, :]).ravel()


class MeshTags:
    def __init__(self, mesh, dim, indices, values):
        self.mesh, self.dim = mesh, dim
        self.indices = np.asarray(indices, dtype=np.int32)
        self.values = np.asarray(values, dtype=np.int32)


def meshtags(mesh, dim, indices, values):
    return MeshTags(mesh, dim, indices, values)


class Measure:
    """ufl.Measure('ds', subdomain_data=facet_tag, metadata=...) -> ds_(tag)."""

    def __init__(self, kind, domain=None, subdomain_data=None, metadata=None, tag=None):
        self.kind, self.domain, self.subdomain_data, self.metadata, self.tag = kind, domain, subdomain_data, metadata, tag

    def __call__(self, tag):
        return Measure(self.kind, self.domain, self.subdomain_data, self.metadata, tag)

    @property
    def sides(self):
        """One-sided (cell, local facet) pairs of this tag on an imported mesh (interior facets: both sides)."""
        return self.subdomain_data.sides(self.tag)

    def facets(self):
        """Indices into the mesh's exterior-facet list carrying this tag."""
        t = self.subdomain_data
        if t is None:
            return None
        idx = t.indices[t.values == self.tag]
        if t.mesh.cell_type == 'interval':
            # facets of an interval are vertices: entity ids are vertex ids (0 or n) -> facet 0 / 1
            idx = np.array([0 if v == 0 else 1 for v in idx], dtype=np.int32)
        return idx


ds = Measure('ds')
dx = Measure('dx')


# --------------------------------------------------------------------------
# forms
# --------------------------------------------------------------------------
class Form:
    """(family instance, kind[, output id, slot]).  kinds: 'residual', 'dRdu',
    'dRdm', 'output', 'output_grad', 'mass', 'mass_rhs'."""

    def __init__(self, fam, kind, out_id=0, slot=0):
        self.fam, self.kind, self.out_id, self.slot = fam, kind, out_id, slot

    def __add__(self, other):
        raise TypeError('femo_b200 forms are closed families; sums of forms are not supported')


def derivative(form, function, direction=None):
    """Gateaux derivative within a family (utils_dolfinx.py:313-314)."""
    fam = form.fam
    slot = fam.slot_of(function)
    if form.kind == 'residual':
        return Form(fam, 'dRdu') if slot == 0 else Form(fam, 'dRdm', slot=slot - 1)
    if form.kind == 'output':
        return Form(fam, 'output_grad', out_id=form.out_id, slot=slot)
    raise TypeError('derivative of a %s form is not available' % form.kind)
