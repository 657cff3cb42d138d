"""Structured meshes and dofmaps for the oracle (TEST INFRASTRUCTURE ONLY).

Restates what the reference gets from dolfinx.mesh.create_unit_square /
create_rectangle / create_interval (/root/reference/femo/fea/utils_dolfinx.py:
136-153).  dolfinx reorders cells and dofs after creation and that reordering
cannot be reproduced offline, so the canonical
numbering is lattice order:

  vertex (ix,iy)        -> iy*(nx+1) + ix
  triangle cells        -> 2*(iy*nx+ix) + {0,1}, "right" diagonal
                           [v0,v1,v3], [v0,v2,v3]   [upstream, from memory]
  quadrilateral cells   -> iy*nx+ix, vertices [v0,v1,v2,v3] in tensor order
  interval cells        -> i, vertices [i, i+1]

Exterior facets are found generically (a facet owned by exactly one cell) and
sorted by (cell, local facet); local facet i of a triangle is opposite vertex i,
facet k of an interval is its vertex k,
quadrilateral facets are (v0,v1),(v0,v2),(v1,v3),(v2,v3) as in basix.
"""
import numpy as np

_SIMPLEX_FACETS = {
    2: np.array([[0], [1]]),                      # interval: facet k is vertex k (basix)
    3: np.array([[1, 2], [0, 2], [0, 1]]),        # triangle
}
_QUAD_FACETS = np.array([[0, 1], [0, 2], [1, 3], [2, 3]])
# basix hexahedron facets (tensor order on each face) [upstream, from memory]
_HEX_FACETS = np.array([[0, 1, 2, 3], [0, 1, 4, 5], [0, 2, 4, 6], [1, 3, 5, 7], [2, 3, 6, 7], [4, 5, 6, 7]])


class Mesh:
    def __init__(self, kind, coords, cells, shape, lo, hi):
        self.kind = kind                  # 'interval' | 'triangle' | 'quadrilateral'
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.shape = tuple(shape)         # (nx,) or (nx, ny)
        self.lo, self.hi = tuple(lo), tuple(hi)
        self.gdim = self.coords.shape[1]
        self.ncells = self.cells.shape[0]
        self.nverts = self.coords.shape[0]
        self._bf = None

    @property
    def local_facets(self):
        if self.kind == 'quadrilateral':
            return _QUAD_FACETS
        if self.kind == 'hexahedron':
            return _HEX_FACETS
        return _SIMPLEX_FACETS[self.cells.shape[1]]

    def exterior_facets(self):
        """(cell, local_facet) of every exterior facet, sorted by (cell, local)."""
        if self._bf is None:
            lf = self.local_facets
            nc, nf = self.ncells, lf.shape[0]
            fv = np.sort(self.cells[:, lf], axis=2).reshape(nc * nf, -1)
            _, inv, cnt = np.unique(fv, axis=0, return_inverse=True, return_counts=True)
            ext = np.nonzero(cnt[inv.ravel()] == 1)[0]
            self._bf = (ext // nf).astype(np.int32), (ext % nf).astype(np.int32)
        return self._bf

    def cell_diameter(self):
        """UFL CellDiameter: largest vertex-to-vertex distance of each cell."""
        x = self.coords[self.cells]
        d = np.zeros(self.ncells)
        nv = x.shape[1]
        for a in range(nv):
            for b in range(a + 1, nv):
                d = np.maximum(d, np.linalg.norm(x[:, a] - x[:, b], axis=1))
        return d


def unit_square_tri(nx, ny=None, lo=(0.0, 0.0), hi=(1.0, 1.0)):
    ny = nx if ny is None else ny
    xs = lo[0] + (hi[0] - lo[0]) * np.arange(nx + 1) / nx
    ys = lo[1] + (hi[1] - lo[1]) * np.arange(ny + 1) / ny
    X, Y = np.meshgrid(xs, ys, indexing='xy')
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing='xy')
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = np.empty((2 * nx * ny, 3), dtype=np.int32)
    cells[0::2] = np.stack([v0, v1, v3], axis=1)
    cells[1::2] = np.stack([v0, v2, v3], axis=1)
    return Mesh('triangle', coords, cells, (nx, ny), lo, hi)


def rectangle_quad(lo, hi, nx, ny):
    xs = lo[0] + (hi[0] - lo[0]) * np.arange(nx + 1) / nx
    ys = lo[1] + (hi[1] - lo[1]) * np.arange(ny + 1) / ny
    X, Y = np.meshgrid(xs, ys, indexing='xy')
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing='xy')
    v0 = (iy * (nx + 1) + ix).ravel()
    cells = np.stack([v0, v0 + 1, v0 + nx + 1, v0 + nx + 2], axis=1)
    return Mesh('quadrilateral', coords, cells, (nx, ny), lo, hi)


def box_hex(lo, hi, nx, ny, nz):
    """dolfinx.mesh.create_box(..., CellType.hexahedron) in lattice order: vertex (ix,iy,iz) ->
    (iz*(ny+1)+iy)*(nx+1)+ix, cell (ix,iy,iz) -> (iz*ny+iy)*nx+ix, vertices in tensor order (x fastest)."""
    xs = lo[0] + (hi[0] - lo[0]) * np.arange(nx + 1) / nx
    ys = lo[1] + (hi[1] - lo[1]) * np.arange(ny + 1) / ny
    zs = lo[2] + (hi[2] - lo[2]) * np.arange(nz + 1) / nz
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing='ij')
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
    sx, sy = nx + 1, ny + 1
    v0 = ((iz * sy + iy) * sx + ix).ravel()
    off = np.array([(k & 1) + ((k >> 1) & 1) * sx + ((k >> 2) & 1) * sx * sy for k in range(8)])
    cells = v0[:, None] + off[None, :]
    return Mesh('hexahedron', coords, cells, (nx, ny, nz), lo, hi)


def interval(n, x0, x1):
    coords = (x0 + (x1 - x0) * np.arange(n + 1) / n).reshape(-1, 1)
    cells = np.stack([np.arange(n), np.arange(n) + 1], axis=1)
    return Mesh('interval', coords, cells, (n,), (x0,), (x1,))


# --------------------------------------------------------------------------
# dofmaps
# --------------------------------------------------------------------------
def triangle_edges(mesh):
    """Edges of a triangle mesh: (edge -> [min vertex, max vertex]) in lexicographic order of that pair, and
    cell -> edges with local edge i opposite vertex i (basix).  P2 dofs: vertices, then nverts + edge."""
    c = mesh.cells.astype(np.int64)
    a = np.stack([c[:, 1], c[:, 0], c[:, 0]], axis=1)
    b = np.stack([c[:, 2], c[:, 2], c[:, 1]], axis=1)
    key = np.minimum(a, b) * (mesh.nverts + 1) + np.maximum(a, b)
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    ev = np.stack([uniq // (mesh.nverts + 1), uniq % (mesh.nverts + 1)], axis=1).astype(np.int32)
    return ev, inv.reshape(-1, 3).astype(np.int32)


def dofmap(mesh, family, degree=1, block=1):
    """cell -> dof table, shape (ncells, ndofs_per_cell), and the space size.

    Blocked spaces interleave components per node (x0,y0,x1,y1,...), the layout
    the reference relies on at utils_dolfinx.py:635-639.  Hermite-3 on an
    interval carries (value, reference derivative) per vertex
    (examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py:218-222).
    """
    if family == 'DG' and degree == 0:
        base = np.arange(mesh.ncells, dtype=np.int32).reshape(-1, 1)
        n = mesh.ncells
    elif family in ('CG', 'Q') and degree == 1:
        base = mesh.cells
        n = mesh.nverts
    elif family == 'CG' and degree == 2:
        assert mesh.kind == 'triangle' and block == 1
        _, ce = triangle_edges(mesh)
        base = np.concatenate([mesh.cells, mesh.nverts + ce], axis=1)
        n = mesh.nverts + int(ce.max()) + 1
    elif family == 'Hermite' and degree == 3:
        assert mesh.kind == 'interval'
        base = mesh.cells
        n = mesh.nverts
        block = 2
    else:
        raise ValueError('oracle: unsupported space %s%d' % (family, degree))
    if block == 1:
        return base.astype(np.int32), n
    cd = (block * base[:, :, None] + np.arange(block)[None, None, :]).reshape(mesh.ncells, -1)
    return cd.astype(np.int32), n * block
