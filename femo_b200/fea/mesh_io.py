"""Unstructured mesh ingest without dolfinx / meshio / h5py.

The reference imports meshes converted by msh2xdmf (`import_mesh`, femo/fea/utils_dolfinx.py:69-123: an XDMF
pair `<prefix>_domain.xdmf` / `<prefix>_boundaries.xdmf` plus `<prefix>_association_table.ini`).  `read_xdmf` reads that
pair with inline XML, raw binary or HDF5 data items (meshio's data_format "XML" / "Binary" / "HDF"; HDF5 through the
pure-Python reader hdf5_lite.py, no h5py needed); if the pair cannot be read `import_mesh` goes back to the Gmsh file the pair was made
from (`<prefix>.msh`, ASCII format 2.2 or 4.1 [upstream layouts, from memory]).  Either way it returns the reference's tuple.  Gmsh's vertex order is converted to the
tensor-product order basix uses for quadrilaterals and hexahedra.
"""
import os
from configparser import ConfigParser

import numpy as np

from .. import engine as _E

# gmsh element type -> (name, nodes, permutation to basix vertex order)
_GMSH = {1: ('line', 2, [0, 1]), 2: ('triangle', 3, [0, 1, 2]), 3: ('quadrilateral', 4, [0, 1, 3, 2]),
         5: ('hexahedron', 8, [0, 1, 3, 2, 4, 5, 7, 6]), 15: ('point', 1, [0])}
_TDIM = {'line': 1, 'triangle': 2, 'quadrilateral': 2, 'hexahedron': 3, 'point': 0}


def read_msh(path):
    """-> points (n,3), {cell type: (connectivity (m,k) 0-based, physical tag (m,))}, {physical name: (tag, dim)}."""
    with open(path) as fh:
        tok = fh.read().split('\n')
    sec = {}
    i = 0
    while i < len(tok):
        line = tok[i].strip()
        if line.startswith('$') and not line.startswith('$End'):
            name = line[1:]
            j = i + 1
            while tok[j].strip() != '$End' + name:
                j += 1
            sec[name] = tok[i + 1:j]
            i = j
        i += 1
    version = float(sec['MeshFormat'][0].split()[0])
    names = {}
    for ln in sec.get('PhysicalNames', [])[1:]:
        d, t, nm = ln.split(None, 2)
        names[nm.strip().strip('"')] = (int(t), int(d))
    blocks = {}
    if version < 4.0:
        n = int(sec['Nodes'][0])
        ids = np.empty(n, dtype=np.int64)
        pts = np.empty((n, 3))
        for k, ln in enumerate(sec['Nodes'][1:1 + n]):
            a = ln.split()
            ids[k] = int(a[0])
            pts[k] = [float(a[1]), float(a[2]), float(a[3])]
        for ln in sec['Elements'][1:1 + int(sec['Elements'][0])]:
            a = [int(v) for v in ln.split()]
            et, ntags = a[1], a[2]
            if et not in _GMSH:
                continue
            phys = a[3] if ntags > 0 else 0
            blocks.setdefault(et, []).append((phys, a[3 + ntags:]))
    else:
        # entities: physical tags per (dim, entity tag)
        ent_phys = {}
        E = sec.get('Entities')
        if E:
            counts = [int(v) for v in E[0].split()]
            row = 1
            for dim, cnt in enumerate(counts):
                for _ in range(cnt):
                    a = E[row].split()
                    row += 1
                    off = 4 if dim == 0 else 7
                    nph = int(a[off])
                    ent_phys[(dim, int(a[0]))] = [int(v) for v in a[off + 1:off + 1 + nph]]
        N = sec['Nodes']
        nb, n = int(N[0].split()[0]), int(N[0].split()[1])
        ids = np.empty(n, dtype=np.int64)
        pts = np.empty((n, 3))
        row, k = 1, 0
        for _ in range(nb):
            cnt = int(N[row].split()[3])
            row += 1
            for q in range(cnt):
                ids[k + q] = int(N[row + q])
            for q in range(cnt):
                pts[k + q] = [float(v) for v in N[row + cnt + q].split()[:3]]
            row += 2 * cnt
            k += cnt
        L = sec['Elements']
        nb = int(L[0].split()[0])
        row = 1
        for _ in range(nb):
            dim, etag, et, cnt = [int(v) for v in L[row].split()]
            row += 1
            ph = ent_phys.get((dim, etag), [0])
            phys = ph[0] if ph else 0
            if et in _GMSH:
                for q in range(cnt):
                    blocks.setdefault(et, []).append((phys, [int(v) for v in L[row + q].split()[1:]]))
            row += cnt
    remap = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
    remap[ids] = np.arange(ids.size)
    cells = {}
    for et, lst in blocks.items():
        name, nn, perm = _GMSH[et]
        conn = remap[np.array([c for _, c in lst], dtype=np.int64).reshape(-1, nn)][:, perm]
        cells[name] = (conn.astype(np.int32), np.array([p for p, _ in lst], dtype=np.int32))
    return pts, cells, names


_XDMF_TOPO = {'triangle': ('triangle', 3), 'polyline': ('line', 2), 'quadrilateral': ('quadrilateral', 4),
              'hexahedron': ('hexahedron', 8), 'polyvertex': ('point', 1)}
# XDMF / VTK vertex order -> basix tensor-product order
_XDMF_PERM = {'quadrilateral': [0, 1, 3, 2], 'hexahedron': [0, 1, 3, 2, 4, 5, 7, 6]}


def _xdmf_item(item, directory):
    fmt = item.get('Format', 'XML').upper()
    dims = [int(d) for d in item.get('Dimensions').split()]
    kind = item.get('DataType', item.get('NumberType', 'Float'))
    prec = int(item.get('Precision', '4'))
    dt = np.dtype({('Float', 4): '<f4', ('Float', 8): '<f8', ('Int', 4): '<i4', ('Int', 8): '<i8', ('UInt', 4): '<u4',
                   ('UInt', 8): '<u8'}[(kind, prec)])
    text = (item.text or '').strip()
    if fmt == 'XML':
        a = np.array(text.split(), dtype=np.float64 if kind == 'Float' else np.int64)
    elif fmt == 'BINARY':
        a = np.fromfile(os.path.join(directory, text), dtype=dt)
    elif fmt in ('HDF', 'H5'):                      # "file.h5:/path/to/dataset" (meshio: /data0 ...; dolfinx: /Mesh/<name>/geometry ...)
        fname, _, dset = text.partition(':')
        from . import hdf5_lite
        a = hdf5_lite.read_dataset(os.path.join(directory, fname.strip()), dset.strip())
    else:
        raise NotImplementedError('XDMF data item format %r (%s)' % (fmt, text))
    return a.reshape(dims)


def read_xdmf(path, attribute=None):
    """One <Grid> of an XDMF 3 file as meshio / msh2xdmf write it -> points (n, gdim), (cell type, connectivity in basix
    vertex order), cell attribute values (the first Cell-centred <Attribute>, or the one named `attribute`), or None."""
    import xml.etree.ElementTree as ET
    directory = os.path.dirname(os.path.abspath(path))
    grid = ET.parse(path).getroot().find('.//Grid')
    if grid is None:
        raise ValueError('%s: no <Grid>' % path)
    pts = _xdmf_item(grid.find('Geometry').find('DataItem'), directory).astype(np.float64)
    topo = grid.find('Topology')
    tname = (topo.get('TopologyType') or topo.get('Type')).lower()
    if tname not in _XDMF_TOPO:
        raise NotImplementedError('XDMF topology %r' % tname)
    kind, nn = _XDMF_TOPO[tname]
    conn = _xdmf_item(topo.find('DataItem'), directory).astype(np.int64).reshape(-1, nn)
    if kind in _XDMF_PERM:
        conn = conn[:, _XDMF_PERM[kind]]
    vals = None
    for att in grid.findall('Attribute'):
        if att.get('Center', 'Cell') == 'Cell' and (attribute is None or att.get('Name') == attribute):
            vals = _xdmf_item(att.find('DataItem'), directory).astype(np.int64).ravel()
            break
    return pts, (kind, conn), vals


class FacetTags:
    """Tagged facets of an imported mesh (`boundaries_mf` of the reference): each tagged line / face element is
    matched to the cell facets with the same vertex set.  `sides(tag)` lists them as one-sided (cell, local facet)
    pairs -- one pair for an exterior facet, two for an interior one (the "+" and "-" restrictions of dS)."""

    def __init__(self, mesh, conn, tags):
        from .fem import _local_facets
        self.mesh = mesh
        self.dim = mesh.topology.dim - 1
        lf = _local_facets(mesh)
        key = {}
        for c in range(mesh.num_cells):
            for l in range(lf.shape[0]):
                key.setdefault(tuple(sorted(mesh.cells[c][lf[l]])), []).append((c, l))
        self._sides = {}
        self.indices, self.values = [], []
        fc, fl = mesh._e.exterior_facets()
        ext = {(int(c), int(l)): k for k, (c, l) in enumerate(zip(fc, fl))}
        for verts, t in zip(conn, tags):
            hit = key.get(tuple(sorted(int(v) for v in verts)), [])
            self._sides.setdefault(int(t), []).extend(hit)
            for s in hit:
                if s in ext:                       # exterior facet: index into the exterior-facet list, as meshtags() uses
                    self.indices.append(ext[s])
                    self.values.append(int(t))
        self.indices = np.array(self.indices, dtype=np.int32)
        self.values = np.array(self.values, dtype=np.int32)

    def sides(self, tag):
        s = sorted(self._sides.get(int(tag), []))
        return (np.array([c for c, _ in s], dtype=np.int32), np.array([l for _, l in s], dtype=np.int32))


def import_mesh(prefix="mesh", subdomains=False, dim=2, directory="."):
    """utils_dolfinx.py:69-123 with the Gmsh file as the source.  Returns (mesh, boundaries_mf, association_table) or
    (mesh, boundaries_mf, subdomains_mf, association_table) exactly like the reference."""
    from .fem import Mesh, meshtags
    names = {}
    dom, bnd = (os.path.join(directory, '%s_%s.xdmf' % (prefix, w)) for w in ('domain', 'boundaries'))
    cells = None
    if os.path.exists(dom) and os.path.exists(bnd):
        try:                                           # the reference's own inputs, when their heavy data is readable
            pts, (kind, conn), ctags = read_xdmf(dom)
            _, (fk, fconn), ftags = read_xdmf(bnd)
            cells = {kind: (conn, np.zeros(conn.shape[0], dtype=np.int32) if ctags is None else ctags.astype(np.int32)),
                     fk: (fconn, np.zeros(fconn.shape[0], dtype=np.int32) if ftags is None else ftags.astype(np.int32))}
        except (NotImplementedError, ValueError, KeyError, OSError):        # e.g. git-LFS pointers instead of the HDF5 files
            if not os.path.exists(os.path.join(directory, prefix + '.msh')):
                raise
    if cells is None:
        pts, cells, names = read_msh(os.path.join(directory, prefix + '.msh'))
    kind = next(k for k in ('triangle', 'quadrilateral') if k in cells) if dim == 2 else 'hexahedron'
    conn, ctags = cells[kind]
    used = np.unique(conn)                          # drop geometry-only points (arc centres ...)
    remap = np.full(pts.shape[0], -1, dtype=np.int64)
    remap[used] = np.arange(used.size)
    mesh = Mesh(_E.EngineMesh.from_arrays(kind, pts[used][:, :dim], remap[conn]), kind)
    fkind = 'line' if dim == 2 else 'quadrilateral'
    fconn, ftags = cells.get(fkind, (np.zeros((0, 2 if dim == 2 else 4), dtype=np.int32), np.zeros(0, dtype=np.int32)))
    boundaries_mf = FacetTags(mesh, remap[fconn], ftags)
    table = {nm.lower(): tag for nm, (tag, _) in names.items()}
    ini = os.path.join(directory, '%s_association_table.ini' % prefix)
    if os.path.exists(ini):
        cp = ConfigParser()
        cp.read(ini)
        table = {k: int(v) for k, v in dict(cp["ASSOCIATION TABLE"]).items()}
    if not subdomains:
        return mesh, boundaries_mf, table
    subdomains_mf = meshtags(mesh, dim, np.arange(mesh.num_cells, dtype=np.int32), ctags)
    return mesh, boundaries_mf, subdomains_mf, table
