"""Unstructured meshes: femo_mesh_create_from_arrays vs the oracle's generic topology (integer ==), the Gmsh
reader behind import_mesh (formats 2.2 and 4.1 written by this test), and -- on the GPU -- assembly parity
on a shuffled, perturbed triangle mesh (no lattice structure anywhere)."""
import numpy as np
import pytest

from femo_b200 import engine as E
from oracle import mesh as om, families as fam, assembly as asm
from _cases import relerr


def shuffled(nx=6, ny=5, seed=0, jitter=0.02):
    m = om.unit_square_tri(nx, ny)
    rng = np.random.default_rng(seed)
    vp, cp = rng.permutation(m.nverts), rng.permutation(m.ncells)
    coords = np.empty_like(m.coords)
    coords[vp] = m.coords
    interior = (coords[:, 0] > 1e-9) & (coords[:, 0] < 1 - 1e-9) & (coords[:, 1] > 1e-9) & (coords[:, 1] < 1 - 1e-9)
    coords[interior] += jitter * rng.standard_normal((interior.sum(), 2))
    cells = vp[m.cells][cp].astype(np.int32)
    return coords, cells, om.Mesh('triangle', coords, cells, (0, 0), (0.0, 0.0), (1.0, 1.0))


@pytest.mark.parametrize('kind', ['triangle', 'quadrilateral', 'hexahedron'])
def test_from_arrays_topology_bit_exact(kind):
    if kind == 'triangle':
        coords, cells, mo = shuffled()
    elif kind == 'quadrilateral':
        mo = om.rectangle_quad((0, 0), (2.0, 1.0), 5, 3)
        coords, cells = mo.coords, mo.cells
    else:
        mo = om.box_hex((0, 0, 0), (3.0, 2.0, 1.0), 3, 2, 2)
        coords, cells = mo.coords, mo.cells
    em = E.EngineMesh.from_arrays(kind, coords, cells)
    fc, fl = mo.exterior_facets()
    ec, el = em.exterior_facets()
    assert np.array_equal(ec, fc) and np.array_equal(el, fl)
    assert np.array_equal(em.coords(), coords) and np.array_equal(em.cells(), cells)
    if kind == 'triangle':
        for famid, F in ((2, fam.NonlinearPoissonP1(mo)), (E.FAMILY_NLPOISSON_P2, fam.NonlinearPoissonP2(mo))):
            p = E.EngineProblem(em, famid)
            rp, col = p.pattern(0)
            orp, ocol = asm.pattern(F.jacobian(np.zeros(F.N), np.zeros(F.M)), (F.N, F.N))
            assert np.array_equal(rp, orp) and np.array_equal(col, ocol)


def _write_msh2(path, pts, tris, tri_tags, lines, line_tags):
    with open(path, 'w') as f:
        f.write('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n3\n1 1000 "Interface"\n2 1 "Steel"\n2 3 "Magnet"\n$EndPhysicalNames\n')
        f.write('$Nodes\n%d\n' % len(pts))
        for i, x in enumerate(pts):
            f.write('%d %.17g %.17g 0\n' % (i + 1, x[0], x[1]))
        f.write('$EndNodes\n$Elements\n%d\n' % (len(tris) + len(lines)))
        k = 1
        for ln, t in zip(lines, line_tags):
            f.write('%d 1 2 %d 7 %d %d\n' % (k, t, ln[0] + 1, ln[1] + 1))
            k += 1
        for tr, t in zip(tris, tri_tags):
            f.write('%d 2 2 %d 9 %d %d %d\n' % (k, t, tr[0] + 1, tr[1] + 1, tr[2] + 1))
            k += 1
        f.write('$EndElements\n')


def _write_msh4(path, pts, tris, tri_tags, lines, line_tags):
    """One curve entity per line tag and one surface entity per cell tag (format 4.1, ASCII)."""
    lt, tt = sorted(set(line_tags)), sorted(set(tri_tags))
    with open(path, 'w') as f:
        f.write('$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Entities\n0 %d %d 0\n' % (len(lt), len(tt)))
        for e, t in enumerate(lt):
            f.write('%d 0 0 0 1 1 0 1 %d 0\n' % (e + 1, t))
        for e, t in enumerate(tt):
            f.write('%d 0 0 0 1 1 0 1 %d 0\n' % (e + 1, t))
        f.write('$EndEntities\n$Nodes\n1 %d 1 %d\n2 1 0 %d\n' % (len(pts), len(pts), len(pts)))
        for i in range(len(pts)):
            f.write('%d\n' % (i + 1))
        for x in pts:
            f.write('%.17g %.17g 0\n' % (x[0], x[1]))
        f.write('$EndNodes\n$Elements\n%d %d 1 %d\n' % (len(lt) + len(tt), len(tris) + len(lines), len(tris) + len(lines)))
        k = 1
        for e, t in enumerate(lt):
            sel = [ln for ln, q in zip(lines, line_tags) if q == t]
            f.write('1 %d 1 %d\n' % (e + 1, len(sel)))
            for ln in sel:
                f.write('%d %d %d\n' % (k, ln[0] + 1, ln[1] + 1))
                k += 1
        for e, t in enumerate(tt):
            sel = [tr for tr, q in zip(tris, tri_tags) if q == t]
            f.write('2 %d 2 %d\n' % (e + 1, len(sel)))
            for tr in sel:
                f.write('%d %d %d %d\n' % (k, tr[0] + 1, tr[1] + 1, tr[2] + 1))
                k += 1
        f.write('$EndElements\n')


@pytest.mark.parametrize('writer', [_write_msh2, _write_msh4])
def test_import_mesh_from_gmsh_file(tmp_path, writer):
    """import_mesh(prefix, subdomains=True) (utils_dolfinx.py:69-123): cell tags, an interior tagged line (both
    sides, as dS(1000) needs them) and the association table."""
    from femo_b200.fea.utils_b200 import import_mesh
    from femo_b200.fea.fem import Measure
    m = om.unit_square_tri(4, 4)
    cx = m.coords[m.cells].mean(axis=1)[:, 0]
    tri_tags = np.where(cx < 0.5, 1, 3)
    # the interior line x = 0.5 (tag 1000): vertical edges between the two materials
    col = np.nonzero(np.isclose(m.coords[:, 0], 0.5))[0]
    col = col[np.argsort(m.coords[col, 1])]
    lines = [(int(a), int(b)) for a, b in zip(col[:-1], col[1:])]
    writer(tmp_path / 'motor.msh', m.coords, m.cells, tri_tags, lines, [1000] * len(lines))
    (tmp_path / 'motor_association_table.ini').write_text('[ASSOCIATION TABLE]\ninterface = 1000\nsteel = 1\nmagnet = 3\n')
    mesh, boundaries_mf, subdomains_mf, table = import_mesh(prefix='motor', subdomains=True, dim=2, directory=str(tmp_path))
    assert table == dict(interface=1000, steel=1, magnet=3)
    assert mesh.num_cells == m.ncells and mesh.num_vertices == m.nverts
    # the reader may reorder cells (format 4.1 groups them by entity): compare through centroids
    cxi = mesh.geometry.x[mesh.cells].mean(axis=1)[:, 0]
    assert np.array_equal(subdomains_mf.values, np.where(cxi < 0.5, 1, 3))
    fc, fl = Measure('dS', domain=mesh, subdomain_data=boundaries_mf)(1000).sides
    assert len(fc) == 2 * len(lines)                       # both sides of every interior facet
    lf = np.array([[1, 2], [0, 2], [0, 1]])
    xs = mesh.geometry.x[mesh.cells[fc][np.arange(len(fc))[:, None], lf[fl]]][:, :, 0]
    assert np.allclose(xs, 0.5)


@pytest.mark.gpu
def test_gpu_assembly_on_unstructured_mesh(cuda_device):
    coords, cells, mo = shuffled(9, 7, seed=3)
    em = E.EngineMesh.from_arrays('triangle', coords, cells)
    rng = np.random.default_rng(1)
    for famid, F in ((2, fam.NonlinearPoissonP1(mo)), (E.FAMILY_NLPOISSON_P2, fam.NonlinearPoissonP2(mo))):
        p = E.EngineProblem(em, famid)
        p.upload(0)
        u, f = rng.standard_normal(F.N), rng.standard_normal(F.M)
        du, df = p.to_device(u), p.to_device(f)
        p.set_coefficient(0, du)
        p.set_coefficient(1, df)
        assert relerr(p.assemble_residual().cpu().numpy(), asm.assemble_vector(F.residual(u, f), F.N)) < 1e-12
        vals, _ = p.assemble_jacobian()
        assert relerr(vals.cpu().numpy(), asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), None).data) < 1e-12
        Jo = asm.assemble_scalar(F.output(0, u, f))
        assert abs(p.assemble_output(0) - Jo) < 1e-12 * abs(Jo)
        # Jacobi-CG solve (no lattice hierarchy on an unstructured mesh)
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rp, col = p.pattern(0)
        A = sp.csr_matrix((vals.cpu().numpy(), col, rp), shape=(F.N, F.N))
        b = rng.standard_normal(F.N)
        x, info = p.linear_solve(vals, p.to_device(b), rtol=1e-12, precond=0, max_it=20000, check_every=20)
        assert info['converged'] and relerr(x.cpu().numpy(), spla.spsolve(A.tocsc(), b)) < 1e-8


def test_gmsh_vertex_order_is_converted_to_tensor_order(tmp_path):
    """Gmsh lists quadrilateral (and hexahedron) vertices counter-clockwise; the engine and basix use tensor-product
    order.  A 2x1 quadrilateral mesh written in Gmsh order must come back with positive Jacobians everywhere."""
    from femo_b200.fea.mesh_io import read_msh
    pts = np.array([[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1]], dtype=float)
    quads_ccw = [(0, 1, 4, 3), (1, 2, 5, 4)]
    with open(tmp_path / 'q.msh', 'w') as f:
        f.write('$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n6\n')
        for i, x in enumerate(pts):
            f.write('%d %g %g 0\n' % (i + 1, x[0], x[1]))
        f.write('$EndNodes\n$Elements\n2\n')
        for k, q in enumerate(quads_ccw):
            f.write('%d 3 2 1 1 %d %d %d %d\n' % (k + 1, *(v + 1 for v in q)))
        f.write('$EndElements\n')
    p, cells, _ = read_msh(str(tmp_path / 'q.msh'))
    conn, tags = cells['quadrilateral']
    assert conn.tolist() == [[0, 1, 3, 4], [1, 2, 4, 5]] and tags.tolist() == [1, 1]
    X = p[conn][:, :, :2]
    e1, e2 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]                      # tensor order: v1 along xi, v2 along eta
    assert np.all(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0] > 0)
    em = E.EngineMesh.from_arrays('quadrilateral', p[:, :2], conn)
    assert em.nbfacets == 6


def _write_xdmf(path, pts, topo, conn, attr_name, values, binary_dir=None, hdf=None):
    """The XDMF 3 layout meshio / msh2xdmf write (one uniform Grid, Geometry XY, one cell attribute): inline XML, raw binary,
    or -- meshio's default -- HDF5 heavy data `<file>.h5:/data<k>` (hdf = keyword arguments of tests/_hdf5_writer.write)."""
    h5 = {}

    def item(a, kind, prec, name):
        dims = ' '.join(str(d) for d in a.shape)
        if hdf is not None:
            key = 'data%d' % len(h5)
            h5[key] = np.asarray(a, dtype='<f8' if kind == 'Float' else '<i8')
            return '<DataItem DataType="%s" Dimensions="%s" Format="HDF" Precision="%d">%s:/%s</DataItem>' % (
                kind, dims, prec, path.with_suffix('.h5').name, key)
        if binary_dir is None:
            body = '\n'.join(' '.join(repr(float(v)) if kind == 'Float' else str(int(v)) for v in row) for row in np.atleast_2d(a))
            return '<DataItem DataType="%s" Dimensions="%s" Format="XML" Precision="%d">\n%s\n</DataItem>' % (kind, dims, prec, body)
        fn = '%s_%s.bin' % (path.name, name)
        np.asarray(a, dtype='<f8' if kind == 'Float' else '<i8').tofile(binary_dir / fn)
        return '<DataItem DataType="%s" Dimensions="%s" Format="Binary" Precision="%d">%s</DataItem>' % (kind, dims, prec, fn)
    path.write_text('<Xdmf Version="3.0"><Domain><Grid Name="Grid">'
                    '<Geometry GeometryType="XY">%s</Geometry>'
                    '<Topology TopologyType="%s" NumberOfElements="%d" NodesPerElement="%d">%s</Topology>'
                    '<Attribute Name="%s" AttributeType="Scalar" Center="Cell">%s</Attribute>'
                    '</Grid></Domain></Xdmf>'
                    % (item(pts[:, :2], 'Float', 8, 'geo'), topo, conn.shape[0], conn.shape[1], item(conn, 'Int', 8, 'topo'),
                       attr_name, item(np.asarray(values).reshape(-1, 1), 'Int', 8, 'attr')))
    if hdf is not None:
        from _hdf5_writer import write
        write(str(path.with_suffix('.h5')), h5, **hdf)


@pytest.mark.parametrize('fmt', ['xml', 'binary', 'hdf', 'hdf-gzip'])
def test_import_mesh_from_the_xdmf_pair(tmp_path, fmt):
    """The reference's own inputs (utils_dolfinx.py:92-107: <prefix>_domain.xdmf with the cell tags, <prefix>_boundaries.xdmf
    with the tagged lines) with inline XML, raw binary or HDF5 heavy data (contiguous, and chunked + shuffled + gzip-compressed
    as meshio writes it), read without h5py; unreadable heavy data falls back to the .msh or raises."""
    from femo_b200.fea.utils_b200 import import_mesh
    from femo_b200.fea.fem import Measure
    m = om.unit_square_tri(4, 4)
    cx = m.coords[m.cells].mean(axis=1)[:, 0]
    tri_tags = np.where(cx < 0.5, 1, 3)
    col = np.nonzero(np.isclose(m.coords[:, 0], 0.5))[0]
    col = col[np.argsort(m.coords[col, 1])]
    lines = np.array([(int(a), int(b)) for a, b in zip(col[:-1], col[1:])])
    bd = tmp_path if fmt == 'binary' else None
    hdf = None if not fmt.startswith('hdf') else (dict() if fmt == 'hdf' else
                                                  dict(chunks=lambda a: (min(7, a.shape[0]), a.shape[1]), gzip=4, shuffle=True))
    _write_xdmf(tmp_path / 'motor_domain.xdmf', m.coords, 'Triangle', m.cells, 'name_to_read', tri_tags, bd, hdf)
    _write_xdmf(tmp_path / 'motor_boundaries.xdmf', m.coords, 'Polyline', lines, 'name_to_read', [1000] * len(lines), bd, hdf)
    (tmp_path / 'motor_association_table.ini').write_text('[ASSOCIATION TABLE]\ninterface = 1000\nsteel = 1\nmagnet = 3\n')
    mesh, boundaries_mf, subdomains_mf, table = import_mesh(prefix='motor', subdomains=True, dim=2, directory=str(tmp_path))
    assert table == dict(interface=1000, steel=1, magnet=3)
    assert np.array_equal(mesh.cells, m.cells) and np.array_equal(mesh.geometry.x[:, :2], m.coords[:, :2])
    assert np.array_equal(subdomains_mf.values, tri_tags)
    fc, fl = Measure('dS', domain=mesh, subdomain_data=boundaries_mf)(1000).sides
    assert len(fc) == 2 * len(lines)
    if fmt != 'hdf':
        return
    # what the reference checkout holds: git-LFS pointers in place of the .h5 files, and no .msh to fall back to
    for w in ('domain', 'boundaries'):
        (tmp_path / ('motor_%s.h5' % w)).write_text('version https://git-lfs.github.com/spec/v1\noid sha256:00\nsize 1\n')
    with pytest.raises(ValueError, match='not an HDF5 file'):
        import_mesh(prefix='motor', subdomains=True, dim=2, directory=str(tmp_path))
