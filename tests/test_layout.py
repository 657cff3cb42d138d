"""Host-side layout (meshes, dofmaps, CSR patterns, gather maps) of the engine
vs the oracle: every integer array compared with ==  (SURVEY.md section 8c)."""
import numpy as np
import pytest

from femo_b200 import engine as E
from oracle import mesh as om, families as fam, assembly as asm


@pytest.mark.parametrize('nx,ny', [(1, 1), (2, 3), (16, 16), (37, 5)])
def test_unit_square_mesh_bit_exact(nx, ny):
    em = E.EngineMesh.unit_square(nx, ny)
    m = om.unit_square_tri(nx, ny)
    assert np.array_equal(em.coords(), m.coords)
    assert np.array_equal(em.cells(), m.cells)
    c, l = em.exterior_facets()
    oc, ol = m.exterior_facets()
    assert np.array_equal(c, oc) and np.array_equal(l, ol)
    assert len(c) == 2 * (nx + ny)


def test_quad_and_interval_mesh_bit_exact():
    em = E.EngineMesh.rectangle_quad((0.0, 0.0), (160.0, 80.0), 8, 4)
    m = om.rectangle_quad((0.0, 0.0), (160.0, 80.0), 8, 4)
    assert np.array_equal(em.coords(), m.coords) and np.array_equal(em.cells(), m.cells)
    c, l = em.exterior_facets()
    oc, ol = m.exterior_facets()
    assert np.array_equal(c, oc) and np.array_equal(l, ol)
    ei = E.EngineMesh.interval(50, 0.0, 1.0)
    mi = om.interval(50, 0.0, 1.0)
    assert np.array_equal(ei.coords(), mi.coords) and np.array_equal(ei.cells(), mi.cells)
    c, l = ei.exterior_facets()
    oc, ol = mi.exterior_facets()
    assert np.array_equal(c, oc) and np.array_equal(l, ol)


def _oracle_gather(blocks, shape, rowptr, col):
    """Reference construction of the sorted gather map: contribution (block b,
    entity e, local a,c) lives at scratch index off_b + (a*nc+c)*ne_b + e."""
    keys, srcs = [], []
    off = 0
    for rd, cd, _ in blocks:
        ne, nr = rd.shape
        nc = cd.shape[1]
        e = np.arange(ne)[:, None, None]
        a = np.arange(nr)[None, :, None]
        c = np.arange(nc)[None, None, :]
        src = off + (a * nc + c) * ne + e
        key = rd[:, :, None].astype(np.int64) * shape[1] + cd[:, None, :]
        keys.append(np.broadcast_to(key, src.shape).ravel())
        srcs.append(src.ravel())
        off += ne * nr * nc
    keys, srcs = np.concatenate(keys), np.concatenate(srcs)
    order = np.lexsort((srcs, keys))
    keys, srcs = keys[order], srcs[order]
    rows = np.repeat(np.arange(shape[0]), np.diff(rowptr))
    nnz_key = rows.astype(np.int64) * shape[1] + col
    ptr = np.searchsorted(keys, nnz_key, side='left')
    ptr = np.append(ptr, keys.size)
    return ptr.astype(np.int32), srcs.astype(np.int32)


@pytest.mark.parametrize('n', [1, 3, 16])
@pytest.mark.parametrize('famid', [1, 2])
def test_patterns_and_gather_maps_bit_exact(n, famid):
    em = E.EngineMesh.unit_square(n)
    m = om.unit_square_tri(n)
    F = fam.PoissonP1(m) if famid == 1 else fam.NonlinearPoissonP1(m)
    p = E.EngineProblem(em, famid)
    assert p.N == F.N and p.M == [F.M]
    u, f = np.zeros(F.N), np.zeros(F.M)
    for which, blocks, shape in ((0, F.jacobian(u, f), (F.N, F.N)), (1, F.dRdm(0, u, f), (F.N, F.M))):
        rp, col = p.pattern(which)
        orp, ocol = asm.pattern(blocks, shape)
        assert np.array_equal(rp, orp)
        assert np.array_equal(col, ocol)
        gp, gs = p.gather_map(which)
        ogp, ogs = _oracle_gather(blocks, shape, orp, ocol)
        assert np.array_equal(gp, ogp)
        assert np.array_equal(gs, ogs)


def test_reference_default_sizes():
    # config 1 at n=16 -> 289 dofs, 1889 nnz, dRdf 289x512 with 1536 nnz
    p = E.EngineProblem(E.EngineMesh.unit_square(16), 1)
    assert p.N == 289 and p.M == [512]
    assert p.pattern_info(0)['nnz'] == 1889
    assert p.pattern_info(1)['nnz'] == 1536


def test_bad_arguments_fail_loudly():
    from femo_b200._lib import FemoError
    with pytest.raises(FemoError):
        E.EngineMesh.unit_square(0)
    with pytest.raises(FemoError):
        E.EngineProblem(E.EngineMesh.unit_square(2), 99)
    p = E.EngineProblem(E.EngineMesh.unit_square(2), 1)
    with pytest.raises(FemoError):
        p.set_bc([np.array([1000])])


def test_no_cpu_fallback():
    """Device entry points must refuse to run without an uploaded CUDA problem."""
    import ctypes as C
    from femo_b200._lib import lib, FemoError, check
    p = E.EngineProblem(E.EngineMesh.unit_square(2), 1)
    buf = np.zeros(p.N)
    with pytest.raises(FemoError) as ei:
        check(lib.femo_assemble_residual(p._h, buf.ctypes.data_as(C.c_void_p)))
    assert ei.value.code == -4
    if E.device_count() == 0:
        with pytest.raises(FemoError):
            p.upload(0)
