"""Layouts of families 3 and 4 vs the oracle (integer arrays ==), no GPU."""
import numpy as np
import pytest

from oracle import assembly as asm
from _cases34 import BeamCase, SimpCase, HexCase


@pytest.mark.parametrize('make', [lambda: BeamCase(50, upload=False), lambda: BeamCase(3, upload=False),
                                  lambda: SimpCase(8, 4, upload=False), lambda: SimpCase(80, 40, upload=False),
                                  lambda: HexCase(4, 3, 2, upload=False), lambda: HexCase(8, 4, 4, upload=False)])
def test_patterns_bit_exact(make):
    c = make()
    F, p = c.F, c.p
    assert p.N == F.N and p.M == [F.M]
    u, m = np.zeros(F.N), np.ones(F.M)
    for which, blocks, shape in ((0, F.jacobian(u, m), (F.N, F.N)), (1, F.dRdm(0, u, m), (F.N, F.M))):
        rp, col = p.pattern(which)
        orp, ocol = asm.pattern(blocks, shape)
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)


def test_reference_sizes():
    c = SimpCase(80, 40, upload=False)
    assert c.p.N == 6642 and c.p.pattern_info(0)['nnz'] == 116644
    assert len(c.tag) == 2                                               # two traction facets around y = 40
    b = BeamCase(50, upload=False)
    assert b.p.N == 102 and b.p.M == [50] and len(b.tag) == 1


def test_hex_mesh_and_sizes():
    """Hexahedral lattice: coordinates / connectivity / exterior facets equal the oracle's (generic facet search),
    81 entries in an interior row, and the node count of SURVEY.md section 8d's C4-3D formula."""
    c = HexCase(6, 4, 3, upload=False)
    assert np.array_equal(c.emesh.coords(), c.omesh.coords)
    assert np.array_equal(c.emesh.cells(), c.omesh.cells)
    fc, fl = c.omesh.exterior_facets()
    ec, el = c.emesh.exterior_facets()
    assert np.array_equal(ec, fc) and np.array_equal(el, fl)
    rp, _ = c.p.pattern(0)
    assert np.diff(rp).max() == 81 and c.p.N == 3 * 7 * 5 * 4
    assert (512 + 1) * (256 + 1) * (128 + 1) == 17007489
