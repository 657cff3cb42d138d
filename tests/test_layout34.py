"""Layouts of families 3 and 4 vs the oracle (integer arrays ==), no GPU."""
import numpy as np
import pytest

from oracle import assembly as asm
from _cases34 import BeamCase, SimpCase


@pytest.mark.parametrize('make', [lambda: BeamCase(50, upload=False), lambda: BeamCase(3, upload=False),
                                  lambda: SimpCase(8, 4, upload=False), lambda: SimpCase(80, 40, upload=False)])
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
