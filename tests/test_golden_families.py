"""Committed oracle dumps of the families beyond 1 and 2 (tests/golden/make_golden_families.py): the oracle must
reproduce them on the CPU, the CUDA path must match them on the GPU (patterns ==, values 1e-12 / motor 1e-10)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden_families as G  # noqa: E402
from _cases import relerr  # noqa: E402

NAMES = sorted(G.cases())


@pytest.mark.parametrize('name', NAMES)
def test_oracle_reproduces_family_fixtures(name):
    z = np.load(os.path.join(HERE, 'golden', 'case_%s.npz' % name))
    c = G.cases()[name](upload=False)
    o = G.oracle_arrays(c)
    assert np.array_equal(o['u'], z['u']) and np.array_equal(o['m'], z['m'])        # same seeded inputs
    for key in ('rowptr', 'col', 'd_rowptr', 'd_col'):
        assert np.array_equal(o[key], z[key]), key
    for key in ['R', 'J', 'D'] + ['out%d%s' % (k, s) for k in range(int(z['n_outputs'])) for s in ('', '_du', '_dm')]:
        assert np.allclose(o[key], z[key], rtol=1e-12, atol=1e-14 * max(1.0, float(np.abs(z[key]).max()))), key
    # and the engine's host-side layout equals the frozen pattern
    rp, col = c.p.pattern(0)
    assert np.array_equal(rp, z['rowptr']) and np.array_equal(col, z['col'])
    rp, col = c.p.pattern(1)
    assert np.array_equal(rp, z['d_rowptr']) and np.array_equal(col, z['d_col'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_gpu_matches_family_fixtures(cuda_device, name):
    z = np.load(os.path.join(HERE, 'golden', 'case_%s.npz' % name))
    c = G.cases()[name]()
    p = c.p
    tol = 1e-10 if name.startswith('motor') else 1e-12
    assert relerr(p.assemble_residual().cpu().numpy(), z['R']) < tol
    vals, _ = p.assemble_jacobian(plain=True, bc=False)
    assert relerr(vals.cpu().numpy(), z['J']) < tol
    assert relerr(p.assemble_dRdm(0).cpu().numpy(), z['D']) < tol
    for k in range(int(z['n_outputs'])):
        ref = float(z['out%d' % k])
        assert abs(p.assemble_output(k) - ref) <= tol * max(abs(ref), 1e-300)
        assert relerr(p.assemble_output_grad(k, 0).cpu().numpy(), z['out%d_du' % k]) < tol
        assert relerr(p.assemble_output_grad(k, 1).cpu().numpy(), z['out%d_dm' % k]) < tol
