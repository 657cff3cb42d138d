"""Pins the oracle (no GPU): quadrature exactness, hand-derived Gateaux
derivatives vs finite differences, analytic known answers of the reference's
examples, reference default sizes, and the committed golden
fixtures.  The reference has no tests / golden vectors and cannot run offline:
parity to dolfinx itself stays UNPINNED (oracle/__init__.py)."""
import glob
import os
from math import factorial

import numpy as np
import pytest

from oracle import mesh as om, families as fam, assembly as asm, solvers, quadrature as quad
from _cases import square_boundary_lists

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('deg', [1, 2, 4, 7, 12])
def test_triangle_rules_exact(deg):
    p, w = quad.triangle(deg)
    assert abs(w.sum() - 0.5) < 1e-15
    for a in range(deg + 1):
        for b in range(deg + 1 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert abs((w * p[:, 0] ** a * p[:, 1] ** b).sum() - exact) < 1e-14 * max(exact, 1e-3)


def test_interval_rule_exact():
    s, w = quad.interval(9)
    assert len(w) == 5
    for k in range(10):
        assert abs((w * s ** k).sum() - 1.0 / (k + 1)) < 1e-15


def _family(famid, n, ny=None):
    m = om.unit_square_tri(n, ny)
    if famid == 1:
        F = fam.PoissonP1(m)
        F.u_ex = np.cos(m.coords[:, 0]) * m.coords[:, 1]
    else:
        F = fam.NonlinearPoissonP1(m)
    return m, F


@pytest.mark.parametrize('famid', [1, 2])
def test_gateaux_derivatives_match_finite_differences(famid):
    """UFL derivative() stand-ins (utils_dolfinx.py:313-314) are hand-derived: check them."""
    m, F = _family(famid, 7, 5)
    rng = np.random.default_rng(3)
    u, f = rng.standard_normal(F.N), rng.standard_normal(F.M)
    du, df = rng.standard_normal(F.N), rng.standard_normal(F.M)
    R = lambda u_, f_: asm.assemble_vector(F.residual(u_, f_), F.N)
    J = lambda u_, f_: asm.assemble_scalar(F.output(0, u_, f_))
    h = 1e-6
    A = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N))
    fd = (R(u + h * du, f) - R(u - h * du, f)) / (2 * h)
    assert np.abs(A @ du - fd).max() < 1e-7 * np.abs(fd).max()
    D = asm.assemble_matrix(F.dRdm(0, u, f), (F.N, F.M))
    fd = (R(u, f + h * df) - R(u, f - h * df)) / (2 * h)
    assert np.abs(D @ df - fd).max() < 1e-7 * np.abs(fd).max()
    gu = asm.assemble_vector(F.output_du(0, u, f), F.N)
    assert abs(gu @ du - (J(u + h * du, f) - J(u - h * du, f)) / (2 * h)) < 1e-7 * abs(gu @ du)
    gm = asm.assemble_vector(F.output_dm(0, 0, u, f), F.M)
    hm = 10.0     # J is quadratic in f (central differences exact); dJ/df ~ alpha is tiny, use a long step
    assert abs(gm @ df - (J(u, f + hm * df) - J(u, f - hm * df)) / (2 * hm)) < 1e-6 * abs(gm @ df)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()          # both Jacobians are symmetric


def test_reference_default_sizes():
    m, F = _family(1, 16)
    u, f = np.zeros(F.N), np.zeros(F.M)
    rp, col = asm.pattern(F.jacobian(u, f), (F.N, F.N))
    assert F.N == 289 and F.M == 512 and col.size == 1889
    rp, col = asm.pattern(F.dRdm(0, u, f), (F.N, F.M))
    assert col.size == 1536 and np.all(np.bincount(col, minlength=F.M) == 3)


def test_poisson_closed_form_fields():
    """examples/poisson_opt/run_poisson_opt.py:78-92: with f = Expression_f the state
    converges to 1/(1+4 alpha pi^4) * Expression_u at O(h^2)."""
    errs = []
    for n in (8, 16, 32):
        m, F = _family(1, n)
        x = m.coords
        c = 1.0 / (1.0 + 1e-6 * 4 * np.pi ** 4)
        uex = c / (2 * np.pi ** 2) * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])
        cent = m.coords[m.cells].mean(axis=1)
        f = c * np.sin(np.pi * cent[:, 0]) * np.sin(np.pi * cent[:, 1])
        bc = asm.DirichletBC(F.N, square_boundary_lists(x), 0.0)
        u, info = solvers.StatePath(F, bc).solve_newton(np.zeros(F.N), [f])
        assert info['iterations'] == 3                                  # quirk B1
        errs.append(np.abs(u - uex).max())
    assert errs[0] / errs[1] > 3.0 and errs[1] / errs[2] > 3.0


def test_nonlinear_poisson_manufactured_solution():
    """run_nonlinear_poisson_opt.py:144-145,167: u_ex = sin(2 pi x) sin(pi y),
    f_ex = -div grad u_ex + u_ex^3; the Nitsche-SNES state converges at O(h^2)."""
    errs = []
    for n in (8, 16, 32):
        m, F = _family(2, n)
        cent = m.coords[m.cells].mean(axis=1)
        ue = fam.u_exact_nlp(cent)
        f = 5 * np.pi ** 2 * ue + ue ** 3
        u, info = solvers.StatePath(F, None).solve_snes(np.zeros(F.N), [f])
        assert info['reason'] in ('ABS', 'REL', 'STOL')
        errs.append(np.abs(u - fam.u_exact_nlp(m.coords)).max())
    assert errs[0] / errs[1] > 3.0 and errs[1] / errs[2] > 3.0


def test_bc_semantics():
    """rows+columns zeroed, diagonal once per dirichletbc object,
    lifting sign, set_bc."""
    m, F = _family(1, 4)
    rng = np.random.default_rng(0)
    u, f = rng.standard_normal(F.N), rng.standard_normal(F.M)
    g = rng.standard_normal(F.N)
    bc = asm.DirichletBC(F.N, square_boundary_lists(m.coords), g)
    A = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), None).toarray()
    Abc = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), bc).toarray()
    free = ~bc.marker
    assert np.allclose(Abc[np.ix_(free, free)], A[np.ix_(free, free)], atol=1e-14)
    assert np.all(Abc[bc.marker][:, free] == 0) and np.all(Abc[free][:, bc.marker] == 0)
    assert np.array_equal(np.diag(Abc)[bc.marker], bc.count[bc.marker].astype(float))
    assert Abc[0, 0] == 2.0                                              # corner dof listed by two bc objects
    b = solvers.StatePath(F, bc).newton_F(u, [f])
    R = asm.assemble_vector(F.residual(u, f), F.N)
    d = np.where(bc.marker, g - u, 0.0)
    expect = R + A @ d
    expect[bc.marker] = (u - g)[bc.marker]
    assert np.allclose(b, expect, rtol=1e-13, atol=1e-13)


def test_total_derivative_vs_finite_differences():
    """Adjoint totals: the BC-consistent variant matches FD; the reference-faithful
    chain (quirk B2) deviates at the Dirichlet rows."""
    m, F = _family(1, 8)
    x = m.coords
    F.u_ex = 1.0 / (2 * np.pi ** 2) * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])
    bc = asm.DirichletBC(F.N, square_boundary_lists(x), 0.0)
    sp = solvers.StatePath(F, bc)
    rng = np.random.default_rng(1)
    f, d = rng.standard_normal(F.M), rng.standard_normal(F.M)

    def Jof(f_):
        u_, _ = sp.solve_newton(np.zeros(F.N), [f_])
        return asm.assemble_scalar(F.output(0, u_, f_))
    u, _ = sp.solve_newton(np.zeros(F.N), [f])
    h = 1e-5
    fd = (Jof(f + h * d) - Jof(f - h * d)) / (2 * h)
    (gq,), _ = sp.total_derivative(0, u, [f])
    (gc,), _ = sp.total_derivative(0, u, [f], consistent_bc=True)
    assert abs(gc @ d - fd) < 1e-7 * abs(fd)
    assert abs(gq @ d - fd) > 1e-5 * abs(fd)


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLD, 'family*_n*.npz'))))
def test_oracle_reproduces_golden_fixtures(path):
    z = np.load(path)
    famid = int(os.path.basename(path)[6])
    n = int(os.path.basename(path).split('_n')[1].split('.')[0])
    m = om.unit_square_tri(n)
    if famid == 1:
        F = fam.PoissonP1(m)
        x = m.coords
        F.u_ex = 1.0 / (2 * np.pi ** 2) * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])
        bc = asm.DirichletBC(F.N, square_boundary_lists(x), 0.0)
    else:
        F, bc = fam.NonlinearPoissonP1(m), None
    u, f = z['u'], z['f']
    A = asm.assemble_matrix(F.jacobian(u, f), (F.N, F.N), bc)
    assert np.array_equal(A.indptr, z['rowptr']) and np.array_equal(A.indices, z['col'])
    assert np.allclose(A.data, z['Jbc'], rtol=1e-13, atol=1e-15)
    assert np.allclose(asm.assemble_vector(F.residual(u, f), F.N), z['R'], rtol=1e-13, atol=1e-15)
    assert np.isclose(asm.assemble_scalar(F.output(0, u, f)), float(z['out']), rtol=1e-13)


def test_hex_oracle_patch_test_and_rigid_modes():
    """Pins for the hexahedral SIMP oracle (no reference fixture exists): rigid-body motions carry no stress, a
    linear displacement field gives constant stress and therefore zero interior nodal forces (patch test), and
    the element matrices are symmetric positive semi-definite with exactly 6 zero modes."""
    m = om.box_hex((0.0, 0.0, 0.0), (3.0, 2.0, 1.5), 4, 3, 2)
    fc, fl = m.exterior_facets()
    F = fam.SimpHex8(m, np.nonzero(fl == 3)[0])
    rho = np.full(F.M, 0.7)
    X = m.coords
    A = asm.assemble_matrix(F.jacobian(np.zeros(F.N), rho), (F.N, F.N), None)
    for u in (np.tile([1.0, -2.0, 0.5], m.nverts),                                   # translation
              np.stack([-X[:, 1], X[:, 0], 0 * X[:, 0]], axis=1).ravel()):           # rotation about z
        assert np.abs(A @ u).max() < 1e-12 * np.abs(A.data).max() * np.abs(u).max()
    G = np.array([[0.3, -0.1, 0.2], [0.05, 0.4, -0.3], [0.1, 0.2, -0.25]])
    u = (X @ G.T).ravel()
    r = A @ u
    on_boundary = np.zeros(m.nverts, dtype=bool)
    on_boundary[np.unique(m.cells[fc][np.arange(fc.size)[:, None], m.local_facets[fl]])] = True
    assert np.abs(r.reshape(-1, 3)[~on_boundary]).max() < 1e-12 * np.abs(r).max()
    K = F._khat()[0]
    assert np.allclose(K, K.T, atol=1e-14)
    w = np.linalg.eigvalsh(K)
    assert (np.abs(w) < 1e-12 * w.max()).sum() == 6 and w.min() > -1e-12 * w.max()
    # total traction equals f times the tagged area
    Fe = asm.assemble_vector([(F.fdofs, None, F._traction())], F.N).reshape(-1, 3).sum(axis=0)
    area = 2.0 * 1.5
    assert np.allclose(Fe, np.array([0.0, -0.25, 0.0]) * area, atol=1e-13)


@pytest.mark.parametrize('Family', [fam.NonlinearPoissonP1, fam.NonlinearPoissonP2])
def test_nonlinear_poisson_oracles_converge_to_the_manufactured_solution(Family):
    """The example's manufactured solution u_ex = sin(2 pi x) sin(pi y) (run_nonlinear_poisson_opt.py:144-145, 165-169) with
    f = -lap(u_ex) + u_ex^3 averaged per cell: the Nitsche-SNES state must converge to u_ex at second order in L2 for P1
    and P2 alike (the DG0 source limits both).  Pins assembly + boundary terms + solver of both oracles at once."""
    errs = []
    for n in (8, 16, 32):
        m = om.unit_square_tri(n)
        F = Family(m)
        pts, w = quad.triangle(6)
        xq = np.einsum('qa,cad->cqd', np.stack([1 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1), m.coords[m.cells])
        ue = np.sin(2 * np.pi * xq[..., 0]) * np.sin(np.pi * xq[..., 1])
        fq = 5 * np.pi ** 2 * ue + ue ** 3                       # -lap(u_ex) + u_ex^3
        f = (fq * w[None, :]).sum(axis=1) / w.sum()
        u, info = solvers.StatePath(F, None).solve_snes(np.zeros(F.N), [f])
        J = asm.assemble_scalar(F.output(0, u, f))
        reg = 0.5 * F.alpha * float((0.5 * F.detJ * f * f).sum())
        errs.append(np.sqrt(2.0 * (J - reg)))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert errs[-1] < 5e-3 and min(rates) > 1.8, (errs, rates)
