"""Pins the oracle's element arithmetic against an INDEPENDENT symbolic restatement of the reference's weak forms (no GPU).

The reference writes its residuals and functionals in UFL (file:line cited per test); dolfinx integrates them with a rule
that is exact for the estimated polynomial degree, so for polynomial integrands its assembled
values equal the EXACT integrals to round-off.  Here the same forms are typed into sympy on a small distorted mesh with
rational vertex coordinates and rational dof values, integrated exactly (monomial formula on the reference simplex /
iterated integrals on boxes), assembled into GLOBAL vectors, and the Gateaux derivatives (`ufl.derivative`,
utils_dolfinx.py:313-314) are taken by `sympy.diff` of the global expressions with respect to the global dof symbols.
Nothing of oracle/ is used on the symbolic side: basis functions, geometry, exterior facets, normals and cell diameters
are re-derived here.  What is compared: oracle R, dR/du, dR/dm, J, dJ/du, dJ/dm (assembled) against the exact values,
1e-13 relative to the largest entry -- the hand-derived derivative blocks, quadrature degrees, scatter and sign
conventions of oracle/families.py all have to be right for that.

The non-polynomial data of config 2 (u_ex = sin 2 pi x sin pi y) is swapped for a cubic polynomial in the exact tests; a
separate test keeps the true u_ex and integrates it with mpmath to show the size of the quadrature error the documented
rules leave on a coarse mesh ('non-polynomial integrands')."""
import os
from fractions import Fraction

import numpy as np
import pytest

sp = pytest.importorskip('sympy')

from oracle import mesh as om, families as fam, assembly as asm   # noqa: E402

XI, ETA, S = sp.symbols('xi eta s')


# ------------------------------------------------------------------ helpers
def _rat(a):
    """numpy floats that are exact dyadic rationals -> sympy Rationals (same numbers on both sides of the comparison)."""
    fr = Fraction(float(a))
    assert fr.denominator <= 1 << 20, 'test data must be dyadic'
    return sp.Rational(fr.numerator, fr.denominator)


def _poly_terms(expr, lead):
    """Terms of a polynomial expression as (exponents of the leading generators, coefficient expression).  `sympy.poly`
    builds the polynomial with sparse rational arithmetic instead of expanding the expression tree (the cubic term of
    the P2 family has ~10^5 monomials in the dof symbols)."""
    gens = list(lead) + sorted(expr.free_symbols - set(lead), key=str)
    p = sp.poly(expr, *gens, domain='QQ')
    k = len(lead)
    out = {}
    for mon, c in p.terms():
        rest = sp.Mul(*[g ** e for g, e in zip(gens[k:], mon[k:]) if e])
        out[mon[:k]] = out.get(mon[:k], 0) + c * rest
    return out.items()


def _tri_int(expr):
    """Exact integral of a polynomial in (xi, eta) over the reference triangle: int xi^a eta^b = a! b! / (a+b+2)!."""
    expr = sp.sympify(expr)
    return sp.Add(*[c * sp.Rational(int(sp.factorial(a) * sp.factorial(b)), int(sp.factorial(a + b + 2)))
                    for (a, b), c in _poly_terms(expr, (XI, ETA))])


def _seg_int(expr):
    return sp.Add(*[c / (k + 1) for (k,), c in _poly_terms(sp.sympify(expr), (S,))])


def _distorted_square(n):
    """n x n right-diagonal triangle lattice whose interior vertices are moved by dyadic offsets (general affine cells)."""
    m = om.unit_square_tri(n)
    rng = np.random.default_rng(11)
    X = m.coords.copy()
    inner = (X[:, 0] > 0) & (X[:, 0] < 1) & (X[:, 1] > 0) & (X[:, 1] < 1)
    X[inner] += rng.integers(-8, 9, size=(int(inner.sum()), 2)) / 128.0
    # stretch the whole mesh so that boundary facets are not axis-aligned unit fractions only
    X = X @ np.array([[1.0, 0.25], [0.125, 0.75]])
    return om.Mesh('triangle', X, m.cells, m.shape, m.lo, m.hi)


def _dyadic(rng, n, scale=8):
    return rng.integers(-2 * scale, 2 * scale + 1, size=n) / float(scale)


def _dyadic_small(rng, n, bound):
    """Dyadic values in [-bound, bound] on a 1/1024 grid (mesh displacements that keep det F well inside (0, 2))."""
    k = int(round(bound * 1024))
    return rng.integers(-k, k + 1, size=n) / 1024.0


class _SymTri:
    """P1 / P2 triangle mesh in sympy: per cell the affine map, physical gradients and |det J|; exterior facets from
    edge counts; everything rational."""

    def __init__(self, mesh):
        self.cells = [tuple(int(v) for v in c) for c in mesh.cells]
        self.X = [sp.Matrix([_rat(x), _rat(y)]) for x, y in mesh.coords]
        self.nv = len(self.X)
        cnt = {}
        for c in self.cells:
            for a in range(3):
                e = tuple(sorted((c[a], c[(a + 1) % 3])))
                cnt[e] = cnt.get(e, 0) + 1
        self.boundary_edges = {e for e, k in cnt.items() if k == 1}
        self.edges = sorted(cnt)                         # lexicographic (min, max): the P2 edge numbering of DESIGN.md

    def geom(self, c):
        v = self.cells[c]
        X0, X1, X2 = (self.X[i] for i in v)
        J = sp.Matrix.hstack(X1 - X0, X2 - X0)
        det = J.det()
        JinvT = J.inv().T
        x = X0 + J * sp.Matrix([XI, ETA])
        lam = [1 - XI - ETA, XI, ETA]
        return v, x, lam, JinvT, sp.Abs(det)

    @staticmethod
    def grad(expr, JinvT):
        return JinvT * sp.Matrix([sp.diff(expr, XI), sp.diff(expr, ETA)])

    def facets(self, c):
        """Exterior facets of cell c: (local vertex pair, opposite local vertex)."""
        v = self.cells[c]
        out = []
        for o in range(3):
            a, b = [k for k in range(3) if k != o]
            if tuple(sorted((v[a], v[b]))) in self.boundary_edges:
                out.append((a, b, o))
        return out

    def diameter(self, c):
        v = self.cells[c]
        d2 = [(self.X[v[a]] - self.X[v[b]]).dot(self.X[v[a]] - self.X[v[b]]) for a, b in ((0, 1), (0, 2), (1, 2))]
        return sp.sqrt(max(d2))


def _close(got, exact, tol=1e-13):
    got = np.asarray(got, dtype=np.float64)
    ex = np.array(sp.Matrix(exact).evalf(30), dtype=np.float64).reshape(got.shape)
    scale = max(np.abs(ex).max(), 1e-300)
    assert np.abs(got - ex).max() <= tol * scale, (np.abs(got - ex).max() / scale)
    return ex


def _jac_at(exprs, syms, table):
    """Jacobian of `exprs` with respect to `syms` at the rational point `table` for expressions too large for `sympy.diff`
    (the rational functions of the mesh-motion family): the exact expressions are evaluated with 80-digit mpmath arithmetic
    and differentiated by central differences of step 1e-25 -- truncation error ~1e-50 x the third derivative, twenty
    orders below the comparison tolerance."""
    import mpmath as mp
    keys = list(table)
    f = sp.lambdify(keys, list(exprs), 'mpmath', cse=True)
    with mp.workdps(80):
        x0 = [mp.mpf(int(table[k].p)) / mp.mpf(int(table[k].q)) for k in keys]
        h = mp.mpf(10) ** -25
        cols = []
        for s_ in syms:
            j = keys.index(s_)
            xp, xm = list(x0), list(x0)
            xp[j] += h
            xm[j] -= h
            fp, fm = f(*xp), f(*xm)
            cols.append([sp.Float(str((a_ - b_) / (2 * h)), 40) for a_, b_ in zip(fp, fm)])
    return sp.Matrix(cols).T


def _subs_all(exprs, table):
    return [sp.sympify(e).xreplace(table) for e in exprs]


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'symbolic')


def _record(name, mesh, state, inputs, exact, meta):
    """The exact values are also the committed fixtures tests/golden/symbolic/<name>.npz (consumed by the GPU parity test
    tests/test_gpu_exact.py, which cannot run sympy-sized jobs per launch and must not need /root/reference):
    FEMO_SYMBOLIC_DUMP=1 rewrites them (scripts/make_symbolic_golden.py); otherwise the values computed now must
    reproduce the committed file, so a fixture can never drift from the derivation in this module."""
    rec = dict(kind=np.array(mesh.kind), coords=mesh.coords, cells=mesh.cells, state=np.asarray(state, dtype=np.float64))
    for k, a in enumerate(inputs):
        rec['input%d' % k] = np.asarray(a, dtype=np.float64)
    for k, v in (meta or {}).items():
        rec['meta_' + k] = np.asarray(v)
    rec.update(exact)
    path = os.path.join(GOLD, name + '.npz')
    if os.environ.get('FEMO_SYMBOLIC_DUMP'):
        os.makedirs(GOLD, exist_ok=True)
        np.savez_compressed(path, **rec)
        return
    assert os.path.exists(path), 'missing fixture %s: run scripts/make_symbolic_golden.py' % path
    with np.load(path) as z:
        assert sorted(z.files) == sorted(rec)
        for k in rec:
            a, b = np.asarray(rec[k]), z[k]
            if a.dtype.kind == 'f':
                assert a.shape == b.shape and np.abs(a - b).max(initial=0.0) <= 1e-14 * max(np.abs(b).max(initial=0.0), 1e-300), k
            else:
                assert np.array_equal(a, b), k


def _check(name, mesh, F, state, inputs, R, Js, Usym, Msyms, tab, tol=1e-13, numeric_jac=False, meta=None):
    """Oracle family F at (state, inputs) against the exact residual expressions R (one per dof), functionals Js and their
    derivatives with respect to the state symbols Usym and the input symbols Msyms[slot]; records the exact values."""
    if numeric_jac:
        jac = lambda ex, sy: _jac_at(ex, sy, tab)                                     # noqa: E731
    else:
        jac = lambda ex, sy: sp.Matrix([sp.sympify(e) for e in ex]).jacobian(list(sy)).xreplace(tab)   # noqa: E731
    N = F.N
    ex = {}
    ex['R'] = _close(asm.assemble_vector(F.residual(state, *inputs), N), _subs_all(R, tab), tol)
    ex['A'] = _close(asm.assemble_matrix(F.jacobian(state, *inputs), (N, N)).toarray(), jac(R, Usym), tol)
    for s_, Ms in enumerate(Msyms):
        ex['D%d' % s_] = _close(asm.assemble_matrix(F.dRdm(s_, state, *inputs), (N, len(Ms))).toarray(), jac(R, Ms), tol)
    for k, Jf in enumerate(Js):
        ex['J%d' % k] = _close([asm.assemble_scalar(F.output(k, state, *inputs))], [sp.sympify(Jf).xreplace(tab)], tol)
        ex['Ju%d' % k] = _close(asm.assemble_vector(F.output_du(k, state, *inputs), N), list(jac([Jf], Usym)), tol)
        for s_, Ms in enumerate(Msyms):
            ex['Jm%d_%d' % (k, s_)] = _close(asm.assemble_vector(F.output_dm(k, s_, state, *inputs), len(Ms)),
                                             list(jac([Jf], Ms)), tol)
    _record(name, mesh, state, inputs, ex, meta)


# ------------------------------------------------------------------ config 1
def test_poisson_p1_against_exact_integrals():
    """examples/poisson_opt/run_poisson_opt.py:32-38 (R = inner(grad u, grad v) dx - f v dx) and :74-76
    (J = 1/2 (u - u_ex)^2 dx + alpha/2 f^2 dx, u_ex a P1 function)."""
    m = _distorted_square(2)
    F = fam.PoissonP1(m)
    rng = np.random.default_rng(5)
    u, f, uex = _dyadic(rng, F.N), _dyadic(rng, F.M), _dyadic(rng, F.N)
    F.u_ex = uex
    alpha = sp.Rational(1, 10 ** 6)
    T = _SymTri(m)
    U = sp.symbols('U0:%d' % F.N)
    Fm = sp.symbols('F0:%d' % F.M)
    R = [0] * F.N
    Jf = 0
    for c in range(len(T.cells)):
        v, x, lam, JinvT, adet = T.geom(c)
        uh = sum(U[v[a]] * lam[a] for a in range(3))
        ue = sum(_rat(uex[v[a]]) * lam[a] for a in range(3))
        gu = T.grad(uh, JinvT)
        for a in range(3):
            R[v[a]] += _tri_int((gu.dot(T.grad(lam[a], JinvT)) - Fm[c] * lam[a]) * adet)
        Jf += _tri_int((sp.Rational(1, 2) * (uh - ue) ** 2 + alpha / 2 * Fm[c] ** 2) * adet)
    tab = {U[i]: _rat(u[i]) for i in range(F.N)}
    tab.update({Fm[i]: _rat(f[i]) for i in range(F.M)})
    _check('poisson_p1', m, F, u, [f], R, [Jf], U, [Fm], tab, meta=dict(u_ex=uex, alpha=1e-6))


# ------------------------------------------------------------------ config 2
def _uex_poly(x, y):
    return x * x * y - 3 * x * y + y * y * y / 4 + sp.Rational(1, 2)


def _nlp_symbolic(T, U, Fm, basis, uex, beta, alpha):
    """examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py:88-95 (interior), :97-116 with sym=True (Nitsche:
    -(grad u . n) v + (u_ex - u)(grad v . n) + beta / h_E (u - u_ex) v on ds), :140-142 (output).  `basis(c)` returns
    the cell's global dofs and its shape functions in (xi, eta)."""
    N = len(U)
    R = [0] * N
    Jf = 0
    for c in range(len(T.cells)):
        v, x, lam, JinvT, adet = T.geom(c)
        dofs, phi = basis(c, lam)
        uh = sum(U[d] * p for d, p in zip(dofs, phi))
        gu = T.grad(uh, JinvT)
        ue = uex(x[0], x[1])
        for d, p in zip(dofs, phi):
            R[d] += _tri_int((gu.dot(T.grad(p, JinvT)) + uh ** 3 * p - Fm[c] * p) * adet)
        Jf += _tri_int((sp.Rational(1, 2) * (uh - ue) ** 2 + alpha / 2 * Fm[c] ** 2) * adet)
        for a, b, o in T.facets(c):
            P, Q, O = T.X[v[a]], T.X[v[b]], T.X[v[o]]
            t = Q - P
            length = sp.sqrt(t.dot(t))
            nl = sp.Matrix([t[1], -t[0]])                 # outward normal x facet length (rational)
            if nl.dot(P - O) < 0:
                nl = -nl
            # restriction to the facet, parametrised from local vertex a (s = 0) to local vertex b (s = 1)
            ref = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
            xa, xb = ref[a], ref[b]
            on = {XI: xa[0] * (1 - S) + xb[0] * S, ETA: xa[1] * (1 - S) + xb[1] * S}
            uf = uh.xreplace(on)
            dudn_l = gu.xreplace(on).dot(nl)
            xs = P + t * S
            uef = uex(xs[0], xs[1])
            h = T.diameter(c)
            for d, p in zip(dofs, phi):
                pf = p.xreplace(on)
                dpdn_l = T.grad(p, JinvT).xreplace(on).dot(nl)
                # ds = length ds_ref; the irrational constants (length, h_E) stay outside the polynomial integrals
                R[d] += _seg_int(-dudn_l * pf + (uef - uf) * dpdn_l) + beta / h * length * _seg_int((uf - uef) * pf)
    return R, Jf


def _p1_basis(T):
    return lambda c, lam: (T.cells[c], lam)


def _compare_family(name, m, F, R, Jf, U, Fm, u, f, tol=1e-13):
    tab = {U[i]: _rat(u[i]) for i in range(F.N)}
    tab.update({Fm[i]: _rat(f[i]) for i in range(F.M)})
    # the Jacobian and dR/df do not involve u_ex, so the recorded A and D0 are exact for the TRUE family as well
    _check(name, m, F, u, [f], R, [Jf], U, [Fm], tab, tol, meta=dict(u_ex=np.array('x^2 y - 3 x y + y^3 / 4 + 1 / 2')))


def test_nonlinear_poisson_p1_against_exact_integrals(monkeypatch):
    m = _distorted_square(2)
    monkeypatch.setattr(fam, 'u_exact_nlp', lambda x: np.asarray(
        x[..., 0] ** 2 * x[..., 1] - 3 * x[..., 0] * x[..., 1] + x[..., 1] ** 3 / 4 + 0.5))
    F = fam.NonlinearPoissonP1(m)
    rng = np.random.default_rng(6)
    u, f = _dyadic(rng, F.N), _dyadic(rng, F.M)
    T = _SymTri(m)
    U = sp.symbols('U0:%d' % F.N)
    Fm = sp.symbols('F0:%d' % F.M)
    R, Jf = _nlp_symbolic(T, U, Fm, _p1_basis(T), _uex_poly, 10, sp.Rational(6, 10 ** 7))
    _compare_family('nlpoisson_p1', m, F, R, Jf, U, Fm, u, f)


def test_nonlinear_poisson_p2_against_exact_integrals(monkeypatch):
    """The synthetic P2 extension of config 2: same forms, quadratic Lagrange basis; dofs = vertices
    then edges in lexicographic (min vertex, max vertex) order, local edge i opposite local vertex i."""
    m = _distorted_square(1)          # two general affine cells, 9 dofs (the cubic term makes larger meshes slow)
    monkeypatch.setattr(fam, 'u_exact_nlp', lambda x: np.asarray(
        x[..., 0] ** 2 * x[..., 1] - 3 * x[..., 0] * x[..., 1] + x[..., 1] ** 3 / 4 + 0.5))
    F = fam.NonlinearPoissonP2(m)
    rng = np.random.default_rng(7)
    u, f = _dyadic(rng, F.N), _dyadic(rng, F.M)
    T = _SymTri(m)
    eid = {e: k for k, e in enumerate(T.edges)}
    assert F.N == T.nv + len(T.edges)

    def basis(c, lam):
        v = T.cells[c]
        dofs = list(v) + [T.nv + eid[tuple(sorted((v[(i + 1) % 3], v[(i + 2) % 3])))] for i in range(3)]
        phi = [l * (2 * l - 1) for l in lam] + [4 * lam[(i + 1) % 3] * lam[(i + 2) % 3] for i in range(3)]
        return dofs, phi

    U = sp.symbols('U0:%d' % F.N)
    Fm = sp.symbols('F0:%d' % F.M)
    R, Jf = _nlp_symbolic(T, U, Fm, basis, _uex_poly, 10, sp.Rational(6, 10 ** 7))
    _compare_family('nlpoisson_p2', m, F, R, Jf, U, Fm, u, f, tol=2e-13)


def test_nonlinear_poisson_true_u_ex_within_quadrature_error():
    """The same comparison with the reference's own u_ex = sin(2 pi x) sin(pi y) (run_nonlinear_poisson_opt.py:144-145):
    the facet and output integrands are no longer polynomial, the oracle's documented rules (5-point Gauss on facets, the
    degree-12 rule in cells) leave a quadrature error -- measured here against 30-digit mpmath integrals on two cells:
    2.6e-10 of the largest residual entry and 1.7e-11 of J at cell diameter 0.3 (it falls with h^10 and faster), and it is the ONLY difference (the polynomial test above is exact to 1e-13)."""
    mp = pytest.importorskip('mpmath')
    m = _distorted_square(1)
    m = om.Mesh('triangle', 0.25 * m.coords, m.cells, m.shape, m.lo, m.hi)       # two cells of diameter ~0.3
    F = fam.NonlinearPoissonP1(m)
    rng = np.random.default_rng(6)
    u, f = _dyadic(rng, F.N), _dyadic(rng, F.M)
    T = _SymTri(m)
    X, Y = sp.symbols('x y')
    uex = sp.sin(2 * sp.pi * X) * sp.sin(sp.pi * Y)
    R = [mp.mpf(0)] * F.N
    Jv = mp.mpf(0)
    mp.mp.dps = 30
    for c in range(len(T.cells)):
        v, x, lam, JinvT, adet = T.geom(c)
        uh = sum(_rat(u[v[a]]) * lam[a] for a in range(3))
        gu = T.grad(uh, JinvT)
        fc = _rat(f[c])
        for a in range(3):
            R[v[a]] += mp.mpf(sp.N(_tri_int((gu.dot(T.grad(lam[a], JinvT)) + uh ** 3 * lam[a] - fc * lam[a]) * adet), 30))
        integrand = sp.lambdify((XI, ETA), (sp.Rational(1, 2) * (uh - uex.subs({X: x[0], Y: x[1]})) ** 2
                                            + sp.Rational(3, 10 ** 7) * fc ** 2) * adet, 'mpmath')
        Jv += mp.quad(lambda a_: mp.quad(lambda b_: integrand(a_, b_), [0, 1 - a_]), [0, 1])
        for a, b, o in T.facets(c):
            P, Q, O = T.X[v[a]], T.X[v[b]], T.X[v[o]]
            t = Q - P
            length = sp.sqrt(t.dot(t))
            nl = sp.Matrix([t[1], -t[0]])
            if nl.dot(P - O) < 0:
                nl = -nl
            ref = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
            on = {XI: ref[a][0] * (1 - S) + ref[b][0] * S, ETA: ref[a][1] * (1 - S) + ref[b][1] * S}
            xs = P + t * S
            uef = uex.subs({X: xs[0], Y: xs[1]})
            uf = uh.xreplace(on)
            h = T.diameter(c)
            for k in range(3):
                pf = lam[k].xreplace(on)
                dpdn_l = T.grad(lam[k], JinvT).dot(nl)
                g = sp.lambdify(S, -gu.dot(nl) * pf + (uef - uf) * dpdn_l + 10 / h * length * (uf - uef) * pf, 'mpmath')
                R[v[k]] += mp.quad(g, [0, 1])
    got = asm.assemble_vector(F.residual(u, f), F.N)
    ex = np.array([float(r) for r in R])
    err = np.abs(got - ex).max() / np.abs(ex).max()
    Jo = asm.assemble_scalar(F.output(0, u, f))
    errJ = abs(Jo - float(Jv)) / abs(float(Jv))
    print('quadrature error: residual %.2e, functional %.2e' % (err, errJ))
    assert 1e-13 < err < 1e-6, err                      # a real quadrature error, and a small one
    assert errJ < 1e-9, errJ


# ------------------------------------------------------------------ config 3
def test_hermite_beam_against_exact_integrals():
    """examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py:64-85: R = inner(div grad v, E b t^3 / 12 div grad u) dx
    - f v ds(100), compliance f u ds(100), volume t b L dx; Hermite-3 on intervals with (value, reference-coordinate
    derivative) dofs per vertex (identity push-forward).  The cubic basis is SOLVED from its
    interpolation conditions here, not typed in."""
    xs = np.array([0.0, 0.125, 0.375, 0.5, 0.875, 1.0])
    m = om.Mesh('interval', xs.reshape(-1, 1), np.stack([np.arange(5), np.arange(5) + 1], axis=1), (5,), (0.0,), (1.0,))
    fc, fl = m.exterior_facets()
    tip = [k for k in range(fc.size) if fc[k] == 4 and fl[k] == 1]
    E, width, L, fload = 2.0, 0.25, 1.0, -1.0
    F = fam.EBBeam(m, np.array(tip), E=E, width=width, L=L, f=fload)
    rng = np.random.default_rng(8)
    u, t = _dyadic(rng, F.N), 0.25 + np.abs(_dyadic(rng, F.M))
    # Hermite cubics on [0, 1]: p_k(xi) with (p(0), p'(0), p(1), p'(1)) = e_k
    a = sp.symbols('a0:4')
    cubic = sum(a[i] * XI ** i for i in range(4))
    conds = [cubic.subs(XI, 0), sp.diff(cubic, XI).subs(XI, 0), cubic.subs(XI, 1), sp.diff(cubic, XI).subs(XI, 1)]
    basis = []
    for k in range(4):
        sol = sp.solve([conds[i] - (1 if i == k else 0) for i in range(4)], a)
        basis.append(cubic.subs(sol))
    U = sp.symbols('U0:%d' % F.N)
    Tm = sp.symbols('T0:%d' % F.M)
    R = [0] * F.N
    for c in range(5):
        h = _rat(xs[c + 1] - xs[c])
        dofs = [2 * c, 2 * c + 1, 2 * c + 2, 2 * c + 3]
        uh = sum(U[d] * p for d, p in zip(dofs, basis))
        EI = _rat(E) * _rat(width) * Tm[c] ** 3 / 12
        for d, p in zip(dofs, basis):
            # d2/dx2 = h^-2 d2/dxi2, dx = h dxi
            R[d] += sp.integrate(sp.expand(sp.diff(p, XI, 2) / h ** 2 * EI * sp.diff(uh, XI, 2) / h ** 2 * h), (XI, 0, 1))
    tipdofs = [2 * 4, 2 * 4 + 1, 2 * 5, 2 * 5 + 1]
    for d, p in zip(tipdofs, basis):
        R[d] -= _rat(fload) * p.subs(XI, 1)
    comp = _rat(fload) * sum(U[d] * p.subs(XI, 1) for d, p in zip(tipdofs, basis))
    vol = sum(Tm[c] * _rat(width) * _rat(L) * _rat(xs[c + 1] - xs[c]) for c in range(5))
    tab = {U[i]: _rat(u[i]) for i in range(F.N)}
    tab.update({Tm[i]: _rat(t[i]) for i in range(F.M)})
    _check('eb_beam', m, F, u, [t], R, [comp, vol], U, [Tm], tab, meta=dict(params=[E, width, L, fload], tagged=tip))


# ------------------------------------------------------------------ config 4 (2-D reference size family and 3-D extension)
def _box_mesh(lines):
    """Axis-aligned tensor-product mesh on non-uniform dyadic grid lines (what create_rectangle / create_box cells are:
    constant Jacobians, so the elasticity integrands are polynomial and the 2-point Gauss rules exact)."""
    d = len(lines)
    n = [len(l) - 1 for l in lines]
    if d == 2:
        m = om.rectangle_quad((0.0, 0.0), (1.0, 1.0), n[0], n[1])
    else:
        m = om.box_hex((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), n[0], n[1], n[2])
    X = m.coords.copy()
    for k in range(d):
        idx = np.rint(X[:, k] * n[k]).astype(int)
        X[:, k] = np.asarray(lines[k])[idx]
    kind = 'quadrilateral' if d == 2 else 'hexahedron'
    return om.Mesh(kind, X, m.cells, m.shape, m.lo, m.hi)


def _simp_symbolic(m, d, U, Rho, fvec, right_cells):
    """examples/beam_topo_opt/run_topo_opt_cantilever_beam.py:62-77: E = rho^3, nu = 0.3, lambda = E nu / (1 + nu) / (1 - 2 nu),
    mu = E / 2 / (1 + nu); res = inner(lambda div(u) I + 2 mu eps(u), eps(v)) dx - dot(f, v) ds(100); :79-86 outputs."""
    xi = sp.symbols('z0:%d' % d)
    nu = sp.Rational(3, 10)
    R = [0] * len(U)
    vol = 0
    avg = 0
    comp = 0

    def box_int(expr, gens):
        return sp.Add(*[c / sp.Mul(*[(e + 1) for e in mon]) for mon, c in _poly_terms(sp.sympify(expr), gens)])

    for c, verts in enumerate(m.cells):
        lo = [_rat(v) for v in m.coords[verts[0]]]
        hi = [_rat(v) for v in m.coords[verts[-1]]]
        hs = [hi[k] - lo[k] for k in range(d)]
        phi = [sp.Mul(*[(xi[k] if (a >> k) & 1 else 1 - xi[k]) for k in range(d)]) for a in range(2 ** d)]
        dofs = [[int(verts[a]) * d + k for k in range(d)] for a in range(2 ** d)]
        uh = [sum(U[dofs[a][k]] * phi[a] for a in range(2 ** d)) for k in range(d)]
        grad = lambda w: sp.Matrix(d, d, lambda i, j: sp.diff(w[i], xi[j]) / hs[j])      # noqa: E731
        Ecell = Rho[c] ** 3
        lam_, mu_ = Ecell * nu / (1 + nu) / (1 - 2 * nu), Ecell / 2 / (1 + nu)
        Gu = grad(uh)
        eps_u = (Gu + Gu.T) / 2
        sig = lam_ * Gu.trace() * sp.eye(d) + 2 * mu_ * eps_u
        cellvol = sp.Mul(*hs)
        for a in range(2 ** d):
            for k in range(d):
                v = [phi[a] if i == k else 0 for i in range(d)]
                Gv = grad(v)
                eps_v = (Gv + Gv.T) / 2
                integrand = sum(sig[i, j] * eps_v[i, j] for i in range(d) for j in range(d))
                R[dofs[a][k]] += box_int(integrand, xi) * cellvol
        vol += cellvol
        avg += Rho[c] * cellvol
        if c in right_cells:                   # traction face x = x_max of this cell: xi_0 = 1
            area = sp.Mul(*hs[1:])
            for a in range(2 ** d):
                pf = phi[a].subs(xi[0], 1)
                w = box_int(pf, xi[1:]) * area
                for k in range(d):
                    R[dofs[a][k]] -= fvec[k] * w
                    comp += fvec[k] * U[dofs[a][k]] * w
    return R, avg / vol, comp


@pytest.mark.parametrize('d', [2, 3])
def test_simp_elasticity_against_exact_integrals(d):
    lines = [[0.0, 0.375, 1.0], [0.0, 0.25, 0.5, 1.0]] if d == 2 else [[0.0, 0.625, 1.0], [0.0, 0.75], [0.0, 0.5]]
    m = _box_mesh(lines)
    fc, fl = m.exterior_facets()
    # tagged facets: faces on x = 1 (what TractionBoundary selects, :45-57; here every such face of the mesh)
    lf = m.local_facets
    on_right = [k for k in range(fc.size) if np.all(m.coords[m.cells[fc[k]][lf[fl[k]]], 0] == 1.0)]
    fvec = (0.0, -0.25) if d == 2 else (0.0, -0.25, 0.125)
    F = (fam.SimpQ1 if d == 2 else fam.SimpHex8)(m, np.array(on_right), f=fvec)
    rng = np.random.default_rng(9)
    u, rho = _dyadic(rng, F.N), 0.125 + np.abs(_dyadic(rng, F.M)) / 4
    U = sp.symbols('U0:%d' % F.N)
    Rho = sp.symbols('R0:%d' % F.M)
    R, avg, comp = _simp_symbolic(m, d, U, Rho, [_rat(v) for v in fvec], {int(c) for c in fc[on_right]})
    tab = {U[i]: _rat(u[i]) for i in range(F.N)}
    tab.update({Rho[i]: _rat(rho[i]) for i in range(F.M)})
    _check('simp_q1' if d == 2 else 'simp_hex8', m, F, u, [rho], R, [avg, comp], U, [Rho], tab, 2e-13,
           meta=dict(params=[0.3] + list(fvec) + [3.0], tagged=on_right))


# ------------------------------------------------------------------ config 5a
def test_mesh_motion_against_exact_integrals():
    """examples/em_motor_opt/motor_pde.py:134-183 (pdeResMM, nitsche and sym True) and :199-210 (area_form), typed in as the
    reference writes them: F = grad(uhat) + I, E = (F^T F - I) / 2, K = mu = det(F)^-3, S = K tr(E) I + 2 mu (E - tr(E) I / 3),
    P = F S; on each tagged one-sided facet -(P n).v + (dP[v] n).(uhat - g) + 5e3 / det(F)^3 / h_E v.(uhat - g) with
    dP[v] = d/d eps P(F + eps grad v) at eps = 0 (`ufl.derivative`).  P1 fields make the cell integrand constant and the facet
    integrands quadratic in the facet parameter, so Simpson's rule IS the exact integral.  The oracle codes dP analytically and
    differentiates the residual by complex steps; here the global rational expressions are differentiated by 80-digit central
    differences (`_jac_at`)."""
    from oracle.motor_mm import MotorMM
    m = _distorted_square(1)
    T = _SymTri(m)
    # tagged one-sided facets: the diagonal 0-3 from both cells ("+" and "-" side of dS(1000)) and the boundary edge of cell 0
    # opposite its local vertex 0 (ds(1000)); local facet i is the edge opposite local vertex i
    fc = np.array([0, 0, 1], dtype=np.int32)
    fl = np.array([0, 1, 1], dtype=np.int32)
    for c, l in zip(fc, fl):
        v = T.cells[c]
        assert (l == 1) == (set(v) - {v[l]} == {0, 3})
    tags = np.array([15, 1])
    F = MotorMM(m, (fc, fl), tags)
    rng = np.random.default_rng(10)
    uh, g = _dyadic_small(rng, F.N, 1 / 16), _dyadic_small(rng, F.N, 1 / 16)
    U = sp.symbols('U0:%d' % F.N)
    Gs = sp.symbols('G0:%d' % F.N)
    eps = sp.Symbol('eps')
    I2 = sp.eye(2)

    def Pk(Fm):
        E = (Fm.T * Fm - I2) / 2
        K = 1 / Fm.det() ** 3
        S_ = K * E.trace() * I2 + 2 * K * (E - E.trace() * I2 / 3)
        return Fm * S_

    R = [0] * F.N
    areas = [0, 0, 0]
    for c in range(2):
        v, x, lam, JinvT, adet = T.geom(c)
        gphi = [T.grad(l, JinvT) for l in lam]                                  # constant physical gradients
        field = lambda W: [sum(W[2 * v[a] + k] * lam[a] for a in range(3)) for k in range(2)]   # noqa: E731
        uhv, gv = field(U), field(Gs)
        Fm = I2 + sp.Matrix(2, 2, lambda i, j: sum(U[2 * v[a] + i] * gphi[a][j] for a in range(3)))
        P = Pk(Fm)
        area = adet / 2
        for a in range(3):
            for k in range(2):
                gradv = sp.Matrix(2, 2, lambda i, j: gphi[a][j] if i == k else 0)
                R[2 * v[a] + k] += area * sum(P[i, j] * gradv[i, j] for i in range(2) for j in range(2))
        for k, ids in enumerate(MotorMM.OUT_IDS):
            if tags[c] in ids:
                areas[k] += Fm.det() * area
        dPs = {}
        for fcell, o in zip(fc, fl):
            if fcell != c:
                continue
            a_, b_ = [k for k in range(3) if k != o]
            Pv, Qv, Ov = T.X[v[a_]], T.X[v[b_]], T.X[v[o]]
            t = Qv - Pv
            length = sp.sqrt(t.dot(t))
            n = sp.Matrix([t[1], -t[0]]) / length
            if n.dot(Pv - Ov) < 0:
                n = -n
            ref = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
            on = lambda s_: {XI: ref[a_][0] * (1 - s_) + ref[b_][0] * s_, ETA: ref[a_][1] * (1 - s_) + ref[b_][1] * s_}   # noqa: E731
            beta = 5000 / Fm.det() ** 3
            h = T.diameter(c)
            for a in range(3):
                for k in range(2):
                    gradv = sp.Matrix(2, 2, lambda i, j: gphi[a][j] if i == k else 0)
                    if (a, k) not in dPs:
                        dPs[a, k] = sp.diff(Pk(Fm + eps * gradv), eps).subs(eps, 0)
                    dP = dPs[a, k]
                    dPn = dP * n
                    Pn = P * n

                    def integrand(s_):
                        tab = on(s_)
                        vv = lam[a].xreplace(tab)
                        d = [(uhv[i] - gv[i]).xreplace(tab) for i in range(2)]
                        return -Pn[k] * vv + dPn[0] * d[0] + dPn[1] * d[1] + beta / h * vv * d[k]

                    simpson = (integrand(0) + 4 * integrand(sp.Rational(1, 2)) + integrand(1)) / 6
                    R[2 * v[a] + k] += simpson * length
    tab = {U[i]: _rat(uh[i]) for i in range(F.N)}
    tab.update({Gs[i]: _rat(g[i]) for i in range(F.N)})
    # beta = 5e3: the penalty rows dominate the scale (tolerance 1e-12)
    _check('motor_mm', m, F, uh, [g], R, areas, U, [Gs], tab, 1e-12, numeric_jac=True,
           meta=dict(params=[5e3], facet_cells=fc, facet_locals=fl, cell_tags=tags))


# ------------------------------------------------------------------ config 5b
def test_magnetostatics_against_exact_integrals():
    """examples/em_motor_opt/motor_pde.py typed in as written: RelativePermeability :12-35 (linear / cubic / exponential
    mu_r(|B|) on ids 1-2 with norm_B = sqrt(B.B + DOLFIN_EPS), 1.05 on the magnet ids 3-14, 1 elsewhere), JS :46-87 (magnet
    H . curl_x v and the three-phase winding currents), pdeResEM :90-130 with nitsche and sym True (Nanson normal
    J F^-T n, beta = 1e4, BOTH boundary_components use the steel curve), B_power_form :186-197; kinematics gradx / J / F of
    femo/fea/utils_dolfinx.py:34-66.  P1 fields: cell integrands are constant up to the factor v, facet integrands quadratic
    in the facet parameter (Simpson exact).  The fit constants are data (femo_b200/forms/bh_fit.json, produced by the
    reference's piecewise_permeability.py); values at 30 digits, derivatives by 80-digit central differences; the oracle
    differentiates by complex steps."""
    import json
    import os
    from oracle.motor import MotorEM
    fit = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'femo_b200', 'forms', 'bh_fit.json')))
    q = lambda x: sp.Rational(Fraction(float(x)).numerator, Fraction(float(x)).denominator)     # noqa: E731
    m = _distorted_square(2)
    T = _SymTri(m)
    tags = np.array([1, 2, 3, 4, 15, 16, 20, 53])
    p_, s_n, Hc, angle, iq, beta = 12, 36, 838e3, 0.3, 282.2 / 0.00016231, 1e4
    Fo = MotorEM(m, tags, Hc=Hc, p=p_, s=s_n, angle=angle, iq=iq, beta=beta)
    rng = np.random.default_rng(16)                      # |B| per cell 0.5 ... 2.4: all three pieces of the curve
    u, uh = _dyadic(rng, Fo.N, 8) / 4, _dyadic_small(rng, Fo.M, 1 / 32)
    U = sp.symbols('U0:%d' % Fo.N)
    W = sp.symbols('W0:%d' % Fo.M)
    tab = {U[i]: _rat(u[i]) for i in range(Fo.N)}
    tab.update({W[i]: _rat(uh[i]) for i in range(Fo.M)})
    EPS = q(3e-16)
    mu0 = q(4e-7 * np.pi)                                # the double the example passes as vacuum_perm
    branches = set()

    def mu_steel(gradu):
        B = sp.Matrix([gradu[1], -gradu[0]])
        nB = sp.sqrt(B.dot(B) + EPS)
        val = float(nB.xreplace(tab))
        if val < fit['x1']:
            branches.add('linear')
            return q(fit['lin'][0]) * nB + q(fit['lin'][1])
        if val < fit['x2']:
            branches.add('cubic')
            a, b, c, d = (q(v) for v in fit['cubic'])
            return a * nB ** 3 + b * nB ** 2 + c * nB + d
        branches.add('exp')
        a, b, c = (q(v) for v in fit['exp'])
        return a * sp.exp(b * nB + c) + 1

    def mu_r(sub, gradu):
        if sub in (1, 2):
            return mu_steel(gradu)
        return sp.Rational(105, 100) if 3 <= sub <= 14 else 1

    ang = q(angle)
    i_abc = [q(iq * np.sin(angle)) + EPS, q(iq * np.sin(angle - 2 * np.pi / 3)) + EPS, q(iq * np.sin(angle + 2 * np.pi / 3)) + EPS]
    R = [0] * Fo.N
    outs = [0, 0]
    for c in range(len(T.cells)):
        v, x, lam, JinvT, adet = T.geom(c)
        area = adet / 2
        gphi = [T.grad(l, JinvT) for l in lam]
        Fm = sp.eye(2) + sp.Matrix(2, 2, lambda i, j: sum(W[2 * v[a] + i] * gphi[a][j] for a in range(3)))
        Jd = Fm.det()
        Finv = Fm.inv()
        gradx = lambda g_: (g_.T * Finv).T                                            # noqa: E731  dot(grad f, inv F)
        gu = gradx(sum((U[v[a]] * gphi[a] for a in range(3)), sp.zeros(2, 1)))
        sub = int(tags[c])
        nu = 1 / mu0 / mu_r(sub, gu)
        for a in range(3):
            gv = gradx(gphi[a])
            R[v[a]] += nu * gu.dot(gv) * Jd * area                                    # cellwise constant integrand
            if 3 <= sub < 3 + p_:                                                      # magnets: dx(i + 2 + 1)
                i = sub - 3
                fa = 2 * np.pi / p_ / 2 + i * (2 * np.pi / p_) + angle * 2 / p_
                H = sp.Matrix([q((-1) ** i * Hc * np.cos(fa)), q((-1) ** i * Hc * np.sin(fa))])
                curl_v = sp.Matrix([gv[1], -gv[0]])
                R[v[a]] -= H.dot(curl_v) * Jd * area
            if p_ + 3 <= sub < p_ + 3 + s_n:                                           # windings: int v J dx = J area / 3
                i, k = divmod(sub - (p_ + 3), 3)
                amp = [i_abc[1] * (-1) ** (i + 1), i_abc[0] * (-1) ** i, i_abc[2] * (-1) ** (i + 1)][k]
                R[v[a]] -= amp * Jd * area / 3
        if sub in (1, 2):
            Bm = sp.sqrt(gu[0] ** 2 + gu[1] ** 2)
            for k, n_ in enumerate((2, q(1.76835))):
                outs[k] += Bm ** n_ * Jd * area
        for a_, b_, o in T.facets(c):
            Pv, Qv, Ov = T.X[v[a_]], T.X[v[b_]], T.X[v[o]]
            t = Qv - Pv
            length = sp.sqrt(t.dot(t))
            n = sp.Matrix([t[1], -t[0]]) / length
            if n.dot(Pv - Ov) < 0:
                n = -n
            nN = Jd * Finv.T * n
            nrm = sp.sqrt(nN.dot(nN))
            h = T.diameter(c)
            ref = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
            uhx = sum(U[v[a]] * lam[a] for a in range(3))
            for comp in (1, 2):                                                        # boundary_components = [0, 1] -> ids 1, 2
                coeff = 1 / mu0 / mu_r(comp, gu)
                for a in range(3):
                    gvn = gradx(gphi[a]).dot(nN)

                    def integrand(s):
                        on = {XI: ref[a_][0] * (1 - s) + ref[b_][0] * s, ETA: ref[a_][1] * (1 - s) + ref[b_][1] * s}
                        vv, uu = lam[a].xreplace(on), uhx.xreplace(on)
                        return coeff * (-gu.dot(nN) * vv - gvn * uu) + sp.Integer(10) ** 4 / h * coeff * nrm * vv * uu

                    R[v[a]] += (integrand(0) + 4 * integrand(sp.Rational(1, 2)) + integrand(1)) / 6 * length
    assert branches == {'linear', 'cubic', 'exp'}, branches   # every piece of the B-H curve is exercised
    _check('motor_em', m, Fo, u, [uh], R, outs, U, [W], tab, 1e-11, numeric_jac=True,
           meta=dict(cell_tags=tags, Hc=Hc, p=p_, s=s_n, angle=angle, iq=iq, beta=beta))


# ------------------------------------------------------------------ Reissner-Mindlin plate
def test_reissner_mindlin_plate_against_the_energy_functional():
    """oracle/rm_plate.py assembles B-matrix element stiffnesses; here the PUBLISHED energy it restates (its docstring: bending
    D(t)/2 [(1 - nu) kappa:kappa + nu tr(kappa)^2], transverse shear ks G t / 2 |grad w - theta|^2 on the reduced 3-point
    interior rule, penalty clamp, load f v) is typed in as ONE scalar functional of the global dofs and the residual, Jacobian,
    dR/dt, dR/df and the output partials are its sympy derivatives.  The reference's own forms live in the un-vendored
    shell_analysis_fenicsx (examples/test_shell_m3l/shell_pde.py:225-253), so this pins the oracle's arithmetic against the
    formulation, not against that package (parity unpinned)."""
    from oracle.rm_plate import RMPlate
    m = _distorted_square(1)
    T = _SymTri(m)
    E_, nu_, pen_, rho_ = 1.0e4, 0.25, 1.0e6, 2.0
    F = RMPlate(m, clamped=None, E=E_, nu=nu_, pen=pen_, rho=rho_)
    nv, ne = T.nv, len(T.edges)
    assert F.N == 3 * nv + ne and F.M == nv
    eid = {e: k for k, e in enumerate(T.edges)}
    rng = np.random.default_rng(13)
    u, t, f = _dyadic(rng, F.N), 0.25 + np.abs(_dyadic(rng, F.M)) / 8, _dyadic(rng, F.M)
    U = sp.symbols('U0:%d' % F.N)
    Ts = sp.symbols('T0:%d' % F.M)
    Fs = sp.symbols('F0:%d' % F.M)
    E, nu, pen, rho = _rat(E_), _rat(nu_), _rat(pen_), _rat(rho_)
    Db, Gs = E / (12 * (1 - nu ** 2)), sp.Rational(5, 6) * E / (2 * (1 + nu))
    shear_pts = [(sp.Rational(1, 6), sp.Rational(1, 6)), (sp.Rational(1, 6), sp.Rational(2, 3)), (sp.Rational(2, 3), sp.Rational(1, 6))]
    energy = 0            # bending + shear
    clamp = 0
    load = 0
    comp = 0
    mass = 0
    for c in range(len(T.cells)):
        v, x, lam, JinvT, adet = T.geom(c)
        wd = list(v) + [nv + eid[tuple(sorted((v[(i + 1) % 3], v[(i + 2) % 3])))] for i in range(3)]
        p2 = [l * (2 * l - 1) for l in lam] + [4 * lam[(i + 1) % 3] * lam[(i + 2) % 3] for i in range(3)]
        w = sum(U[d] * p for d, p in zip(wd, p2))
        th = [sum(U[nv + ne + 2 * v[a] + k] * lam[a] for a in range(3)) for k in range(2)]
        th_ = sum(Ts[v[a]] * lam[a] for a in range(3))
        ff = sum(Fs[v[a]] * lam[a] for a in range(3))
        gth = [T.grad(th[k], JinvT) for k in range(2)]
        kxx, kyy, kxy = gth[0][0], gth[1][1], (gth[0][1] + gth[1][0]) / 2
        bend = Db * th_ ** 3 / 2 * ((1 - nu) * (kxx ** 2 + kyy ** 2 + 2 * kxy ** 2) + nu * (kxx + kyy) ** 2)
        energy += _tri_int(bend * adet)
        gw = T.grad(w, JinvT)
        shear = Gs * th_ / 2 * ((gw[0] - th[0]) ** 2 + (gw[1] - th[1]) ** 2)
        energy += sum(shear.xreplace({XI: a_, ETA: b_}) for a_, b_ in shear_pts) * adet / 6
        load += _tri_int(ff * w * adet)
        comp += _tri_int(w ** 2 / 2 * adet)
        mass += _tri_int(rho * th_ * adet)
        for a_, b_, o in T.facets(c):
            t_ = T.X[v[b_]] - T.X[v[a_]]
            length = sp.sqrt(t_.dot(t_))
            ref = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
            on = {XI: ref[a_][0] * (1 - S) + ref[b_][0] * S, ETA: ref[a_][1] * (1 - S) + ref[b_][1] * S}
            clamp += pen / 2 * length * _seg_int(w.xreplace(on) ** 2 + th[0].xreplace(on) ** 2 + th[1].xreplace(on) ** 2)
    total = energy + clamp - load
    R = [sp.diff(total, s) for s in U]
    tab = {U[i]: _rat(u[i]) for i in range(F.N)}
    tab.update({Ts[i]: _rat(t[i]) for i in range(F.M)})
    tab.update({Fs[i]: _rat(f[i]) for i in range(F.M)})
    # the 1e6 penalty rows set the scale of the matrix (tolerance 1e-12)
    _check('rm_plate', m, F, u, [t, f], R, [comp, mass, energy], U, [Ts, Fs], tab, 1e-12,
           meta=dict(params=[E_, nu_, pen_, rho_]))
