"""The reference's OWN form code, executed: `pdeRes`, `outputForm`, `compliance`, `volume`, `averageFunc` ... are lifted out of
the example scripts under /root/reference/examples (tests/_ufl_sympy.py: `ast` + a sympy implementation of the UFL operators
they use) and evaluated cell by cell on the meshes of the exact-integral fixtures.  The global residuals / functionals that come
out -- and their `sympy.diff` derivatives -- must equal tests/golden/symbolic/*.npz, i.e. the values the oracle (1e-13) and the
CUDA kernels (1e-12, tests/test_gpu_exact.py) are held to.  This closes the last hand-typed link: the fixtures are what the
reference's code says, integrated exactly.  Needs the reference checkout (skips on the GPU box); no GPU."""
import os

import numpy as np
import pytest

sp = pytest.importorskip('sympy')

import _ufl_sympy as U   # noqa: E402

REF = '/root/reference/examples'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'symbolic')
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='no reference checkout')
TOL = 1e-13


def _num(exprs, tab):
    return np.array([float(sp.N(U.exactify(e).xreplace(tab), 30)) for e in exprs], dtype=np.float64)


def _same(got, exact, what):
    exact = np.asarray(exact, dtype=np.float64).reshape(np.shape(got))
    scale = max(np.abs(exact).max(), 1e-300)
    assert np.abs(np.asarray(got) - exact).max() <= TOL * scale, (what, np.abs(np.asarray(got) - exact).max() / scale)


def _compare(z, R, Js, Usym, Msym, tab):
    _same(_num(R, tab), z['R'], 'R')
    Rm = sp.Matrix([U.exactify(r) for r in R])
    _same(np.array(Rm.jacobian(list(Usym)).xreplace(tab).evalf(30), dtype=np.float64), z['A'], 'dR/du')
    _same(np.array(Rm.jacobian(list(Msym)).xreplace(tab).evalf(30), dtype=np.float64), z['D0'], 'dR/dm')
    for k, Jf in enumerate(Js):
        Jf = U.exactify(Jf)
        _same(_num([Jf], tab), z['J%d' % k], 'J%d' % k)
        _same(_num([sp.diff(Jf, s) for s in Usym], tab), z['Ju%d' % k], 'dJ%d/du' % k)
        _same(_num([sp.diff(Jf, s) for s in Msym], tab), z['Jm%d_0' % k], 'dJ%d/dm' % k)


def _table(z, Usym, Msym):
    tab = {Usym[i]: U.rat(z['state'][i]) for i in range(len(Usym))}
    tab.update({Msym[i]: U.rat(z['input0'][i]) for i in range(len(Msym))})
    return tab


def _tri_cells(z):
    """Cells of a triangle fixture with their exterior facets (edges that belong to one cell)."""
    cells = [tuple(int(v) for v in c) for c in z['cells']]
    X = [sp.Matrix([U.rat(a), U.rat(b)]) for a, b in z['coords']]
    cnt = {}
    for c in cells:
        for a in range(3):
            e = tuple(sorted((c[a], c[(a + 1) % 3])))
            cnt[e] = cnt.get(e, 0) + 1
    out = []
    for c in cells:
        ext = []
        for o in range(3):
            a, b = [k for k in range(3) if k != o]
            if cnt[tuple(sorted((c[a], c[b])))] == 1:
                ext.append(((a, b, o), None))
        out.append((c, U.Cell('triangle', [X[v] for v in c], ext)))
    return out, sorted(cnt)


def test_poisson_forms_as_written_in_the_example():
    """examples/poisson_opt/run_poisson_opt.py: pdeRes(u, v, f) :136 (strong BCs are not part of the form), outputForm(u, f, u_ex) :115."""
    ns = U.load_defs(os.path.join(REF, 'poisson_opt', 'run_poisson_opt.py'))
    z = np.load(os.path.join(GOLD, 'poisson_p1.npz'))
    cells, _ = _tri_cells(z)
    N, M = z['state'].size, z['input0'].size
    Us, Fs = sp.symbols('U0:%d' % N), sp.symbols('F0:%d' % M)
    R, J = [0] * N, 0
    for c, (v, cell) in enumerate(cells):
        U.CTX = cell
        u = U.Field(sum(Us[v[a]] * cell.lam[a] for a in range(3)), coeffs=[Us[i] for i in v])
        uex = U.Field(sum(U.rat(z['meta_u_ex'][v[a]]) * cell.lam[a] for a in range(3)))
        f = U.Field(Fs[c])
        for a in range(3):
            test = U.Field(cell.lam[a], values=[1 if b == a else 0 for b in range(3)])
            R[v[a]] += ns['pdeRes'](u, test, f).value()
        J += ns['outputForm'](u, f, uex).value()
    _compare(z, R, [J], Us, Fs, _table(z, Us, Fs))


@pytest.mark.parametrize('degree', [1, 2])
def test_nonlinear_poisson_forms_as_written_in_the_example(degree):
    """examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py: pdeRes(u, v, f, u_exact=..., weak_bc=True, sym=True) :196-197,
    outputForm(u, f, u_ex) :172, with the fixture's polynomial u_exact (the example's sin field is not polynomial)."""
    ns = U.load_defs(os.path.join(REF, 'nonlinear_poisson_opt', 'run_nonlinear_poisson_opt.py'))
    z = np.load(os.path.join(GOLD, 'nlpoisson_p%d.npz' % degree))
    cells, edges = _tri_cells(z)
    nv = z['coords'].shape[0]
    eid = {e: k for k, e in enumerate(edges)}
    N, M = z['state'].size, z['input0'].size
    Us, Fs = sp.symbols('U0:%d' % N), sp.symbols('F0:%d' % M)
    R, J = [0] * N, 0
    for c, (v, cell) in enumerate(cells):
        U.CTX = cell
        lam = cell.lam
        if degree == 1:
            dofs, phi = list(v), lam
        else:
            dofs = list(v) + [nv + eid[tuple(sorted((v[(i + 1) % 3], v[(i + 2) % 3])))] for i in range(3)]
            phi = [l * (2 * l - 1) for l in lam] + [4 * lam[(i + 1) % 3] * lam[(i + 2) % 3] for i in range(3)]
        u = U.Field(sum(Us[d] * p for d, p in zip(dofs, phi)), coeffs=[Us[d] for d in dofs])
        x, y = cell.x
        uex = x * x * y - 3 * x * y + y ** 3 / 4 + sp.Rational(1, 2)
        f = U.Field(Fs[c])
        for a, d in enumerate(dofs):
            test = U.Field(phi[a], values=[1 if b == a else 0 for b in range(len(dofs))])
            R[d] += ns['pdeRes'](u, test, f, u_exact=uex, weak_bc=True, sym=True).value()
        J += ns['outputForm'](u, f, uex).value()
    _compare(z, R, [J], Us, Fs, _table(z, Us, Fs))


def test_beam_forms_as_written_in_the_example():
    """examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py: pdeRes(u, v, t, f, ds_(100), E, width) :127-131,
    compliance(u, f, ds_(100)) :135, volume(t, width, L) :137; Hermite-3 with reference-derivative slope dofs."""
    ns = U.load_defs(os.path.join(REF, 'beam_thickness_opt', 'run_thickness_opt_cantilever_beam.py'))
    z = np.load(os.path.join(GOLD, 'eb_beam.npz'))
    Eb, width, L, fload = (float(v) for v in z['meta_params'])
    xs = z['coords'][:, 0]
    nc = z['cells'].shape[0]
    N, M = z['state'].size, z['input0'].size
    Us, Ts = sp.symbols('U0:%d' % N), sp.symbols('T0:%d' % M)
    xi = sp.Symbol('xi_0')
    a_ = sp.symbols('a0:4')
    cubic = sum(a_[i] * xi ** i for i in range(4))
    conds = [cubic.subs(xi, 0), sp.diff(cubic, xi).subs(xi, 0), cubic.subs(xi, 1), sp.diff(cubic, xi).subs(xi, 1)]
    basis = [cubic.subs(sp.solve([conds[i] - (1 if i == k else 0) for i in range(4)], a_)) for k in range(4)]
    f = U.Constant(None, fload)
    ds100 = U.NAMESPACE['ds'](100)
    R, comp, vol = [0] * N, 0, 0
    for c in range(nc):
        ext = [((0, 1), 100)] if c == nc - 1 else ([((0, 0), 0)] if c == 0 else [])
        U.CTX = U.Cell('interval', [sp.Matrix([U.rat(xs[c])]), sp.Matrix([U.rat(xs[c + 1])])], ext)
        dofs = [2 * c, 2 * c + 1, 2 * c + 2, 2 * c + 3]
        u = U.Field(sum(Us[d] * p for d, p in zip(dofs, basis)), coeffs=[Us[d] for d in dofs])
        t = U.Field(Ts[c])
        for a, d in enumerate(dofs):
            test = U.Field(basis[a], values=[1 if b == a else 0 for b in range(4)])
            R[d] += ns['pdeRes'](u, test, t, f, ds100, Eb, width).value()
        comp += ns['compliance'](u, f, ds100).value()
        vol += ns['volume'](t, width, L).value()
    _compare(z, R, [comp, vol], Us, Ts, _table(z, Us, Ts))


@pytest.mark.parametrize('name', ['simp_q1', 'simp_hex8'])
def test_simp_forms_as_written_in_the_example(name):
    """examples/beam_topo_opt/run_topo_opt_cantilever_beam.py: pdeRes(u, v, rho, f, dss=ds_(100), method='SIMP') :102-107,
    averageFunc(rho) :111, compliance(u, f, dss=ds_(100)) :113-115.  The form is written for any dimension (`len(u)`,
    `Identity(d)`), so the same reference code defines the 3-D hexahedral extension."""
    z = np.load(os.path.join(GOLD, name + '.npz'))
    d = z['coords'].shape[1]
    prm = [float(v) for v in z['meta_params']]
    assert prm[0] == 0.3 and prm[-1] == 3.0                     # the script hard-codes nu = 0.3 and the cubic SIMP law
    cells = z['cells']
    X = z['coords']
    lo, hi = X.min(axis=0), X.max(axis=0)
    total = float(np.prod(hi - lo))
    ns = U.load_defs(os.path.join(REF, 'beam_topo_opt', 'run_topo_opt_cantilever_beam.py'),
                     extra=dict(mesh=None, assemble=lambda form: U.rat(total), Function=lambda V: U.Field(sp.Integer(0))))
    N, M = z['state'].size, z['input0'].size
    Us, Rs = sp.symbols('U0:%d' % N), sp.symbols('R0:%d' % M)
    fvec = U.Constant(None, tuple(prm[1:1 + d]))
    ds100 = U.NAMESPACE['ds'](100)
    R, avg, comp = [0] * N, 0, 0
    for c, verts in enumerate(cells):
        Xc = [sp.Matrix([U.rat(v) for v in X[k]]) for k in verts]
        ext = []
        for k in range(d):
            if X[verts[0]][k] == lo[k]:
                ext.append(((k, 0), 0))
            if X[verts[-1]][k] == hi[k]:
                ext.append(((k, 1), 100 if k == 0 else 0))          # the traction faces of the fixture: x = x_max
        cell = U.CTX = U.Cell('box', Xc, ext)
        phi = [sp.Mul(*[(cell.xi[k] if (a >> k) & 1 else 1 - cell.xi[k]) for k in range(d)]) for a in range(2 ** d)]
        dofs = [[int(verts[a]) * d + k for k in range(d)] for a in range(2 ** d)]
        flat = [dofs[a][k] for a in range(2 ** d) for k in range(d)]
        u = U.Field(sp.Matrix([sum(Us[dofs[a][k]] * phi[a] for a in range(2 ** d)) for k in range(d)]), coeffs=[Us[i] for i in flat])
        rho = U.Field(Rs[c])
        for a in range(2 ** d):
            for k in range(d):
                test = U.Field(sp.Matrix([phi[a] if i == k else 0 for i in range(d)]))
                R[dofs[a][k]] += ns['pdeRes'](u, test, rho, fvec, dss=ds100, method='SIMP').value()
        avg += ns['averageFunc'](rho).value()
        comp += ns['compliance'](u, fvec, dss=ds100).value()
    _compare(z, R, [avg, comp], Us, Rs, _table(z, Us, Rs))
