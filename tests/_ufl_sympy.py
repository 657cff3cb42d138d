"""A slice of UFL on top of sympy, just large enough to EXECUTE the form-defining functions of the reference's example scripts
(`pdeRes`, `interiorResidual`, `boundaryResidual`, `outputForm`, `compliance`, `volume`, `averageFunc`, ...) exactly as they
are written there -- the function definitions are lifted out of the scripts with `ast` (the scripts themselves build meshes
and run optimisations at import time) and run in a namespace where `grad`, `inner`, `dot`, `div`, `derivative`, `dx`,
`ds(tag)`, `FacetNormal`, `CellDiameter`, `SpatialCoordinate`, `Constant`, `Identity` ... are the sympy versions below.

One cell at a time: the "functions" handed to the reference code are polynomials in the reference coordinates of the current
cell with the global dof symbols as coefficients, the test function is one basis function, and a form `expr*dx + expr*ds(100)`
evaluates to the exact integral over the cell and its tagged exterior facets.  Summing over cells and test functions gives the
global residual as sympy expressions in the dof symbols -- produced by the REFERENCE'S code, not by a restatement of it.
TEST INFRASTRUCTURE ONLY (tests/test_reference_forms.py); needs a reference checkout, so those tests skip elsewhere."""
import ast
import types
from fractions import Fraction

import sympy as sp

CTX = None                      # the current cell
NSYM = sp.symbols('n_0:3')      # facet normal components, substituted per facet
HSYM = sp.Symbol('h_E')         # CellDiameter


def rat(v):
    fr = Fraction(float(v))
    return sp.Rational(fr.numerator, fr.denominator)


# ------------------------------------------------------------------------------------------------- cells
class Facet:
    def __init__(self, restrict, params, normal, measure, tag=None, interior=False):
        self.restrict, self.params, self.normal, self.measure, self.tag, self.interior = restrict, params, normal, measure, tag, interior


def exactify(expr):
    """Python floats that entered through the reference code (0.5, 0.3, 6E-7 ...) become sympy Floats; replace them by the
    exact rational value of the double so that the polynomial arithmetic below stays exact."""
    expr = sp.sympify(expr)
    return expr.xreplace({f: rat(f) for f in expr.atoms(sp.Float)})


POLYNOMIAL_COEFFICIENTS = True      # False for the motor families: coefficients are rational / sqrt / exp expressions of the dofs


def _rule_int(expr, gens, triangle):
    """Integration by substitution for integrands that are polynomials of LOW degree in the reference coordinates with
    arbitrary coefficient expressions (P1 fields on affine cells): the edge-midpoint rule on the reference triangle (exact to
    degree 2) and Simpson's rule per direction elsewhere (exact to degree 3).  The degree is checked at the numeric point."""
    half = sp.Rational(1, 2)
    if triangle:
        rule, maxdeg = [((half, 0), sp.Rational(1, 6)), ((0, half), sp.Rational(1, 6)), ((half, half), sp.Rational(1, 6))], 2
    else:
        rule, maxdeg = [((), sp.Integer(1))], 3
        for _ in gens:
            rule = [(p + (x,), w * wx) for p, w in rule for x, wx in ((0, sp.Rational(1, 6)), (half, sp.Rational(2, 3)), (1, sp.Rational(1, 6)))]
    num = sp.Poly(sp.N(expr.xreplace(NUMERIC), 30), *gens)
    assert num.total_degree() <= maxdeg, 'integrand of degree %d in the reference coordinates' % num.total_degree()
    return sp.Add(*[w * expr.xreplace(dict(zip(gens, p))) for p, w in rule])


def _poly_int(expr, gens, weight, triangle=False):
    """Exact integral of a polynomial in `gens`; weight(monomial exponents) is the integral of that monomial."""
    expr = exactify(expr)
    if not gens:
        return expr
    # irrational constants (facet lengths, cell diameters: square roots of rationals) ride along as extra generators
    irr = {a: sp.Dummy('c') for a in expr.atoms(sp.Pow) if a.is_number and not a.is_Rational}
    expr = expr.xreplace(irr)
    others = sorted(expr.free_symbols - set(gens), key=str)
    if not POLYNOMIAL_COEFFICIENTS:
        return _rule_int(expr.xreplace({v: a for a, v in irr.items()}), gens, triangle)
    p = sp.poly(expr, *gens, *others, domain='QQ')
    k = len(gens)
    tot = 0
    for mon, c in p.terms():
        tot += sp.sympify(c) * weight(mon[:k]) * sp.Mul(*[g ** e for g, e in zip(others, mon[k:]) if e])
    return sp.sympify(tot).xreplace({v: a for a, v in irr.items()})


class Cell:
    """kind 'triangle' | 'interval' | 'box'; X = vertex coordinates (sympy Matrices); tags = {facet key: tag}."""

    def __init__(self, kind, X, exterior=(), cell_tag=None, interior=()):
        """exterior / interior: [(facet key, tag)] of this cell's tagged facets on the boundary / inside the mesh (an interior
        facet is listed by BOTH cells that share it: each contributes its own side of a `dS` integral)."""
        self.kind, self.X, self.cell_tag = kind, X, cell_tag
        d = X[0].rows
        self.d = d
        self.xi = sp.symbols('xi_0:%d' % d)
        if kind == 'triangle':
            J = sp.Matrix.hstack(X[1] - X[0], X[2] - X[0])
            self.x = X[0] + J * sp.Matrix(self.xi)
            self.lam = [1 - self.xi[0] - self.xi[1], self.xi[0], self.xi[1]]
            self.diameter = sp.sqrt(max((X[a] - X[b]).dot(X[a] - X[b]) for a, b in ((0, 1), (0, 2), (1, 2))))
        else:                                                     # interval / axis-aligned box, tensor vertex order
            h = [X[-1][k] - X[0][k] for k in range(d)]
            J = sp.diag(*h)
            self.x = X[0] + J * sp.Matrix(self.xi)
            self.h = h
            self.diameter = sp.sqrt(sum(v * v for v in h))
        self.JinvT = J.inv().T
        self.absdet = sp.Abs(J.det())
        self.facets = []
        s = sp.Symbol('s')
        ref = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
        for key, tag, inside in [(k, t, False) for k, t in exterior] + [(k, t, True) for k, t in interior]:
            if kind == 'triangle':
                a, b, o = key                                     # local vertices of the facet, opposite vertex
                t = X[b] - X[a]
                length = sp.sqrt(t.dot(t))
                n = sp.Matrix([t[1], -t[0]]) / length
                if n.dot(X[a] - X[o]) < 0:
                    n = -n
                restrict = {self.xi[0]: ref[a][0] * (1 - s) + ref[b][0] * s, self.xi[1]: ref[a][1] * (1 - s) + ref[b][1] * s}
                self.facets.append(Facet(restrict, (s,), n, length, tag, inside))
            else:
                k, side = key                                     # axis, 0 / 1
                n = sp.zeros(d, 1)
                n[k] = 1 if side else -1
                params = tuple(x for j, x in enumerate(self.xi) if j != k)
                self.facets.append(Facet({self.xi[k]: side}, params, n, sp.Mul(*[h[j] for j in range(d) if j != k]), tag, inside))

    def integrate(self, expr):
        if self.kind == 'triangle':
            w = lambda m: sp.Rational(int(sp.factorial(m[0]) * sp.factorial(m[1])), int(sp.factorial(m[0] + m[1] + 2)))   # noqa: E731
        else:
            w = lambda m: sp.Rational(1, int(sp.Mul(*[e + 1 for e in m])))                                              # noqa: E731
        return _poly_int(sp.sympify(expr).xreplace({HSYM: self.diameter}) * self.absdet, self.xi, w, self.kind == 'triangle')

    def integrate_facet(self, f, expr):
        tab = {NSYM[i]: f.normal[i] for i in range(self.d)}
        tab[HSYM] = self.diameter
        e = sp.sympify(expr).xreplace(tab)
        e = e.xreplace(f.restrict)
        return _poly_int(e, f.params, lambda m: sp.Rational(1, int(sp.Mul(*[k + 1 for k in m])))) * f.measure


# ------------------------------------------------------------------------------------------------- fields and forms
def E(x):
    """Unwrap to sympy: Field -> its expression, sequences -> column Matrix, numbers -> exact rationals."""
    if isinstance(x, Field):
        return x.e
    if isinstance(x, (list, tuple)):
        return sp.Matrix([E(v) for v in x])
    if isinstance(x, float):
        return rat(x)
    return x if isinstance(x, sp.MatrixBase) else sp.sympify(x)


class _Vec:
    def __init__(self, f):
        self.f = f

    def set(self, v):
        self.f.e = E(v)


class Field:
    """What the reference code receives as a dolfinx Function / TestFunction: arithmetic yields plain sympy objects."""

    def __init__(self, e, coeffs=None, values=None, mesh=None):
        self.e, self.coeffs, self.values = e, coeffs, values
        self.function_space = types.SimpleNamespace(mesh=mesh)
        self.vector = _Vec(self)

    def __len__(self):
        return self.e.rows

    def __getitem__(self, k):
        return self.e[k]

    def _b(op):
        def f(self, o):
            return getattr(E(self), op)(E(o))
        return f
    __add__, __radd__, __sub__, __rsub__ = _b('__add__'), _b('__radd__'), _b('__sub__'), _b('__rsub__')
    __mul__, __rmul__, __truediv__, __rtruediv__, __pow__ = _b('__mul__'), _b('__rmul__'), _b('__truediv__'), _b('__rtruediv__'), _b('__pow__')

    def __neg__(self):
        return -self.e


class Restricted:
    """`expr("+")` / `expr("-")` of UFL (load_defs rewrites that call syntax into restricted(expr, side))."""

    def __init__(self, e, side):
        self.e, self.side = E(e), side

    def __mul__(self, measure):
        assert measure.kind == 'dS'
        return Form([(Measure('dS' + self.side, measure.tag), self.e)])


class Measure:
    def __init__(self, kind, tag=None):
        self.kind, self.tag = kind, tag

    def __call__(self, tag):
        return Measure(self.kind, tag)

    def __rmul__(self, expr):
        return Form([(self, E(expr))])


class Form:
    def __init__(self, terms):
        self.terms = terms

    def __add__(self, o):
        if isinstance(o, Form):
            return Form(self.terms + o.terms)
        if o == 0:
            return self
        return NotImplemented

    __radd__ = __add__

    def __neg__(self):
        return Form([(m, -e) for m, e in self.terms])

    def __sub__(self, o):
        return self + (-o)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, c):
        return Form([(m, E(c) * e) for m, e in self.terms])

    __rmul__ = __mul__

    def value(self):
        """Exact integral over the current cell (dx terms) and its tagged exterior facets (ds terms)."""
        tot = 0
        for m, e in self.terms:
            if m.kind == 'dx':
                if m.tag is None or m.tag == CTX.cell_tag:
                    tot += CTX.integrate(e)
            elif m.kind == 'ds':
                for f in CTX.facets:
                    if not f.interior and (m.tag is None or m.tag == f.tag):
                        tot += CTX.integrate_facet(f, e)
            elif m.kind == 'dS+':
                # an interior facet is visited from both of its cells; the current cell plays the "+" side on its visit and
                # the other cell's visit supplies the "-" term (the reference's forms restrict the SAME expression both ways)
                for f in CTX.facets:
                    if f.interior and (m.tag is None or m.tag == f.tag):
                        tot += CTX.integrate_facet(f, e)
            elif m.kind == 'dS-':
                pass
            else:
                raise NotImplementedError(m.kind)
        return tot


# ------------------------------------------------------------------------------------------------- UFL operators
def grad(f):
    e = E(f)
    g = lambda s: CTX.JinvT * sp.Matrix([sp.diff(s, x) for x in CTX.xi])      # noqa: E731
    if isinstance(e, sp.MatrixBase):
        return sp.Matrix.vstack(*[g(e[i]).T for i in range(e.rows)])
    return g(e)


def div(f):
    e = E(f)
    if e.cols == 1:
        G = grad(e)
        return sum(G[i, i] for i in range(e.rows))
    return sp.Matrix([sum(grad(e[i, j])[j] for j in range(e.cols)) for i in range(e.rows)])


def inner(a, b):
    a, b = E(a), E(b)
    if isinstance(a, sp.MatrixBase):
        return sum(a[i] * b[i] for i in range(len(a)))
    return a * b


def dot(a, b):
    a, b = E(a), E(b)
    am, bm = isinstance(a, sp.MatrixBase), isinstance(b, sp.MatrixBase)
    if not am or not bm:
        return a * b
    if a.cols == 1 and b.cols == 1:
        return a.dot(b)
    if a.cols == 1:
        return (a.T * b).T
    return a * b


def derivative(expr, u, v=None):
    e = E(expr)
    d = lambda s: sum(sp.diff(s, c) * w for c, w in zip(u.coeffs, v.values))  # noqa: E731
    return e.applyfunc(d) if isinstance(e, sp.MatrixBase) else d(e)


def Constant(mesh, value):
    return E(value) if not isinstance(value, (list, tuple)) else sp.Matrix([E(v) for v in value])


NUMERIC = {}                    # symbol -> value: where `conditional` picks its branch
_MEMO = {}


def _memo(op, m):
    """The motor forms invert / take the determinant of the same deformation gradient hundreds of times per cell."""
    key = (op, sp.ImmutableMatrix(m))
    if key not in _MEMO:
        if len(_MEMO) > 4096:
            _MEMO.clear()
        _MEMO[key] = getattr(m, op)()
    return _MEMO[key]



def conditional(cond, a, b):
    op, l, r = cond
    assert op == 'lt'
    return E(a) if float(sp.N((E(l) - E(r)).xreplace(NUMERIC), 30)) < 0 else E(b)


NAMESPACE = dict(
    conditional=conditional, lt=lambda a, b: ('lt', a, b), restricted=Restricted,
    grad=grad, div=div, inner=inner, dot=dot, derivative=derivative, Constant=Constant,
    dx=Measure('dx'), ds=Measure('ds'), dS=Measure('dS'),
    FacetNormal=lambda mesh: sp.Matrix(NSYM[:CTX.d]), CellDiameter=lambda mesh: HSYM, SpatialCoordinate=lambda mesh: CTX.x,
    Identity=lambda d: sp.eye(d), tr=lambda a: E(a).trace(), det=lambda a: _memo('det', E(a)), inv=lambda a: _memo('inv', E(a)),
    sqrt=lambda a: sp.sqrt(E(a)), exp=lambda a: sp.exp(E(a)), as_vector=lambda a: sp.Matrix([E(v) for v in a]),
    ufl=types.SimpleNamespace(ds=Measure('ds'), dx=Measure('dx'), pi=sp.pi, sin=lambda a: sp.sin(E(a)), cos=lambda a: sp.cos(E(a)),
                              sqrt=lambda a: sp.sqrt(E(a)), dot=dot),
)


def load_defs(path, extra=None):
    """The top-level function definitions and plain constant assignments of a reference script, executed in the sympy-UFL
    namespace; nothing else of the script runs."""
    import numpy as np
    ns = dict(NAMESPACE, np=np)
    ns.update(extra or {})
    class Restrict(ast.NodeTransformer):          # X("+") -> restricted(X, "+"): sympy expressions are not callable
        def visit_Call(self, node):
            self.generic_visit(node)
            if (len(node.args) == 1 and not node.keywords and isinstance(node.args[0], ast.Constant)
                    and node.args[0].value in ('+', '-')):
                return ast.copy_location(ast.Call(ast.Name('restricted', ast.Load()), [node.func, node.args[0]], []), node)
            return node
    tree = ast.fix_missing_locations(Restrict().visit(ast.parse(open(path).read())))
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            exec(compile(ast.Module([node], []), path, 'exec'), ns)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id.isupper() for t in node.targets):
            try:
                exec(compile(ast.Module([node], []), path, 'exec'), ns)
            except Exception:
                pass
    return ns


def jac_numeric(exprs, syms, table):
    """d exprs / d syms at the rational point `table` by 80-digit central differences of step 1e-25 (the expressions of the
    motor families are too large for sympy.diff): truncation error ~1e-50 x the third derivative."""
    import mpmath as mp
    keys = list(table)
    f = sp.lambdify(keys, [exactify(e) for e in exprs], 'mpmath', cse=True)
    with mp.workdps(80):
        x0 = [mp.mpf(int(table[k].p)) / mp.mpf(int(table[k].q)) for k in keys]
        h = mp.mpf(10) ** -25
        cols = []
        for s_ in syms:
            j = keys.index(s_)
            xp, xm = list(x0), list(x0)
            xp[j] += h
            xm[j] -= h
            cols.append([float((a - b) / (2 * h)) for a, b in zip(f(*xp), f(*xm))])
    import numpy as np
    return np.array(cols, dtype=np.float64).T
