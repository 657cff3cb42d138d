"""Stand-ins for the slice of CSDL / python_csdl_backend the femo operations use.

csdl, python_csdl_backend and modopt are not installable offline (SURVEY.md
section 8c).  If csdl is importable its classes are used; otherwise these
minimal bases provide `parameters.declare`, `add_input/add_output`,
`declare_derivatives`, `Model.add/create_input/...` and a `Simulator` that
calls the five operation callbacks in the backend's order:
run() -> solve_residual_equations / compute; compute_totals() ->
compute_derivatives, apply_inverse_jacobian('rev'), compute_jacvec_product('rev').
"""
import numpy as np

from .. import _hostops as _H

try:                                    # pragma: no cover - csdl is absent in this image
    from csdl import Model, CustomImplicitOperation, CustomExplicitOperation   # noqa: F401
    import csdl                                                                 # noqa: F401
    HAVE_CSDL = True
except Exception:
    HAVE_CSDL = False

    class _Parameters(dict):
        def declare(self, name, default=None, types=None, **kw):
            self.setdefault(name, default)

    class _Op:
        def __init__(self, **kwargs):
            self.parameters = _Parameters()
            self.initialize()
            self.parameters.update(kwargs)
            self.input_meta, self.output_meta = {}, {}
            self.define()

        def initialize(self):
            pass

        def add_input(self, name, shape=(1,), val=1.0):
            self.input_meta[name] = dict(shape=shape, val=val)

        def add_output(self, name, shape=(1,), val=1.0):
            self.output_meta[name] = dict(shape=shape, val=val)

        def declare_derivatives(self, of, wrt, **kw):
            pass

    class CustomImplicitOperation(_Op):
        pass

    class CustomExplicitOperation(_Op):
        pass

    class _Var:
        def __init__(self, name, shape, val):
            self.name, self.shape, self.val = name, shape, val

    class Model:
        def __init__(self, **kwargs):
            self.parameters = _Parameters()
            self.initialize()
            self.parameters.update(kwargs)
            self.ops = []                # (op, [input names], [output names]) in execution order
            self.inputs = {}             # created inputs: name -> value
            self.declared = {}
            self.design_variables, self.objectives, self.constraints = {}, {}, {}
            self._defined = False

        def initialize(self):
            pass

        def define(self):
            pass

        def _ensure_defined(self):
            if not self._defined:
                self._defined = True
                self.define()

        def declare_variable(self, name, shape=(1,), val=1.0):
            v = _Var(name, shape, val)
            self.declared[name] = v
            return v

        def create_input(self, name, shape=(1,), val=1.0):
            shape = (shape,) if np.isscalar(shape) else tuple(shape)
            self.inputs[name] = np.broadcast_to(np.asarray(val, dtype=np.float64), shape).copy()
            return _Var(name, shape, val)

        def register_output(self, name, var):
            if isinstance(var, _CustomResult):
                self.ops.append((var.op, list(var.arg_names), name))
            return _Var(name, None, None)

        def print_var(self, var):
            pass

        def add(self, submodel, name=None):
            submodel._ensure_defined()
            self.ops += submodel.ops
            for k, v in submodel.inputs.items():
                self.inputs.setdefault(k, v)
            for k, v in submodel.declared.items():
                self.declared.setdefault(k, v)

        def add_design_variable(self, name, lower=None, upper=None, scaler=None):
            self.design_variables[name] = dict(lower=lower, upper=upper, scaler=scaler)

        def add_objective(self, name, scaler=None):
            self.objectives[name] = dict(scaler=scaler)

        def add_constraint(self, name, lower=None, upper=None, equals=None, scaler=None):
            self.constraints[name] = dict(lower=lower, upper=upper, equals=equals, scaler=scaler)

    class _CustomResult:
        def __init__(self, op, arg_names):
            self.op, self.arg_names = op, arg_names

    class _Csdl:
        """csdl.custom(*args, op=...) -> handle consumed by Model.register_output."""
        @staticmethod
        def custom(*args, op=None):
            return _CustomResult(op, [a.name for a in args])

    csdl = _Csdl()


class _Plain(np.ndarray):
    """Plain pageable variable storage (no version tracking): `bump` is a no-op so the simulator code is shared."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def bump(self):
        pass


class Simulator:
    """python_csdl_backend.Simulator stand-in for models made of femo operations.

    Variables live in one flat (promoted) namespace, as in the reference's
    examples; `sim['submodel.var']` resolves by the last path component.
    """

    def __init__(self, model, analytics=False, pinned=True, **kw):
        """pinned=False keeps every variable in plain pageable, untracked numpy arrays -- what python_csdl_backend
        hands to the operations -- instead of page-locked, version-tracked storage (bench.py reports both)."""
        model._ensure_defined()
        self.model = model
        self.pinned = bool(pinned)
        self.vars = {k: self._store(np.asarray(v, dtype=np.float64)) for k, v in model.inputs.items()}
        for op, args, out in model.ops:
            for a in args:
                if a not in self.vars:
                    shape = op.input_meta[a]['shape']
                    d = model.declared.get(a)
                    val = d.val if d is not None else 1.0
                    self.vars[a] = self._store(np.broadcast_to(np.asarray(val, dtype=np.float64), shape))
            if out not in self.vars:
                self.vars[out] = self._store(np.zeros(op.output_meta[out]['shape']))
        self._work = {}

    def _store(self, value):
        """Variable storage: page-locked for large vectors so the operations' uploads are plain DMA."""
        if not getattr(self, 'pinned', True):
            return _Plain(np.array(value, dtype=np.float64, copy=True))
        buf = _H.pinned_empty(value.size).reshape(value.shape)
        np.copyto(buf, value)
        return _H.tracked(buf)

    def _buffer(self, tag, like):
        buf = self._work.get(tag)
        if buf is None or buf.shape != like.shape:
            buf = self._work[tag] = (_H.pinned_empty(like.size) if self.pinned else np.empty(like.size)).reshape(like.shape)
        return buf

    def _scratch(self, tag, like):
        """Zeroed work array reused across calls (the backend's d_inputs / d_residuals vectors)."""
        buf = self._buffer(tag, like)
        _H.fill(buf, 0.0)
        return buf

    @staticmethod
    def _key(name):
        return name.split('.')[-1]

    def __getitem__(self, name):
        v = self.vars[self._key(name)]
        v.bump()                                 # a writable reference leaves the simulator: assume it is written
        return v

    def __setitem__(self, name, value):
        k = self._key(name)
        value = np.asarray(value, dtype=np.float64)
        if value.size == self.vars[k].size:
            _H.copy(self.vars[k], value)
        else:
            np.copyto(self.vars[k], np.broadcast_to(value, self.vars[k].shape))
        self.vars[k].bump()

    def run(self):
        for op, args, out in self.model.ops:
            inputs = {a: self.vars[a] for a in args}
            outputs = {out: self.vars[out]}          # the operation reads the initial guess, then assigns
            if isinstance(op, CustomImplicitOperation):
                op.solve_residual_equations(inputs, outputs)
            else:
                op.compute(inputs, outputs)
            res = np.asarray(outputs[out], dtype=np.float64).reshape(self.vars[out].shape)
            if res is not self.vars[out]:
                _H.copy(self.vars[out], res)         # CSDL copies on assignment
            self.vars[out].bump()

    def compute_totals(self, of, wrt):
        """Reverse-mode totals d(of)/d(wrt) for scalar `of` (the adjoint chain of
        SURVEY.md section 3.2-3.4).  The returned arrays are the simulator's own accumulation buffers:
        valid until the next compute_totals call."""
        of = [of] if isinstance(of, str) else list(of)
        wrt = [wrt] if isinstance(wrt, str) else list(wrt)
        totals = {}
        lin = {}
        for o in of:
            bar = {self._key(o): np.ones(1)}
            for op, args, out in reversed(self.model.ops):
                if out not in bar:
                    continue
                inputs = {a: self.vars[a] for a in args}
                if isinstance(op, CustomImplicitOperation):
                    outputs = {out: self.vars[out]}
                    if id(op) not in lin:
                        op.compute_derivatives(inputs, outputs, {})
                        lin[id(op)] = True
                    d_res = {}                               # assigned by the operation
                    op.apply_inverse_jacobian({out: bar[out]}, d_res, 'rev')
                    d_in = {a: self._scratch('in:%s:%s' % (o, a), self.vars[a]) for a in args}
                    op.compute_jacvec_product(inputs, outputs, d_in, {}, {out: np.asarray(d_res[out])}, 'rev')
                    for a in args:
                        self._accumulate(bar, o, a, d_in[a], -1.0)
                elif hasattr(op, 'vjp'):
                    # vector-valued explicit operation with a constant Jacobian (declared sparse in CSDL)
                    for a in args:
                        self._accumulate(bar, o, a, op.vjp(out, a, np.asarray(bar[out])), 1.0)
                else:
                    derivs = {}
                    op.compute_derivatives(inputs, derivs)
                    seed = float(np.ravel(bar[out])[0])
                    for a in args:
                        self._accumulate(bar, o, a, np.ravel(derivs[out, a]), seed)
            for w in wrt:
                k = self._key(w)
                totals[(o, w)] = bar[k] if k in bar else np.zeros_like(self.vars[k])
        return totals

    def _accumulate(self, bar, o, a, g, scale):
        """bar[a] (+)= scale * g in a reusable buffer (one per (output, variable))."""
        g = np.ravel(np.asarray(g, dtype=np.float64))
        if a in bar:
            _H.iadd(bar[a], g, scale)
        else:
            buf = self._buffer('bar:%s:%s' % (o, a), g)
            _H.scaled_copy(buf, g, scale)
            bar[a] = buf

    def check_totals(self, of, wrt, step=1e-6, directions=3, seed=0, compact_print=True):
        """Directional finite-difference check of compute_totals (the reference's
        only derivative validation, sim.check_totals)."""
        rng = np.random.default_rng(seed)
        self.run()
        tot = self.compute_totals(of, wrt)
        report = {}
        for (o, w), g in tot.items():
            x0 = self.vars[self._key(w)].copy()
            errs = []
            for _ in range(directions):
                d = rng.standard_normal(x0.shape)
                self[w] = x0 + step * d
                self.run()
                fp = float(np.ravel(self[o])[0])
                self[w] = x0 - step * d
                self.run()
                fm = float(np.ravel(self[o])[0])
                fd = (fp - fm) / (2 * step)
                an = float(g.ravel() @ d.ravel())
                errs.append(abs(fd - an) / max(abs(fd), 1e-300))
            self[w] = x0
            self.run()
            report[(o, w)] = max(errs)
            if compact_print:
                print('check_totals d%s/d%s: max rel err %.3e' % (o, w, report[(o, w)]))
        return report
