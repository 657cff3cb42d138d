"""Tiny stand-in for modopt's SLSQP driver over a femo_b200 Simulator (modopt and
python_csdl_backend are not installable offline): scipy.optimize SLSQP with the
adjoint totals of the CSDL operations."""
import numpy as np
import scipy.optimize as so


def slsqp(sim, model, maxiter=100, ftol=1e-9, verbose=True):
    (dv, dv_opts), = model.design_variables.items()
    (obj, obj_opts), = model.objectives.items()
    scale = obj_opts.get('scaler') or 1.0
    x0 = np.array(sim[dv], dtype=np.float64)

    def f(x):
        sim[dv] = x
        sim.run()
        g = sim.compute_totals(obj, dv)[(obj, dv)]
        return scale * float(sim[obj][0]), scale * g

    cons = []
    for name, c in model.constraints.items():
        def cval(x, name=name):
            sim[dv] = x
            sim.run()
            return float(sim[name][0])

        def cjac(x, name=name):
            return sim.compute_totals(name, dv)[(name, dv)]
        if c.get('equals') is not None:
            cons.append(dict(type='eq', fun=lambda x, c=c, cval=cval: cval(x) - c['equals'], jac=cjac))
        if c.get('upper') is not None:
            cons.append(dict(type='ineq', fun=lambda x, c=c, cval=cval: c['upper'] - cval(x), jac=lambda x, cjac=cjac: -cjac(x)))
        if c.get('lower') is not None:
            cons.append(dict(type='ineq', fun=lambda x, c=c, cval=cval: cval(x) - c['lower'], jac=cjac))
    lo, hi = dv_opts.get('lower'), dv_opts.get('upper')
    bounds = None if lo is None and hi is None else [(lo, hi)] * x0.size
    res = so.minimize(f, x0, jac=True, method='SLSQP', bounds=bounds, constraints=cons,
                      options=dict(maxiter=maxiter, ftol=ftol, disp=verbose))
    sim[dv] = res.x
    sim.run()
    return res
