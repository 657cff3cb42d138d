"""Forms of examples/poisson_opt/run_poisson_opt.py (config 1).

  pdeRes(u, v, f)              :64-71   int grad(u).grad(v) - f v dx
  outputForm(u, f, u_exact)    :74-76   int 1/2 (u-u_exact)^2 + ALPHA/2 f^2 dx
"""
from ..fea.fem import Form
from ..fea.family import FormFamily
from .. import engine as _E

ALPHA = 1e-6


def _family(u, f):
    return FormFamily.get(_E.FAMILY_POISSON_P1, u.function_space.mesh, u, [f], params=[ALPHA])


def pdeRes(u, v, f, u_exact=None, weak_bc=False, sym=False):
    if weak_bc:
        raise NotImplementedError('poisson family: the weak (Nitsche) variant lives in forms.nonlinear_poisson')
    return Form(_family(u, f), 'residual')


def outputForm(u, f, u_exact, alpha=ALPHA):
    fam = _family(u, f)
    fam.set_param(0, alpha)
    fam.set_aux(0, u_exact)
    return Form(fam, 'output', out_id=0)
