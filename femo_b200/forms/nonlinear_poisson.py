"""Forms of examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py (config 2).

  pdeRes(u, v, f, u_exact, weak_bc=True, sym=True)   :118-125 (interior :88-95, Nitsche :97-116)
  outputForm(u, f, u_exact)                           :140-142
u_exact is the UFL expression sin(2 pi x) sin(pi y) of :144-145; the engine
evaluates it analytically at the quadrature points (token `U_EX_UFL`).
"""
from ..fea.fem import Form
from ..fea.family import FormFamily
from .. import engine as _E

ALPHA_1 = 6E-7
BETA = 1e1
U_EX_UFL = 'sin(2*pi*x[0])*sin(pi*x[1])'

# expressions for project(): u_ex and f_ex = -div(grad(u_ex)) + u_ex**3 (run_nonlinear_poisson_opt.py:165-169)
from ..fea.fem import Expr  # noqa: E402
u_ex_ufl = Expr('u_ex')
f_ex_ufl = Expr('f_ex')


def _family(u, f):
    # V = CG1 as in the example script, or CG2 (the "P2" variant of BASELINE.json configs[1])
    fid = _E.FAMILY_NLPOISSON_P2 if u.function_space.degree == 2 else _E.FAMILY_NLPOISSON_P1
    return FormFamily.get(fid, u.function_space.mesh, u, [f], params=[ALPHA_1, BETA])


def pdeRes(u, v, f, u_exact=U_EX_UFL, weak_bc=True, sym=True, overPenalize=False):
    if not (weak_bc and sym) or u_exact != U_EX_UFL:
        raise NotImplementedError('nonlinear Poisson family: only the symmetric-Nitsche form of the example '
                                  '(weak_bc=True, sym=True, u_exact=U_EX_UFL) has device kernels')
    return Form(_family(u, f), 'residual')


def outputForm(u, f, u_exact=U_EX_UFL, alpha=ALPHA_1):
    fam = _family(u, f)
    fam.set_param(0, alpha)
    return Form(fam, 'output', out_id=0)
