"""Forms of examples/em_motor_opt/motor_pde.py (config 5b): nonlinear magnetostatics on a moving mesh.

  pdeResEM(u, v, uhat, iq, dx, p, s, Hc, vacuum_perm, angle, g, nitsche=True, sym=True, ds_=ds)   :90-130
  JS(v, uhat, iq, p, s, Hc, angle)                                                              :46-87
  B_power_form(A_z, uhat, n, dx, subdomains)                                                     :186-197
`dx` is the Measure whose subdomain_data carries the cell tags (meshtags of the cells).  The piecewise
mu_r(|B|) of RelativePermeability (:12-35) uses the coefficients in bh_fit.json (tests/golden/make_bh_fit.py).
The mesh-motion family (pdeResMM :134-183) is not implemented yet.
"""
import json
import os

import numpy as np

from ..fea.fem import Form
from ..fea.family import FormFamily
from .. import engine as _E

_FIT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'bh_fit.json')))
JS_SCALE = 6       # parameter slot of the source scaling (csrc/families4.cuh, EM_JS_SCALE)


def em_params(Hc, p, s, vacuum_perm, angle, iq, beta=1e4, js_scale=1.0, exponents=(2.0, 1.76835)):
    return [vacuum_perm, Hc, iq, angle, p, s, js_scale, beta, _FIT['x1'], _FIT['x2']] + list(_FIT['lin']) + \
        list(_FIT['cubic']) + list(_FIT['exp']) + list(exponents)


def pdeResEM(u, v, uhat, iq, dx, p, s, Hc, vacuum_perm, angle, g=None, nitsche=False, sym=False, overpenalty=False,
             ds_=None):
    if not (nitsche and sym):
        raise NotImplementedError('motor EM family: only the symmetric-Nitsche form of run_motor_opt.py:279-281')
    tags = dx.subdomain_data
    if tags is None:
        raise ValueError('pdeResEM: dx must carry the subdomain (cell) tags')
    cell_tags = np.zeros(u.function_space.mesh.num_cells, dtype=np.int32)
    cell_tags[tags.indices] = tags.values
    fam = FormFamily.get(_E.FAMILY_MOTOR_EM, u.function_space.mesh, u, [uhat],
                         params=em_params(Hc, p, s, vacuum_perm, angle, iq))
    fam.cell_tags = cell_tags
    return Form(fam, 'residual')


class JS:
    """Handle on the source term of a residual: `set_scale(t)` ramps magnets and windings
    (solveIncrementalEM, run_motor_opt.py:231-250)."""

    def __init__(self, residual_form):
        self.fam = residual_form.fam

    def set_scale(self, t):
        self.fam.params[JS_SCALE] = float(t)
        if self.fam._prob is not None:
            self.fam._prob.set_param(JS_SCALE, float(t))


def B_power_form(A_z, uhat, n, dx, subdomains):
    fam = next(iter(A_z.__dict__.get('_femo_families', {}).values()), None)
    if fam is None:
        raise ValueError('B_power_form: build pdeResEM(...) first')
    if sorted(subdomains) != [1, 2]:
        raise NotImplementedError('B_power_form: the engine integrates over the steel subdomains [1, 2]')
    if abs(n - fam.params[19]) < 1e-12:
        return Form(fam, 'output', out_id=0)
    fam.set_param(20, float(n))
    return Form(fam, 'output', out_id=1)
