"""Forms of examples/em_motor_opt/motor_pde.py (config 5b): nonlinear magnetostatics on a moving mesh.

  pdeResEM(u, v, uhat, iq, dx, p, s, Hc, vacuum_perm, angle, g, nitsche=True, sym=True, ds_=ds)   :90-130
  JS(v, uhat, iq, p, s, Hc, angle)                                                              :46-87
  B_power_form(A_z, uhat, n, dx, subdomains)                                                     :186-197
`dx` is the Measure whose subdomain_data carries the cell tags (meshtags of the cells).  The piecewise
mu_r(|B|) of RelativePermeability (:12-35) uses the coefficients in bh_fit.json (tests/golden/make_bh_fit.py).
The mesh-motion family (pdeResMM :134-183, area_form :199-210) follows below (csrc/families5.cuh).
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
    fam = next((f for f in A_z.__dict__.get('_femo_families', {}).values() if f.family_id == _E.FAMILY_MOTOR_EM), None)
    if fam is None:
        raise ValueError('B_power_form: build pdeResEM(...) first')
    if sorted(subdomains) != [1, 2]:
        raise NotImplementedError('B_power_form: the engine integrates over the steel subdomains [1, 2]')
    if abs(n - fam.params[19]) < 1e-12:
        return Form(fam, 'output', out_id=0)
    fam.set_param(20, float(n))
    return Form(fam, 'output', out_id=1)


# --------------------------------------------------------------------------------------------
# mesh motion (config 5a): pdeResMM :134-183, area_form :199-210
# --------------------------------------------------------------------------------------------
class SideMeasure:
    """dS(tag) / ds(tag) stand-in carrying the tagged facets as ONE-SIDED (cell, local facet) pairs:
    an interior facet contributes one pair per side, exactly how the "+" / "-" restrictions of the
    reference's Nitsche terms act (motor_pde.py:170-178)."""

    def __init__(self, cells, locals_):
        self.sides = (np.asarray(cells, dtype=np.int32), np.asarray(locals_, dtype=np.int32))

    def __call__(self, tag):
        return self


def annulus_circle_sides(mesh, ir):
    """One-sided facets on the circle of radial node index `ir` of an annulus mesh (both sides inside)."""
    nr, nth = mesh._e.shape
    cells, locs = [], []
    for it in range(nth):
        if ir < nr:
            cells.append(2 * (ir * nth + it) + 1)
            locs.append(2)
        if ir > 0:
            cells.append(2 * ((ir - 1) * nth + it))
            locs.append(0)
    return np.asarray(cells, dtype=np.int32), np.asarray(locs, dtype=np.int32)


def pdeResMM(uhat, duhat, g=None, nitsche=False, sym=False, overpenalty=False, dS_=None, ds_=None, cell_tags=None):
    if not (nitsche and sym) or g is None:
        raise NotImplementedError('mesh-motion family: the symmetric-Nitsche form of run_motor_opt.py:190-192')
    sides = [m.sides for m in (dS_, ds_) if m is not None]
    if not sides:
        raise ValueError('pdeResMM: pass the tagged facet measures dS_(1000) / ds_(1000)')
    fc = np.concatenate([s[0] for s in sides])
    fl = np.concatenate([s[1] for s in sides])
    o = np.lexsort((fl, fc))
    fam = FormFamily.get(_E.FAMILY_MOTOR_MM, uhat.function_space.mesh, uhat, [g], params=[5e3])
    fam.facets = (fc[o], fl[o])
    if cell_tags is not None:
        fam.cell_tags = np.asarray(cell_tags, dtype=np.int32)
    return Form(fam, 'residual')


def area_form(uhat, dx, subdomains):
    """int det(F) dx over a subdomain group: [15] winding, [3] magnet, [1, 2] steel (run_motor_opt.py:176-181)."""
    fam = next((f for f in uhat.__dict__.get('_femo_families', {}).values() if f.family_id == _E.FAMILY_MOTOR_MM), None)
    if fam is None:
        raise ValueError('area_form: build pdeResMM(...) first')
    ids = sorted([subdomains] if isinstance(subdomains, int) else list(subdomains))
    table = {(15,): 0, (3,): 1, (1, 2): 2}
    if tuple(ids) not in table:
        raise NotImplementedError('area_form: subdomain groups [15], [3] and [1, 2] have kernels')
    if fam.cell_tags is None and dx is not None and dx.subdomain_data is not None:
        tags = np.zeros(uhat.function_space.mesh.num_cells, dtype=np.int32)
        tags[dx.subdomain_data.indices] = dx.subdomain_data.values
        fam.cell_tags = tags
    return Form(fam, 'output', out_id=table[tuple(ids)])


def synthetic_motor_tags(mesh, p=12, s=36):
    """Subdomain ids of the synthetic annulus standing in for the reference's motor mesh,
    with the ids of its association table: 51 shaft | 1 rotor core | magnets 3..14 with air (52) between them |
    53 air gap | windings 15..50 with stator teeth (2) between them | 2 stator yoke; bands are fractions of the
    radial extent."""
    X = mesh.geometry.x[:, :2]
    c = X[mesh.cells].mean(axis=1)
    r = np.hypot(c[:, 0], c[:, 1])
    th = np.mod(np.arctan2(c[:, 1], c[:, 0]), 2 * np.pi)
    r0, r1 = mesh._e.lo[0], mesh._e.hi[0]
    f = (r - r0) / (r1 - r0)
    tag = np.full(mesh.num_cells, 52, dtype=np.int32)
    tag[f < 0.15] = 51
    tag[(f >= 0.15) & (f < 0.35)] = 1
    band = (f >= 0.35) & (f < 0.5)
    sec = th / (2 * np.pi / p)
    k = np.floor(sec).astype(int)
    sel = band & (sec - k < 0.75)
    tag[sel] = (3 + k)[sel]
    tag[(f >= 0.5) & (f < 0.58)] = 53
    band = (f >= 0.58) & (f < 0.8)
    sec = th / (2 * np.pi / s)
    k = np.floor(sec).astype(int)
    tag[band] = 2
    sel = band & (sec - k < 0.6)
    tag[sel] = (15 + k)[sel]
    tag[f >= 0.8] = 2
    return tag
