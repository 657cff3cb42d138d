"""Forms of examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py (config 3):
Euler-Bernoulli cantilever, Hermite-3 state, DG0 thickness.

  pdeRes(u, v, t, f, dss, E, width)   :71-73   int v'' (E b t^3/12) u'' dx - f v dss
  volume(t, width, L)                 :75-76   int t b L dx
  compliance(u, f, dss)               :78-79   f u dss
`dss` is the tagged point measure ds_(100) of the tip (:113-124).
"""
from ..fea.fem import Form
from ..fea.family import FormFamily
from .. import engine as _E


def _fval(f):
    return float(getattr(f, 'value', f))


def _family(u, t, f=None, dss=None, E=1.0, width=0.1, L=1.0):
    tagged = dss.facets() if dss is not None else None
    fam = FormFamily.get(_E.FAMILY_EB_BEAM, u.function_space.mesh, u, [t],
                         params=[float(E), float(width), float(L), -1.0 if f is None else _fval(f)], tagged=tagged)
    return fam


def pdeRes(u, v, t, f, dss, E, width):
    fam = _family(u, t, f, dss, E, width)
    fam.set_param(0, float(E))
    fam.set_param(1, float(width))
    fam.set_param(3, _fval(f))
    t.__dict__['_femo_family_of_input'] = fam
    return Form(fam, 'residual')


def compliance(u, f, dss=None, t=None):
    """Output 0.  `t` is only needed when the residual form was not built first."""
    fam = next(iter(u.__dict__.get('_femo_families', {}).values()), None) if t is None else _family(u, t, f, dss)
    if fam is None:
        raise ValueError('compliance: build pdeRes(...) first or pass t=')
    return Form(fam, 'output', out_id=0)


def volume(t, width, L, u=None):
    """Output 1: needs the state to find the family (`u=`), or uses the family registered on `t`."""
    fam = t.__dict__.get('_femo_family_of_input')
    if u is not None:
        fam = next(iter(u.__dict__.get('_femo_families', {}).values()), fam)
    if fam is None:
        raise ValueError('volume: build pdeRes(...) first')
    fam.set_param(1, float(width))
    fam.set_param(2, float(L))
    return Form(fam, 'output', out_id=1)
