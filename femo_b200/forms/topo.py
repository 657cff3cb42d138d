"""Forms of examples/beam_topo_opt/run_topo_opt_cantilever_beam.py (config 4):
SIMP linear elasticity on Q1 quadrilaterals (or, as the 3-D extension, trilinear hexahedra), vector CG1
state, DG0 density.

  pdeRes(u, v, rho_e, f, dss=ds_(100), method='SIMP')   :62-77
  averageFunc(func)                                       :79-83   int rho/|Omega| dx
  compliance(u, f, dss)                                   :85-86   int u.f dss
"""
import numpy as np

from ..fea.fem import Form
from ..fea.family import FormFamily
from .. import engine as _E

NU = 0.3


def _fvec(f):
    return np.asarray(getattr(f, 'value', f), dtype=np.float64).ravel()


def pdeRes(u, v, rho_e, f, E=1, dss=None, method='SIMP'):
    if method != 'SIMP':
        raise NotImplementedError("topology family: only method='SIMP' (E = rho^3) has device kernels")
    fv = _fvec(f)
    mesh = u.function_space.mesh
    if mesh.cell_type == 'hexahedron':          # 3-D extension
        fam = FormFamily.get(_E.FAMILY_SIMP_HEX8, mesh, u, [rho_e], params=[NU, fv[0], fv[1], fv[2], 3.0],
                             tagged=None if dss is None else dss.facets())
    else:
        fam = FormFamily.get(_E.FAMILY_SIMP_Q1, mesh, u, [rho_e],
                             params=[NU, fv[0], fv[1], 3.0], tagged=None if dss is None else dss.facets())
    rho_e.__dict__['_femo_family_of_input'] = fam
    return Form(fam, 'residual')


def averageFunc(func):
    fam = func.__dict__.get('_femo_family_of_input')
    if fam is None:
        raise ValueError('averageFunc: build pdeRes(...) with this density first')
    return Form(fam, 'output', out_id=0)


def compliance(u, f, dss=None):
    fam = next(iter(u.__dict__.get('_femo_families', {}).values()), None)
    if fam is None:
        raise ValueError('compliance: build pdeRes(...) first')
    return Form(fam, 'output', out_id=1)
