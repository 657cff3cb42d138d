"""Forms of the reference's shell examples (examples/test_shell_m3l/shell_pde.py:219-311) for a FLAT mid-surface:
Reissner-Mindlin plate, CG2 deflection x CG1^2 rotations, reduced shear integration, penalty clamp.

The reference builds these forms from the un-vendored package shell_analysis_fenicsx (ShellElement "CG2CG1",
MaterialModel, ElasticModel.weakFormResidual, shell_pde.py:225-253); the engine implements the published
formulation restated in oracle/rm_plate.py (parity unpinned).  `ShellPDE` keeps the names the example uses:

  pde.W, pde.VT, pde.VF                              :229-232   state / thickness / load spaces
  pde.pdeRes(h, w, f, E, nu, penalty=True, dss=...)  :246-253   residual (linear problem)
  pde.compliance(w, h)                               :281-282   1/2 int w^2 dx
  pde.mass(h, rho)                                   :287-288   rho int h dx
  pde.elastic_energy(w, h, E)                        :290-293   bending + shear energy
"""
from ..fea.fem import Form, FunctionSpace
from ..fea.family import FormFamily
from .. import engine as _E

PENALTY = 1.0e8


class ShellPDE:
    def __init__(self, mesh):
        if mesh.cell_type != 'triangle':
            raise NotImplementedError('ShellPDE: the plate family has kernels for triangle meshes')
        self.mesh = mesh
        self.W = FunctionSpace(mesh, ('RMPlate', 1))       # [w vertices | w edge midpoints | (theta_x, theta_y) per vertex]
        self.VT = FunctionSpace(mesh, ('CG', 1))
        self.VF = FunctionSpace(mesh, ('CG', 1))           # transverse load (the plate's component of the example's force)
        self._fam = None

    def pdeRes(self, h, w, f, E, nu, penalty=True, dss=None, dSS=None, g=None, pen=PENALTY):
        if g is not None:
            raise NotImplementedError('ShellPDE.pdeRes: only the homogeneous clamp (g=None)')
        tagged = None
        if dss is not None and getattr(dss, 'subdomain_data', None) is not None:
            tagged = dss.facets()
        elif not penalty:
            tagged = []                                    # no clamped facets at all
        self._fam = FormFamily.get(_E.FAMILY_RM_PLATE, self.mesh, w, [h, f],
                                   params=[float(E), float(nu), float(pen), 1.0], tagged=tagged)
        return Form(self._fam, 'residual')

    def _family(self):
        if self._fam is None:
            raise ValueError('ShellPDE: build pdeRes(...) first')
        return self._fam

    def compliance(self, u_mid=None, h=None, dxx=None):
        return Form(self._family(), 'output', out_id=0)

    def mass(self, h, rho):
        fam = self._family()
        fam.set_param(3, float(rho))
        return Form(fam, 'output', out_id=1)

    def elastic_energy(self, w=None, h=None, E=None):
        return Form(self._family(), 'output', out_id=2)
