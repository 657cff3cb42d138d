"""L2 projection (`project`, /root/reference/femo/fea/utils_dolfinx.py:549-583) restated
for the oracle (TEST INFRASTRUCTURE ONLY): find u in V with int u w dx = int g w dx.

Targets: CG1 or DG0 on the triangle mesh.  Sources g: the analytic fields of
examples/nonlinear_poisson_opt (u_ex :144-145, f_ex = -div grad u_ex + u_ex^3 :167),
a DG0 function raised to a power (examples/beam_topo_opt:264-268) or a CG1 function.
Non-polynomial sources use the degree-12 collapsed Gauss rule (the oracle's contract,
SURVEY.md Appendix A.3); the reference solves with PETSc's default KSP (rtol 1e-5), the
oracle solves exactly.
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import quadrature as quad
from .families import _TriP1, u_exact_nlp
from . import assembly as asm


def f_exact_nlp(x):
    ue = u_exact_nlp(x)
    return 5.0 * np.pi ** 2 * ue + ue ** 3


class MassProjection(_TriP1):
    """residual R(u) = int (u - g) w dx, Jacobian = mass matrix."""

    def __init__(self, mesh, target='CG', source='u_ex', power=1.0):
        super().__init__(mesh)
        self.target, self.source, self.power = target, source, power
        self.tdofs = self.cell_dofs if target == 'CG' else self.dg_dofs
        self.N = self.cell_dofs.max() + 1 if target == 'CG' else self.M
        self.n_state = self.N

    def _basis(self, ph):
        return ph if self.target == 'CG' else np.ones((ph.shape[0], 1))

    def _g(self, xq, ph_q, src):
        if self.source == 'u_ex':
            return u_exact_nlp(xq)
        if self.source == 'f_ex':
            return f_exact_nlp(xq)
        if self.source == 'dg_pow':
            return src ** self.power
        return src[self.cell_dofs] @ ph_q                     # CG1 function

    def residual(self, u, src):
        pts, w = quad.triangle(12)
        ph = self.phi(pts)
        B = self._basis(ph)
        xq = self.xq(pts)
        ue = u[self.tdofs]
        Re = np.zeros((self.mesh.ncells, B.shape[1]))
        for q in range(len(w)):
            uq = ue @ B[q]
            Re += (w[q] * self.detJ * (uq - self._g(xq[:, q], ph[q], src)))[:, None] * B[q][None, :]
        return [(self.tdofs, None, Re)]

    def jacobian(self, u, src):
        pts, w = quad.triangle(2)
        B = self._basis(self.phi(pts))
        Ae = np.zeros((self.mesh.ncells, B.shape[1], B.shape[1]))
        for q in range(len(w)):
            Ae += (w[q] * self.detJ)[:, None, None] * np.outer(B[q], B[q])[None]
        return [(self.tdofs, self.tdofs, Ae)]

    def project(self, src=None):
        n = self.N
        src = np.zeros(self.M) if src is None else src
        A = asm.assemble_matrix(self.jacobian(np.zeros(n), src), (n, n))
        b = -asm.assemble_vector(self.residual(np.zeros(n), src), n)
        return spla.spsolve(A.tocsc(), b)
