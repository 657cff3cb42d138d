"""Form families of the reference's examples, restated as explicit quadrature
loops the way FFCx would generate them (TEST INFRASTRUCTURE ONLY).

Every family exposes integral blocks (see assembly.py) for
  residual(u, m...)          R                       state_model.py:85
  jacobian(u, m...)          dR/du                   state_model.py:129-132
  dRdm(slot, u, m...)        dR/dm                   state_model.py:136-141
  output(k, ...) / output_du / output_dm             output_model.py:69-87
with the quadrature degrees of SURVEY.md Appendix A.3.  Gateaux derivatives
(UFL `derivative`, utils_dolfinx.py:313-314) are derived by hand and checked
against finite differences / sympy in tests/test_oracle_*.py.
"""
import numpy as np
from . import quadrature as quad
from .mesh import dofmap


# --------------------------------------------------------------------------
# P1 triangle geometry
# --------------------------------------------------------------------------
class _TriP1:
    """Affine triangle geometry + P1 tabulation shared by the 2-D families."""

    def __init__(self, mesh):
        assert mesh.kind == 'triangle'
        self.mesh = mesh
        X = mesh.coords[mesh.cells]                       # (nc,3,2)
        self.X = X
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)   # (nc,2,2) columns = edges
        det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        self.detJ = np.abs(det)                           # = 2*area
        Jinv = np.empty_like(J)
        Jinv[:, 0, 0], Jinv[:, 0, 1] = J[:, 1, 1] / det, -J[:, 0, 1] / det
        Jinv[:, 1, 0], Jinv[:, 1, 1] = -J[:, 1, 0] / det, J[:, 0, 0] / det
        gref = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])        # d phi_a / d xi
        # physical gradients g[c,a,:] = Jinv^T gref[a]
        self.G = np.einsum('ckd,ak->cad', Jinv, gref)
        self.cell_dofs, self.N = dofmap(mesh, 'CG', 1)
        self.dg_dofs, self.M = dofmap(mesh, 'DG', 0)

    @staticmethod
    def phi(pts):
        return np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1)   # (nq,3)

    def xq(self, pts):
        return np.einsum('qa,cad->cqd', self.phi(pts), self.X)                          # (nc,nq,2)


# --------------------------------------------------------------------------
# config 1: Poisson source-term optimisation
# --------------------------------------------------------------------------
class PoissonP1(_TriP1):
    """examples/poisson_opt/run_poisson_opt.py:32-38 (residual), :74-76 (output).

    R = int grad(u).grad(v) - f v dx ;  J = int 1/2 (u-u_ex)^2 + alpha/2 f^2 dx
    u, u_ex in CG1, f in DG0, alpha = 1e-6 (:29,108).
    """
    name = 'poisson_p1'
    n_outputs = 1

    def __init__(self, mesh, alpha=1e-6, u_ex=None):
        super().__init__(mesh)
        self.alpha = alpha
        self.u_ex = np.zeros(self.N) if u_ex is None else np.asarray(u_ex, dtype=np.float64)

    def residual(self, u, f):
        pts, w = quad.triangle(1)
        ph = self.phi(pts)
        ue = u[self.cell_dofs]
        gu = np.einsum('ca,cad->cd', ue, self.G)
        Re = np.zeros((self.mesh.ncells, 3))
        for q in range(len(w)):
            wq = w[q] * self.detJ
            Re += wq[:, None] * (np.einsum('cd,cad->ca', gu, self.G) - f[:, None] * ph[q][None, :])
        return [(self.cell_dofs, None, Re)]

    def jacobian(self, u, f):
        pts, w = quad.triangle(0)
        Ae = np.zeros((self.mesh.ncells, 3, 3))
        for q in range(len(w)):
            Ae += (w[q] * self.detJ)[:, None, None] * np.einsum('cad,cbd->cab', self.G, self.G)
        return [(self.cell_dofs, self.cell_dofs, Ae)]

    def dRdm(self, slot, u, f):
        assert slot == 0
        pts, w = quad.triangle(1)
        ph = self.phi(pts)
        De = np.zeros((self.mesh.ncells, 3, 1))
        for q in range(len(w)):
            De[:, :, 0] -= (w[q] * self.detJ)[:, None] * ph[q][None, :]
        return [(self.cell_dofs, self.dg_dofs, De)]

    def output(self, k, u, f):
        pts, w = quad.triangle(2)
        ph = self.phi(pts)
        e = (u - self.u_ex)[self.cell_dofs]
        val = np.zeros(self.mesh.ncells)
        for q in range(len(w)):
            eq = e @ ph[q]
            val += w[q] * self.detJ * (0.5 * eq * eq + 0.5 * self.alpha * f * f)
        return [val]

    def output_du(self, k, u, f):
        pts, w = quad.triangle(2)
        ph = self.phi(pts)
        e = (u - self.u_ex)[self.cell_dofs]
        ge = np.zeros((self.mesh.ncells, 3))
        for q in range(len(w)):
            ge += (w[q] * self.detJ * (e @ ph[q]))[:, None] * ph[q][None, :]
        return [(self.cell_dofs, None, ge)]

    def output_dm(self, k, slot, u, f):
        pts, w = quad.triangle(0)
        ge = np.zeros((self.mesh.ncells, 1))
        for q in range(len(w)):
            ge[:, 0] += w[q] * self.detJ * self.alpha * f
        return [(self.dg_dofs, None, ge)]


# --------------------------------------------------------------------------
# config 2: nonlinear Poisson with symmetric Nitsche boundary terms
# --------------------------------------------------------------------------
def u_exact_nlp(x):
    """examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py:144-145."""
    return np.sin(2.0 * np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1])


class NonlinearPoissonP1(_TriP1):
    """examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py.

    interior (:88-95):  int grad(u).grad(v) + u^3 v - f v dx        (degree 4, 6 pts)
    boundary (:97-116, sym=True, beta=10):
        - int (grad(u).n) v ds + int (u_ex-u)(grad(v).n) ds
        + beta/h_E int (u-u_ex) v ds                                (degree 9, 5-pt Gauss)
    output (:140-142): int 1/2 (u-u_ex)^2 + alpha_1/2 f^2 dx, u_ex a UFL expression
        (degree 12 -> collapsed Gauss rule; same rule for dJ/du)
    """
    name = 'nlpoisson_p1'
    n_outputs = 1

    def __init__(self, mesh, alpha=6e-7, beta=10.0):
        super().__init__(mesh)
        self.alpha, self.beta = alpha, beta
        self.h = mesh.cell_diameter()
        fc, fl = mesh.exterior_facets()
        self.fc, self.fl = fc, fl
        lf = mesh.local_facets[fl]                                   # (nf,2) local vertex ids
        P = self.X[fc, lf[:, 0]]
        Q = self.X[fc, lf[:, 1]]
        O = self.X[fc, fl]                                           # opposite vertex
        t = Q - P
        self.flen = np.linalg.norm(t, axis=1)
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / self.flen[:, None]
        sgn = np.sign(np.einsum('fd,fd->f', nrm, P - O))
        self.fn = nrm * sgn[:, None]                                 # outward unit normal
        self.fP, self.fQ, self.flv = P, Q, lf
        self.fdofs = self.cell_dofs[fc]

    # -- facet tabulation -------------------------------------------------
    def _facet_q(self):
        s, w = quad.interval(9)
        nf = self.fc.size
        ph = np.zeros((nf, len(w), 3))
        ar = np.arange(nf)
        for q in range(len(w)):
            ph[ar, q, self.flv[:, 0]] = 1.0 - s[q]
            ph[ar, q, self.flv[:, 1]] = s[q]
        xq = self.fP[:, None, :] + s[None, :, None] * (self.fQ - self.fP)[:, None, :]
        return ph, xq, w

    def residual(self, u, f):
        pts, w = quad.triangle(4)
        ph = self.phi(pts)
        ue = u[self.cell_dofs]
        gu = np.einsum('ca,cad->cd', ue, self.G)
        Re = np.zeros((self.mesh.ncells, 3))
        for q in range(len(w)):
            uq = ue @ ph[q]
            wq = w[q] * self.detJ
            Re += wq[:, None] * (np.einsum('cd,cad->ca', gu, self.G)
                                 + (uq ** 3 - f)[:, None] * ph[q][None, :])
        # exterior facets
        phf, xq, wf = self._facet_q()
        uf = u[self.fdofs]
        Gf = self.G[self.fc]
        gn = np.einsum('fad,fd->fa', Gf, self.fn)                     # grad(phi_a).n
        dudn = np.einsum('fa,fa->f', uf, gn)
        bh = self.beta / self.h[self.fc]
        Rf = np.zeros((self.fc.size, 3))
        for q in range(len(wf)):
            wq = wf[q] * self.flen
            uq = np.einsum('fa,fa->f', uf, phf[:, q])
            ex = u_exact_nlp(xq[:, q])
            Rf += wq[:, None] * (-dudn[:, None] * phf[:, q] + (ex - uq)[:, None] * gn
                                 + (bh * (uq - ex))[:, None] * phf[:, q])
        return [(self.cell_dofs, None, Re), (self.fdofs, None, Rf)]

    def jacobian(self, u, f):
        pts, w = quad.triangle(4)
        ph = self.phi(pts)
        ue = u[self.cell_dofs]
        K = np.einsum('cad,cbd->cab', self.G, self.G)
        Ae = np.zeros((self.mesh.ncells, 3, 3))
        for q in range(len(w)):
            uq = ue @ ph[q]
            wq = w[q] * self.detJ
            Ae += wq[:, None, None] * (K + (3.0 * uq * uq)[:, None, None]
                                       * np.outer(ph[q], ph[q])[None])
        phf, xq, wf = self._facet_q()
        Gf = self.G[self.fc]
        gn = np.einsum('fad,fd->fa', Gf, self.fn)
        bh = self.beta / self.h[self.fc]
        Af = np.zeros((self.fc.size, 3, 3))
        for q in range(len(wf)):
            wq = wf[q] * self.flen
            p = phf[:, q]
            Af += wq[:, None, None] * (-p[:, :, None] * gn[:, None, :] - gn[:, :, None] * p[:, None, :]
                                       + bh[:, None, None] * p[:, :, None] * p[:, None, :])
        return [(self.cell_dofs, self.cell_dofs, Ae), (self.fdofs, self.fdofs, Af)]

    def dRdm(self, slot, u, f):
        assert slot == 0
        pts, w = quad.triangle(1)
        ph = self.phi(pts)
        De = np.zeros((self.mesh.ncells, 3, 1))
        for q in range(len(w)):
            De[:, :, 0] -= (w[q] * self.detJ)[:, None] * ph[q][None, :]
        return [(self.cell_dofs, self.dg_dofs, De)]

    def output(self, k, u, f):
        pts, w = quad.triangle(12)
        ph = self.phi(pts)
        xq = self.xq(pts)
        ue = u[self.cell_dofs]
        val = np.zeros(self.mesh.ncells)
        for q in range(len(w)):
            eq = ue @ ph[q] - u_exact_nlp(xq[:, q])
            val += w[q] * self.detJ * 0.5 * eq * eq
        val += 0.5 * self.detJ * 0.5 * self.alpha * f * f
        return [val]

    def output_du(self, k, u, f):
        pts, w = quad.triangle(12)
        ph = self.phi(pts)
        xq = self.xq(pts)
        ue = u[self.cell_dofs]
        ge = np.zeros((self.mesh.ncells, 3))
        for q in range(len(w)):
            eq = ue @ ph[q] - u_exact_nlp(xq[:, q])
            ge += (w[q] * self.detJ * eq)[:, None] * ph[q][None, :]
        return [(self.cell_dofs, None, ge)]

    def output_dm(self, k, slot, u, f):
        ge = (0.5 * self.detJ * self.alpha * f)[:, None]
        return [(self.dg_dofs, None, ge)]


# --------------------------------------------------------------------------
# config 3: Euler-Bernoulli cantilever, Hermite-3 on an interval
# --------------------------------------------------------------------------
def _hermite_d2(xi):
    """Second reference derivatives of the Hermite cubics, dof order
    (value@v0, slope@v0, value@v1, slope@v1); identity push-forward as in basix 0.5
    [upstream, SURVEY.md section 7 "Hermite push-forward"]: slope dofs are reference derivatives."""
    return np.stack([-6.0 + 12.0 * xi, -4.0 + 6.0 * xi, 6.0 - 12.0 * xi, -2.0 + 6.0 * xi], axis=-1)


def _hermite(xi):
    return np.stack([1 - 3 * xi ** 2 + 2 * xi ** 3, xi - 2 * xi ** 2 + xi ** 3,
                     3 * xi ** 2 - 2 * xi ** 3, -xi ** 2 + xi ** 3], axis=-1)


class EBBeam:
    """examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py.

    R = int v'' (E b t^3/12) u'' dx - f v|_{ds(100)}   (:64-79,127-131), degree 2 -> 2-pt Gauss
    outputs: 0 compliance = f u|_{ds(100)} (:84-85), 1 volume = int t b L dx (:81-82)
    `tagged` indexes the mesh's exterior-facet list (the tip, :115-124).
    """
    name = 'eb_beam'
    n_outputs = 2

    def __init__(self, mesh, tagged, E=1.0, width=0.1, L=1.0, f=-1.0):
        assert mesh.kind == 'interval'
        self.mesh = mesh
        self.E, self.width, self.L, self.f = E, width, L, f
        self.cell_dofs, self.N = dofmap(mesh, 'Hermite', 3)
        self.dg_dofs, self.M = dofmap(mesh, 'DG', 0)
        X = mesh.coords[mesh.cells][:, :, 0]
        self.h = X[:, 1] - X[:, 0]
        fc, fl = mesh.exterior_facets()
        self.fc, self.fl = fc[tagged], fl[tagged]
        self.fdofs = self.cell_dofs[self.fc]
        self.fphi = _hermite(self.fl.astype(np.float64))          # basis at the facet point (xi = 0 or 1)

    def _khat(self):
        s, w = quad.interval(2)
        d2 = _hermite_d2(s)                                       # (nq,4)
        return np.einsum('q,qa,qb->ab', w, d2, d2)

    def _bend(self, u):
        return np.einsum('ab,cb->ca', self._khat(), u[self.cell_dofs]) / self.h[:, None] ** 3

    def residual(self, u, t):
        EI = self.E * self.width * t ** 3 / 12.0
        Re = EI[:, None] * self._bend(u)
        Rf = -self.f * self.fphi
        return [(self.cell_dofs, None, Re), (self.fdofs, None, Rf)]

    def jacobian(self, u, t):
        EI = self.E * self.width * t ** 3 / 12.0
        Ae = (EI / self.h ** 3)[:, None, None] * self._khat()[None]
        return [(self.cell_dofs, self.cell_dofs, Ae)]

    def dRdm(self, slot, u, t):
        dEI = self.E * self.width * 3.0 * t ** 2 / 12.0
        return [(self.cell_dofs, self.dg_dofs, (dEI[:, None] * self._bend(u))[:, :, None])]

    def output(self, k, u, t):
        if k == 0:
            return [self.f * np.einsum('fa,fa->f', self.fphi, u[self.fdofs])]
        return [t * self.width * self.L * self.h]

    def output_du(self, k, u, t):
        if k == 0:
            return [(self.fdofs, None, self.f * self.fphi)]
        return [(self.cell_dofs, None, np.zeros((self.mesh.ncells, 4)))]

    def output_dm(self, k, slot, u, t):
        if k == 0:
            return [(self.dg_dofs, None, np.zeros((self.mesh.ncells, 1)))]
        return [(self.dg_dofs, None, (self.width * self.L * self.h)[:, None])]


# --------------------------------------------------------------------------
# config 4: SIMP linear elasticity, Q1 quadrilaterals, vector CG1 state
# --------------------------------------------------------------------------
class SimpQ1:
    """examples/beam_topo_opt/run_topo_opt_cantilever_beam.py.

    R = int sigma(u):eps(v) dx - int_{ds(100)} f.v ds          (:62-77)
        E = rho^3, nu = 0.3, lambda = E nu/((1+nu)(1-2nu)), mu = E/(2(1+nu)); f = (0,-1/4) (:101)
        cells degree 2 -> 2x2 Gauss (no degree reduction on quads); ds(100) degree 4 -> 3-pt Gauss
    outputs: 0 avg_density = int rho/|Omega| dx (:79-83), 1 compliance = int_{ds(100)} u.f ds (:85-86)
    """
    name = 'simp_q1'
    n_outputs = 2

    def __init__(self, mesh, tagged, nu=0.3, f=(0.0, -0.25), penal=3.0):
        assert mesh.kind == 'quadrilateral'
        self.mesh, self.nu, self.f, self.penal = mesh, nu, np.asarray(f, dtype=np.float64), penal
        self.cell_dofs, self.N = dofmap(mesh, 'Q', 1, block=2)
        self.dg_dofs, self.M = dofmap(mesh, 'DG', 0)
        self.X = mesh.coords[mesh.cells]                           # (nc,4,2)
        fc, fl = mesh.exterior_facets()
        self.fc, self.fl = fc[tagged], fl[tagged]
        self.fdofs = self.cell_dofs[self.fc]
        lf = mesh.local_facets[self.fl]
        P, Q = self.X[self.fc, lf[:, 0]], self.X[self.fc, lf[:, 1]]
        self.flen = np.linalg.norm(Q - P, axis=1)
        self.flv = lf
        self.area = self._cell_area()
        self.volume = self.area.sum()

    @staticmethod
    def _N(p):
        x, y = p[..., 0], p[..., 1]
        return np.stack([(1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y], axis=-1)

    @staticmethod
    def _dN(p):
        x, y = p[..., 0], p[..., 1]
        return np.stack([np.stack([-(1 - y), (1 - y), -y, y], axis=-1),
                         np.stack([-(1 - x), -x, (1 - x), x], axis=-1)], axis=-1)      # (...,4,2)

    def _geom(self, pt):
        dN = self._dN(pt)                                           # (4,2) d/dxi
        J = np.einsum('cad,ak->cdk', self.X, dN)                    # J[c,d,k] = dx_d/dxi_k
        det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        Jinv = np.empty_like(J)
        Jinv[:, 0, 0], Jinv[:, 0, 1] = J[:, 1, 1] / det, -J[:, 0, 1] / det
        Jinv[:, 1, 0], Jinv[:, 1, 1] = -J[:, 1, 0] / det, J[:, 0, 0] / det
        G = np.einsum('ckd,ak->cad', Jinv, dN)                      # physical gradients (nc,4,2)
        return G, np.abs(det)

    def _cell_area(self):
        pts, w = quad.square(2)
        return sum(w[q] * self._geom(pts[q])[1] for q in range(len(w)))

    def _khat(self):
        """Unit-modulus element stiffness (nc,8,8), local dof a*2+c."""
        lam = self.nu / ((1 + self.nu) * (1 - 2 * self.nu))
        mu = 1.0 / (2 * (1 + self.nu))
        pts, w = quad.square(2)
        K = np.zeros((self.mesh.ncells, 4, 2, 4, 2))
        I2 = np.eye(2)
        for q in range(len(w)):
            G, det = self._geom(pts[q])
            wq = (w[q] * det)[:, None, None, None, None]
            K += wq * (lam * np.einsum('cai,cbj->caibj', G, G)
                       + mu * (np.einsum('caj,cbi->caibj', G, G)
                               + np.einsum('cad,cbd->cab', G, G)[:, :, None, :, None] * I2[None, None, :, None, :]))
        return K.reshape(self.mesh.ncells, 8, 8)

    def _traction(self):
        s, w = quad.interval(4)
        nf = self.fc.size
        Fe = np.zeros((nf, 4, 2))
        ar = np.arange(nf)
        for q in range(len(w)):
            ph = np.zeros((nf, 4))
            ph[ar, self.flv[:, 0]] = 1.0 - s[q]
            ph[ar, self.flv[:, 1]] = s[q]
            Fe += (w[q] * self.flen)[:, None, None] * ph[:, :, None] * self.f[None, None, :]
        return Fe.reshape(nf, 8)

    def residual(self, u, rho):
        Re = (rho ** self.penal)[:, None] * np.einsum('cab,cb->ca', self._khat(), u[self.cell_dofs])
        return [(self.cell_dofs, None, Re), (self.fdofs, None, -self._traction())]

    def jacobian(self, u, rho):
        return [(self.cell_dofs, self.cell_dofs, (rho ** self.penal)[:, None, None] * self._khat())]

    def dRdm(self, slot, u, rho):
        De = (self.penal * rho ** (self.penal - 1))[:, None] * np.einsum('cab,cb->ca', self._khat(), u[self.cell_dofs])
        return [(self.cell_dofs, self.dg_dofs, De[:, :, None])]

    def output(self, k, u, rho):
        if k == 0:
            return [rho * self.area / self.volume]
        return [np.einsum('fa,fa->f', self._traction(), u[self.fdofs])]

    def output_du(self, k, u, rho):
        if k == 0:
            return [(self.cell_dofs, None, np.zeros((self.mesh.ncells, 8)))]
        return [(self.fdofs, None, self._traction())]

    def output_dm(self, k, slot, u, rho):
        if k == 0:
            return [(self.dg_dofs, None, (self.area / self.volume)[:, None])]
        return [(self.dg_dofs, None, np.zeros((self.mesh.ncells, 1)))]


class SimpHex8:
    """3-D extension of SimpQ1 on trilinear hexahedra: same residual
    R = int sigma(u):eps(v) dx - int_{ds(100)} f.v ds with E = rho^p, 2x2x2 Gauss points in the cells
    and 2x2 Gauss points on the traction faces; outputs 0 average density, 1 compliance.
    """
    name = 'simp_hex8'
    n_outputs = 2

    def __init__(self, mesh, tagged, nu=0.3, f=(0.0, -0.25, 0.0), penal=3.0):
        assert mesh.kind == 'hexahedron'
        self.mesh, self.nu, self.f, self.penal = mesh, nu, np.asarray(f, dtype=np.float64), penal
        self.cell_dofs, self.N = dofmap(mesh, 'Q', 1, block=3)
        self.dg_dofs, self.M = dofmap(mesh, 'DG', 0)
        self.X = mesh.coords[mesh.cells]                           # (nc,8,3)
        fc, fl = mesh.exterior_facets()
        self.fc, self.fl = fc[tagged], fl[tagged]
        self.fdofs = self.cell_dofs[self.fc]
        self.flv = mesh.local_facets[self.fl]                      # (nf,4)
        g = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])
        self.qp = np.array([[g[i], g[j], g[k]] for k in range(2) for j in range(2) for i in range(2)])
        self.qw = np.full(8, 0.125)
        self.vol = sum(self.qw[q] * self._geom(self.qp[q])[1] for q in range(8))
        self.volume = self.vol.sum()

    @staticmethod
    def _dN(p):
        out = np.zeros((8, 3))
        for a in range(8):
            b = [(a >> d) & 1 for d in range(3)]
            l = [p[d] if b[d] else 1.0 - p[d] for d in range(3)]
            s = [1.0 if b[d] else -1.0 for d in range(3)]
            out[a] = [s[0] * l[1] * l[2], s[1] * l[0] * l[2], s[2] * l[0] * l[1]]
        return out

    def _geom(self, pt):
        dN = self._dN(pt)
        J = np.einsum('cad,ak->cdk', self.X, dN)                    # dx_d/dxi_k
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
        G = np.einsum('ckd,ak->cad', Jinv, dN)                      # (nc,8,3)
        return G, np.abs(det)

    def _khat(self):
        lam = self.nu / ((1 + self.nu) * (1 - 2 * self.nu))
        mu = 1.0 / (2 * (1 + self.nu))
        K = np.zeros((self.mesh.ncells, 8, 3, 8, 3))
        I3 = np.eye(3)
        for q in range(8):
            G, det = self._geom(self.qp[q])
            wq = (self.qw[q] * det)[:, None, None, None, None]
            K += wq * (lam * np.einsum('cai,cbj->caibj', G, G)
                       + mu * (np.einsum('caj,cbi->caibj', G, G)
                               + np.einsum('cad,cbd->cab', G, G)[:, :, None, :, None] * I3[None, None, :, None, :]))
        return K.reshape(self.mesh.ncells, 24, 24)

    def _traction(self):
        nf = self.fc.size
        ar = np.arange(nf)
        Xf = self.X[self.fc[:, None], self.flv]                     # (nf,4,3)
        g = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])
        m = np.zeros((nf, 4))
        for t in g:
            for s_ in g:
                N = np.array([(1 - s_) * (1 - t), s_ * (1 - t), (1 - s_) * t, s_ * t])
                dNs = np.array([-(1 - t), (1 - t), -t, t])
                dNt = np.array([-(1 - s_), -s_, (1 - s_), s_])
                xs = np.einsum('fkd,k->fd', Xf, dNs)
                xt = np.einsum('fkd,k->fd', Xf, dNt)
                da = 0.25 * np.linalg.norm(np.cross(xs, xt), axis=1)
                m += da[:, None] * N[None, :]
        Fe = np.zeros((nf, 8, 3))
        for k in range(4):
            Fe[ar, self.flv[:, k]] += m[:, k, None] * self.f[None, :]
        return Fe.reshape(nf, 24)

    def residual(self, u, rho):
        Re = (rho ** self.penal)[:, None] * np.einsum('cab,cb->ca', self._khat(), u[self.cell_dofs])
        return [(self.cell_dofs, None, Re), (self.fdofs, None, -self._traction())]

    def jacobian(self, u, rho):
        return [(self.cell_dofs, self.cell_dofs, (rho ** self.penal)[:, None, None] * self._khat())]

    def dRdm(self, slot, u, rho):
        De = (self.penal * rho ** (self.penal - 1))[:, None] * np.einsum('cab,cb->ca', self._khat(), u[self.cell_dofs])
        return [(self.cell_dofs, self.dg_dofs, De[:, :, None])]

    def output(self, k, u, rho):
        if k == 0:
            return [rho * self.vol / self.volume]
        return [np.einsum('fa,fa->f', self._traction(), u[self.fdofs])]

    def output_du(self, k, u, rho):
        if k == 0:
            return [(self.cell_dofs, None, np.zeros((self.mesh.ncells, 24)))]
        return [(self.fdofs, None, self._traction())]

    def output_dm(self, k, slot, u, rho):
        if k == 0:
            return [(self.dg_dofs, None, (self.vol / self.volume)[:, None])]
        return [(self.dg_dofs, None, np.zeros((self.mesh.ncells, 1)))]


class NonlinearPoissonP2(_TriP1):
    """Config 2 with a quadratic Lagrange state (BASELINE.json configs[1] "P1/P2"):
    the forms of examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py:88-116,140-145 with V = CG2.

    cells:  int grad(u).grad(v) + u^3 v - f v dx       (degree 8 -> collapsed Gauss, exact to 9)
    facets: symmetric Nitsche terms, beta = 10          (6-pt Gauss)
    output: int 1/2 (u-u_ex)^2 + alpha/2 f^2 dx         (degree-12 rule, u_ex analytic)
    Basis in barycentric coordinates: phi_i = l_i(2 l_i - 1), phi_{3+i} = 4 l_j l_k (edge opposite vertex i).
    """
    name = 'nlpoisson_p2'
    n_outputs = 1

    def __init__(self, mesh, alpha=6e-7, beta=10.0):
        super().__init__(mesh)
        self.alpha, self.beta = alpha, beta
        self.cell_dofs, self.N = dofmap(mesh, 'CG', 2)
        self.h = mesh.cell_diameter()
        fc, fl = mesh.exterior_facets()
        self.fc, self.fl = fc, fl
        lf = mesh.local_facets[fl]
        P, Q, O = self.X[fc, lf[:, 0]], self.X[fc, lf[:, 1]], self.X[fc, fl]
        t = Q - P
        self.flen = np.linalg.norm(t, axis=1)
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / self.flen[:, None]
        self.fn = nrm * np.sign(np.einsum('fd,fd->f', nrm, P - O))[:, None]
        self.fP, self.fQ, self.flv = P, Q, lf
        self.fdofs = self.cell_dofs[fc]

    @staticmethod
    def _bary(pts):
        return np.stack([1.0 - pts[..., 0] - pts[..., 1], pts[..., 0], pts[..., 1]], axis=-1)

    @staticmethod
    def _phi(lam):
        """(...,3) barycentric -> (...,6) basis values."""
        out = [lam[..., i] * (2.0 * lam[..., i] - 1.0) for i in range(3)]
        out += [4.0 * lam[..., (i + 1) % 3] * lam[..., (i + 2) % 3] for i in range(3)]
        return np.stack(out, axis=-1)

    @staticmethod
    def _grad(lam, G):
        """lam (...,3) broadcastable against cells, G (nc,3,2) grad lambda -> (nc,...,6,2)."""
        g = []
        for i in range(3):
            g.append((4.0 * lam[..., i] - 1.0)[..., None] * G[:, i])
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            g.append(4.0 * (lam[..., j][..., None] * G[:, k] + lam[..., k][..., None] * G[:, j]))
        return np.stack(g, axis=-2)

    def _cells(self, u, f, want):
        pts, w = quad.triangle(8)
        ue = u[self.cell_dofs]
        nc = self.mesh.ncells
        Re, Ae = np.zeros((nc, 6)), np.zeros((nc, 6, 6))
        for q in range(len(w)):
            lam = self._bary(pts[q])
            ph = self._phi(lam)                                     # (6,)
            gp = self._grad(np.broadcast_to(lam, (nc, 3)), self.G)  # (nc,6,2)
            uq = ue @ ph
            wq = w[q] * self.detJ
            if want == 'R':
                gu = np.einsum('ca,cad->cd', ue, gp)
                Re += wq[:, None] * (np.einsum('cd,cad->ca', gu, gp) + (uq ** 3 - f)[:, None] * ph[None, :])
            else:
                Ae += wq[:, None, None] * (np.einsum('cad,cbd->cab', gp, gp)
                                           + (3.0 * uq * uq)[:, None, None] * np.outer(ph, ph)[None])
        return Re if want == 'R' else Ae

    def _facets(self, u, want):
        s, w = quad.interval(11)
        nf = self.fc.size
        ar = np.arange(nf)
        uf = u[self.fdofs]
        Gf = self.G[self.fc]
        bh = self.beta / self.h[self.fc]
        Rf, Af = np.zeros((nf, 6)), np.zeros((nf, 6, 6))
        for q in range(len(w)):
            lam = np.zeros((nf, 3))
            lam[ar, self.flv[:, 0]] = 1.0 - s[q]
            lam[ar, self.flv[:, 1]] = s[q]
            ph = self._phi(lam)                                     # (nf,6)
            gn = np.einsum('fad,fd->fa', self._grad(lam, Gf), self.fn)
            wq = w[q] * self.flen
            if want == 'R':
                uq = np.einsum('fa,fa->f', uf, ph)
                dudn = np.einsum('fa,fa->f', uf, gn)
                ex = u_exact_nlp(self.fP + s[q] * (self.fQ - self.fP))
                Rf += wq[:, None] * (-dudn[:, None] * ph + (ex - uq)[:, None] * gn + (bh * (uq - ex))[:, None] * ph)
            else:
                Af += wq[:, None, None] * (-ph[:, :, None] * gn[:, None, :] - gn[:, :, None] * ph[:, None, :]
                                           + bh[:, None, None] * ph[:, :, None] * ph[:, None, :])
        return Rf if want == 'R' else Af

    def residual(self, u, f):
        return [(self.cell_dofs, None, self._cells(u, f, 'R')), (self.fdofs, None, self._facets(u, 'R'))]

    def jacobian(self, u, f):
        return [(self.cell_dofs, self.cell_dofs, self._cells(u, f, 'J')), (self.fdofs, self.fdofs, self._facets(u, 'J'))]

    def dRdm(self, slot, u, f):
        pts, w = quad.triangle(2)
        De = np.zeros((self.mesh.ncells, 6, 1))
        for q in range(len(w)):
            De[:, :, 0] -= (w[q] * self.detJ)[:, None] * self._phi(self._bary(pts[q]))[None, :]
        return [(self.cell_dofs, self.dg_dofs, De)]

    def _out(self, u, f, want):
        pts, w = quad.triangle(12)
        ue = u[self.cell_dofs]
        xq = self.xq(pts)
        val, ge = np.zeros(self.mesh.ncells), np.zeros((self.mesh.ncells, 6))
        for q in range(len(w)):
            ph = self._phi(self._bary(pts[q]))
            eq = ue @ ph - u_exact_nlp(xq[:, q])
            val += w[q] * self.detJ * 0.5 * eq * eq
            ge += (w[q] * self.detJ * eq)[:, None] * ph[None, :]
        return val if want == 'v' else ge

    def output(self, k, u, f):
        return [self._out(u, f, 'v') + 0.5 * self.detJ * 0.5 * self.alpha * f * f]

    def output_du(self, k, u, f):
        return [(self.cell_dofs, None, self._out(u, f, 'g'))]

    def output_dm(self, k, slot, u, f):
        return [(self.dg_dofs, None, (0.5 * self.detJ * self.alpha * f)[:, None])]
