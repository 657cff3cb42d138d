"""Reissner-Mindlin plate family restated in numpy (TEST INFRASTRUCTURE ONLY).

The reference's shell examples (/root/reference/examples/test_shell_m3l/shell_pde.py:219-311) build their forms from
the un-vendored package `shell_analysis_fenicsx` (ShellElement "CG2CG1", MaterialModel, ElasticModel.weakFormResidual,
shell_pde.py:225-253): a Reissner-Mindlin formulation with quadratic mid-surface displacements, linear rotations,
reduced integration of the transverse-shear energy and penalty boundary conditions.  That package is absent from
/root/reference, so this file restates the PUBLISHED formulation for a flat mid-surface (the plate limit of the shell):

    U(w, theta; t) = 1/2 int D(t) [ (1-nu) kappa:kappa + nu tr(kappa)^2 ] dx          bending, kappa = sym grad(theta)
                   + 1/2 int ks G t |grad(w) - theta|^2 dx_reduced                      transverse shear, ks = 5/6
                   + 1/2 pen int_{Gamma_c} ( w^2 + theta.theta ) ds                     penalty clamp (shell_pde.py:41,244)
    R(v, eta)      = dU[(v, eta)] - int f v dx                                          (linear_problem = True, :41)
    D(t) = E t^3 / (12 (1 - nu^2)),  G = E / (2 (1 + nu))

with w in CG2 (scalar), theta in CG1^2, thickness t in CG1 (shell_pde.py:231 `pde.VT`), load f in CG1 (:232, the
transverse component of `pde.VF`).  Outputs follow shell_pde.py:281-311: compliance 1/2 int w^2 dx (:281-282), mass
rho int t dx (:287-288), elastic energy = the bending + shear energy (:290-293).

PARITY UNPINNED: neither dolfinx nor shell_analysis_fenicsx is installable offline and the reference holds no
vectors for this path; quadrature choices are this file's contract (bending + load: 6-point degree-4 rule; shear:
3-point degree-2 rule = "reduced" against the degree-3 integrand; penalty: 5-point Gauss).  The oracle is pinned by
finite differences of its own energy, symmetry / rigid-body checks and the Kirchhoff thin-plate limit of a clamped
square plate (tests/test_rm_plate.py).

Local dof order (12): w at vertices 0..2, w at the midpoints of the edges opposite vertices 0..2, then
(theta_x, theta_y) per vertex.  Global: [w vertices | w edges | theta interleaved per vertex].
"""
import numpy as np

from . import quadrature as quad
from .mesh import triangle_edges
from .families import _TriP1

KS = 5.0 / 6.0


class RMPlate(_TriP1):
    name = 'rm_plate'
    n_outputs = 3

    def __init__(self, mesh, clamped=None, E=1.0e4, nu=0.3, pen=1.0e8, rho=1.0):
        """clamped: indices into mesh.exterior_facets() carrying the penalty clamp (None = all exterior facets)."""
        super().__init__(mesh)
        self.E, self.nu, self.pen, self.rho = float(E), float(nu), float(pen), float(rho)
        nv = mesh.nverts
        ev, ce = triangle_edges(mesh)
        self.nedges = ev.shape[0]
        c = mesh.cells
        th = (nv + self.nedges + 2 * c).astype(np.int64)
        self.cell_dofs = np.concatenate([c, nv + ce, np.stack([th[:, 0], th[:, 0] + 1, th[:, 1], th[:, 1] + 1,
                                                               th[:, 2], th[:, 2] + 1], axis=1)], axis=1).astype(np.int32)
        self.N = 3 * nv + self.nedges
        self.in_dofs = c.astype(np.int32)            # CG1 inputs (thickness, load)
        self.M = nv
        fc, fl = mesh.exterior_facets()
        if clamped is not None:
            idx = np.asarray(clamped, dtype=np.int64)
            fc, fl = fc[idx], fl[idx]
        self.fc, self.fl = fc, fl

    # -- tabulation ---------------------------------------------------------------------------------------
    @staticmethod
    def _p2(l):
        """P2 basis values (nq,6) at barycentric points l (nq,3)."""
        ph = np.empty((l.shape[0], 6))
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            ph[:, i] = l[:, i] * (2.0 * l[:, i] - 1.0)
            ph[:, 3 + i] = 4.0 * l[:, j] * l[:, k]
        return ph

    def _p2_grad(self, l):
        """Physical gradients (nc,nq,6,2) of the P2 basis."""
        G = self.G                                   # (nc,3,2) P1 gradients
        out = np.empty((G.shape[0], l.shape[0], 6, 2))
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            out[:, :, i] = (4.0 * l[:, i] - 1.0)[None, :, None] * G[:, None, i]
            out[:, :, 3 + i] = 4.0 * (l[None, :, j, None] * G[:, None, k] + l[None, :, k, None] * G[:, None, j])
        return out

    @staticmethod
    def _bary(pts):
        return np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1)

    # -- internal force K(t) u (and its thickness derivative) ----------------------------------------------
    def _bending_B(self):
        """kappa = (k_xx, k_yy, 2 k_xy) = B theta_e with theta_e the 6 rotation dofs; B is (nc,3,6), constant per cell."""
        G = self.G
        B = np.zeros((G.shape[0], 3, 6))
        for a in range(3):
            B[:, 0, 2 * a] = G[:, a, 0]
            B[:, 1, 2 * a + 1] = G[:, a, 1]
            B[:, 2, 2 * a] = G[:, a, 1]
            B[:, 2, 2 * a + 1] = G[:, a, 0]
        return B

    def _Cb(self):
        nu = self.nu
        return np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]])

    def _element_matrix(self, t, order=0):
        """K_e(t) (order 0) or its derivative wrt the nodal thickness t_a (order 1 -> (nc,3,12,12))."""
        nc = self.mesh.ncells
        te = t[self.in_dofs]                                     # (nc,3)
        Db = self.E / (12.0 * (1.0 - self.nu ** 2))
        Gs = KS * self.E / (2.0 * (1.0 + self.nu))
        B = self._bending_B()
        BCB = np.einsum('cki,kl,clj->cij', B, self._Cb(), B)     # (nc,6,6)
        K = np.zeros((nc, 12, 12)) if order == 0 else np.zeros((nc, 3, 12, 12))
        # bending, 6-point rule: D(t_q) = Db t_q^3
        pts, w = quad.triangle(4)
        l = self._bary(pts)
        for q in range(len(w)):
            tq = te @ l[q]
            wq = w[q] * self.detJ
            if order == 0:
                K[:, 6:, 6:] += (wq * Db * tq ** 3)[:, None, None] * BCB
            else:
                for a in range(3):
                    K[:, a, 6:, 6:] += (wq * Db * 3.0 * tq ** 2 * l[q, a])[:, None, None] * BCB
        # shear, 3-point rule: gamma = grad(w) - theta = Bs u_e, Bs (nc,2,12)
        pts, w = quad.triangle(2)
        l = self._bary(pts)
        gp = self._p2_grad(l)                                    # (nc,nq,6,2)
        for q in range(len(w)):
            Bs = np.zeros((nc, 2, 12))
            Bs[:, 0, :6] = gp[:, q, :, 0]
            Bs[:, 1, :6] = gp[:, q, :, 1]
            for a in range(3):
                Bs[:, 0, 6 + 2 * a] = -l[q, a]
                Bs[:, 1, 7 + 2 * a] = -l[q, a]
            BB = np.einsum('cki,ckj->cij', Bs, Bs)
            tq = te @ l[q]
            wq = w[q] * self.detJ
            if order == 0:
                K += (wq * Gs * tq)[:, None, None] * BB
            else:
                for a in range(3):
                    K[:, a] += (wq * Gs * l[q, a])[:, None, None] * BB
        return K

    def _load_matrix(self):
        """L_e[i,b] = int phi^w_i phi^f_b dx (6 x 3), degree-4 rule (exact)."""
        pts, w = quad.triangle(4)
        l = self._bary(pts)
        ph = self._p2(l)
        L = np.einsum('q,qi,qb->ib', w, ph, l)
        return self.detJ[:, None, None] * L[None]

    def _penalty(self):
        """Facet blocks: pen int (w v + theta.eta) ds on the clamped facets -> (dofs (nf,12), K_f (nf,12,12))."""
        s, ws = quad.gauss_legendre_01(5)
        fc, fl = self.fc, self.fl
        lf = self.mesh.local_facets[fl]
        P, Q = self.X[fc, lf[:, 0]], self.X[fc, lf[:, 1]]
        flen = np.linalg.norm(Q - P, axis=1)
        nf = fc.size
        Kf = np.zeros((nf, 12, 12))
        ar = np.arange(nf)
        for q in range(len(ws)):
            l = np.zeros((nf, 3))
            l[ar, lf[:, 0]] = 1.0 - s[q]
            l[ar, lf[:, 1]] = s[q]
            ph = np.zeros((nf, 6))
            for i in range(3):
                j, k = (i + 1) % 3, (i + 2) % 3
                ph[:, i] = l[:, i] * (2.0 * l[:, i] - 1.0)
                ph[:, 3 + i] = 4.0 * l[:, j] * l[:, k]
            N = np.zeros((nf, 3, 12))                            # rows: w, theta_x, theta_y
            N[:, 0, :6] = ph
            for a in range(3):
                N[:, 1, 6 + 2 * a] = l[:, a]
                N[:, 2, 7 + 2 * a] = l[:, a]
            Kf += (ws[q] * flen * self.pen)[:, None, None] * np.einsum('fki,fkj->fij', N, N)
        return self.cell_dofs[fc], Kf

    # -- family interface -------------------------------------------------------------------------------------
    def residual(self, u, t, f):
        ue = u[self.cell_dofs]
        Re = np.einsum('cij,cj->ci', self._element_matrix(t), ue)
        Re[:, :6] -= np.einsum('cib,cb->ci', self._load_matrix(), f[self.in_dofs])
        fd, Kf = self._penalty()
        return [(self.cell_dofs, None, Re), (fd, None, np.einsum('fij,fj->fi', Kf, u[fd]))]

    def jacobian(self, u, t, f):
        fd, Kf = self._penalty()
        return [(self.cell_dofs, self.cell_dofs, self._element_matrix(t)), (fd, fd, Kf)]

    def dRdm(self, slot, u, t, f):
        if slot == 0:                                            # thickness
            dK = self._element_matrix(t, order=1)                # (nc,3,12,12)
            De = np.einsum('caij,cj->cia', dK, u[self.cell_dofs])
            return [(self.cell_dofs, self.in_dofs, De)]
        De = np.zeros((self.mesh.ncells, 12, 3))                 # load
        De[:, :6, :] = -self._load_matrix()
        return [(self.cell_dofs, self.in_dofs, De)]

    def output(self, k, u, t, f):
        if k == 0:                                               # compliance 1/2 int w^2 dx
            pts, w = quad.triangle(4)
            ph = self._p2(self._bary(pts))
            wq = u[self.cell_dofs[:, :6]] @ ph.T                 # (nc,nq)
            return [0.5 * self.detJ * (wq ** 2 @ w)]
        if k == 1:                                               # mass rho int t dx
            return [self.rho * self.detJ / 6.0 * t[self.in_dofs].sum(axis=1)]
        ue = u[self.cell_dofs]                                   # elastic energy 1/2 u K u
        return [0.5 * np.einsum('ci,cij,cj->c', ue, self._element_matrix(t), ue)]

    def output_du(self, k, u, t, f):
        nc = self.mesh.ncells
        ge = np.zeros((nc, 12))
        if k == 0:
            pts, w = quad.triangle(4)
            ph = self._p2(self._bary(pts))
            wq = u[self.cell_dofs[:, :6]] @ ph.T
            ge[:, :6] = self.detJ[:, None] * np.einsum('q,cq,qi->ci', w, wq, ph)
        elif k == 2:
            ge = np.einsum('cij,cj->ci', self._element_matrix(t), u[self.cell_dofs])
        return [(self.cell_dofs, None, ge)]

    def output_dm(self, k, slot, u, t, f):
        nc = self.mesh.ncells
        ge = np.zeros((nc, 3))
        if slot == 0 and k == 1:
            ge[:] = (self.rho * self.detJ / 6.0)[:, None]
        elif slot == 0 and k == 2:
            ue = u[self.cell_dofs]
            ge = 0.5 * np.einsum('ci,caij,cj->ca', ue, self._element_matrix(t, order=1), ue)
        return [(self.in_dofs, None, ge)]
