"""Hyperelastic mesh-motion family (config 5a) restated for the oracle (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/examples/em_motor_opt/motor_pde.py:134-183 (pdeResMM) and :199-210 (area_form):

    F = I + grad(uhat),  E = (F^T F - I)/2,  K = mu = det(F)^-3,
    S = K tr(E) I + 2 mu (E - tr(E) I/3) = det(F)^-3 (2 E + tr(E)/3 I),   P = F S
    R = int P : grad(v) dx                                   (f0 = -div P(g) vanishes for P1 g)
      + sum over the tagged facets dS(1000) ("+" and "-" side separately) and ds(1000):
          - (P n).v  +  (dP[v] n).(uhat - g)  +  beta/h_E v.(uhat - g),   beta = 5e3 / det(F)^3
    with dP[v] = derivative(P, uhat, v), n the outward normal of the side's own cell.

Each side of an interior facet only involves that side's cell, so a tagged interior facet is two
ONE-SIDED facets (cell, local facet) -- the same entity the exterior-facet integrals use.
dP[v] is coded analytically; dR/duhat and dR/dg are taken by complex-step differentiation of R
(independent of the engine's nested dual numbers).  Outputs: int det(F) dx over id sets.
"""
import numpy as np

from .mesh import dofmap


def circle_facets(mesh, ir):
    """One-sided facets (cell, local) on the circle of radial node index `ir` of an annulus mesh
    (oracle.motor.annulus_tri): both sides for interior circles, one side on the boundary."""
    nr, nth = mesh.shape
    cells, locs = [], []
    for it in range(nth):
        if ir < nr:        # cell above the circle: c1 = [v0, v2, v3] of (ir, it) has the edge v0-v2 (local facet 2)
            cells.append(2 * (ir * nth + it) + 1)
            locs.append(2)
        if ir > 0:         # cell below: c0 = [v0, v1, v3] of (ir-1, it) has the edge v1-v3 (local facet 0)
            cells.append(2 * ((ir - 1) * nth + it))
            locs.append(0)
    order = np.lexsort((locs, cells))
    return np.asarray(cells, dtype=np.int32)[order], np.asarray(locs, dtype=np.int32)[order]


class MotorMM:
    name = 'motor_mm'
    n_outputs = 3
    OUT_IDS = ((15,), (3,), (1, 2))          # winding_id, magnet_id, steel_id (run_motor_opt.py:68-70)

    def __init__(self, mesh, facets, tags, beta0=5e3):
        self.mesh, self.tags, self.beta0 = mesh, np.asarray(tags), beta0
        self.cell_dofs, self.N = dofmap(mesh, 'CG', 1, block=2)
        self.in_dofs, self.M = self.cell_dofs, self.N
        X = mesh.coords[mesh.cells]
        self.X = X
        Jm = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
        det = Jm[:, 0, 0] * Jm[:, 1, 1] - Jm[:, 0, 1] * Jm[:, 1, 0]
        self.area = 0.5 * np.abs(det)
        Ji = np.empty_like(Jm)
        Ji[:, 0, 0], Ji[:, 0, 1] = Jm[:, 1, 1] / det, -Jm[:, 0, 1] / det
        Ji[:, 1, 0], Ji[:, 1, 1] = -Jm[:, 1, 0] / det, Jm[:, 0, 0] / det
        gref = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
        self.G = np.einsum('ckd,ak->cad', Ji, gref)
        self.h = mesh.cell_diameter()
        fc, fl = facets
        self.fc, self.fl = np.asarray(fc), np.asarray(fl)
        lf = mesh.local_facets[self.fl]
        P, Q, O = X[self.fc, lf[:, 0]], X[self.fc, lf[:, 1]], X[self.fc, self.fl]
        t = Q - P
        self.flen = np.linalg.norm(t, axis=1)
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / self.flen[:, None]
        self.fn = nrm * np.sign(np.einsum('fd,fd->f', nrm, P - O))[:, None]
        self.flv = lf
        self.fdofs = self.cell_dofs[self.fc]

    # -- constitutive law ------------------------------------------------------
    @staticmethod
    def _F(G, uhe):
        return np.eye(2)[None] + np.einsum('cai,caj->cij', uhe, G)

    @staticmethod
    def _det(F):
        return F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]

    @classmethod
    def _P(cls, F):
        I = np.eye(2)[None]
        E = 0.5 * (np.einsum('cki,ckj->cij', F, F) - I)
        trE = E[:, 0, 0] + E[:, 1, 1]
        J = cls._det(F)
        S = (2.0 * E + (trE / 3.0)[:, None, None] * I) / (J ** 3)[:, None, None]
        return np.einsum('cik,ckj->cij', F, S)

    @classmethod
    def _dP(cls, F, dF):
        """Gateaux derivative of P at F in the direction dF (analytic)."""
        I = np.eye(2)[None]
        E = 0.5 * (np.einsum('cki,ckj->cij', F, F) - I)
        trE = E[:, 0, 0] + E[:, 1, 1]
        J = cls._det(F)
        Fi = np.empty_like(F)
        Fi[:, 0, 0], Fi[:, 0, 1] = F[:, 1, 1] / J, -F[:, 0, 1] / J
        Fi[:, 1, 0], Fi[:, 1, 1] = -F[:, 1, 0] / J, F[:, 0, 0] / J
        dE = 0.5 * (np.einsum('cki,ckj->cij', dF, F) + np.einsum('cki,ckj->cij', F, dF))
        dtr = dE[:, 0, 0] + dE[:, 1, 1]
        trFidF = np.einsum('cij,cji->c', Fi, dF)
        Jm3 = 1.0 / J ** 3
        core = 2.0 * E + (trE / 3.0)[:, None, None] * I
        S = Jm3[:, None, None] * core
        dS = (-3.0 * Jm3 * trFidF)[:, None, None] * core + Jm3[:, None, None] * (2.0 * dE + (dtr / 3.0)[:, None, None] * I)
        return np.einsum('cik,ckj->cij', dF, S) + np.einsum('cik,ckj->cij', F, dS)

    # -- element residuals ----------------------------------------------------------
    def _cell_res(self, uhe, ge):
        P = self._P(self._F(self.G, uhe))
        R = self.area[:, None, None] * np.einsum('cij,caj->cai', P, self.G)          # P : grad(phi_a e_i)
        return R.reshape(R.shape[0], 6)                                         # local dof 2a+i

    def _facet_res(self, uhe, ge):
        G = self.G[self.fc]
        F = self._F(G, uhe)
        P = self._P(F)
        J = self._det(F)
        n = self.fn
        nf = self.fc.size
        ar = np.arange(nf)
        Pn = np.einsum('fij,fj->fi', P, n)
        d = uhe - ge                                                             # (nf,3,2) nodal uhat - g
        la, lb = self.flv[:, 0], self.flv[:, 1]
        on = np.zeros((nf, 3))
        on[ar, la] = 1.0
        on[ar, lb] = 1.0
        wint = (0.5 * self.flen)[:, None] * (d[ar, la] + d[ar, lb])             # int (uhat - g) ds
        R = np.zeros((nf, 3, 2), dtype=P.dtype)
        R -= (0.5 * self.flen)[:, None, None] * on[:, :, None] * Pn[:, None, :]  # - (P n).v
        for a in range(3):
            for c in range(2):
                dF = np.zeros((nf, 2, 2), dtype=P.dtype)
                dF[:, c, :] = G[:, a, :]                                         # grad(phi_a e_c)
                dPn = np.einsum('fij,fj->fi', self._dP(F, dF), n)
                R[:, a, c] += np.einsum('fi,fi->f', dPn, wint)                   # (dP[v] n).(uhat - g)
        bh = self.beta0 / (J ** 3) / self.h[self.fc]
        # penalty: facet mass matrix [[1/3,1/6],[1/6,1/3]] * len on the two facet nodes
        m_aa = (self.flen / 3.0)[:, None]
        m_ab = (self.flen / 6.0)[:, None]
        pen = np.zeros((nf, 3, 2), dtype=P.dtype)
        pen[ar, la] = m_aa * d[ar, la] + m_ab * d[ar, lb]
        pen[ar, lb] = m_ab * d[ar, la] + m_aa * d[ar, lb]
        R += bh[:, None, None] * pen
        return R.reshape(nf, 6)

    def _loc(self, v, dofs):
        return v[dofs].reshape(-1, 3, 2)

    def residual(self, uh, g):
        return [(self.cell_dofs, None, self._cell_res(self._loc(uh, self.cell_dofs), None)),
                (self.fdofs, None, self._facet_res(self._loc(uh, self.fdofs), self._loc(g, self.fdofs)))]

    def _cs(self, fun, uhe, ge, wrt):
        hstep = 1e-30
        out = []
        for k in range(6):
            a = uhe.astype(complex)
            b = None if ge is None else ge.astype(complex)
            if wrt == 'u':
                a[:, k // 2, k % 2] += 1j * hstep
            else:
                b[:, k // 2, k % 2] += 1j * hstep
            out.append(np.asarray(fun(a, b)).imag / hstep)
        return np.stack(out, axis=-1)

    def jacobian(self, uh, g):
        return [(self.cell_dofs, self.cell_dofs, self._cs(self._cell_res, self._loc(uh, self.cell_dofs), None, 'u')),
                (self.fdofs, self.fdofs, self._cs(self._facet_res, self._loc(uh, self.fdofs), self._loc(g, self.fdofs), 'u'))]

    def dRdm(self, slot, uh, g):
        return [(self.fdofs, self.fdofs, self._cs(self._facet_res, self._loc(uh, self.fdofs), self._loc(g, self.fdofs), 'g'))]

    # -- outputs: areas of subdomain groups in the deformed configuration ------------------
    def _area_cells(self, k, uhe, ge=None):
        sel = np.isin(self.tags, self.OUT_IDS[k])
        return np.where(sel, self._det(self._F(self.G, uhe)) * self.area, 0.0)

    def output(self, k, uh, g):
        return [self._area_cells(k, self._loc(uh, self.cell_dofs)).real]

    def output_du(self, k, uh, g):
        return [(self.cell_dofs, None, self._cs(lambda a, b: self._area_cells(k, a), self._loc(uh, self.cell_dofs), None, 'u'))]

    def output_dm(self, k, slot, uh, g):
        return [(self.cell_dofs, None, np.zeros((self.mesh.ncells, 6)))]
