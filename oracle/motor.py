"""Nonlinear magnetostatic motor family (config 5b) restated for the oracle
(TEST INFRASTRUCTURE ONLY).

Follows /root/reference/examples/em_motor_opt/motor_pde.py:
  RelativePermeability :12-35 (piecewise linear / cubic / exponential mu_r(|B|) for the steel ids 1,2;
                                magnets 1.05; everything else 1), coefficients from
                                permeability/piecewise_permeability.py (femo_b200/forms/bh_fit.json)
  JS                   :46-87  (magnet H.curl(v) and three-phase winding current sources)
  pdeResEM             :90-130 (sum_i int nu_i gradx(u).gradx(v) J dx(i) - JS + symmetric Nitsche with the
                                Nanson-transformed normal on the exterior boundary, beta = 1e4, both
                                `boundary_components` use the steel curve, so the term enters twice)
  B_power_form         :186-197, area_form :199-210
and the ALE kinematics gradx / J / F of femo/fea/utils_dolfinx.py:34-66.

Everything is P1, so the cell integrand is constant except v*J (degree 1); the facet integrand is
quadratic in the facet coordinate (2-pt Gauss is exact).  The reference's meshes are git-LFS
pointers; the mesh here is a synthetic annulus (r in [0.06, 0.12]) with a band/sector tag layout of
the same ids.  Gateaux derivatives are taken by COMPLEX-STEP differentiation of the residual /
functionals (exact to round-off, independent of the engine's forward-mode dual numbers).
"""
import json
import os

import numpy as np

from .mesh import Mesh, dofmap
from . import quadrature as quad

_FIT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'femo_b200', 'forms', 'bh_fit.json')))
DOLFIN_EPS = 3e-16


def annulus_tri(nr, nth, r0=0.06, r1=0.12):
    """Periodic polar lattice: node (ir, ith) -> ir*nth + ith; cells 2*(ir*nth+ith)+{0,1} =
    [v0,v1,v3], [v0,v2,v3] with v1 = (ir+1,ith), v2 = (ir,ith+1), v3 = (ir+1,ith+1)."""
    ir, it = np.meshgrid(np.arange(nr + 1), np.arange(nth), indexing='ij')
    r = r0 + (r1 - r0) * ir / nr
    th = 2.0 * np.pi * it / nth
    coords = np.stack([(r * np.cos(th)).ravel(), (r * np.sin(th)).ravel()], axis=1)
    i, t = np.meshgrid(np.arange(nr), np.arange(nth), indexing='ij')
    tp = (t + 1) % nth
    v0, v1, v2, v3 = (i * nth + t).ravel(), ((i + 1) * nth + t).ravel(), (i * nth + tp).ravel(), ((i + 1) * nth + tp).ravel()
    cells = np.empty((2 * nr * nth, 3), dtype=np.int32)
    cells[0::2] = np.stack([v0, v1, v3], axis=1)
    cells[1::2] = np.stack([v0, v2, v3], axis=1)
    return Mesh('triangle', coords, cells, (nr, nth), (r0, 0.0), (r1, 2 * np.pi))


def motor_tags(mesh, p=12, s=36):
    """Synthetic subdomain layout with the ids of the reference's association table:
    51 shaft | 1 rotor core | magnets 3..14 (air 52 between them) | 53 air gap | windings 15..50 with
    stator teeth (2) between them | 2 stator yoke.  Bands are fractions of the radial extent."""
    c = mesh.coords[mesh.cells].mean(axis=1)
    r = np.hypot(c[:, 0], c[:, 1])
    th = np.mod(np.arctan2(c[:, 1], c[:, 0]), 2 * np.pi)
    r0, r1 = mesh.lo[0], mesh.hi[0]
    f = (r - r0) / (r1 - r0)
    tag = np.full(mesh.ncells, 52, dtype=np.int32)
    tag[f < 0.15] = 51
    tag[(f >= 0.15) & (f < 0.35)] = 1
    band = (f >= 0.35) & (f < 0.5)
    sec = th / (2 * np.pi / p)
    k = np.floor(sec).astype(int)
    tag[band & (sec - k < 0.75)] = (3 + k)[band & (sec - k < 0.75)]
    tag[(f >= 0.5) & (f < 0.58)] = 53
    band = (f >= 0.58) & (f < 0.8)
    sec = th / (2 * np.pi / s)
    k = np.floor(sec).astype(int)
    tag[band] = 2
    tag[band & (sec - k < 0.6)] = (15 + k)[band & (sec - k < 0.6)]
    tag[f >= 0.8] = 2
    return tag


def mu_r_steel(nb):
    """RelativePermeability for subdomains 1 and 2 (motor_pde.py:15-26); nb may be complex."""
    lin = _FIT['lin'][0] * nb + _FIT['lin'][1]
    a, b, c, d = _FIT['cubic']
    cub = a * nb ** 3 + b * nb ** 2 + c * nb + d
    ea, eb, ec = _FIT['exp']
    ex = ea * np.exp(eb * nb + ec) + 1.0
    return np.where(nb.real < _FIT['x1'], lin, np.where(nb.real < _FIT['x2'], cub, ex))


class MotorEM:
    name = 'motor_em'
    n_outputs = 2

    def __init__(self, mesh, tags, Hc=838e3, p=12, s=36, vacuum_perm=4e-7 * np.pi, angle=0.0, iq=282.2 / 0.00016231,
                 beta=1e4, js_scale=1.0, exponents=(2.0, 1.76835)):
        self.mesh, self.tags = mesh, np.asarray(tags)
        self.Hc, self.p, self.s, self.mu0, self.angle, self.iq, self.beta = Hc, p, s, vacuum_perm, angle, iq, beta
        self.js_scale, self.exponents = js_scale, exponents
        self.cell_dofs, self.N = dofmap(mesh, 'CG', 1)
        self.in_dofs, self.M = dofmap(mesh, 'CG', 1, block=2)
        X = mesh.coords[mesh.cells]
        self.X = X
        Jm = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
        det = Jm[:, 0, 0] * Jm[:, 1, 1] - Jm[:, 0, 1] * Jm[:, 1, 0]
        self.area = 0.5 * np.abs(det)
        Ji = np.empty_like(Jm)
        Ji[:, 0, 0], Ji[:, 0, 1] = Jm[:, 1, 1] / det, -Jm[:, 0, 1] / det
        Ji[:, 1, 0], Ji[:, 1, 1] = -Jm[:, 1, 0] / det, Jm[:, 0, 0] / det
        gref = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
        self.G = np.einsum('ckd,ak->cad', Ji, gref)                    # reference-configuration gradients
        self.h = mesh.cell_diameter()
        fc, fl = mesh.exterior_facets()
        self.fc, self.fl = fc, fl
        lf = mesh.local_facets[fl]
        P, Q, O = X[fc, lf[:, 0]], X[fc, lf[:, 1]], X[fc, fl]
        t = Q - P
        self.flen = np.linalg.norm(t, axis=1)
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / self.flen[:, None]
        self.fn = nrm * np.sign(np.einsum('fd,fd->f', nrm, P - O))[:, None]
        self.flv = lf
        self.fdofs, self.fin_dofs = self.cell_dofs[fc], self.in_dofs[fc]
        # sources per cell
        tg = self.tags
        self.H = np.zeros((mesh.ncells, 2))
        mag = (tg >= 3) & (tg < 3 + p)
        i = (tg - 3)[mag]
        fa = 2 * np.pi / p / 2 + i * (2 * np.pi / p) + angle * 2 / p
        self.H[mag, 0] = (-1.0) ** i * Hc * np.cos(fa)
        self.H[mag, 1] = (-1.0) ** i * Hc * np.sin(fa)
        JA, JB, JC = (iq * np.sin(angle) + DOLFIN_EPS, iq * np.sin(angle - 2 * np.pi / 3) + DOLFIN_EPS,
                      iq * np.sin(angle + 2 * np.pi / 3) + DOLFIN_EPS)
        self.cur = np.zeros(mesh.ncells)
        wnd = (tg >= 15) & (tg < 15 + s)
        w = (tg - 15)[wnd]
        pole, k = w // 3, w % 3
        amp = np.choose(k, [JB, JA, JC])
        sgn = np.where(k == 1, (-1.0) ** pole, (-1.0) ** (pole + 1))
        self.cur[wnd] = amp * sgn

    # -- kinematics (utils_dolfinx.py:34-66) ------------------------------------
    def _kin(self, G, ue, uhe):
        gradu = np.einsum('ca,cad->cd', ue, G)
        F = np.eye(2)[None] + np.einsum('cai,caj->cij', uhe, G)
        det = F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]
        Fi = np.empty_like(F)
        Fi[:, 0, 0], Fi[:, 0, 1] = F[:, 1, 1] / det, -F[:, 0, 1] / det
        Fi[:, 1, 0], Fi[:, 1, 1] = -F[:, 1, 0] / det, F[:, 0, 0] / det
        gx = np.einsum('ci,cij->cj', gradu, Fi)                        # gradx(u) = grad(u) F^-1
        gv = np.einsum('cai,cij->caj', G, Fi)                          # gradx(phi_a)
        return gx, gv, det, Fi

    def _nu(self, tags, B2, steel_only=False):
        nb = np.sqrt(B2 + DOLFIN_EPS)
        mur = mu_r_steel(nb)
        if steel_only:
            return 1.0 / (self.mu0 * mur)
        steel = (tags == 1) | (tags == 2)
        magnet = (tags >= 3) & (tags <= 14)
        return 1.0 / (self.mu0 * np.where(steel, mur, np.where(magnet, 1.05, 1.0)))

    def _cell_res(self, ue, uhe):
        gx, gv, det, _ = self._kin(self.G, ue, uhe)
        nu = self._nu(self.tags, gx[:, 0] ** 2 + gx[:, 1] ** 2)
        Re = (self.area * nu * det)[:, None] * np.einsum('cj,caj->ca', gx, gv)
        Jm = (self.H[:, 0, None] * gv[:, :, 1] - self.H[:, 1, None] * gv[:, :, 0]) * (det * self.area)[:, None]
        Jw = (self.cur * det * self.area / 3.0)[:, None] * np.ones((1, 3))
        return Re - self.js_scale * (Jm + Jw)

    def _facet_res(self, uf, uhf):
        gx, gv, det, Fi = self._kin(self.G[self.fc], uf, uhf)
        nN = det[:, None] * np.einsum('fji,fj->fi', Fi, self.fn)       # Nanson: J F^-T n
        nrmN = np.sqrt(nN[:, 0] ** 2 + nN[:, 1] ** 2)
        coeff = self._nu(None, gx[:, 0] ** 2 + gx[:, 1] ** 2, steel_only=True)
        gxn = np.einsum('fj,fj->f', gx, nN)
        gvn = np.einsum('faj,fj->fa', gv, nN)
        bh = self.beta / self.h[self.fc]
        s, w = quad.interval(2)
        nf = self.fc.size
        ar = np.arange(nf)
        Rf = np.zeros((nf, 3), dtype=uf.dtype)
        for q in range(len(w)):
            ph = np.zeros((nf, 3))
            ph[ar, self.flv[:, 0]] = 1.0 - s[q]
            ph[ar, self.flv[:, 1]] = s[q]
            uq = np.einsum('fa,fa->f', uf, ph)                          # g = 0 (ubc_em, run_motor_opt.py:277-278)
            Rf += (w[q] * self.flen)[:, None] * (coeff[:, None] * (-gxn[:, None] * ph - gvn * uq[:, None])
                                                 + (bh * coeff * nrmN * uq)[:, None] * ph)
        return 2.0 * Rf                                                  # boundary_components = [0, 1]

    def residual(self, u, uh):
        return [(self.cell_dofs, None, self._cell_res(u[self.cell_dofs], uh[self.in_dofs].reshape(-1, 3, 2))),
                (self.fdofs, None, self._facet_res(u[self.fdofs], uh[self.fin_dofs].reshape(-1, 3, 2)))]

    # -- complex-step derivatives -------------------------------------------------
    def _cs(self, fun, ue, uhe, wrt):
        hstep = 1e-30
        n = 3 if wrt == 'u' else 6
        out = []
        for k in range(n):
            a, b = ue.astype(complex), uhe.astype(complex)
            if wrt == 'u':
                a[:, k] += 1j * hstep
            else:
                b[:, k // 2, k % 2] += 1j * hstep
            out.append(fun(a, b).imag / hstep)
        return np.stack(out, axis=-1)

    def jacobian(self, u, uh):
        ue, uhe = u[self.cell_dofs], uh[self.in_dofs].reshape(-1, 3, 2)
        uf, uhf = u[self.fdofs], uh[self.fin_dofs].reshape(-1, 3, 2)
        return [(self.cell_dofs, self.cell_dofs, self._cs(self._cell_res, ue, uhe, 'u')),
                (self.fdofs, self.fdofs, self._cs(self._facet_res, uf, uhf, 'u'))]

    def dRdm(self, slot, u, uh):
        ue, uhe = u[self.cell_dofs], uh[self.in_dofs].reshape(-1, 3, 2)
        uf, uhf = u[self.fdofs], uh[self.fin_dofs].reshape(-1, 3, 2)
        return [(self.cell_dofs, self.in_dofs, self._cs(self._cell_res, ue, uhe, 'uh')),
                (self.fdofs, self.fin_dofs, self._cs(self._facet_res, uf, uhf, 'uh'))]

    # -- outputs: int |B|^n J dx(1,2)  (B_power_form) -----------------------------------
    def _out_cells(self, k, ue, uhe):
        gx, _, det, _ = self._kin(self.G, ue, uhe)
        sel = (self.tags == 1) | (self.tags == 2)
        Bm = np.sqrt(gx[:, 0] ** 2 + gx[:, 1] ** 2)
        return np.where(sel, Bm ** self.exponents[k] * det * self.area, 0.0)

    def output(self, k, u, uh):
        return [self._out_cells(k, u[self.cell_dofs], uh[self.in_dofs].reshape(-1, 3, 2)).real]

    def output_du(self, k, u, uh):
        ue, uhe = u[self.cell_dofs], uh[self.in_dofs].reshape(-1, 3, 2)
        return [(self.cell_dofs, None, self._cs(lambda a, b: self._out_cells(k, a, b), ue, uhe, 'u'))]

    def output_dm(self, k, slot, u, uh):
        ue, uhe = u[self.cell_dofs], uh[self.in_dofs].reshape(-1, 3, 2)
        return [(self.in_dofs, None, self._cs(lambda a, b: self._out_cells(k, a, b), ue, uhe, 'uh'))]
