"""Builder for the motor magnetostatics family (config 5b): engine problem + oracle on the same
synthetic annulus, tags and seeded inputs."""
import json
import os

import numpy as np
import scipy.sparse as sp

from femo_b200 import engine as E
from oracle import motor, assembly as asm, solvers

_FIT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'femo_b200', 'forms', 'bh_fit.json')))


def em_params(Hc=838e3, p=12, s=36, mu0=4e-7 * np.pi, angle=0.0, iq=282.2 / 0.00016231, beta=1e4, js_scale=1.0,
              exponents=(2.0, 1.76835)):
    """Parameter vector of FEMO_FAMILY_MOTOR_EM (csrc/families4.cuh, enum EmParam)."""
    return [mu0, Hc, iq, angle, p, s, js_scale, beta, _FIT['x1'], _FIT['x2']] + list(_FIT['lin']) + list(_FIT['cubic']) + \
        list(_FIT['exp']) + list(exponents)


class MotorCase:
    def __init__(self, nr=12, nth=48, seed=0, upload=True, uscale=1e-2, uhscale=2e-4):
        self.emesh = E.EngineMesh.annulus(nr, nth)
        self.omesh = motor.annulus_tri(nr, nth)
        self.tags = motor.motor_tags(self.omesh)
        self.F = motor.MotorEM(self.omesh, self.tags)
        self.p = E.EngineProblem(self.emesh, E.FAMILY_MOTOR_EM, em_params(), cell_tags=self.tags)
        rng = np.random.default_rng(seed)
        self.u = uscale * rng.standard_normal(self.F.N)
        self.m = uhscale * rng.standard_normal(self.F.M)
        self.bc = None
        self.sp = solvers.StatePath(self.F, None)
        if upload:
            self.p.upload(0)
            self.d_u = self.p.to_device(self.u)
            self.d_m = self.p.to_device(self.m)
            self.p.set_coefficient(0, self.d_u)
            self.p.set_coefficient(1, self.d_m)

    def csr(self, which, vals):
        rp, col = self.p.pattern(which)
        i = self.p.pattern_info(which)
        return sp.csr_matrix((vals.cpu().numpy(), col, rp), shape=(i['rows'], i['cols']))


class MotorMMCase:
    """Mesh-motion family (config 5a) on the synthetic annulus: Nitsche facets on the inner and outer
    boundary circles and on one interior circle (both sides)."""

    def __init__(self, nr=8, nth=24, seed=0, upload=True, scale=2e-4):
        from oracle import motor_mm
        self.emesh = E.EngineMesh.annulus(nr, nth)
        self.omesh = motor.annulus_tri(nr, nth)
        self.tags = motor.motor_tags(self.omesh)
        self.mid = nr // 2
        parts = [motor_mm.circle_facets(self.omesh, k) for k in (0, self.mid, nr)]
        fc = np.concatenate([p[0] for p in parts])
        fl = np.concatenate([p[1] for p in parts])
        o = np.lexsort((fl, fc))
        self.facets = (fc[o], fl[o])
        self.F = motor_mm.MotorMM(self.omesh, self.facets, self.tags)
        self.p = E.EngineProblem(self.emesh, E.FAMILY_MOTOR_MM, [5e3], facets=self.facets, cell_tags=self.tags)
        rng = np.random.default_rng(seed)
        self.u = scale * rng.standard_normal(self.F.N)
        self.m = scale * rng.standard_normal(self.F.M)
        self.bc = None
        self.sp = solvers.StatePath(self.F, None)
        self.nr, self.nth = nr, nth
        if upload:
            self.p.upload(0)
            self.d_u = self.p.to_device(self.u)
            self.d_m = self.p.to_device(self.m)
            self.p.set_coefficient(0, self.d_u)
            self.p.set_coefficient(1, self.d_m)

    def radial_bc(self, frac=0.02):
        """Prescribed displacement: the interior circle moves radially by `frac` of its radius."""
        g = np.zeros(self.F.N)
        nodes = self.mid * self.nth + np.arange(self.nth)
        xy = self.omesh.coords[nodes]
        g[2 * nodes], g[2 * nodes + 1] = frac * xy[:, 0], frac * xy[:, 1]
        return g

    csr = MotorCase.csr
