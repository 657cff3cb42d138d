"""Builders for families 3 (Euler-Bernoulli beam) and 4 (SIMP Q1 elasticity):
engine problem + oracle family on the same mesh, tags and seeded inputs."""
import numpy as np
import scipy.sparse as sp

from femo_b200 import engine as E
from oracle import mesh as om, families as fam, assembly as asm, solvers


def _tagged(omesh, marker):
    """Indices (into the exterior-facet list) of facets whose vertices all satisfy marker."""
    fc, fl = omesh.exterior_facets()
    lf = omesh.local_facets[fl]
    X = omesh.coords[omesh.cells]
    return np.array([k for k in range(fc.size) if all(marker(X[fc[k], v]) for v in lf[k])], dtype=np.int32)


class BeamCase:
    """examples/beam_thickness_opt: 50 cells, L=1, tip load at x=L, clamp at x=0."""

    def __init__(self, n=50, seed=0, upload=True):
        self.emesh = E.EngineMesh.interval(n, 0.0, 1.0)
        self.omesh = om.interval(n, 0.0, 1.0)
        self.tag = _tagged(self.omesh, lambda x: abs(x[0] - 1.0) < 1e-12)
        self.F = fam.EBBeam(self.omesh, self.tag)
        self.p = E.EngineProblem(self.emesh, E.FAMILY_EB_BEAM, [1.0, 0.1, 1.0, -1.0], tagged=self.tag)
        lists = [np.array([0]), np.array([1])]                      # run_thickness_opt_cantilever_beam.py:157-162
        self.bc = asm.DirichletBC(self.F.N, lists, 0.0)
        self.p.set_bc(lists)
        rng = np.random.default_rng(seed)
        self.u = rng.standard_normal(self.F.N)
        self.m = 0.05 + 0.1 * rng.random(self.F.M)
        self.sp = solvers.StatePath(self.F, self.bc)
        if upload:
            _upload(self)


class SimpCase:
    """examples/beam_topo_opt: Q1 on [0,160]x[0,80], traction on the right edge, clamp x=0."""

    def __init__(self, nx=80, ny=40, seed=0, upload=True, rho_lo=0.2):
        lo, hi = (0.0, 0.0), (160.0, 80.0)
        self.emesh = E.EngineMesh.rectangle_quad(lo, hi, nx, ny)
        self.omesh = om.rectangle_quad(lo, hi, nx, ny)
        eps = 3e-16 * 1e10

        def traction(x):                                              # run_topo_opt_cantilever_beam.py:45-47
            return abs(x[1] - 40.0) < 80.0 / ny + eps and abs(x[0] - 160.0) < eps
        self.tag = _tagged(self.omesh, traction)
        self.F = fam.SimpQ1(self.omesh, self.tag)
        self.p = E.EngineProblem(self.emesh, E.FAMILY_SIMP_Q1, [0.3, 0.0, -0.25, 3.0], tagged=self.tag)
        nodes = np.nonzero(np.isclose(self.omesh.coords[:, 0], 0.0, atol=1e-6))[0]
        lists = [np.stack([2 * nodes, 2 * nodes + 1], axis=1).ravel()]  # one dirichletbc object, :139-144
        self.bc = asm.DirichletBC(self.F.N, lists, 0.0)
        self.p.set_bc(lists)
        rng = np.random.default_rng(seed)
        self.u = rng.standard_normal(self.F.N)
        self.m = rho_lo + (1.0 - rho_lo) * rng.random(self.F.M)
        self.sp = solvers.StatePath(self.F, self.bc)
        if upload:
            _upload(self)


def _upload(c):
    c.p.upload(0)
    c.d_u = c.p.to_device(c.u)
    c.d_m = c.p.to_device(c.m)
    c.p.set_coefficient(0, c.d_u)
    c.p.set_coefficient(1, c.d_m)


def set_state(c, u):
    c.u = np.asarray(u, dtype=np.float64)
    c.d_u.copy_(c.p.to_device(c.u))


def set_input(c, m):
    c.m = np.asarray(m, dtype=np.float64)
    c.d_m.copy_(c.p.to_device(c.m))


def csr(c, which, vals):
    rp, col = c.p.pattern(which)
    i = c.p.pattern_info(which)
    return sp.csr_matrix((vals.cpu().numpy(), col, rp), shape=(i['rows'], i['cols']))


class HexCase:
    """3-D SIMP cantilever (scaled down): box [0,Lx]x[0,Ly]x[0,Lz], clamp x=0,
    traction on the x=Lx face near mid-height, random density; optionally a sheared (non-axis-aligned) box."""

    def __init__(self, nx=8, ny=4, nz=2, seed=0, upload=True, rho_lo=0.2, shear=0.0):
        lo, hi = (0.0, 0.0, 0.0), (2.0 * nx, 2.0 * ny, 2.0 * nz)
        self.emesh = E.EngineMesh.box_hex(lo, hi, nx, ny, nz)
        self.omesh = om.box_hex(lo, hi, nx, ny, nz)
        assert shear == 0.0
        eps = 1e-9

        def traction(x):
            return abs(x[0] - hi[0]) < eps and abs(x[1] - 0.5 * hi[1]) < 2.0 + eps
        self.tag = _tagged(self.omesh, traction)
        self.F = fam.SimpHex8(self.omesh, self.tag)
        self.p = E.EngineProblem(self.emesh, E.FAMILY_SIMP_HEX8, [0.3, 0.0, -0.25, 0.0, 3.0], tagged=self.tag)
        nodes = np.nonzero(np.isclose(self.omesh.coords[:, 0], 0.0, atol=1e-9))[0]
        lists = [np.stack([3 * nodes, 3 * nodes + 1, 3 * nodes + 2], axis=1).ravel()]
        self.bc = asm.DirichletBC(self.F.N, lists, 0.0)
        self.p.set_bc(lists)
        rng = np.random.default_rng(seed)
        self.u = rng.standard_normal(self.F.N)
        self.m = rho_lo + (1.0 - rho_lo) * rng.random(self.F.M)
        self.sp = solvers.StatePath(self.F, self.bc)
        if upload:
            _upload(self)
