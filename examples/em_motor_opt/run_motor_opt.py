"""Coupled mesh-motion -> magnetostatics chain of the reference's examples/em_motor_opt/run_motor_opt.py (config 5):
edge displacement uhat_bc -> hyperelastic mesh motion uhat (incremental SNES, :131-166) -> nonlinear B-H
magnetostatics A_z on the moved mesh (5-step load ramp, :231-250) -> flux-density influence outputs and subdomain
areas (:176-181, :281-291), with the chained adjoint checked by finite differences.

The reference reads its motor mesh from git-LFS blobs that are not part of the checkout; this script runs on the
synthetic annulus with the same 216-subdomain tag layout.  With a Gmsh file at hand,
`import_mesh(prefix, subdomains=True)` + `Measure('dS', subdomain_data=boundaries_mf)(1000)` replace the synthetic
mesh, tags and facet measures below.  The FFD / power-loss CSDL models around the FEA are outside the engine's scope.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from femo_b200.fea.fea_b200 import *                                        # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator                          # noqa: E402
from femo_b200.forms import motor as pde                                    # noqa: E402
from femo_b200 import engine as E                                           # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--nr', default='12')
parser.add_argument('--nth', default='72')
args = parser.parse_args()
nr, nth = int(args.nr), int(args.nth)
mesh = Mesh(E.EngineMesh.annulus(nr, nth), 'triangle')
tags = pde.synthetic_motor_tags(mesh)
dx = Measure('dx', domain=mesh, subdomain_data=meshtags(mesh, 2, np.arange(mesh.num_cells), tags))
mid = nr // 2
dS = pde.SideMeasure(*pde.annulus_circle_sides(mesh, mid))
c0, c1 = pde.annulus_circle_sides(mesh, 0), pde.annulus_circle_sides(mesh, nr)
ds = pde.SideMeasure(np.concatenate([c0[0], c1[0]]), np.concatenate([c0[1], c1[1]]))
Hc, p, s, vacuum_perm, angle, iq = 838.e3, 12, 36, 4e-7 * np.pi, 0., 282.2 / 0.00016231     # run_motor_opt.py:84-92
winding_id, magnet_id, steel_id = [15], [3], [1, 2]                                           # :68-70

# ---- mesh motion subproblem (:94-208) ---------------------------------------------------------------------------
VV = VectorFunctionSpace(mesh, ('CG', 1))
fea_mm = FEA(mesh)
fea_mm.PDE_SOLVER, fea_mm.REPORT = 'SNES', False
uhat_bc, uhat = Function(VV), Function(VV)
res_mm = pde.pdeResMM(uhat, TestFunction(VV), g=uhat_bc, nitsche=True, sym=True, dS_=dS(1000), ds_=ds(1000), cell_tags=tags)


def solveIncremental(res, func, bc, report=False):
    vec = np.copy(getFuncArray(uhat_bc))
    STEPS = 2
    func.vector.set(0.0)
    for i in range(STEPS):
        uhat_bc.vector.setArray(vec * (i + 1) / STEPS)
        solveNonlinear(res, func, bc, 'SNES', False, False)
    uhat_bc.vector.setArray(vec)


fea_mm.custom_solve = solveIncremental
fea_mm.add_input('uhat_bc', uhat_bc, init_val=0.0)
fea_mm.add_state(name='uhat', function=uhat, residual_form=res_mm, arguments=['uhat_bc'])
fea_mm.add_output(name='winding_area', type='scalar', form=pde.area_form(uhat, dx, winding_id), arguments=['uhat'])
fea_mm.add_output(name='magnet_area', type='scalar', form=pde.area_form(uhat, dx, magnet_id), arguments=['uhat'])
fea_mm.add_output(name='steel_area', type='scalar', form=pde.area_form(uhat, dx, steel_id), arguments=['uhat'])

# ---- electromagnetic subproblem (:212-317) ----------------------------------------------------------------------
fea_em = FEA(mesh)
fea_em.PDE_SOLVER, fea_em.REPORT = 'SNES', False
V = FunctionSpace(mesh, ('CG', 1))
A_z = Function(V)
res_em = pde.pdeResEM(A_z, TestFunction(V), uhat, iq, dx, p, s, Hc, vacuum_perm, angle, g=Function(V), nitsche=True, sym=True)
js = pde.JS(res_em)


def solveIncrementalEM(res, func, bc, report=False):
    func.vector.set(0.0)
    for i in range(5):
        js.set_scale((i + 1) / 5)
        solveNonlinear(res, func, bc, 'SNES', False, False)


fea_em.custom_solve = solveIncrementalEM
fea_em.add_input('uhat', uhat, init_val=0.0)
fea_em.add_state(name='A_z', function=A_z, residual_form=res_em, arguments=['uhat'])
fea_em.add_output(name='B_influence_eddy_current', type='scalar', form=pde.B_power_form(A_z, uhat, 2, dx, steel_id),
                  arguments=['A_z', 'uhat'])
fea_em.add_output(name='B_influence_hysteresis', type='scalar', form=pde.B_power_form(A_z, uhat, 1.76835, dx, steel_id),
                  arguments=['A_z', 'uhat'])

fea_model = FEAModel(fea=[fea_mm, fea_em], debug_mode=False)
xy = mesh.geometry.x[:, :2]
g0 = np.zeros(2 * mesh.num_vertices)
nodes = mid * nth + np.arange(nth)                           # the interior circle moves radially by 1 %
g0[2 * nodes], g0[2 * nodes + 1] = 0.01 * xy[nodes, 0], 0.01 * xy[nodes, 1]
fea_model.create_input('uhat_bc', shape=g0.size, val=g0)
sim = Simulator(fea_model)
sim.run()
for name in ('winding_area', 'magnet_area', 'steel_area', 'B_influence_eddy_current', 'B_influence_hysteresis'):
    print('%-26s %.10e' % (name, sim[name][0]))
sim.check_totals('B_influence_eddy_current', 'uhat_bc', step=1e-7, directions=2)
sim.check_totals('steel_area', 'uhat_bc', step=1e-7, directions=2)
