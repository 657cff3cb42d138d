"""SIMP topology optimisation set-up of a 3-D cantilever on trilinear hexahedra: the 3-D extension of the
reference's examples/beam_topo_opt/run_topo_opt_cantilever_beam.py named by BASELINE.json configs[3] / SURVEY.md 8d
(C4-3D).  Forward solve and adjoint totals, GMG-preconditioned CG in place of LU."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from femo_b200.fea.fea_b200 import *                                        # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator                          # noqa: E402
from femo_b200.forms.topo import pdeRes, averageFunc, compliance            # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--nelx', default='64')
parser.add_argument('--nely', default='32')
parser.add_argument('--nelz', default='16')
args = parser.parse_args()
num_el_x, num_el_y, num_el_z = int(args.nelx), int(args.nely), int(args.nelz)
LENGTH_X, LENGTH_Y, LENGTH_Z = 160., 80., 40.
mesh = createBoxMesh(np.zeros(3), np.array([LENGTH_X, LENGTH_Y, LENGTH_Z]), num_el_x, num_el_y, num_el_z)


def TractionBoundary(x):
    return np.logical_and(abs(x[1] - LENGTH_Y / 2) < LENGTH_Y / num_el_y + DOLFIN_EPS * 1e10,
                          abs(x[0] - LENGTH_X) < DOLFIN_EPS * 1e10)


traction_facets = locate_entities_boundary(mesh, 2, TractionBoundary)
facet_tag = meshtags(mesh, 2, traction_facets, np.full(len(traction_facets), 100, dtype=np.int32))
ds_ = Measure('ds', domain=mesh, subdomain_data=facet_tag, metadata={"quadrature_degree": 4})

fea = FEA(mesh)
input_name, state_name = 'density', 'displacements'
input_function = Function(FunctionSpace(mesh, ('DG', 0)))
state_function_space = VectorFunctionSpace(mesh, ('CG', 1))
state_function = Function(state_function_space)
v = TestFunction(state_function_space)
f = Constant(mesh, (0, -1 / 4, 0))
residual_form = pdeRes(state_function, v, input_function, f, dss=ds_(100), method='SIMP')
fea.add_input(input_name, input_function)
fea.add_state(name=state_name, function=state_function, residual_form=residual_form, arguments=[input_name])
fea.add_output(name='avg_density', type='scalar', form=averageFunc(input_function), arguments=[input_name])
fea.add_output(name='compliance', type='scalar', form=compliance(state_function, f, dss=ds_(100)),
               arguments=[state_name])
ubc = Function(state_function_space)
locate_BC1 = locate_dofs_geometrical((state_function_space, state_function_space),
                                     lambda x: np.isclose(x[0], 0., atol=1e-6))
fea.add_strong_bc(ubc, [locate_BC1], state_function_space)
fea.REPORT = False

fea_model = FEAModel(fea=[fea], debug_mode=False)
np.random.seed(0)
nel = mesh.num_cells
fea_model.create_input('density', shape=nel, val=0.3 + 0.5 * np.random.random(nel))
fea_model.add_design_variable('density', upper=1.0, lower=1e-4)
fea_model.add_objective('compliance')
fea_model.add_constraint('avg_density', upper=0.40)
sim = Simulator(fea_model)
sim.run()
print("Compliance value: ", sim['compliance'])
print("Constraint value: ", sim['avg_density'])
g = sim.compute_totals('compliance', 'density')[('compliance', 'density')]
print("|d compliance / d density|_2 =", np.linalg.norm(g))
