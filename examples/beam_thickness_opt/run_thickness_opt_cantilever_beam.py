"""Thickness optimisation of an Euler-Bernoulli cantilever (config 3): counterpart of the
reference's examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from femo_b200.fea.fea_b200 import *                                  # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator                    # noqa: E402
from femo_b200.forms.beam import pdeRes, compliance, volume           # noqa: E402
from _slsqp import slsqp                                               # noqa: E402

E, L, b, h, nel = 1., 1., 0.1, 0.1, 50
mesh = createIntervalMesh(nel, 0., L)
fea = FEA(mesh)
input_name, state_name = 'thickness', 'displacements'
input_function = Function(FunctionSpace(mesh, ('DG', 0)))
state_function_space = FunctionSpace(mesh, ('Hermite', 3))
state_function = Function(state_function_space)
v = TestFunction(state_function_space)
f = Constant(mesh, -1.)

endpoint_node = locate_entities_boundary(mesh, 0, lambda x: np.isclose(x[0], L))
facet_tag = meshtags(mesh, 0, endpoint_node, np.full(len(endpoint_node), 100, dtype=np.int32))
ds_ = Measure('ds', domain=mesh, subdomain_data=facet_tag, metadata={"quadrature_degree": 4})
residual_form = pdeRes(state_function, v, input_function, f, ds_(100), E, b)
fea.add_input(input_name, input_function)
fea.add_state(name=state_name, function=state_function, residual_form=residual_form, arguments=[input_name])
fea.add_output(name='compliance', type='scalar', form=compliance(state_function, f, ds_(100)),
               arguments=[input_name, state_name])
fea.add_output(name='volume', type='scalar', form=volume(input_function, b, L), arguments=[input_name])
ubc = Function(state_function_space)
startpt = locate_entities_boundary(mesh, 0, lambda x: np.isclose(x[0], 0))
locate_BC1 = locate_dofs_topological(state_function_space, 0, startpt)
fea.add_strong_bc(ubc, [locate_BC1[0:1], locate_BC1[1:2]])
fea.REPORT = False

fea_model = FEAModel(fea=[fea], debug_mode=False)
fea_model.create_input('thickness', shape=nel, val=h)
fea_model.add_design_variable('thickness', upper=10., lower=1e-2)
fea_model.add_objective('compliance')
fea_model.add_constraint('volume', equals=b * h * L)
sim = Simulator(fea_model)
sim.run()
print("Compliance value: ", sim['compliance'], " tip deflection:", sim['displacements'][-2])
res = slsqp(sim, fea_model, maxiter=300, ftol=1e-12)
print("optimised thickness:", np.round(sim['thickness'], 6))
