"""Nonlinear Poisson (u^3) control problem with symmetric Nitsche boundary terms (config 2):
counterpart of the reference's examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from femo_b200.fea.fea_b200 import *                                                     # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator                                       # noqa: E402
from femo_b200.forms.nonlinear_poisson import pdeRes, outputForm, u_ex_ufl, f_ex_ufl     # noqa: E402
from _slsqp import slsqp                                                                  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--nel', dest='nel', default='16', help='Number of elements')
parser.add_argument('--maxiter', default='30')
args = parser.parse_args()
mesh = createUnitSquareMesh(int(args.nel))

fea = FEA(mesh)
input_name, state_name, output_name = 'f', 'u', 'l2_functional'
input_function_space = FunctionSpace(mesh, ('DG', 0))
input_function = Function(input_function_space)
state_function_space = FunctionSpace(mesh, ('CG', 1))
state_function = Function(state_function_space)
v = TestFunction(state_function_space)
u_ex = Function(state_function_space)
project(u_ex_ufl, u_ex)
f_ex = Function(input_function_space)
project(f_ex_ufl, f_ex)
output_form = outputForm(state_function, input_function)
residual_form = pdeRes(state_function, v, input_function, weak_bc=True, sym=True)

fea.add_input(input_name, input_function)
fea.add_state(name=state_name, function=state_function, residual_form=residual_form, arguments=[input_name])
fea.add_output(name=output_name, type='scalar', form=output_form, arguments=[input_name, state_name])
fea.PDE_SOLVER = 'SNES'
fea.REPORT = False

fea_model = FEAModel(fea=[fea], debug_mode=False)
fea_model.create_input(input_name, shape=fea.inputs_dict[input_name]['shape'], val=0.1)
fea_model.add_design_variable(input_name)
fea_model.add_objective(output_name)
sim = Simulator(fea_model)
sim.run()
res = slsqp(sim, fea_model, maxiter=int(args.maxiter), ftol=1e-10)
print("=" * 40)
print("Objective value: ", sim[output_name])
print("Error in controls:", errorNorm(f_ex, input_function))
print("Error in states:", errorNorm(u_ex, state_function))
