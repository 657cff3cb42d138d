"""Source-term optimisation of the Poisson problem on the unit square (config 1):
the femo_b200 counterpart of the reference's examples/poisson_opt/run_poisson_opt.py.
Needs a B200 (the engine has no CPU path)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from femo_b200.fea.fea_b200 import *                         # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator           # noqa: E402
from femo_b200.forms.poisson import pdeRes, outputForm       # noqa: E402
from _slsqp import slsqp                                      # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--nel', dest='nel', default='16', help='Number of elements')
args = parser.parse_args()
num_el = int(args.nel)
mesh = createUnitSquareMesh(num_el)
PI = np.pi


class Expression_f:
    def eval(self, x):
        return 1 / (1 + 1e-6 * 4 * np.power(PI, 4)) * np.sin(PI * x[0]) * np.sin(PI * x[1])


class Expression_u:
    def eval(self, x):
        return 1 / (2 * np.power(PI, 2)) * np.sin(PI * x[0]) * np.sin(PI * x[1])


fea = FEA(mesh)
input_name, state_name, output_name = 'f', 'u', 'l2_functional'
input_function = Function(FunctionSpace(mesh, ('DG', 0)))
state_function_space = FunctionSpace(mesh, ('CG', 1))
state_function = Function(state_function_space)
v = TestFunction(state_function_space)
u_ex = fea.add_exact_solution(Expression_u, state_function_space)
f_ex = fea.add_exact_solution(Expression_f, input_function.function_space)
output_form = outputForm(state_function, input_function, u_ex)

ubc = Function(state_function_space)
ubc.vector.set(0.0)
locate_BC_list = [locate_dofs_geometrical((state_function_space, state_function_space),
                                          lambda x, a=a, b=b: np.isclose(x[a], b, atol=1e-6))
                  for a, b in ((0, 0.), (0, 1.), (1, 0.), (1, 1.))]
fea.add_strong_bc(ubc, locate_BC_list, state_function_space)
residual_form = pdeRes(state_function, v, input_function)

fea.add_input(input_name, input_function)
fea.add_state(name=state_name, function=state_function, residual_form=residual_form, arguments=[input_name])
fea.add_output(name=output_name, type='scalar', form=output_form, arguments=[input_name, state_name])
fea.PDE_SOLVER = 'Newton'
fea.REPORT = False

fea_model = FEAModel(fea=[fea], debug_mode=False)
fea_model.create_input(input_name, shape=fea.inputs_dict[input_name]['shape'],
                       val=0.1 * np.ones(fea.inputs_dict[input_name]['shape']) * 0.86)
fea_model.add_design_variable(input_name)
fea_model.add_objective(output_name, scaler=1e5)
sim = Simulator(fea_model)
sim.run()
sim.check_totals(output_name, input_name, step=1e-3)
res = slsqp(sim, fea_model, maxiter=20, ftol=1e-13)
print("=" * 40)
print("Objective value: ", sim[output_name])
print("Error in controls:", errorNorm(f_ex, input_function))
print("Error in states:", errorNorm(u_ex, state_function))
