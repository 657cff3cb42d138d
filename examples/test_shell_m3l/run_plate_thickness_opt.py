"""Thickness optimisation of a clamped Reissner-Mindlin plate: the flat-mid-surface counterpart of the reference's shell
examples (examples/test_shell_m3l/shell_pde.py:219-311 -- ShellPDE with thickness and load inputs, a linear state solve,
compliance / mass / elastic-energy outputs handed to the optimiser through the CSDL operations).
Minimise the elastic energy under a mass bound:  python run_plate_thickness_opt.py --nel 16"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from femo_b200.fea.fea_b200 import *                       # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator          # noqa: E402
from femo_b200.forms.shell import ShellPDE                  # noqa: E402
from _slsqp import slsqp                                    # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--nel', default='16')
parser.add_argument('--maxiter', default='40')
args = parser.parse_args()
mesh = createUnitSquareMesh(int(args.nel))
pde = ShellPDE(mesh)
fea = FEA(mesh)
fea.PDE_SOLVER, fea.REPORT, fea.linear_problem = 'Newton', False, True
h, f, w = Function(pde.VT), Function(pde.VF), Function(pde.W)
E, nu, rho = 2.0e4, 0.3, 2.7
res = pde.pdeRes(h, w, f, E, nu, penalty=True, pen=1.0e6)           # clamped on the whole boundary (penalty)
fea.add_input('thicknesses', h)
fea.add_input('F_solid', f)
fea.add_state(name='disp_solid', function=w, residual_form=res, arguments=['thicknesses', 'F_solid'])
fea.add_output(name='elastic_energy', type='scalar', form=pde.elastic_energy(), arguments=['thicknesses', 'disp_solid'])
fea.add_output(name='mass', type='scalar', form=pde.mass(h, rho), arguments=['thicknesses'])
model = FEAModel(fea=[fea], debug_mode=False)
nT = fea.inputs_dict['thicknesses']['shape']
model.create_input('thicknesses', shape=nT, val=0.08)
model.create_input('F_solid', shape=nT, val=1.0)
model.add_design_variable('thicknesses', lower=0.02, upper=0.2)
model.add_objective('elastic_energy')
model.add_constraint('mass', upper=rho * 0.08)                     # no heavier than the uniform initial plate
sim = Simulator(model)
sim.run()
print('initial: elastic energy %.6e, mass %.4f' % (float(np.ravel(sim['elastic_energy'])[0]), float(np.ravel(sim['mass'])[0])))
slsqp(sim, model, maxiter=int(args.maxiter), ftol=1e-10)
t = np.asarray(sim['thicknesses'])
print('optimised: elastic energy %.6e, mass %.4f, thickness in [%.4f, %.4f]'
      % (float(np.ravel(sim['elastic_energy'])[0]), float(np.ravel(sim['mass'])[0]), t.min(), t.max()))
