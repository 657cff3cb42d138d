"""Fits the piecewise relative-permeability curve of the motor example and writes the
coefficients to femo_b200/forms/bh_fit.json.

Restates examples/em_motor_opt/permeability/piecewise_permeability.py:26-94 of the reference
(linear fit on data rows 1:3, exponential fit from row 4, C1 cubic bridge between x1 = 0.8 T
and x2 = 1.4 T) on the reference's data table
"examples/em_motor_opt/permeability/Magnetic alloy, silicon core iron C.tab".  Run in the
build container (the reference tree is not available on the GPU box); the JSON is committed.
"""
import json
import os
import sys

import numpy as np
from scipy.optimize import curve_fit

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
TAB = os.path.join(REF, 'examples/em_motor_opt/permeability/Magnetic alloy, silicon core iron C.tab')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'femo_b200', 'forms', 'bh_fit.json')


def linfun(x, a, b):
    return a * x + b


def expfun(x, a, b, c):
    return a * np.exp(b * x + c) + 1


data = np.genfromtxt(TAB, skip_header=1, delimiter='\t')
H, B = data[:, 0], data[:, 1]
with np.errstate(divide='ignore', invalid='ignore'):
    mu = B / H / (4e-7 * np.pi)
popt_lin, _ = curve_fit(linfun, B[1:3], mu[1:3])
popt_exp, _ = curve_fit(expfun, B[4:], mu[4:])
x1, x2 = 0.8, 1.4
lin_f, lin_d = linfun(x1, *popt_lin), popt_lin[0]
exp_f = expfun(x2, *popt_exp)
exp_d = (exp_f - 1) * popt_exp[1]
A = np.array([[3 * x1 ** 2, 2 * x1, 1, 0], [3 * x2 ** 2, 2 * x2, 1, 0], [x1 ** 3, x1 ** 2, x1, 1], [x2 ** 3, x2 ** 2, x2, 1]])
cub = np.linalg.solve(A, np.array([lin_d, exp_d, lin_f, exp_f]))
fit = dict(x1=x1, x2=x2, lin=[float(v) for v in popt_lin], cubic=[float(v) for v in cub], exp=[float(v) for v in popt_exp],
           rows=int(len(H)), source='Magnetic alloy, silicon core iron C.tab')
json.dump(fit, open(OUT, 'w'), indent=1)
print(json.dumps(fit))
