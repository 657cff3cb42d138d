"""CSDL implicit operation for one PDE state: the upper face of the drop-in
boundary.  Callback names, argument meaning, assignment-vs-accumulation and the
stateful hand-over between callbacks follow the reference's
femo/csdl_opt/state_model.py (line numbers cited per method).
"""
import numpy as np

from ..fea.fea_b200 import FEA
from ..fea.utils_b200 import (update, getFuncArray, assembleVector, assembleMatrix, assembleSystem, computePartials,
                              createFunction, computeMatVecProductFwd, computeMatVecProductBwd, setUpKSP_MUMPS)
from ._csdl_compat import Model, CustomImplicitOperation, csdl
from .. import _hostops as _H


def _iadd(container, key, values):
    """`container[key] += values` (state_model.py:180-199) without a single-threaded pass over large vectors."""
    dst = container[key]
    if isinstance(dst, np.ndarray):
        _H.iadd(dst, values)
    else:
        container[key] = dst + values


class StateModel(Model):
    """state_model.py:7-38"""

    def initialize(self):
        self.parameters.declare('debug_mode', default=False)
        self.parameters.declare('fea', types=FEA)
        self.parameters.declare('state_name', types=str)
        self.parameters.declare('arg_name_list', types=list)

    def define(self):
        fea = self.fea = self.parameters['fea']
        state_name = self.parameters['state_name']
        args_dict, args_list = dict(), []
        for arg_name in self.parameters['arg_name_list']:
            entry = args_dict[arg_name] = fea.inputs_dict[arg_name]
            args_list.append(self.declare_variable(arg_name, shape=(entry['shape'],),
                                                   val=getFuncArray(entry['function'])))
        op = StateOperation(fea=fea, args_dict=args_dict, state_name=state_name,
                            debug_mode=self.parameters['debug_mode'])
        self.register_output(state_name, csdl.custom(*args_list, op=op))


class StateOperation(CustomImplicitOperation):
    """inputs: design/input fields; output: the PDE state (state_model.py:41-218)."""

    def initialize(self):
        for key in ('debug_mode', 'fea', 'args_dict', 'state_name'):
            self.parameters.declare(key)

    def _banner(self, what):
        if self.debug_mode == True:            # noqa: E712
            print(str(self.state_name) + "=" * 40)
            print("CSDL: Running %s..." % what)
            print("=" * 40)

    def define(self):                          # :52-73
        self.debug_mode = self.parameters['debug_mode']
        self.fea = self.parameters['fea']
        self.state_name = self.parameters['state_name']
        self.args_dict = self.parameters['args_dict']
        self._banner('define()')
        for arg_name, arg in self.args_dict.items():
            self.add_input(arg_name, shape=(arg['shape'],))
        self.state = self.fea.states_dict[self.state_name]
        self.add_output(self.state_name, shape=(self.state['shape'],))
        self.declare_derivatives('*', '*')
        self.bcs = self.fea.bc
        self.linear = self.fea.linear_problem
        self.ksp = None

    def _push(self, inputs, outputs):
        for arg_name in inputs:
            update(self.args_dict[arg_name]['function'], inputs[arg_name])
        update(self.state['function'], outputs[self.state_name])

    def evaluate_residuals(self, inputs, outputs, residuals):          # :75-85
        self._banner('evaluate_residuals()')
        self._push(inputs, outputs)
        residuals[self.state_name] = assembleVector(self.state['residual_form'])

    def solve_residual_equations(self, inputs, outputs):              # :87-115
        self._banner('solve_residual_equations()')
        self.fea.opt_iter += 1
        for arg_name in inputs:
            arg = self.args_dict[arg_name]
            update(arg['function'], inputs[arg_name])
            if arg['record']:
                arg['recorder'].write_function(arg['function'], self.fea.opt_iter)
        update(self.state['function'], outputs[self.state_name])
        self.fea.solve(self.state['residual_form'], self.state['function'], self.bcs)
        outputs[self.state_name] = getFuncArray(self.state['function'])
        if self.fea.record and self.state['recorder'] is not None:
            self.state['recorder'].write_function(self.state['function'], self.fea.opt_iter)

    def compute_derivatives(self, inputs, outputs, derivatives):      # :117-158
        self._banner('compute_derivatives()')
        self._push(inputs, outputs)
        state, args_dict = self.state, self.args_dict
        dR_du = state['dR_du']
        if dR_du is None:
            dR_du = computePartials(state['residual_form'], state['function'])
        # one element pass yields both the un-BC'd dR/du (:132) and the BC'd system matrix (:149-151)
        self.A, _ = assembleSystem(dR_du, state['residual_form'], bcs=self.bcs, rhs=False)
        self.dRdu = self.A.plain
        dRdf_dict = dict()
        dR_df_list = state['dR_df_list']
        for arg_ind, arg_name in enumerate(state['arguments']):
            if dR_df_list is None:
                dRdf = assembleMatrix(computePartials(state['residual_form'], args_dict[arg_name]['function']))
            else:
                dRdf = dR_df_list[arg_ind]
            dRdf_dict[arg_name] = dict(dRdf=dRdf, df=createFunction(args_dict[arg_name]['function']))
        self.dRdf_dict = dRdf_dict
        self.dR = state['d_residual']
        self.du = state['d_state']
        if self.linear is True:
            self.ksp = setUpKSP_MUMPS(self.A)

    def compute_jacvec_product(self, inputs, outputs, d_inputs, d_outputs, d_residuals, mode):   # :161-200
        self._banner('compute_jacvec_product()...mode ' + str(mode))
        self._push(inputs, outputs)
        name = self.state_name
        if mode == 'fwd':
            if name in d_residuals:
                if name in d_outputs:
                    update(self.du, d_outputs[name])
                    _iadd(d_residuals, name, computeMatVecProductFwd(self.dRdu, self.du))
                for arg_name, entry in self.dRdf_dict.items():
                    if arg_name in d_inputs:
                        update(entry['df'], d_inputs[arg_name])
                        _iadd(d_residuals, name, computeMatVecProductFwd(entry['dRdf'], entry['df']))
        if mode == 'rev':
            if name in d_residuals:
                update(self.dR, d_residuals[name])
                if name in d_outputs:
                    _iadd(d_outputs, name, computeMatVecProductBwd(self.dRdu, self.dR))
                for arg_name, entry in self.dRdf_dict.items():
                    if arg_name in d_inputs:
                        _iadd(d_inputs, arg_name, computeMatVecProductBwd(entry['dRdf'], self.dR))

    def apply_inverse_jacobian(self, d_outputs, d_residuals, mode):     # :202-218
        self._banner('apply_inverse_jacobian()...mode ' + str(mode))
        name = self.state_name
        if mode == 'fwd':
            d_outputs[name] = self.fea.solveLinearFwd(self.du, self.A, self.dR, d_residuals[name], self.ksp)
        else:
            d_residuals[name] = self.fea.solveLinearBwd(self.dR, self.A, self.du, d_outputs[name], self.ksp)
