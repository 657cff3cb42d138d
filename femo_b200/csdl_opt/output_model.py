"""CSDL explicit operations for scalar outputs and projected field outputs;
mirrors femo/csdl_opt/output_model.py of the reference (lines cited)."""
import numpy as np

from ..fea.fea_b200 import FEA
from ..fea.utils_b200 import update, getFuncArray, assemble, computePartials
from ._csdl_compat import Model, CustomExplicitOperation, csdl


def _collect_args(model, fea, arg_name_list):
    args_dict, args_list = dict(), []
    for arg_name in arg_name_list:
        if arg_name in fea.inputs_dict:
            args_dict[arg_name] = fea.inputs_dict[arg_name]
        elif arg_name in fea.states_dict:
            args_dict[arg_name] = fea.states_dict[arg_name]
        args_list.append(model.declare_variable(arg_name, shape=(args_dict[arg_name]['shape'],), val=1.0))
    return args_dict, args_list


class OutputModel(Model):
    """output_model.py:7-40"""

    def initialize(self):
        self.parameters.declare('fea', types=FEA)
        self.parameters.declare('output_name', types=str)
        self.parameters.declare('arg_name_list', types=list)

    def define(self):
        fea = self.fea = self.parameters['fea']
        output_name = self.parameters['output_name']
        args_dict, args_list = _collect_args(self, fea, self.parameters['arg_name_list'])
        op = OutputOperation(fea=fea, args_dict=args_dict, output_name=output_name)
        self.print_var(self.register_output(output_name, csdl.custom(*args_list, op=op)))


class OutputOperation(CustomExplicitOperation):
    """output_model.py:42-87"""

    def initialize(self):
        for key in ('fea', 'args_dict', 'output_name'):
            self.parameters.declare(key)

    def define(self):
        self.fea = self.parameters['fea']
        self.output_name = self.parameters['output_name']
        self.args_dict = self.parameters['args_dict']
        for arg_name, arg in self.args_dict.items():
            self.add_input(arg_name, shape=(arg['shape'],))
        self.output = self.fea.outputs_dict[self.output_name]
        self.output_size = self.output['shape']
        self.output_dim = 0 if self.output_size == 1 else 1     # scalar vs field
        self.add_output(self.output_name, shape=(self.output_size,))
        self.declare_derivatives('*', '*')

    def _push(self, inputs):
        for arg_name in inputs:
            update(self.args_dict[arg_name]['function'], inputs[arg_name])

    def compute(self, inputs, outputs):                                  # :69-75
        self._push(inputs)
        outputs[self.output_name] = np.array(assemble(self.output['form'], dim=self.output_dim))

    def compute_derivatives(self, inputs, derivatives):                  # :77-87
        self._push(inputs)
        # vectors come back in a ring of three page-locked staging buffers per size (utils_b200._download): with more than
        # three partials a fourth result could land on the first one's buffer while a backend still holds it -- own the data then
        own = len(self.args_dict) > 3
        for arg_name, arg in self.args_dict.items():
            g = assemble(computePartials(self.output['form'], arg['function']), dim=self.output_dim + 1)
            derivatives[self.output_name, arg_name] = np.array(g) if own else g


class OutputFieldModel(Model):
    """output_model.py:91-119"""

    def initialize(self):
        self.parameters.declare('fea', types=FEA)
        self.parameters.declare('output_name', types=str)
        self.parameters.declare('arg_name_list', types=list)

    def define(self):
        fea = self.fea = self.parameters['fea']
        output_name = self.parameters['output_name']
        args_dict, args_list = _collect_args(self, fea, self.parameters['arg_name_list'])
        op = OutputFieldOperation(fea=fea, args_dict=args_dict, output_name=output_name)
        self.register_output(output_name, csdl.custom(*args_list, op=op))


class OutputFieldOperation(CustomExplicitOperation):
    """L2-projected field output, no derivatives (output_model.py:121-159)."""

    def initialize(self):
        for key in ('fea', 'args_dict', 'output_name'):
            self.parameters.declare(key)

    def define(self):
        self.fea = self.parameters['fea']
        self.output_name = self.parameters['output_name']
        self.args_dict = self.parameters['args_dict']
        for arg_name, arg in self.args_dict.items():
            self.add_input(arg_name, shape=(arg['shape'],))
        self.output = self.fea.outputs_field_dict[self.output_name]
        self.output_size = self.output['shape']
        self.output_dim = 1
        self.add_output(self.output_name, shape=(self.output_size,))

    def compute(self, inputs, outputs):
        for arg_name in inputs:
            update(self.args_dict[arg_name]['function'], inputs[arg_name])
        self.fea.projectFieldOutput(self.output['form'], self.output['func'])
        if self.output['record']:
            self.output['recorder'].write_function(self.output['func'], self.fea.opt_iter)
        outputs[self.output_name] = getFuncArray(self.output['func'])
