"""Aggregates one StateModel per state and one Output(Field)Model per output of
every FEA in `fea` (femo/csdl_opt/fea_model.py:5-38)."""
from ._csdl_compat import Model
from .state_model import StateModel
from .output_model import OutputModel, OutputFieldModel


class FEAModel(Model):
    def initialize(self):
        self.parameters.declare('fea')
        self.parameters.declare('debug_mode', default=True)      # the reference hard-codes True (quirk B10)

    def define(self):
        self.fea_list = self.parameters['fea']
        debug = self.parameters['debug_mode']
        for fea in self.fea_list:
            for state_name, state in fea.states_dict.items():
                self.add(StateModel(fea=fea, debug_mode=debug, state_name=state_name,
                                    arg_name_list=state['arguments']),
                         name='{}_state_model'.format(state_name))
            for output_name, out in fea.outputs_dict.items():
                self.add(OutputModel(fea=fea, output_name=output_name, arg_name_list=out['arguments']),
                         name='{}_output_model'.format(output_name))
            for output_name, out in fea.outputs_field_dict.items():
                self.add(OutputFieldModel(fea=fea, output_name=output_name, arg_name_list=out['arguments']),
                         name='{}_output_model'.format(output_name))
