"""The FEA registry: same public surface as the reference's
femo/fea/fea_dolfinx.py:70-234 (attributes :87-98, methods cited below), with
the dolfinx/PETSc work re-pointed at the CUDA engine.
"""
import os

import numpy as np

from .utils_b200 import *          # noqa: F401,F403  (the reference star-imports its utils too, fea_dolfinx.py:5)
from .utils_b200 import (getFuncArray, setFuncArray, derivative, solveNonlinear, solveKSP_mumps, transpose,
                         Function, FunctionSpace, dirichletbc)


class FEA(object):
    """Registry of inputs / states / outputs of one PDE problem plus the solve
    entry points the CSDL operations call."""

    def __init__(self, mesh):
        self.mesh = mesh

        self.inputs_dict = dict()
        self.states_dict = dict()
        self.outputs_dict = dict()
        self.outputs_field_dict = dict()
        self.bc = []

        # solver flags, fea_dolfinx.py:87-98
        self.PDE_SOLVER = "Newton"
        self.REPORT = True
        self.ubc = None
        self.custom_solve = None
        self.opt_iter = 0
        self.initial_solve = True
        self.initialize = False
        self.record = False
        self.recorder_path = "records"
        self.linear_problem = False

    # -- registry (fea_dolfinx.py:100-176) -------------------------------------
    def add_input(self, name, function, init_val=1.0, record=False):
        if name in self.inputs_dict:
            raise ValueError('name has already been used for an input')
        function.x.array[:] = init_val                         # quirk B5: overwrites the function
        self.inputs_dict[name] = dict(
            function=function,
            function_space=function.function_space,
            shape=len(getFuncArray(function)),
            recorder=self.createRecorder(name, record),
            record=record)

    def add_state(self, name, function, residual_form, arguments, dR_du=None, dR_df_list=None, record=False):
        self.states_dict[name] = dict(
            function=function,
            residual_form=residual_form,
            function_space=function.function_space,
            shape=len(getFuncArray(function)),
            d_residual=Function(function.function_space),
            d_state=Function(function.function_space),
            dR_du=dR_du,
            dR_df_list=dR_df_list,
            arguments=arguments,
            recorder=self.createRecorder(name, record),
            record=record)

    def add_output(self, name, type, form, arguments):
        if type == 'field':
            raise NotImplementedError("field outputs of form type: use add_field_output")
        elif type == 'scalar':
            shape = 1
        partials = []
        for argument in arguments:
            if argument in self.inputs_dict:
                partial = derivative(form, self.inputs_dict[argument]['function'])
            elif argument in self.states_dict:
                partial = derivative(form, self.states_dict[argument]['function'])
            partials.append(partial)
        self.outputs_dict[name] = dict(form=form, shape=shape, arguments=arguments, partials=partials)

    def add_field_output(self, name, form, arguments, record=False):
        V = FunctionSpace(self.mesh, ("CG", 1))
        output_func = Function(V)
        self.outputs_field_dict[name] = dict(
            form=form, func=output_func, shape=len(getFuncArray(output_func)), arguments=arguments,
            partials=[], recorder=self.createRecorder(name, record), record=record)

    def add_exact_solution(self, Expression, function_space):
        f_analytic = Expression()
        f_ex = Function(function_space)
        f_ex.interpolate(f_analytic.eval)
        return f_ex

    def add_strong_bc(self, ubc, locate_BC_list, function_space=None):
        for locate_BC in locate_BC_list:
            self.bc.append(dirichletbc(ubc, locate_BC, function_space))

    # -- solves (fea_dolfinx.py:178-222) --------------------------------------------
    def solve(self, res, func, bc):
        solver_type = self.PDE_SOLVER
        report = self.REPORT
        initialize = self.initialize
        if self.custom_solve is not None and self.initial_solve == True:   # noqa: E712  sticky, quirk B9
            self.custom_solve(res, func, bc, report)
        else:
            solveNonlinear(res, func, bc, solver_type, report, initialize)

    def solveLinearFwd(self, du, A, dR, dR_array, ksp=None):
        """Solve A du = dR.  The reference's ksp=None branch passes the transposed
        operator with swapped vectors and returns zeros (quirk B4); this is the
        intended forward solve."""
        setFuncArray(dR, dR_array)
        du.vector.set(0.0)
        if ksp is None:
            solveKSP_mumps(A, dR.vector, du.vector)
        else:
            ksp.solve(dR.vector, du.vector)
        return du.vector.getArray()

    def solveLinearBwd(self, dR, A, du, du_array, ksp=None):
        """Adjoint solve A^T dR = du (fea_dolfinx.py:208-222)."""
        setFuncArray(du, du_array)
        dR.vector.set(0.0)
        if ksp is None:
            solveKSP_mumps(transpose(A), du.vector, dR.vector)
        else:
            # the reference reuses the KSP of A here, valid only for symmetric A (quirk B3);
            # the engine always applies the true transpose
            solveKSP_mumps(transpose(ksp.A), du.vector, dR.vector)
        return dR.vector.getArray()

    def projectFieldOutput(self, form, func):
        from .utils_b200 import project
        project(form, func, lump_mass=False)

    # -- recording (fea_dolfinx.py:228-234): raw .npy dumps instead of XDMF ---------------
    def createRecorder(self, name, record=False):
        recorder = None
        if record or self.record:
            recorder = NpyRecorder(os.path.join(self.recorder_path, "record_" + name))
        return recorder


class NpyRecorder:
    """XDMF-free recorder: one .npy per write."""

    def __init__(self, prefix):
        self.prefix = prefix

    def write_function(self, function, t=0):
        os.makedirs(os.path.dirname(self.prefix) or '.', exist_ok=True)
        np.save('%s_%05d.npy' % (self.prefix, int(t)), getFuncArray(function))

    def write_mesh(self, mesh):
        pass
