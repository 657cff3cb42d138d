"""ctypes binding of libfemo_b200.so (see include/femo_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C
femo_b200/csrc`.  There is no Python or CPU fallback: if the shared library is
missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libfemo_b200.so')


class FemoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libfemo_b200 error %d: %s' % (code, msg))
        self.code = code


class KrylovOpts(C.Structure):
    _fields_ = [('rtol', C.c_double), ('atol', C.c_double), ('max_it', C.c_int), ('precond', C.c_int),
                ('cheb_degree', C.c_int), ('method', C.c_int), ('restart', C.c_int), ('check_every', C.c_int),
                ('cheb_ratio', C.c_double), ('mg_precision', C.c_int), ('forcing', C.c_double)]


class KrylovInfo(C.Structure):
    _fields_ = [('iterations', C.c_int), ('converged', C.c_int), ('rnorm', C.c_double), ('bnorm', C.c_double),
                ('spmv_count', C.c_int)]


class AmgOpts(C.Structure):
    _fields_ = [('theta', C.c_double), ('theta_decay', C.c_double), ('max_levels', C.c_int), ('coarse_size', C.c_int),
                ('block', C.c_int), ('omega_scale_set', C.c_int), ('omega_scale', C.c_double)]


class NewtonOpts(C.Structure):
    _fields_ = [('kind', C.c_int), ('atol', C.c_double), ('rtol', C.c_double), ('stol', C.c_double),
                ('max_it', C.c_int), ('krylov', KrylovOpts)]


class NewtonInfo(C.Structure):
    _fields_ = [('iterations', C.c_int), ('converged', C.c_int), ('fnorm0', C.c_double), ('fnorm', C.c_double),
                ('krylov_iterations', C.c_int), ('spmv_count', C.c_int)]


_P = C.c_void_p
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol declared in include/femo_b200.h
SIGNATURES = {
    'femo_version': (C.c_int, []),
    'femo_last_error': (C.c_char_p, []),
    'femo_device_count': (C.c_int, []),
    'femo_mesh_create_unit_square': (C.c_int, [C.c_int, C.c_int, _DP, _DP, C.POINTER(_P)]),
    'femo_mesh_create_rectangle_quad': (C.c_int, [C.c_int, C.c_int, _DP, _DP, C.POINTER(_P)]),
    'femo_mesh_create_from_arrays': (C.c_int, [C.c_int, C.c_int, C.c_int64, _DP, C.c_int64, _I32P, C.POINTER(_P)]),
    'femo_mesh_create_box_hex': (C.c_int, [C.c_int, C.c_int, C.c_int, _DP, _DP, C.POINTER(_P)]),
    'femo_mesh_create_interval': (C.c_int, [C.c_int, C.c_double, C.c_double, C.POINTER(_P)]),
    'femo_mesh_sizes': (C.c_int, [_P, _I64P]),
    'femo_mesh_copy': (C.c_int, [_P, C.c_int, _P]),
    'femo_mesh_destroy': (None, [_P]),
    'femo_problem_create': (C.c_int, [_P, C.c_int, _DP, C.c_int, C.POINTER(_P)]),
    'femo_problem_create_tagged': (C.c_int, [_P, C.c_int, _DP, C.c_int, _P, C.c_int, C.POINTER(_P)]),
    'femo_problem_create_ex': (C.c_int, [_P, C.c_int, _DP, C.c_int, _P, _P, C.c_int, _P, C.POINTER(_P)]),
    'femo_problem_set_param': (C.c_int, [_P, C.c_int, C.c_double]),
    'femo_mesh_create_annulus': (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(_P)]),
    'femo_problem_destroy': (None, [_P]),
    'femo_problem_sizes': (C.c_int, [_P, _I64P]),
    'femo_problem_pattern_info': (C.c_int, [_P, C.c_int, _I64P]),
    'femo_problem_pattern': (C.c_int, [_P, C.c_int, _P, _P]),
    'femo_problem_gather_map': (C.c_int, [_P, C.c_int, _P, _P]),
    'femo_problem_set_bc': (C.c_int, [_P, _P, _P, C.c_int, _P]),
    'femo_comm_unique_id': (C.c_int, [C.c_char_p]),
    'femo_comm_init': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int]),
    'femo_comm_finalize': (C.c_int, []),
    'femo_link_create': (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.c_char_p]),
    'femo_link_open': (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    'femo_link_error': (C.c_int, []),
    'femo_comm_stats': (C.c_int, [C.POINTER(C.c_longlong)]),
    'femo_problem_create_slab': (C.c_int, [C.c_int, _DP, C.c_int, C.c_int, C.c_int, _DP, _DP, C.c_int, C.c_int, C.POINTER(_P)]),
    'femo_problem_create_slab_hex': (C.c_int, [C.c_int, _DP, C.c_int, C.c_int, C.c_int, C.c_int, _DP, _DP, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(_P)]),
    'femo_problem_slab_info': (C.c_int, [_P, _I64P]),
    'femo_halo_exchange': (C.c_int, [_P, _P, C.c_int]),
    'femo_problem_set_partition': (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, _P, _P, _I64P]),
    'femo_problem_mesh_sizes': (C.c_int, [_P, _I64P]),
    'femo_problem_mesh_copy': (C.c_int, [_P, C.c_int, _P]),
    'femo_problem_enable_multigrid': (C.c_int, [_P]),
    'femo_problem_mg_levels': (C.c_int, [_P]),
    'femo_problem_device_bytes': (C.c_int, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    'femo_problem_upload': (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t]),
    'femo_set_coefficient': (C.c_int, [_P, C.c_int, _P, C.c_int64]),
    'femo_problem_launch_count': (C.c_int, [_P, C.POINTER(C.c_longlong)]),
    'femo_problem_graph_replays': (C.c_int, [_P, C.POINTER(C.c_longlong)]),
    'femo_assemble_residual': (C.c_int, [_P, _P]),
    'femo_assemble_jacobian': (C.c_int, [_P, _P, _P]),
    'femo_assemble_dRdm': (C.c_int, [_P, C.c_int, _P]),
    'femo_newton_rhs': (C.c_int, [_P, _P, _P]),
    'femo_assemble_system_rhs': (C.c_int, [_P, _P, _P]),
    'femo_assemble_output': (C.c_int, [_P, C.c_int, _DP]),
    'femo_assemble_output_grad': (C.c_int, [_P, C.c_int, C.c_int, _P]),
    'femo_assemble_output_and_grad': (C.c_int, [_P, C.c_int, _DP, _P]),
    'femo_spmv': (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int]),
    'femo_spmv_bsr3': (C.c_int, [_P, _P, _P, _P, C.c_int]),
    'femo_axpy': (C.c_int, [_P, C.c_double, _P, _P, C.c_int64]),
    'femo_filter_apply': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _P, _P, _P, C.c_int]),
    'femo_filter_apply3': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P, C.c_int]),
    'femo_host_scaled_copy': (None, [_P, _P, C.c_int64, C.c_double]),
    'femo_host_axpy': (None, [_P, _P, C.c_int64, C.c_double]),
    'femo_host_fill': (None, [_P, C.c_int64, C.c_double]),
    'femo_pointwise_divide': (C.c_int, [_P, C.c_double, _P, _P, _P, C.c_int64]),
    'femo_vcycle_op_probe': (C.c_int, [_P, C.c_int, _I64P]),
    'femo_amg_symbolic': (C.c_int, [_P, _P, C.POINTER(AmgOpts), _I64P]),
    'femo_amg_attach': (C.c_int, [_P, _P, C.c_int64]),
    'femo_amg_numeric': (C.c_int, [_P, _P]),
    'femo_amg_level_info': (C.c_int, [_P, C.c_int, _I64P, _DP]),
    'femo_amg_level_array': (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64]),
    'femo_amg_level_values': (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int64]),
    'femo_linear_solve': (C.c_int, [_P, _P, _P, _P, C.c_int, C.POINTER(KrylovOpts), C.POINTER(KrylovInfo)]),
    'femo_newton_solve': (C.c_int, [_P, C.POINTER(NewtonOpts), C.POINTER(NewtonInfo)]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'femo_b200: %s not found. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C femo_b200/csrc`; there is no Python/CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise FemoError(rc, lib.femo_last_error().decode('utf-8', 'replace'))
    return rc
