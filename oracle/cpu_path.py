"""ctypes binding of oracle/libfemo_cpu.so (cpu_path.cpp): the C++/OpenMP restatement of the benched hot path.
TEST / BASELINE INFRASTRUCTURE ONLY -- imported by tests/ and by bench.py's cpu legs, never by femo_b200/."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libfemo_cpu.so')
_DP = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s not found: build it with `make -C oracle` (or __graft_entry__.build())' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.femo_cpu_nlp_residual.argtypes = [C.c_int, _DP, _DP, _DP]
    lib.femo_cpu_nlp_residual.restype = None
    lib.femo_cpu_nlp_jacobian_dia.argtypes = [C.c_int, _DP, _DP]
    lib.femo_cpu_nlp_jacobian_dia.restype = None
    lib.femo_cpu_nlp_output.argtypes = [C.c_int, _DP, _DP, C.POINTER(C.c_double), _DP, _DP]
    lib.femo_cpu_nlp_output.restype = None
    lib.femo_cpu_nlp_step.argtypes = [C.c_int, _DP, _DP, _DP, _DP, C.POINTER(C.c_double), C.c_double, C.c_int,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.femo_cpu_nlp_step.restype = C.c_int
    return lib


lib = _load()


def residual(n, u, f):
    R = np.empty((n + 1) ** 2)
    lib.femo_cpu_nlp_residual(n, np.ascontiguousarray(u, dtype=np.float64), np.ascontiguousarray(f, dtype=np.float64), R)
    return R


def jacobian(n, u):
    """scipy CSR of the un-BC'd Jacobian (built from the 7 DIA planes the C++ code assembles)."""
    import scipy.sparse as sp
    N, w = (n + 1) ** 2, n + 1
    planes = np.empty(7 * N)
    lib.femo_cpu_nlp_jacobian_dia(n, np.ascontiguousarray(u, dtype=np.float64), planes)
    planes = planes.reshape(7, N)
    offs = [-w - 1, -w, -1, 0, 1, w, w + 1]
    rows, cols, vals = [], [], []
    i = np.arange(N)
    for s, o in enumerate(offs):
        j = i + o
        ok = (j >= 0) & (j < N)
        rows.append(i[ok]); cols.append(j[ok]); vals.append(planes[s][ok])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return A


def output(n, u, f):
    J = C.c_double()
    dJdu, dJdf = np.empty((n + 1) ** 2), np.empty(2 * n * n)
    lib.femo_cpu_nlp_output(n, np.ascontiguousarray(u, dtype=np.float64), np.ascontiguousarray(f, dtype=np.float64),
                            C.byref(J), dJdu, dJdf)
    return J.value, dJdu, dJdf


def step(n, f, krylov_rtol=1e-10, nthreads=0):
    """One state+adjoint solve from u = 0.  Returns dict(u, lam, grad, J, newton_its, krylov_its, adjoint_its, ...)."""
    N, M = (n + 1) ** 2, 2 * n * n
    f = np.ascontiguousarray(np.broadcast_to(np.asarray(f, dtype=np.float64), (M,)))
    u, lam, grad = np.empty(N), np.zeros(N), np.empty(M)
    J = C.c_double()
    info = (C.c_int64 * 8)()
    fn = (C.c_double * 2)()
    rc = lib.femo_cpu_nlp_step(n, f, u, lam, grad, C.byref(J), krylov_rtol, nthreads, info, fn)
    return dict(u=u, lam=lam, grad=grad, J=J.value, newton_its=int(info[0]), krylov_its=int(info[1]),
                adjoint_its=int(info[2]), reason=int(info[3]), vcycles=int(info[4]), threads=int(info[5]),
                fnorm0=fn[0], fnorm=fn[1], converged=(rc == 0))
