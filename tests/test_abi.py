"""The drop-in boundary: libfemo_b200.so loads without a GPU, exports every entry point include/femo_b200.h declares,
the ctypes table of femo_b200/_lib.py (the binding INTEGRATION.md shows to a femo maintainer) covers exactly that set,
and the device entry points fail loudly -- with an error code and a message, never a CPU fallback -- when no CUDA
device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'femo_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)                      # comments mention functions too
    names = re.findall(r'^\s*(?:const\s+)?(?:int|void|char|double|size_t)\s*\**\s*(femo_[A-Za-z0-9_]+)\s*\(', src, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 45 and 'femo_newton_solve' in names and 'femo_problem_create_slab_hex' in names
    lib = C.CDLL(os.path.join(ROOT, 'femo_b200', 'libfemo_b200.so'))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from femo_b200 import _lib
    names = declared_symbols()
    assert sorted(_lib.SIGNATURES) == names, (sorted(set(names) - set(_lib.SIGNATURES)), sorted(set(_lib.SIGNATURES) - set(names)))
    # struct layouts the header declares
    assert C.sizeof(_lib.KrylovOpts) == 64 and C.sizeof(_lib.KrylovInfo) == 32       # 56 + the forcing field
    assert [f[0] for f in _lib.KrylovOpts._fields_] == ['rtol', 'atol', 'max_it', 'precond', 'cheb_degree', 'method', 'restart',
                                                        'check_every', 'cheb_ratio', 'mg_precision', 'forcing']


def test_host_side_entry_points_work_without_gpu():
    from femo_b200 import engine as E
    assert E.lib.femo_version() >= 100
    m = E.EngineMesh.unit_square(4)
    p = E.EngineProblem(m, E.FAMILY_NLPOISSON_P1)
    assert p.N == 25 and p.pattern_info(0)['nnz'] > 0
    ptr, src = p.gather_map(0)
    assert ptr[-1] == src.size == p.pattern_info(0)['ncontrib']


def test_device_entry_points_fail_loudly_without_gpu():
    from femo_b200 import engine as E
    from femo_b200._lib import FemoError, lib
    if E.device_count() > 0:
        pytest.skip('CUDA device present')
    p = E.EngineProblem(E.EngineMesh.unit_square(2), E.FAMILY_POISSON_P1)
    out = np.zeros(p.N)
    rc = lib.femo_assemble_residual(p._h, out.ctypes.data_as(C.c_void_p))
    assert rc != 0 and lib.femo_last_error()                         # error code + message, no silent CPU path
    with pytest.raises(FemoError):
        p.upload(0)
