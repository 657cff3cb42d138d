"""Host-side bulk array helpers for the API boundary.

The femo callbacks exchange dense fp64 numpy vectors; at 16M dofs a single-threaded
numpy pass over one of them costs as much as a multigrid cycle on the GPU.  These helpers run the large
copies / accumulations through torch's multi-threaded CPU kernels on the SAME memory (zero-copy views) and
hand out page-locked storage so the host<->device DMA needs no staging copy.  Small arrays stay on numpy.
"""
import numpy as np

BIG = 1 << 18


class TrackedArray(np.ndarray):
    """Variable storage whose OWNER bumps `.version` on every write it performs or permits.  `update()`
    skips re-uploading a tracked source it already holds at the same version (each femo callback pushes all
    its inputs again, state_model.py:75-200); untracked arrays are always copied."""
    version = None

    def __array_finalize__(self, obj):
        self.version = None                     # views / copies are not tracked

    def bump(self):
        self.version = (self.version or 0) + 1


def tracked(a):
    t = a.view(TrackedArray)
    t.version = 1
    return t


def _t(a):
    import torch
    return torch.from_numpy(a)


def _ptr(a):
    return a.ctypes.data


def _lib():
    from ._lib import lib
    return lib


def _big(a):
    return (isinstance(a, np.ndarray) and a.size >= BIG and a.dtype == np.float64 and a.flags.c_contiguous
            and a.flags.writeable)


def _big_src(a):
    return isinstance(a, np.ndarray) and a.size >= BIG and a.dtype == np.float64 and a.flags.c_contiguous


def pinned_empty(n):
    """Page-locked fp64 host array when a CUDA device exists, plain numpy otherwise."""
    n = int(n)
    if n >= BIG:
        import torch
        if torch.cuda.is_available():
            return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    return np.empty(n)


def is_pinned(a):
    if not _big(a):
        return False
    import torch
    return torch.cuda.is_available() and _t(a).is_pinned()


def copy(dst, src):
    src = np.asarray(src, dtype=np.float64)
    if _big(dst) and _big_src(src) and src.size == dst.size:
        _lib().femo_host_scaled_copy(_ptr(dst), _ptr(src), dst.size, 1.0)       # all host cores (csrc/hostops.cpp)
    elif src.size == 1:
        dst.fill(float(src.ravel()[0]))
    else:
        np.copyto(dst, src.reshape(dst.shape))


def fill(dst, value):
    if _big(dst):
        _lib().femo_host_fill(_ptr(dst), dst.size, float(value))
    else:
        dst.fill(value)


def iadd(dst, src, alpha=1.0):
    """dst += alpha * src in place."""
    src = np.asarray(src, dtype=np.float64)
    if _big(dst) and _big_src(src) and src.size == dst.size:
        _lib().femo_host_axpy(_ptr(dst), _ptr(src), dst.size, float(alpha))
    elif alpha == 1.0:
        dst += src.reshape(dst.shape)
    else:
        dst += alpha * src.reshape(dst.shape)


def scaled_copy(dst, src, alpha=1.0):
    """dst = alpha * src."""
    src = np.asarray(src, dtype=np.float64)
    if alpha == 1.0:
        copy(dst, src)
    elif _big(dst) and _big_src(src) and src.size == dst.size:
        _lib().femo_host_scaled_copy(_ptr(dst), _ptr(src), dst.size, float(alpha))
    else:
        np.multiply(src.reshape(dst.shape), alpha, out=dst)
