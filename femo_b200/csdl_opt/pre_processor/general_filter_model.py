"""Density filter of the topology-optimisation example, on the device.

Mirrors examples/beam_topo_opt/pre_processor/general_filter_model.py of the
reference (GeneralFilterModel :8-30, GeneralFilterOperation :33-90): the output is
W x with the cone weights W_ij = (R - d_ij) / sum_k (R - d_ik), R = beta * h_avg.
The reference rebuilds a KD-tree for every cell (O(n^2 log n)); here the cell
centres form a lattice, so the neighbour search is index arithmetic inside the
CUDA kernel and the constant Jacobian is applied matrix-free (W^T for reverse mode).
"""
import ctypes as C

import numpy as np

from .._csdl_compat import Model, CustomExplicitOperation, csdl, HAVE_CSDL
from ..._lib import lib, check


def _lattice(coords):
    """(nx, ny, nz, dx, dy, dz) of cell centres given in lattice order (x fastest, then y, then z; nz = 1 in 2-D)."""
    c = np.asarray(coords)
    x, y = c[:, 0], c[:, 1]
    dec = np.nonzero(x[1:] <= x[:-1])[0]                 # first wrap-around of x ends the first lattice row
    nx = int(dec[0]) + 1 if dec.size else x.size
    if nx <= 0 or x.size % nx:
        raise ValueError('general filter: cell centres are not a lattice in x-fastest order')
    rows = x.size // nx
    ny, nz = rows, 1
    if c.shape[1] > 2 and np.ptp(c[:, 2]) > 0:           # 3-D: y wraps around at the end of every z-layer
        yr = y[::nx]
        decy = np.nonzero(yr[1:] <= yr[:-1])[0]
        ny = int(decy[0]) + 1 if decy.size else rows
        if rows % ny:
            raise ValueError('general filter: cell centres are not a lattice in x-fastest order')
        nz = rows // ny
    dx = (x[nx - 1] - x[0]) / max(nx - 1, 1) if nx > 1 else 1.0
    dy = (y[nx * (ny - 1)] - y[0]) / max(ny - 1, 1) if ny > 1 else 1.0
    dz = (c[-1, 2] - c[0, 2]) / max(nz - 1, 1) if nz > 1 else 1.0
    return nx, ny, nz, float(dx), float(dy), float(dz)


class GeneralFilterModel(Model):
    def initialize(self):
        self.parameters.declare('nel')
        self.parameters.declare('beta', default=2.)
        self.parameters.declare('coordinates')
        self.parameters.declare('h_avg')

    def define(self):
        nel = self.parameters['nel']
        density_unfiltered = self.declare_variable('density_unfiltered', shape=(nel,), val=1.0)
        e = GeneralFilterOperation(nel=nel, beta=self.parameters['beta'], coordinates=self.parameters['coordinates'],
                                   h_avg=self.parameters['h_avg'])
        self.register_output('density', csdl.custom(density_unfiltered, op=e))


class GeneralFilterOperation(CustomExplicitOperation):
    """input: unfiltered density; output: filtered density."""

    def initialize(self):
        self.parameters.declare('nel')
        self.parameters.declare('beta', default=2.)
        self.parameters.declare('coordinates')
        self.parameters.declare('h_avg')

    def define(self):
        nel = self.parameters['nel']
        coords = self.parameters['coordinates']
        self.radius = float(self.parameters['beta'] * self.parameters['h_avg'])
        self.nx, self.ny, self.nz, self.dx, self.dy, self.dz = _lattice(coords)
        if self.nx * self.ny * self.nz != nel:
            raise ValueError('general filter: nel does not match the coordinates')
        self.add_input('density_unfiltered', shape=(nel,), val=0.0)
        self.add_output('density', shape=(nel,))
        if HAVE_CSDL:                       # real CSDL wants the constant sparse Jacobian up front
            rows, cols, vals = self.weight_triplets()
            self.declare_derivatives('density', 'density_unfiltered', rows=rows, cols=cols, val=vals)
        else:
            self.declare_derivatives('density', 'density_unfiltered')
        self._buf = None

    def _apply(self, x, transpose):
        import torch
        if not torch.cuda.is_available():
            from ..._lib import FemoError
            raise FemoError(-2, 'general filter needs a CUDA device; there is no CPU path')
        idev = torch.cuda.current_device()               # the rank's device (torch.cuda.set_device(LOCAL_RANK))
        dev = torch.device('cuda', idev)
        n = self.nx * self.ny * self.nz
        if self._buf is None or self._buf[0].device != dev:
            self._buf = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
        din, dout, den = self._buf
        din.copy_(torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)))
        check(lib.femo_filter_apply3(idev, C.c_void_p(torch.cuda.current_stream().cuda_stream), self.nx, self.ny, self.nz,
                                     self.dx, self.dy, self.dz, self.radius, C.c_void_p(din.data_ptr()),
                                     C.c_void_p(dout.data_ptr()), C.c_void_p(den.data_ptr()), 1 if transpose else 0))
        return dout.cpu().numpy()

    def compute(self, inputs, outputs):
        outputs['density'] = self._apply(inputs['density_unfiltered'], False)

    def vjp(self, of, wrt, bar):
        """W^T bar: the reverse-mode action of the constant Jacobian."""
        return self._apply(bar, True)

    def weight_triplets(self):
        """(rows, cols, weights) of W from lattice offsets (what real CSDL wants as the constant sparse Jacobian)."""
        nx, ny, nz, R = self.nx, self.ny, self.nz, self.radius
        kx, ky = int(np.floor(R / self.dx)), int(np.floor(R / self.dy))
        kz = int(np.floor(R / self.dz)) if nz > 1 else 0
        K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
        dist = lambda a, b, c: np.sqrt((a * self.dx) ** 2 + (b * self.dy) ** 2 + (c * self.dz) ** 2)
        offs = [(a, b, c) for c in range(-kz, kz + 1) for b in range(-ky, ky + 1) for a in range(-kx, kx + 1)
                if dist(a, b, c) <= R]
        inside = lambda a, b, c: ((I + a >= 0) & (I + a < nx) & (J + b >= 0) & (J + b < ny) & (K + c >= 0) & (K + c < nz))
        den = np.zeros((nz, ny, nx))
        for a, b, c in offs:
            den += inside(a, b, c) * (R - dist(a, b, c))
        rows, cols, vals = [], [], []
        for a, b, c in offs:
            ok = inside(a, b, c)
            w = (R - dist(a, b, c)) / den
            rows.append((((K * ny) + J) * nx + I)[ok])
            cols.append(((((K + c) * ny) + (J + b)) * nx + (I + a))[ok])
            vals.append(w[ok])
        return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
