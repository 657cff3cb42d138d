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
    """(nx, ny, dx, dy) of cell centres given in lattice order (x fastest)."""
    x, y = np.asarray(coords)[:, 0], np.asarray(coords)[:, 1]
    dec = np.nonzero(x[1:] <= x[:-1])[0]                 # first wrap-around of x ends the first lattice row
    nx = int(dec[0]) + 1 if dec.size else x.size
    if nx <= 0 or x.size % nx:
        raise ValueError('general filter: cell centres are not a lattice in x-fastest order')
    ny = x.size // nx
    dx = (x[nx - 1] - x[0]) / max(nx - 1, 1) if nx > 1 else 1.0
    dy = (y[-1] - y[0]) / max(ny - 1, 1) if ny > 1 else 1.0
    return nx, ny, float(dx), float(dy)


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
        self.nx, self.ny, self.dx, self.dy = _lattice(coords)
        if self.nx * self.ny != nel:
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
        dev = torch.device('cuda', 0)
        n = self.nx * self.ny
        if self._buf is None:
            self._buf = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
        din, dout, den = self._buf
        din.copy_(torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)))
        check(lib.femo_filter_apply(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), self.nx, self.ny, self.dx,
                                    self.dy, self.radius, C.c_void_p(din.data_ptr()), C.c_void_p(dout.data_ptr()),
                                    C.c_void_p(den.data_ptr()), 1 if transpose else 0))
        return dout.cpu().numpy()

    def compute(self, inputs, outputs):
        outputs['density'] = self._apply(inputs['density_unfiltered'], False)

    def vjp(self, of, wrt, bar):
        """W^T bar: the reverse-mode action of the constant Jacobian."""
        return self._apply(bar, True)

    def weight_triplets(self):
        """(rows, cols, weights) of W, by applying the filter to unit vectors of one lattice window."""
        n, nx, ny, R = self.nx * self.ny, self.nx, self.ny, self.radius
        kx, ky = int(np.floor(R / self.dx)), int(np.floor(R / self.dy))
        rows, cols, vals = [], [], []
        I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing='xy')
        den = np.zeros((ny, nx))
        offs = [(a, b) for b in range(-ky, ky + 1) for a in range(-kx, kx + 1)
                if np.hypot(a * self.dx, b * self.dy) <= R]
        for a, b in offs:
            ok = (I + a >= 0) & (I + a < nx) & (J + b >= 0) & (J + b < ny)
            den += ok * (R - np.hypot(a * self.dx, b * self.dy))
        for a, b in offs:
            ok = (I + a >= 0) & (I + a < nx) & (J + b >= 0) & (J + b < ny)
            w = (R - np.hypot(a * self.dx, b * self.dy)) / den
            rows.append((J * nx + I)[ok])
            cols.append(((J + b) * nx + (I + a))[ok])
            vals.append(w[ok])
        return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
