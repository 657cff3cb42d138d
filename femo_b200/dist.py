"""Multi-GPU plumbing on the Python side: one process per GPU (torchrun).
torch.distributed only ships 64-byte IPC handles (or the 128-byte NCCL id); all data-path communication is issued by
libfemo_b200 on its own stream (include/femo_b200.h, "multi-GPU"): by default through the engine's own peer-memory
transport (csrc/link.cuh; FEMO_COMM=link), optionally through NCCL (FEMO_COMM=nccl)."""
import ctypes as C
import os

import numpy as np

from ._lib import lib, check
from . import engine as _E

_state = dict(initialised=False, rank=0, nranks=1)


def init(device=None, backend=None):
    """Create the engine's communicator from the torch.distributed world (any torch backend: it only carries the
    handles).  backend 'link' (default): peer-memory windows, also valid for several ranks on one device;
    'nccl': ncclSend/Recv + ncclAllReduce."""
    import torch
    import torch.distributed as dist
    if _state['initialised']:
        return _state['rank'], _state['nranks']
    rank, nranks = dist.get_rank(), dist.get_world_size()
    device = int(os.environ.get('LOCAL_RANK', rank)) if device is None else device
    backend = backend or os.environ.get('FEMO_COMM', 'link')
    on_gpu = dist.get_backend() == 'nccl'
    tdev = torch.device('cuda', device) if on_gpu else torch.device('cpu')
    if backend == 'link':
        h = C.create_string_buffer(64)
        check(lib.femo_link_create(device, 0, 0, h))
        mine = torch.tensor(list(h.raw), dtype=torch.uint8, device=tdev)
        parts = [torch.empty_like(mine) for _ in range(nranks)]
        dist.all_gather(parts, mine)
        raw = b''.join(bytes(t.cpu().tolist()) for t in parts)
        check(lib.femo_link_open(raw, rank, nranks))
        _state.update(initialised=True, rank=rank, nranks=nranks, backend='link')
        return rank, nranks
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib.femo_comm_unique_id(buf))
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=tdev)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    check(lib.femo_comm_init(raw, rank, nranks, device))
    _state.update(initialised=True, rank=rank, nranks=nranks, backend='nccl')
    return rank, nranks


def finalize():
    if _state['initialised']:
        lib.femo_comm_finalize()
        _state['initialised'] = False


def stats():
    s = (C.c_longlong * 2)()
    check(lib.femo_comm_stats(s))
    return dict(halo_exchanges=int(s[0]), allreduces=int(s[1]), backend=_state.get('backend'),
                link_error=int(lib.femo_link_error()))


class SlabProblem(_E.EngineProblem):
    """The local problem of one rank: y-slab of an (nx x gny)-cell triangle lattice
    on [lo, hi] with a one-cell ghost layer (host-side layout needs no GPU)."""

    def __init__(self, family, nx, gny, rank, nranks, lo=(0.0, 0.0), hi=(1.0, 1.0), params=(), ny=None, face_mask=0):
        """Triangle families: y-slab of the (nx x gny) lattice.  Hexahedral SIMP family (pass ny): z-slab of the
        (nx x ny x gny) box, lo / hi with three entries, face_mask = box faces carrying the traction."""
        h = C.c_void_p()
        pa = (C.c_double * max(1, len(params)))(*params)
        if ny is None:
            check(lib.femo_problem_create_slab(int(family), pa, len(params), int(nx), int(gny), (C.c_double * 2)(*lo),
                                               (C.c_double * 2)(*hi), int(rank), int(nranks), C.byref(h)))
        else:
            check(lib.femo_problem_create_slab_hex(int(family), pa, len(params), int(nx), int(ny), int(gny),
                                                   (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), int(rank), int(nranks),
                                                   int(face_mask), C.byref(h)))
        self.mesh = None
        self.family = family
        self._h = h
        s = (C.c_int64 * 16)()
        check(lib.femo_problem_sizes(self._h, s))
        self.N, self.nin, self.naux, self.nout = int(s[0]), int(s[1]), int(s[2]), int(s[3])
        self.M = [int(s[4 + i]) for i in range(self.nin)]
        self.aux_sizes = [int(s[8 + i]) for i in range(self.naux)]
        self.uploaded = False
        self.h2d_bytes = self.d2h_bytes = 0
        self._keep, self._pinned = {}, {}
        self.device = None
        i = (C.c_int64 * 16)()
        check(lib.femo_problem_slab_info(self._h, i))
        keys = ('active', 'rank', 'nranks', 'gny', 'crow0', 'ncrows', 'own0', 'own1', 'cown0', 'cown1', 'own_off',
                'own_n', 'cown_off', 'cown_n', 'nx', 'ny_local')
        self.slab = {k: int(v) for k, v in zip(keys, i)}

    def local_coords(self):
        s = (C.c_int64 * 6)()
        check(lib.femo_problem_mesh_sizes(self._h, s))
        out = np.empty((int(s[1]), int(s[3])), dtype=np.float64)
        check(lib.femo_problem_mesh_copy(self._h, 0, out.ctypes.data_as(C.c_void_p)))
        return out

    def local_cells(self):
        s = (C.c_int64 * 6)()
        check(lib.femo_problem_mesh_sizes(self._h, s))
        out = np.empty((int(s[0]), int(s[2])), dtype=np.int32)
        check(lib.femo_problem_mesh_copy(self._h, 1, out.ctypes.data_as(C.c_void_p)))
        return out

    def halo(self, tensor, kind=0):
        check(lib.femo_halo_exchange(self._h, C.c_void_p(tensor.data_ptr()), kind))
        return tensor

    def owned(self, tensor):
        """The owned block of a local state-space vector."""
        return tensor[self.slab['own_off']:self.slab['own_off'] + self.slab['own_n']]


class PartProblem(_E.EngineProblem):
    """One rank's problem on an unstructured partition (femo_b200/partition.py): the local mesh -- owned cells + the
    ghost cells around the owned vertices, owned vertices numbered first -- as an ordinary engine problem plus the
    scatter lists of the ghost refresh (femo_problem_set_partition).  Assembly needs no communication; SpMV / Krylov /
    Newton exchange halos and all-reduce inside the engine exactly as on slabs."""

    def __init__(self, family, views, rank, kind='triangle', params=(), cell_tags=None, facets=None):
        """cell_tags: one subdomain id per GLOBAL cell.  facets: (global cells, local facet ids) of the one-sided facets that
        carry facet integrals -- pass the GLOBAL mesh's exterior_facets() for forms with boundary terms (the cut faces of a
        local mesh are not boundaries); they are restricted to this rank's local cells here."""
        from . import partition as _P
        self.view = v = views[rank]
        self.views = views
        self._layout = _P.gather_layout(views)
        mesh = _E.EngineMesh.from_arrays(kind, v.coords, v.cells)
        lt = None if cell_tags is None else np.asarray(cell_tags, dtype=np.int32)[v.cells_global]
        lf = None
        if facets is not None:
            gc, gl = np.asarray(facets[0], dtype=np.int64), np.asarray(facets[1], dtype=np.int32)
            ncg = 1 + max(int(m.cells_global.max()) for m in views)
            g2l = np.full(ncg, -1, dtype=np.int64)
            g2l[v.cells_global] = np.arange(v.cells_global.size)
            keep = g2l[gc] >= 0
            lc, ll = g2l[gc[keep]], gl[keep]
            o = np.lexsort((ll, lc))
            lf = (lc[o].astype(np.int32), ll[o])
        super().__init__(mesh, family, params, facets=lf, cell_tags=lt)

    def upload(self, device=0):
        import torch
        super().upload(device)
        v = self.view
        blk, send_nodes, ghost_src = self._layout
        sn = np.ascontiguousarray(send_nodes[v.rank], dtype=np.int32)
        gs = np.ascontiguousarray(ghost_src[v.rank], dtype=np.int32)
        nbytes = C.c_int64(0)
        args = (self._h, int(v.n_owned_verts), int(v.n_owned_cells), int(blk), sn.size, sn.ctypes.data_as(C.c_void_p), gs.size,
                gs.ctypes.data_as(C.c_void_p))
        check(lib.femo_problem_set_partition(*args, None, C.byref(nbytes)))
        self._part_buf = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        check(lib.femo_problem_set_partition(*args, C.c_void_p(self._part_buf.data_ptr()), C.byref(nbytes)))
        return self

    def owned(self, tensor):
        b = self.N // self.view.verts_global.size
        return tensor[:self.view.n_owned_verts * b]

    def halo(self, tensor, kind=0):
        check(lib.femo_halo_exchange(self._h, C.c_void_p(tensor.data_ptr()), 0))
        return tensor
