"""Cell partition of an unstructured mesh with a one-cell ghost layer (host side, numpy).

The reference creates its meshes on MPI.COMM_WORLD and lets dolfinx partition them and build the index maps
(/root/reference/femo/fea/utils_dolfinx.py:32,69-123,140-153; ghost mode "shared_facet" [upstream, from memory]).
This module is the engine-side counterpart for meshes that are not lattices (Gmsh / array meshes), with the same
ownership rules the slab partition of csrc/dist.cuh uses, so the same argument gives assembly without communication:

  * cells are assigned to ranks by recursive coordinate bisection of their centroids (deterministic: ties are broken
    by the global cell index);
  * a vertex is owned by the LOWEST rank among the cells that contain it;
  * a rank's local mesh = its owned cells + every other cell that touches one of its owned vertices (ghost cells): all
    cells around an owned vertex are local, so the owned rows of every assembled vector and matrix are complete;
  * local numbering: owned vertices first (ascending global id), then ghost vertices grouped by owner rank (ascending
    rank, ascending global id inside a group); owned cells first, then ghost cells (both ascending).  Reductions run
    over [0, n_owned); the ghosts owned by one neighbour form ONE contiguous range, so a halo exchange needs a
    gather (pack) on the sender only and receives in place;
  * per neighbour q: `send[q]` = local indices of the owned vertices q holds as ghosts, in exactly the order of q's
    receive range `recv[q] = (start, stop)`.

Every rank can build every rank's view from the replicated mesh (no communication in the set-up).
"""
import numpy as np


def rcb(centroids, nparts):
    """Recursive coordinate bisection -> part[cell] in [0, nparts).  Splits the longest extent at the k/nparts quantile,
    so any nparts works; equal coordinates are ordered by the global cell index."""
    centroids = np.asarray(centroids, dtype=np.float64)
    part = np.zeros(centroids.shape[0], dtype=np.int32)

    def rec(idx, p0, npart):
        if npart == 1:
            part[idx] = p0
            return
        k = npart // 2
        c = centroids[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = np.lexsort((idx, c[:, axis]))
        cut = (idx.size * k) // npart
        rec(idx[order[:cut]], p0, k)
        rec(idx[order[cut:]], p0 + k, npart - k)

    rec(np.arange(centroids.shape[0]), 0, int(nparts))
    return part


class RankMesh:
    """One rank's local mesh and its scatter lists (all indices int32 / int64 numpy arrays)."""

    def __init__(self, rank, nranks):
        self.rank, self.nranks = rank, nranks
        self.cells_global = self.verts_global = None     # local -> global
        self.n_owned_cells = self.n_owned_verts = 0
        self.cells = self.coords = None                  # local connectivity / coordinates
        self.ghost_owner = None                          # owner rank of each ghost vertex (local order)
        self.neighbours = []
        self.recv, self.send = {}, {}

    def dof_lists(self, block=1):
        """Send lists / receive ranges for a vertex space with `block` interleaved dofs per vertex."""
        send = {q: (idx[:, None] * block + np.arange(block)[None, :]).ravel() for q, idx in self.send.items()}
        recv = {q: (a * block, b * block) for q, (a, b) in self.recv.items()}
        return send, recv


def vertex_owner(cells, part, nverts, nranks):
    owner = np.full(nverts, nranks, dtype=np.int32)
    np.minimum.at(owner, cells.ravel(), np.repeat(part.astype(np.int32), cells.shape[1]))
    return owner


def partition_mesh(coords, cells, nranks, part=None):
    """-> (part, owner, [RankMesh for every rank]).  coords (nverts, gdim), cells (ncells, k) global connectivity."""
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    nverts = coords.shape[0]
    if part is None:
        part = rcb(coords[cells].mean(axis=1), nranks)
    part = np.asarray(part, dtype=np.int32)
    owner = vertex_owner(cells, part, nverts, nranks)
    if owner.max() >= nranks:
        raise ValueError('partition_mesh: a vertex belongs to no cell')
    views = []
    for r in range(nranks):
        m = RankMesh(r, nranks)
        owned_c = np.nonzero(part == r)[0]
        ghost_c = np.nonzero((part != r) & (owner[cells] == r).any(axis=1))[0]
        m.cells_global = np.concatenate([owned_c, ghost_c])
        m.n_owned_cells = owned_c.size
        used = np.unique(cells[m.cells_global])
        own_v = used[owner[used] == r]
        gh_v = used[owner[used] != r]
        gh_v = gh_v[np.lexsort((gh_v, owner[gh_v]))]
        m.verts_global = np.concatenate([own_v, gh_v])
        m.n_owned_verts = own_v.size
        m.ghost_owner = owner[gh_v]
        g2l = np.full(nverts, -1, dtype=np.int64)
        g2l[m.verts_global] = np.arange(m.verts_global.size)
        m.cells = g2l[cells[m.cells_global]].astype(np.int32)
        m.coords = coords[m.verts_global]
        m.neighbours = sorted(int(q) for q in np.unique(m.ghost_owner))
        for q in m.neighbours:
            sel = np.nonzero(m.ghost_owner == q)[0]
            m.recv[q] = (int(m.n_owned_verts + sel[0]), int(m.n_owned_verts + sel[-1] + 1))
        m._g2l = g2l
        views.append(m)
    # send lists: what q receives from r, in q's order, as r's local indices
    for q, mq in enumerate(views):
        for r, (a, b) in mq.recv.items():
            views[r].send[q] = views[r]._g2l[mq.verts_global[a:b]].astype(np.int32)
    for m in views:
        m.neighbours = sorted(set(m.neighbours) | set(m.send))
        del m._g2l
    return part, owner, views


def exchange(views_or_view, vec, block=1, group=None):
    """Refresh the ghost entries of a local vertex vector.

    With a list of all ranks' views and a list of their vectors: done in-process (tests, single process).  With ONE view
    and a torch.distributed process group: point-to-point exchange between the ranks (gloo on CPU, one rank per process)."""
    if isinstance(views_or_view, (list, tuple)):
        views, vecs = views_or_view, vec
        for m in views:
            send, _ = m.dof_lists(block)
            for q, idx in send.items():
                a, b = views[q].dof_lists(block)[1][m.rank]
                vecs[q][a:b] = vecs[m.rank][idx]
        return vecs
    import torch
    import torch.distributed as dist
    m = views_or_view
    send, recv = m.dof_lists(block)
    t = torch.from_numpy(vec)
    reqs, bufs = [], {}
    for q in sorted(recv):
        a, b = recv[q]
        bufs[q] = torch.empty(b - a, dtype=t.dtype)
        reqs.append(dist.irecv(bufs[q], src=q, group=group))
    for q in sorted(send):
        reqs.append(dist.isend(t[torch.from_numpy(send[q].astype(np.int64))].contiguous(), dst=q, group=group))
    for rq in reqs:
        rq.wait()
    for q, (a, b) in recv.items():
        t[a:b] = bufs[q]
    return vec


def gather_layout(views):
    """Index lists of the engine's ghost refresh (csrc/dist_ops.cuh, gpart_halo): every rank packs the owned values its
    neighbours need into a send buffer (segments ordered by destination rank), the buffers of all ranks -- padded to the
    longest, `blk` nodes -- are all-gathered, and ghost j of rank r reads entry q*blk + offset(q -> r) + position.
    -> blk, [send_nodes of rank r], [ghost_src of rank r] (node level; the engine expands by the block size)."""
    R = len(views)
    send_nodes, offsets = [], []
    for m in views:
        off, parts, o = {}, [], 0
        for q in sorted(m.send):
            off[q] = o
            parts.append(m.send[q])
            o += m.send[q].size
        send_nodes.append(np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, dtype=np.int32))
        offsets.append(off)
    blk = max([s.size for s in send_nodes] + [1])
    ghost_src = []
    for m in views:
        g = np.zeros(m.verts_global.size - m.n_owned_verts, dtype=np.int32)
        for q, (a, b) in m.recv.items():
            g[a - m.n_owned_verts:b - m.n_owned_verts] = q * blk + offsets[q][m.rank] + np.arange(b - a)
        ghost_src.append(g)
    return blk, send_nodes, ghost_src
