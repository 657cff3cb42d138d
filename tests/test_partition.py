"""femo_b200/partition.py: recursive-coordinate-bisection cell partition of unstructured meshes, lowest-rank vertex
ownership, one-cell ghost layer, owned-first local numbering and the send / receive lists -- what dolfinx's mesh
partitioner and index maps provide the reference under MPI.COMM_WORLD
(/root/reference/femo/fea/utils_dolfinx.py:32,69-123,140-153).  Integer results are checked exactly against a
brute-force construction; a world_size-2 gloo run exchanges ghosts with the lists."""
import multiprocessing as mp
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from femo_b200 import partition as P
from oracle import mesh as om


def _mesh(n=12, seed=0):
    m = om.unit_square_tri(n, n + 3)
    x = m.coords.copy()
    rng = np.random.default_rng(seed)
    inner = (x[:, 0] > 1e-9) & (x[:, 0] < 1 - 1e-9) & (x[:, 1] > 1e-9) & (x[:, 1] < 1 - 1e-9)
    x[inner] += 0.3 / n * (rng.random((inner.sum(), 2)) - 0.5)
    perm = rng.permutation(m.cells.shape[0])               # no lattice order left in the cell numbering
    return x, m.cells[perm].astype(np.int64)


def _stiffness(x, cells, n):
    """P1 Laplacian by explicit element loops (enough to check which rows a local assembly completes)."""
    rows, cols, vals = [], [], []
    for c in cells:
        X = x[c]
        B = np.array([X[1] - X[0], X[2] - X[0]]).T
        area = 0.5 * abs(np.linalg.det(B))
        G = np.linalg.solve(B.T, np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]).T).T
        K = area * G @ G.T
        for a in range(3):
            for b in range(3):
                rows.append(c[a]); cols.append(c[b]); vals.append(K[a, b])
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


@pytest.mark.parametrize('R', [2, 3, 5, 8])
def test_rcb_is_balanced_and_deterministic(R):
    x, cells = _mesh()
    cen = x[cells].mean(axis=1)
    part = P.rcb(cen, R)
    counts = np.bincount(part, minlength=R)
    assert counts.min() >= cells.shape[0] // R - 1 and counts.max() <= cells.shape[0] // R + R
    assert np.array_equal(part, P.rcb(cen, R))
    # parts are separated by axis-aligned cuts: bounding boxes of two parts overlap in at most a thin strip per axis
    if R == 2:
        axis = int(np.argmax(cen.max(axis=0) - cen.min(axis=0)))
        assert cen[part == 0][:, axis].max() <= cen[part == 1][:, axis].min()


@pytest.mark.parametrize('R', [2, 3, 4])
def test_views_against_brute_force(R):
    x, cells = _mesh(10, seed=R)
    nv = x.shape[0]
    part, owner, views = P.partition_mesh(x, cells, R)
    # every cell and every vertex is owned exactly once
    oc = np.concatenate([m.cells_global[:m.n_owned_cells] for m in views])
    ov = np.concatenate([m.verts_global[:m.n_owned_verts] for m in views])
    assert np.array_equal(np.sort(oc), np.arange(cells.shape[0])) and np.array_equal(np.sort(ov), np.arange(nv))
    for v in range(nv):                                     # owner = lowest rank among the cells around the vertex
        assert owner[v] == min(part[c] for c in range(cells.shape[0]) if v in cells[c])
    A = _stiffness(x, cells, nv)
    for m in views:
        r = m.rank
        local_cells = set(m.cells_global.tolist())
        for v in m.verts_global[:m.n_owned_verts]:          # all cells around an owned vertex are local
            assert all(c in local_cells for c in range(cells.shape[0]) if v in cells[c])
        # ghost cells: exactly the foreign cells touching an owned vertex
        want = sorted(c for c in range(cells.shape[0]) if part[c] != r and any(owner[v] == r for v in cells[c]))
        assert m.cells_global[m.n_owned_cells:].tolist() == want
        assert np.all(np.diff(m.cells_global[:m.n_owned_cells]) > 0) and np.all(np.diff(m.verts_global[:m.n_owned_verts]) > 0)
        assert np.array_equal(m.verts_global[m.cells], cells[m.cells_global]) and np.array_equal(m.coords, x[m.verts_global])
        # ghosts grouped by owner, ascending global id inside a group, ranges = recv
        go = m.ghost_owner
        assert np.all(np.diff(go) >= 0)
        for q, (a, b) in m.recv.items():
            assert np.all(owner[m.verts_global[a:b]] == q) and np.all(np.diff(m.verts_global[a:b]) > 0)
            s = views[q].send[r]
            assert np.all(s < views[q].n_owned_verts) and np.array_equal(views[q].verts_global[s], m.verts_global[a:b])
        # local assembly completes the owned rows: rows of the global matrix, columns in local numbering
        Al = _stiffness(m.coords, m.cells.astype(np.int64), m.verts_global.size).tocsr()
        g = m.verts_global
        Ag = A[g][:, g].tocsr()
        d = (Al - Ag)[:m.n_owned_verts]
        assert abs(d).max() < 1e-13 * abs(A).max()
    # in-process exchange: every ghost ends up with its owner's value (block 2)
    vecs = []
    for m in views:
        v = np.full(2 * m.verts_global.size, -1.0)
        no = m.n_owned_verts
        v[:2 * no:2] = m.verts_global[:no]
        v[1:2 * no:2] = m.verts_global[:no] + 0.5
        vecs.append(v)
    P.exchange(views, vecs, block=2)
    for m, v in zip(views, vecs):
        assert np.array_equal(v[0::2], m.verts_global) and np.array_equal(v[1::2], m.verts_global + 0.5)


@pytest.mark.parametrize('R', [2, 3, 4])
def test_gather_layout_reproduces_the_exchange(R):
    """The engine's ghost refresh (csrc/dist_ops.cuh gpart_halo) = pack, all-gather of equal blocks, scatter: emulated
    here with the index lists handed to femo_problem_set_partition, block size 2."""
    x, cells = _mesh(10, seed=10 + R)
    _, _, views = P.partition_mesh(x, cells, R)
    blk, send_nodes, ghost_src = P.gather_layout(views)
    b = 2
    gathered = np.full(R * blk * b, np.nan)
    vecs = []
    for m in views:
        v = np.full(b * m.verts_global.size, -7.0)
        no = m.n_owned_verts
        v[0:b * no:b] = m.verts_global[:no]
        v[1:b * no:b] = -m.verts_global[:no] - 0.25
        vecs.append(v)
        assert send_nodes[m.rank].size <= blk and np.all(send_nodes[m.rank] < no)
        sd = (send_nodes[m.rank][:, None] * b + np.arange(b)[None, :]).ravel()          # the engine's expansion by the block size
        gathered[m.rank * blk * b:m.rank * blk * b + sd.size] = v[sd]
    for m, v in zip(views, vecs):
        q, pos = ghost_src[m.rank] // blk, ghost_src[m.rank] % blk
        gd = ((q * blk * b + pos * b)[:, None] + np.arange(b)[None, :]).ravel()
        v[b * m.n_owned_verts:] = gathered[gd]
        assert np.array_equal(v[0::b], m.verts_global) and np.array_equal(v[1::b], -m.verts_global - 0.25)


def _worker(rank, R, port, q):
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=R)
    try:
        x, cells = _mesh(14, seed=7)
        _, _, views = P.partition_mesh(x, cells, R)
        m = views[rank]
        A = _stiffness(x, cells, x.shape[0])
        xg = np.sin(3 * x[:, 0]) + x[:, 1] ** 2
        v = np.full(m.verts_global.size, np.nan)
        v[:m.n_owned_verts] = xg[m.verts_global[:m.n_owned_verts]]
        P.exchange(m, v)                                     # ghost refresh over gloo
        Al = _stiffness(m.coords, m.cells.astype(np.int64), m.verts_global.size).tocsr()
        y = (Al @ v)[:m.n_owned_verts]                        # distributed SpMV: owned rows only
        err = np.abs(y - (A @ xg)[m.verts_global[:m.n_owned_verts]]).max()
        q.put((rank, float(err), bool(np.array_equal(v, xg[m.verts_global]))))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_ghost_exchange_and_spmv():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    R = 2
    procs = [ctx.Process(target=_worker, args=(r, R, port, q)) for r in range(R)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(R)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, ok in res:
        assert ok and err < 1e-12, (rank, err)
