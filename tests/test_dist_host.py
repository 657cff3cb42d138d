"""Host-side logic of the multi-GPU path without GPUs: slab partition, ghost maps
and owned ranges of the engine vs the oracle's global numbering (integer ==),
plus a world_size-2 gloo run that exchanges halos according to those maps and
checks SpMV / dot against the unpartitioned result."""
import os
import socket
import sys

import numpy as np
import pytest

from femo_b200.dist import SlabProblem
from femo_b200 import engine as E
from oracle import mesh as om, families as fam, assembly as asm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('R', [2, 3, 4])
def test_slab_partition_covers_global_numbering(R):
    nx, gny = 6, 4 * R
    m = om.unit_square_tri(nx, gny)
    F = fam.NonlinearPoissonP1(m)
    rp, col = asm.pattern(F.jacobian(np.zeros(F.N), np.zeros(F.M)), (F.N, F.N))
    owner_nodes = np.full(F.N, -1)
    owner_cells = np.full(F.M, -1)
    for r in range(R):
        p = SlabProblem(2, nx, gny, r, R)
        s = p.slab
        # coordinates and connectivity are the global ones restricted to the slab (bit for bit)
        g0 = s['crow0'] * (nx + 1)
        assert np.array_equal(p.local_coords(), m.coords[g0:g0 + p.N])
        c0 = s['crow0'] * 2 * nx
        assert np.array_equal(p.local_cells() + g0, m.cells[c0:c0 + p.M[0]])
        gl = g0 + np.arange(p.N)                              # local -> global dof
        own = gl[s['own_off']:s['own_off'] + s['own_n']]
        assert np.all(owner_nodes[own] == -1)
        owner_nodes[own] = r
        cown = c0 + np.arange(s['cown_off'], s['cown_off'] + s['cown_n'])
        assert np.all(owner_cells[cown] == -1)
        owner_cells[cown] = r
        # owned rows of the local pattern are the global rows (columns shifted by the slab offset)
        lrp, lcol = p.pattern(0)
        for i in range(s['own_off'], s['own_off'] + s['own_n']):
            gi = g0 + i
            assert np.array_equal(lcol[lrp[i]:lrp[i + 1]] + g0, col[rp[gi]:rp[gi + 1]])
        # ghost rows: one below (except rank 0), one above (except the last rank)
        assert s['own0'] == (0 if r == 0 else 1)
        assert (s['ncrows'] + 1) - s['own1'] == (0 if r == R - 1 else 1)
        # exterior facets only on true boundaries
        nb = 2 * s['ncrows'] + (nx if r == 0 else 0) + (nx if r == R - 1 else 0)
        assert p.pattern_info(0)['ncontrib'] == 9 * p.M[0] + 9 * nb
    assert np.all(owner_nodes >= 0) and np.all(owner_cells >= 0)          # every dof / cell owned exactly once


def test_partitioned_multigrid_needs_powers_of_two(monkeypatch):
    from femo_b200._lib import FemoError
    monkeypatch.setenv('FEMO_DIST_MIN_ROWS', '16')
    p = SlabProblem(2, 64, 128, 0, 2)
    assert p.enable_multigrid() >= 3
    assert SlabProblem(2, 60, 100, 0, 2).enable_multigrid() >= 2   # 50 rows per rank: replicated from the 30 x 50 level on
    q = SlabProblem(2, 60, 50, 0, 2)                          # 25 rows per rank: no 2:1 nested coarse level at all
    with pytest.raises(FemoError):
        q.enable_multigrid()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, R, port, nx, gny, q, hex_ny=None):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=R)
    try:
        rng = np.random.default_rng(3)
        if hex_ny is None:
            m = om.unit_square_tri(nx, gny)
            F = fam.NonlinearPoissonP1(m)
            u = rng.standard_normal(F.N)
            A = asm.assemble_matrix(F.jacobian(u, np.zeros(F.M)), (F.N, F.N)).tocsr()
            p = SlabProblem(2, nx, gny, rank, R)
            rowlen = nx + 1
        else:                                   # hexahedral z-slabs: a "row" is a whole plane of nodes x 3 components
            lo, hi = (0.0, 0.0, 0.0), (float(nx), float(hex_ny), float(gny))
            m = om.box_hex(lo, hi, nx, hex_ny, gny)
            fl = m.exterior_facets()[1]
            F = fam.SimpHex8(m, np.nonzero(fl == 3)[0])
            A = asm.assemble_matrix(F.jacobian(np.zeros(F.N), 0.2 + rng.random(F.M)), (F.N, F.N)).tocsr()
            p = SlabProblem(E.FAMILY_SIMP_HEX8, nx, gny, rank, R, lo=lo, hi=hi, ny=hex_ny, face_mask=8)
            rowlen = 3 * (nx + 1) * (hex_ny + 1)
        x = rng.standard_normal(F.N)
        s = p.slab
        g0 = s['crow0'] * rowlen
        # local matrix = the slab's own assembly (oracle on the local sub-mesh would need the same ghost rules);
        # here: global rows of the owned dofs, columns restricted to the local window
        Aloc = A[g0:g0 + p.N, g0:g0 + p.N]
        xl = x[g0:g0 + p.N].copy()
        # poison ghosts, then exchange exactly the rows femo_halo_exchange would
        if s['own0'] > 0:
            xl[:rowlen] = np.nan
        if rank < R - 1:
            xl[s['own1'] * rowlen:] = np.nan
        t = torch.from_numpy(xl)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(t[s['own0'] * rowlen:(s['own0'] + 1) * rowlen].clone(), rank - 1))
            reqs.append(dist.irecv(t[(s['own0'] - 1) * rowlen:s['own0'] * rowlen], rank - 1))
        if rank < R - 1:
            reqs.append(dist.isend(t[(s['own1'] - 1) * rowlen:s['own1'] * rowlen].clone(), rank + 1))
            reqs.append(dist.irecv(t[s['own1'] * rowlen:(s['own1'] + 1) * rowlen], rank + 1))
        for r_ in reqs:
            r_.wait()
        y = Aloc @ xl
        o = slice(s['own_off'], s['own_off'] + s['own_n'])
        yg = (A @ x)[g0:g0 + p.N]
        ok = bool(np.allclose(y[o], yg[o], rtol=1e-13, atol=1e-13))
        d = torch.tensor([float(xl[o] @ y[o])], dtype=torch.float64)
        dist.all_reduce(d)
        ok = ok and abs(d.item() - float(x @ (A @ x))) < 1e-10 * abs(float(x @ (A @ x)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('hex_ny', [None, 3])
def test_gloo_world2_halo_and_allreduce(hex_ny):
    import torch.multiprocessing as mp
    R, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    args = (8, 16, q) if hex_ny is None else (4, 6, q, hex_ny)
    procs = [ctx.Process(target=_gloo_worker, args=(r, R, port) + args) for r in range(R)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in range(R)]
    for p_ in procs:
        p_.join(timeout=60)
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize('R', [2, 4])
def test_hex_slab_partition_covers_global_numbering(R):
    """z-slabs of the hexahedral box: local coordinates / connectivity / owned pattern rows are the global ones
    restricted to the slab (integer ==), every dof, cell and traction facet is owned exactly once."""
    nx, ny, gnz = 4, 3, 2 * R
    lo, hi = (0.0, 0.0, 0.0), (4.0, 3.0, float(gnz))
    m = om.box_hex(lo, hi, nx, ny, gnz)
    fc, fl = m.exterior_facets()
    tag = np.nonzero(fl == 3)[0]
    F = fam.SimpHex8(m, tag)
    rp, col = asm.pattern(F.jacobian(np.zeros(F.N), np.ones(F.M)), (F.N, F.N))
    owner_nodes, owner_cells = np.full(F.N, -1), np.full(F.M, -1)
    pl, cl = (nx + 1) * (ny + 1), nx * ny
    nfacets_owned = 0
    for r in range(R):
        p = SlabProblem(E.FAMILY_SIMP_HEX8, nx, gnz, r, R, lo=lo, hi=hi, ny=ny, face_mask=1 << 3)
        s = p.slab
        g0, c0 = s['crow0'] * pl, s['crow0'] * cl
        assert np.array_equal(p.local_coords(), m.coords[g0:g0 + p.N // 3])
        assert np.array_equal(p.local_cells() + g0, m.cells[c0:c0 + p.M[0]])
        gl = 3 * g0 + np.arange(p.N)
        own = gl[s['own_off']:s['own_off'] + s['own_n']]
        assert np.all(owner_nodes[own] == -1)
        owner_nodes[own] = r
        cown = c0 + np.arange(s['cown_off'], s['cown_off'] + s['cown_n'])
        assert np.all(owner_cells[cown] == -1)
        owner_cells[cown] = r
        lrp, lcol = p.pattern(0)
        for i in range(s['own_off'], s['own_off'] + s['own_n'], 7):
            gi = 3 * g0 + i
            assert np.array_equal(lcol[lrp[i]:lrp[i + 1]] + 3 * g0, col[rp[gi]:rp[gi + 1]])
        nfacets_owned += ny * (s['cown1'] - s['cown0'])
    assert np.all(owner_nodes >= 0) and np.all(owner_cells >= 0)
    assert nfacets_owned == len(tag)


def test_hex_partitioned_multigrid_levels(monkeypatch):
    monkeypatch.setenv('FEMO_DIST_MIN_ROWS', '16')
    p = SlabProblem(E.FAMILY_SIMP_HEX8, 16, 32, 0, 2, lo=(0, 0, 0), hi=(16., 8., 32.), ny=8, face_mask=8)
    assert p.enable_multigrid() >= 4
