// dist_ops.cuh -- halo exchange, gathers and scalar all-reduces on a problem's slab.
// All calls enqueue on the problem's stream; nothing synchronises the host.
#pragma once
#include "dist.cuh"

using namespace femo;

// Refresh the ghost node rows of a state-space vector (ghostUpdate FORWARD of the reference,
// utils_dolfinx.py:167): whole lattice rows are contiguous, so rows are sent in place.
static int halo_nodes(femo_problem *p, double *v) {
    const SlabInfo &s = p->slab;
    if (p->skip_next_halo) {
        p->skip_next_halo = false;
        return FEMO_OK;
    }
    if (!s.active || !g_comm.active) return FEMO_OK;
    const size_t len = (size_t)(p->mesh.n[0] + 1) * (p->mesh.kind == MESH_HEX ? (size_t)(p->mesh.n[1] + 1) : 1) * p->state.block;
    NcclApi &a = g_comm.api;
    FEMO_NCCL(a.GroupStart());
    if (s.rank > 0) {
        FEMO_NCCL(a.Send(v + (size_t)s.own0 * len, len, ncclDouble, s.rank - 1, g_comm.comm, p->stream));
        FEMO_NCCL(a.Recv(v + (size_t)(s.own0 - 1) * len, len, ncclDouble, s.rank - 1, g_comm.comm, p->stream));
    }
    if (s.rank < s.nranks - 1) {
        FEMO_NCCL(a.Send(v + (size_t)(s.own1 - 1) * len, len, ncclDouble, s.rank + 1, g_comm.comm, p->stream));
        FEMO_NCCL(a.Recv(v + (size_t)s.own1 * len, len, ncclDouble, s.rank + 1, g_comm.comm, p->stream));
    }
    FEMO_NCCL(a.GroupEnd());
    g_comm.halo_exchanges++;
    return FEMO_OK;
}

// Same for an fp32 node vector (the diagonal plane of a DIA level: ghost rows are only partially assembled locally)
static int halo_nodes_f32(femo_problem *p, float *v) {
    const SlabInfo &s = p->slab;
    if (!s.active || !g_comm.active) return FEMO_OK;
    const size_t len = (size_t)(p->mesh.n[0] + 1) * (p->mesh.kind == MESH_HEX ? (size_t)(p->mesh.n[1] + 1) : 1) * p->state.block;
    NcclApi &a = g_comm.api;
    FEMO_NCCL(a.GroupStart());
    if (s.rank > 0) {
        FEMO_NCCL(a.Send(v + (size_t)s.own0 * len, len, ncclFloat, s.rank - 1, g_comm.comm, p->stream));
        FEMO_NCCL(a.Recv(v + (size_t)(s.own0 - 1) * len, len, ncclFloat, s.rank - 1, g_comm.comm, p->stream));
    }
    if (s.rank < s.nranks - 1) {
        FEMO_NCCL(a.Send(v + (size_t)(s.own1 - 1) * len, len, ncclFloat, s.rank + 1, g_comm.comm, p->stream));
        FEMO_NCCL(a.Recv(v + (size_t)s.own1 * len, len, ncclFloat, s.rank + 1, g_comm.comm, p->stream));
    }
    FEMO_NCCL(a.GroupEnd());
    g_comm.halo_exchanges++;
    return FEMO_OK;
}

// Refresh the ghost cell row (below) of a cell-wise (DG0) vector.
static int halo_cells(femo_problem *p, double *v) {
    const SlabInfo &s = p->slab;
    if (!s.active || !g_comm.active) return FEMO_OK;
    const size_t len = (size_t)(p->mesh.ncells / s.ncrows) * p->in[0].block;   // cells per lattice row
    NcclApi &a = g_comm.api;
    FEMO_NCCL(a.GroupStart());
    if (s.rank < s.nranks - 1)
        FEMO_NCCL(a.Send(v + (size_t)(s.cown1 - 1) * len, len, ncclDouble, s.rank + 1, g_comm.comm, p->stream));
    if (s.rank > 0)
        FEMO_NCCL(a.Recv(v + (size_t)(s.cown0 - 1) * len, len, ncclDouble, s.rank - 1, g_comm.comm, p->stream));
    FEMO_NCCL(a.GroupEnd());
    g_comm.halo_exchanges++;
    return FEMO_OK;
}

// Sum (or max) `count` device scalars starting at `slot` over all ranks, in place.
static int allreduce_scalars(femo_problem *p, int slot, int count, bool is_max = false) {
    if (!g_comm.active || (!p->slab.active)) return FEMO_OK;
    FEMO_NCCL(g_comm.api.AllReduce(p->d_scalars + slot, p->d_scalars + slot, (size_t)count, ncclDouble,
                                   is_max ? ncclMax : ncclSum, g_comm.comm, p->stream));
    g_comm.allreduces++;
    return FEMO_OK;
}

// Assemble a replicated global node vector from the rows each rank owns: `g` holds the global
// lattice (gny+1 rows of len doubles); rank r has filled rows [r*rows, (r+1)*rows) (the last rank
// also row gny).  In-place all-gather of the equal blocks, then the top row from the last rank.
static int gather_rows(femo_problem *p, double *g, size_t len, int gny) {
    if (!g_comm.active) return FEMO_OK;
    const int R = g_comm.nranks, rows = gny / R;
    const size_t blk = (size_t)rows * len;
    FEMO_NCCL(g_comm.api.AllGather(g + (size_t)g_comm.rank * blk, g, blk, ncclDouble, g_comm.comm, p->stream));
    FEMO_NCCL(g_comm.api.Broadcast(g + (size_t)gny * len, g + (size_t)gny * len, len, ncclDouble, R - 1, g_comm.comm,
                                   p->stream));
    return FEMO_OK;
}

// Same for a replicated cell-wise vector (no extra top row): rank r has filled cell rows [r*rows, (r+1)*rows).
static int gather_cell_rows(femo_problem *p, double *g, size_t len, int gn) {
    if (!g_comm.active) return FEMO_OK;
    const size_t blk = (size_t)(gn / g_comm.nranks) * len;
    FEMO_NCCL(g_comm.api.AllGather(g + (size_t)g_comm.rank * blk, g, blk, ncclDouble, g_comm.comm, p->stream));
    return FEMO_OK;
}
