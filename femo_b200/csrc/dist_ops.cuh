// dist_ops.cuh -- halo exchange, gathers and scalar all-reduces on a problem's slab.
// All calls enqueue on the problem's stream; nothing synchronises the host.
#pragma once
#include "dist.cuh"
#include "link.cuh"

using namespace femo;

// ---- peer-memory transport (link.cuh) ------------------------------------------------------------------------
static inline int link_chunks(size_t len) {
    const size_t c = (len + 2047) / 2048;
    return (int)std::max<size_t>(1, std::min<size_t>(c, kLinkMaxChunks));
}

// symmetric exchange of slab rows of T: send [send_lo, +len_lo_s) down / [send_hi, +len_hi_s) up, receive into
// [recv_lo, ...) from below / [recv_hi, ...) from above (lengths in elements of T; zero-length directions still flag)
template <class T>
static int link_halo(femo_problem *p, T *v, size_t send_lo, size_t recv_lo, size_t send_hi, size_t recv_hi, size_t slen_lo,
                     size_t rlen_lo, size_t slen_hi, size_t rlen_hi, cudaStream_t st = nullptr) {
    if (!st) st = p->stream;
    const size_t mx = std::max(std::max(slen_lo, rlen_lo), std::max(slen_hi, rlen_hi));
    if ((mx * sizeof(T) + 7) / 8 + 1 > g_link.lay.halo_cap)
        return set_err(FEMO_ELIMIT, "halo row exceeds the link window (femo_link_create halo capacity)");
    const int nch = link_chunks(mx);
    const unsigned int seq = ++g_link.seq_halo;
    k_link_halo<T><<<2 * nch, kThreads, 0, st>>>(g_link.dev(), v, send_lo, recv_lo, send_hi, recv_hi, slen_lo, rlen_lo,
                                                         slen_hi, rlen_hi, nch, seq);
    p->launches++;
    g_comm.halo_exchanges++;
    FEMO_CHECK_LAUNCH();
    return FEMO_OK;
}

static int link_allreduce(femo_problem *p, double *scalars, int count, bool is_max) {
    if (count > kLinkArMax) return set_err(FEMO_ELIMIT, "all-reduce of more than 16 scalars");
    const unsigned int seq = ++g_link.seq_ar;
    k_link_allreduce<<<1, kThreads, 0, p->stream>>>(g_link.dev(), scalars, count, is_max ? 1 : 0, nullptr, nullptr, 0, 0, 0, seq);
    p->launches++;
    g_comm.allreduces++;
    FEMO_CHECK_LAUNCH();
    return FEMO_OK;
}

static int link_gather(femo_problem *p, double *g, size_t blk, size_t tail) {
    const size_t total = blk * g_link.nranks + tail;
    if (total > g_link.lay.gather_cap) return set_err(FEMO_ELIMIT, "gathered level exceeds the link window (femo_link_create gather capacity)");
    const unsigned int seq = ++g_link.seq_gather;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((blk + tail + kThreads - 1) / kThreads, 296));
    k_link_gather<<<grid, kThreads, 0, p->stream>>>(g_link.dev(), g, blk, tail, seq);
    p->launches++;
    FEMO_CHECK_LAUNCH();
    return FEMO_OK;
}

namespace femo {
__global__ void __launch_bounds__(kThreads)
    k_gpart_pack(const double *__restrict__ v, const int32_t *__restrict__ idx, int64_t n, double *__restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = v[idx[t]];
}
__global__ void __launch_bounds__(kThreads)
    k_gpart_unpack(const double *__restrict__ gathered, const int32_t *__restrict__ src, int64_t n, double *__restrict__ ghosts) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) ghosts[t] = gathered[src[t]];
}
}  // namespace femo

static inline bool partitioned(const femo_problem *p) { return p->slab.active || p->gpart.active; }
static int gather_cell_rows(femo_problem *p, double *g, size_t len, int gn);

// ghost refresh of an unstructured partition: pack -> all-gather of the ranks' send buffers -> scatter into the ghost range
static int gpart_halo(femo_problem *p, double *v) {
    GenPart &G = p->gpart;
    if (!g_comm.active) return FEMO_OK;
    if (!G.d_gather) return set_err(FEMO_ESTATE, "partitioned problem: upload before exchanging halos");
    const int64_t ns = (int64_t)G.send_idx.size(), ng = (int64_t)G.ghost_src.size();
    if (ns) k_gpart_pack<<<(int)((ns + kThreads - 1) / kThreads), kThreads, 0, p->stream>>>(v, G.d_send_idx, ns, G.d_gather + (size_t)g_comm.rank * G.blk);
    int rc = gather_cell_rows(p, G.d_gather, (size_t)G.blk, g_comm.nranks);
    if (rc) return rc;
    if (ng) k_gpart_unpack<<<(int)((ng + kThreads - 1) / kThreads), kThreads, 0, p->stream>>>(G.d_gather, G.d_ghost_src, ng, v + G.n_owned_dofs);
    p->launches += (ns ? 1 : 0) + (ng ? 1 : 0);
    g_comm.halo_exchanges++;
    FEMO_CHECK_LAUNCH();
    return FEMO_OK;
}

// Refresh the ghost node rows of a state-space vector (ghostUpdate FORWARD of the reference,
// utils_dolfinx.py:167): whole lattice rows are contiguous, so rows are sent in place.
static int halo_nodes(femo_problem *p, double *v, cudaStream_t st = nullptr) {
    const SlabInfo &s = p->slab;
    if (p->skip_next_halo) {
        p->skip_next_halo = false;
        return FEMO_OK;
    }
    if (p->gpart.active) return gpart_halo(p, v);
    if (!s.active || !g_comm.active) return FEMO_OK;
    if (!st) st = p->stream;
    const size_t len = (size_t)(p->mesh.n[0] + 1) * (p->mesh.kind == MESH_HEX ? (size_t)(p->mesh.n[1] + 1) : 1) * p->state.block;
    if (g_link.active)
        return link_halo<double>(p, v, (size_t)s.own0 * len, (size_t)(s.own0 - 1) * len, (size_t)(s.own1 - 1) * len,
                                 (size_t)s.own1 * len, len, len, len, len, st);
    NcclApi &a = g_comm.api;
    FEMO_NCCL(a.GroupStart());
    if (s.rank > 0) {
        FEMO_NCCL(a.Send(v + (size_t)s.own0 * len, len, ncclDouble, s.rank - 1, g_comm.comm, st));
        FEMO_NCCL(a.Recv(v + (size_t)(s.own0 - 1) * len, len, ncclDouble, s.rank - 1, g_comm.comm, st));
    }
    if (s.rank < s.nranks - 1) {
        FEMO_NCCL(a.Send(v + (size_t)(s.own1 - 1) * len, len, ncclDouble, s.rank + 1, g_comm.comm, st));
        FEMO_NCCL(a.Recv(v + (size_t)s.own1 * len, len, ncclDouble, s.rank + 1, g_comm.comm, st));
    }
    FEMO_NCCL(a.GroupEnd());
    g_comm.halo_exchanges++;
    return FEMO_OK;
}

// Overlap of a slab level's halo exchange with its interior rows (K11 of SURVEY.md section 8a): rows of the first and
// last OWNED lattice row read ghost values, all other owned rows do not.  `overlap_begin` forks the side stream behind
// everything enqueued so far and runs the exchange there; the caller then launches the boundary rows on the side stream
// and the interior rows on the main stream, and `overlap_end` joins.  Returns false when the level is not split
// (one GPU, small level, ghost rows provably fresh) -- the caller then takes the plain path.
static int64_t overlap_min_rows() { return g_env.overlap_min_rows; }   // owned dofs below which a level is not split (FEMO_OVERLAP_MIN_ROWS)
struct RowSplit {
    int b0 = 0, nb0 = 0, b1 = 0, nb1 = 0;   // boundary segments (first / last owned lattice row)
    int i0 = 0, ni = 0;                     // interior rows
};
static bool overlap_begin(femo_problem *p, double *v, RowSplit &R, int *rc) {
    *rc = FEMO_OK;
    const SlabInfo &s = p->slab;
    if (!s.active || !g_comm.active || p->skip_next_halo || !p->stream2 || p->mesh.kind == MESH_HEX || g_env.no_overlap) return false;
    const int64_t len = (int64_t)(p->mesh.n[0] + 1) * p->state.block;
    const int64_t rows = s.own1 - s.own0;
    if (rows < 4 || rows * len < overlap_min_rows()) return false;
    R.b0 = (int)(s.own0 * len); R.nb0 = (int)len;
    R.b1 = (int)((s.own1 - 1) * len); R.nb1 = (int)len;
    R.i0 = (int)((s.own0 + 1) * len); R.ni = (int)((rows - 2) * len);
    cudaError_t e = cudaEventRecord(p->ev_fork, p->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->stream2, p->ev_fork, 0);
    if (e != cudaSuccess) { *rc = set_err(FEMO_ECUDA, cudaGetErrorString(e)); return false; }
    *rc = halo_nodes(p, v, p->stream2);
    return *rc == FEMO_OK;
}
static int overlap_end(femo_problem *p) {
    FEMO_CUDA(cudaEventRecord(p->ev_join, p->stream2));
    FEMO_CUDA(cudaStreamWaitEvent(p->stream, p->ev_join, 0));
    return FEMO_OK;
}

// Same for an fp32 node vector (the diagonal plane of a DIA level: ghost rows are only partially assembled locally)
static int halo_nodes_f32(femo_problem *p, float *v) {
    const SlabInfo &s = p->slab;
    if (!s.active || !g_comm.active) return FEMO_OK;
    const size_t len = (size_t)(p->mesh.n[0] + 1) * (p->mesh.kind == MESH_HEX ? (size_t)(p->mesh.n[1] + 1) : 1) * p->state.block;
    if (g_link.active)
        return link_halo<float>(p, v, (size_t)s.own0 * len, (size_t)(s.own0 - 1) * len, (size_t)(s.own1 - 1) * len,
                                (size_t)s.own1 * len, len, len, len, len);
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
    if (g_link.active)     // cell rows travel upwards only; the downward direction exchanges flags (symmetry)
        return link_halo<double>(p, v, 0, (size_t)(s.cown0 - 1) * len, (size_t)(s.cown1 - 1) * len, 0, 0, len, len, 0);
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
    if (!g_comm.active || !partitioned(p)) return FEMO_OK;
    if (g_link.active) {
        // the link all-reduce carries up to kLinkArMax scalars per collective: GMRES' batched dot products (up to restart + 1)
        // go in chunks
        for (int o = 0; o < count; o += kLinkArMax) {
            int rc = link_allreduce(p, p->d_scalars + slot + o, std::min(kLinkArMax, count - o), is_max);
            if (rc) return rc;
        }
        return FEMO_OK;
    }
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
    if (g_link.active) return link_gather(p, g, blk, len);       // the top row rides with the last rank's block
    FEMO_NCCL(g_comm.api.AllGather(g + (size_t)g_comm.rank * blk, g, blk, ncclDouble, g_comm.comm, p->stream));
    FEMO_NCCL(g_comm.api.Broadcast(g + (size_t)gny * len, g + (size_t)gny * len, len, ncclDouble, R - 1, g_comm.comm,
                                   p->stream));
    return FEMO_OK;
}

// Same for a replicated cell-wise vector (no extra top row): rank r has filled cell rows [r*rows, (r+1)*rows).
static int gather_cell_rows(femo_problem *p, double *g, size_t len, int gn) {
    if (!g_comm.active) return FEMO_OK;
    const size_t blk = (size_t)(gn / g_comm.nranks) * len;
    if (g_link.active) return link_gather(p, g, blk, 0);
    FEMO_NCCL(g_comm.api.AllGather(g + (size_t)g_comm.rank * blk, g, blk, ncclDouble, g_comm.comm, p->stream));
    return FEMO_OK;
}
