// link.cuh -- the engine's own multi-GPU transport over peer memory (NVLink / NVSwitch; also two ranks on ONE device).
//
// Every rank owns a "window" (cudaMalloc'ed, exported with cudaIpcGetMemHandle, mapped by all peers).  A collective is
// ONE kernel per rank on the problem's stream that (1) stores its payload straight into the peers' windows, (2) raises
// a sequence-numbered flag there (st.release.sys), (3) spins on its own flags (ld.acquire.sys) and (4) consumes what the
// peers stored.  No host involvement, no NCCL proxy / channel set-up per message: a halo exchange of one lattice row
// or a 2-scalar all-reduce costs a few microseconds instead of NCCL's 30-50 us, which is what bounded the weak
// scaling of round 1 (about 550 collectives per state+adjoint step).  Replaces what the reference would obtain from
// PETSc VecScatter / MPI_Allreduce (femo/fea/utils_dolfinx.py:167,236,354-358).
//
// Protocol invariants: all ranks issue the same sequence of collectives (SPMD), numbered by a host-side counter; every
// collective is symmetric between communicating pairs (both sides store and both wait), payload buffers are double
// buffered by the parity of the sequence number -- so a rank can only be overwriting buffer k%2 after its peer has
// started collective k-1, i.e. finished consuming collective k-2.  Reductions add in rank order on every rank:
// results are bit-identical across ranks and run to run.  Spins give up after kLinkTimeoutNs and raise an error word.
#pragma once
#include "common.cuh"

namespace femo {

constexpr int kLinkMaxRanks = 16;
constexpr int kLinkMaxChunks = 64;       // CTAs per direction of a halo exchange
constexpr int kLinkArMax = 16;           // scalars per all-reduce
constexpr unsigned long long kLinkTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

// window layout (byte offsets from the window base); flags are uint64 sequence numbers
struct LinkLayout {
    size_t halo_cap = 0, gather_cap = 0;   // doubles
    static constexpr size_t kErr = 0;                                   // int: set when a spin timed out
    static constexpr size_t kCounters = 64;                             // uint32[2]: intra-kernel CTA counters (gather)
    static constexpr size_t kHaloFlags = 256;                           // uint64[2][kLinkMaxChunks]
    static constexpr size_t kArFlags = kHaloFlags + 8 * 2 * kLinkMaxChunks;   // uint64[kLinkMaxRanks]
    static constexpr size_t kGatherFlags = kArFlags + 8 * kLinkMaxRanks;      // uint64[kLinkMaxRanks]
    static constexpr size_t kArVals = kGatherFlags + 8 * kLinkMaxRanks;       // double[2][kLinkMaxRanks][kLinkArMax]
    static constexpr size_t kPayload = (kArVals + 8 * 2 * kLinkMaxRanks * kLinkArMax + 255) & ~size_t(255);
    __host__ __device__ size_t halo_off(int parity, int slot) const { return kPayload + 8 * ((size_t)(parity * 2 + slot) * halo_cap); }
    __host__ __device__ size_t gather_off(int parity) const { return kPayload + 8 * (4 * halo_cap + (size_t)parity * gather_cap); }
    size_t bytes() const { return kPayload + 8 * (4 * halo_cap + 2 * gather_cap); }
};

struct LinkDev {           // kernel argument: peer-mapped window bases
    char *win[kLinkMaxRanks];
    int rank, nranks;
    LinkLayout lay;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// thread 0 of the CTA spins until *flag >= seq (or the timeout raises the window's error word); all threads leave together
__device__ __forceinline__ void link_wait(const LinkDev &L, const unsigned long long *flag, unsigned long long seq) {
    if (threadIdx.x == 0) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) < seq) {
            if (global_ns() - t0 > kLinkTimeoutNs) {
                *reinterpret_cast<volatile int *>(L.win[L.rank] + LinkLayout::kErr) = 1;
                break;
            }
        }
    }
    __syncthreads();
}

// Halo exchange with the lower (dir 0, rank-1) and upper (dir 1, rank+1) neighbour.  CTA (dir, chunk) stores its chunk of
// the boundary row into the neighbour's slot (we are the neighbour's upper / lower side), flags it, then waits for the
// same chunk from that neighbour and copies it into the ghost row.  A direction with send_len == 0 / recv_len == 0
// still exchanges flags (the symmetry the buffer-reuse argument needs).
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_link_halo(LinkDev L, T *v, size_t send_lo, size_t recv_lo, size_t send_hi, size_t recv_hi, size_t send_len_lo,
                size_t recv_len_lo, size_t send_len_hi, size_t recv_len_hi, int nchunks, unsigned long long seq) {
    const int dir = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
    const int peer = dir == 0 ? L.rank - 1 : L.rank + 1;
    if (peer < 0 || peer >= L.nranks) return;
    const int parity = (int)(seq & 1);
    const size_t slen = dir == 0 ? send_len_lo : send_len_hi, rlen = dir == 0 ? recv_len_lo : recv_len_hi;
    const size_t soff = dir == 0 ? send_lo : send_hi, roff = dir == 0 ? recv_lo : recv_hi;
    // store: we are the peer's upper neighbour (slot 1) when it is our lower one, and vice versa
    {
        T *dst = reinterpret_cast<T *>(L.win[peer] + L.lay.halo_off(parity, dir == 0 ? 1 : 0));
        const size_t per = (slen + nchunks - 1) / nchunks, a = (size_t)chunk * per, b = a + per < slen ? a + per : slen;
        for (size_t k = a + threadIdx.x; k < b; k += kThreads) dst[k] = v[soff + k];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0)
            st_release_sys(reinterpret_cast<unsigned long long *>(L.win[peer] + LinkLayout::kHaloFlags) +
                               (dir == 0 ? 1 : 0) * kLinkMaxChunks + chunk, seq);
    }
    // wait for the peer's chunk in our slot `dir`, then unpack
    link_wait(L, reinterpret_cast<const unsigned long long *>(L.win[L.rank] + LinkLayout::kHaloFlags) + dir * kLinkMaxChunks + chunk, seq);
    {
        const T *src = reinterpret_cast<const T *>(L.win[L.rank] + L.lay.halo_off(parity, dir));
        const size_t per = (rlen + nchunks - 1) / nchunks, a = (size_t)chunk * per, b = a + per < rlen ? a + per : rlen;
        for (size_t k = a + threadIdx.x; k < b; k += kThreads) v[roff + k] = __ldcg(src + k);
    }
}

// All-reduce (sum or max) of `count` <= kLinkArMax doubles at scalars[0..count): every rank stores its values into
// every window (its own included), flags, waits for all ranks, reduces in rank order.  With `partials` the block sums
// of up to two partial arrays are taken first (the Krylov kernels' finalize step and its all-reduce in ONE launch).
__global__ void __launch_bounds__(kThreads)
    k_link_allreduce(LinkDev L, double *scalars, int count, int is_max, const double *pa, const double *pb, int np,
                     int slotA, int slotB, unsigned long long seq) {
    __shared__ double mine[kLinkArMax];
    if (pa) {               // fused finalize: fixed-order block sums of the partial arrays
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            a += pa[i];
            if (pb) b += pb[i];
        }
        a = block_sum(a);
        b = block_sum(b);
        if (threadIdx.x == 0) {
            scalars[slotA] = a;
            if (pb) scalars[slotB] = b;
        }
        __syncthreads();
        scalars += slotA;
    }
    const int parity = (int)(seq & 1);
    if (threadIdx.x < count) mine[threadIdx.x] = scalars[threadIdx.x];
    __syncthreads();
    // thread (r, k): value k into rank r's window
    for (int t = threadIdx.x; t < L.nranks * count; t += blockDim.x) {
        const int r = t / count, k = t % count;
        double *dst = reinterpret_cast<double *>(L.win[r] + LinkLayout::kArVals) + ((size_t)parity * kLinkMaxRanks + L.rank) * kLinkArMax;
        dst[k] = mine[k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < L.nranks)
        st_release_sys(reinterpret_cast<unsigned long long *>(L.win[threadIdx.x] + LinkLayout::kArFlags) + L.rank, seq);
    // wait for every rank's contribution
    if (threadIdx.x < L.nranks) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(L.win[L.rank] + LinkLayout::kArFlags) + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) < seq) {
            if (global_ns() - t0 > kLinkTimeoutNs) {
                *reinterpret_cast<volatile int *>(L.win[L.rank] + LinkLayout::kErr) = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < count) {
        const double *src = reinterpret_cast<const double *>(L.win[L.rank] + LinkLayout::kArVals) + (size_t)parity * kLinkMaxRanks * kLinkArMax;
        double acc = __ldcg(src + threadIdx.x);
        for (int r = 1; r < L.nranks; ++r) {
            const double v = __ldcg(src + (size_t)r * kLinkArMax + threadIdx.x);
            acc = is_max ? fmax(acc, v) : acc + v;
        }
        scalars[threadIdx.x] = acc;
    }
}

// All-gather of contiguous blocks into a replicated vector g[0..total): rank r contributes g[off_r .. off_r + len_r)
// (blocks are the equal shares off_r = r*blk, len_r = blk, the last rank extended by `tail`).  Every rank stores its
// block into all windows; the last CTA to finish storing raises the flags; after all ranks arrived the window is
// copied into g (own block included: g may hold stale data outside the own block).
__global__ void __launch_bounds__(kThreads)
    k_link_gather(LinkDev L, double *g, size_t blk, size_t tail, unsigned long long seq) {
    const int parity = (int)(seq & 1);
    const size_t myoff = (size_t)L.rank * blk, mylen = blk + (L.rank == L.nranks - 1 ? tail : 0);
    const size_t total = blk * L.nranks + tail;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < L.nranks; ++r) {
        double *dst = reinterpret_cast<double *>(L.win[r] + L.lay.gather_off(parity)) + myoff;
        for (size_t k = tid; k < mylen; k += nth) dst[k] = g[myoff + k];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        unsigned int *ctr = reinterpret_cast<unsigned int *>(L.win[L.rank] + LinkLayout::kCounters) + parity;
        const unsigned int done = atomicAdd(ctr, 1u);
        last = done == gridDim.x - 1;
        if (last) {
            *ctr = 0;      // ready for the next use of this parity (two collectives later)
            __threadfence_system();
            for (int r = 0; r < L.nranks; ++r)
                st_release_sys(reinterpret_cast<unsigned long long *>(L.win[r] + LinkLayout::kGatherFlags) + L.rank, seq);
        }
    }
    __syncthreads();
    if (threadIdx.x < L.nranks) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(L.win[L.rank] + LinkLayout::kGatherFlags) + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) < seq) {
            if (global_ns() - t0 > kLinkTimeoutNs) {
                *reinterpret_cast<volatile int *>(L.win[L.rank] + LinkLayout::kErr) = 1;
                break;
            }
        }
    }
    __syncthreads();
    const double *src = reinterpret_cast<const double *>(L.win[L.rank] + L.lay.gather_off(parity));
    for (size_t k = tid; k < total; k += nth) g[k] = __ldcg(src + k);
}

// host-side state of the transport
struct Link {
    bool active = false;
    int rank = 0, nranks = 1, device = -1;
    char *local = nullptr;                 // this rank's window (cudaMalloc)
    char *win[kLinkMaxRanks] = {nullptr};  // mapped peers
    LinkLayout lay;
    unsigned long long seq = 0;            // collectives issued so far (same on every rank)
    LinkDev dev() const {
        LinkDev d;
        for (int r = 0; r < kLinkMaxRanks; ++r) d.win[r] = win[r];
        d.rank = rank;
        d.nranks = nranks;
        d.lay = lay;
        return d;
    }
};

static Link g_link;

}  // namespace femo
