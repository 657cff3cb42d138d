// link.cuh -- the engine's own multi-GPU transport over peer memory (NVLink / NVSwitch; also two ranks on ONE device).
//
// Every rank owns a "window" (cudaMalloc'ed, exported with cudaIpcGetMemHandle, mapped by all peers).  A collective is
// ONE kernel per rank on the problem's stream that (1) stores its payload straight into the peers' windows, (2) raises
// the sequence number of the collective inside every 16-byte line, (3) polls the lines of its own window and (4) consumes what the
// peers stored.  No host involvement, no NCCL proxy / channel set-up per message: a halo exchange of one lattice row
// or a 2-scalar all-reduce costs a few microseconds instead of NCCL's 30-50 us, which is what bounded the weak
// scaling of round 1 (about 550 collectives per state+adjoint step).  Replaces what the reference would obtain from
// PETSc VecScatter / MPI_Allreduce (femo/fea/utils_dolfinx.py:167,236,354-358).
//
// Protocol invariants: all ranks issue the same sequence of collectives (SPMD); each KIND of collective (halo, scalar
// all-reduce, gather) has its own host-side counter and its own double-buffered region of the window, selected by the
// parity of that counter.  Every collective is symmetric between the ranks that share a buffer (neighbours for halos,
// all ranks for the other two): a rank passes collective m-1 of a kind only after its partners have STARTED m-1, i.e.
// finished consuming m-2 (stream order) -- so when it overwrites buffer m%2 nobody is still reading it.  Reductions
// add in rank order on every rank: bit-identical across ranks and run to run.  Polls give up after kLinkTimeoutNs and
// raise an error word instead of hanging the device.
#pragma once
#include "common.cuh"

namespace femo {

constexpr int kLinkMaxRanks = 16;
constexpr int kLinkMaxChunks = 64;       // CTAs per direction of a halo exchange
constexpr int kLinkArMax = 16;           // scalars per all-reduce
constexpr unsigned long long kLinkTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

// A "line" is 16 bytes: two 32-bit payload words, each next to the 32-bit sequence number of the collective
// (st.volatile.v4.u32; 8-byte halves arrive atomically, the scheme of NCCL's LL protocol).  The receiver polls the line
// until both flags carry the expected sequence number: data and "ready" travel in ONE store, so a message costs one
// NVLink traversal -- no __threadfence_system, no separate flag round trip.  A double occupies one line.
struct LinkLine {
    uint32_t d0, f0, d1, f1;
};

// window layout (byte offsets from the window base)
struct LinkLayout {
    size_t halo_cap = 0, gather_cap = 0;   // lines (= doubles) per halo slot / per gather buffer
    static constexpr size_t kErr = 0;                                   // int: set when a spin timed out
    static constexpr size_t kAr = 4096;                                 // LinkLine[2][kLinkMaxRanks][kLinkArMax]
    static constexpr size_t kPayload = kAr + 16 * 2 * kLinkMaxRanks * kLinkArMax;
    __host__ __device__ size_t halo_off(int parity, int slot) const { return kPayload + 16 * ((size_t)(parity * 2 + slot) * halo_cap); }
    __host__ __device__ size_t gather_off(int parity) const { return kPayload + 16 * (4 * halo_cap + (size_t)parity * gather_cap); }
    size_t bytes() const { return kPayload + 16 * (4 * halo_cap + 2 * gather_cap); }
};

struct LinkDev {           // kernel argument: peer-mapped window bases
    char *win[kLinkMaxRanks];
    int rank, nranks;
    LinkLayout lay;
};

__device__ __forceinline__ void ll_store(LinkLine *p, uint32_t w0, uint32_t w1, uint32_t seq) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w0), "r"(seq), "r"(w1), "r"(seq) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until the line carries `seq` in both flags; false after the timeout (the window's error word is raised)
__device__ __forceinline__ bool ll_load(const LinkDev &L, const LinkLine *p, uint32_t seq, uint32_t &w0, uint32_t &w1) {
    uint32_t f0, f1;
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(f0), "=r"(w1), "=r"(f1) : "l"(p) : "memory");
        if (f0 == seq && f1 == seq) return true;
        if ((spins & 1023u) == 1023u) {
            const unsigned long long t = global_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > kLinkTimeoutNs) {
                *reinterpret_cast<volatile int *>(L.win[L.rank] + LinkLayout::kErr) = 1;
                return false;
            }
        }
    }
}
__device__ __forceinline__ void ll_store_double(LinkLine *p, double v, uint32_t seq) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    ll_store(p, (uint32_t)u, (uint32_t)(u >> 32), seq);
}
__device__ __forceinline__ double ll_load_double(const LinkDev &L, const LinkLine *p, uint32_t seq) {
    uint32_t w0 = 0, w1 = 0;
    ll_load(L, p, seq, w0, w1);
    return __longlong_as_double((long long)(((unsigned long long)w1 << 32) | w0));
}

// Halo exchange with the lower (dir 0, rank-1) and upper (dir 1, rank+1) neighbour.  The payload is handled as 32-bit
// words (a double = 2 words = 1 line, two floats share a line).  CTA (dir, chunk): every thread stores its lines of the
// boundary row into the neighbour's slot (we are the neighbour's upper / lower side), then polls the same lines of our
// own slot and writes the ghost row.  Zero-length directions exchange nothing (a rank is never ahead by more than one
// collective of a neighbour it exchanges data with; see the reuse argument in the header).
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_link_halo(LinkDev L, T *v, size_t send_lo, size_t recv_lo, size_t send_hi, size_t recv_hi, size_t send_len_lo,
                size_t recv_len_lo, size_t send_len_hi, size_t recv_len_hi, int nchunks, unsigned int seq) {
    const int dir = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
    const int peer = dir == 0 ? L.rank - 1 : L.rank + 1;
    if (peer < 0 || peer >= L.nranks) return;
    const int parity = (int)(seq & 1u);
    constexpr size_t WPE = sizeof(T) / 4;            // 32-bit words per element
    const size_t swords = (dir == 0 ? send_len_lo : send_len_hi) * WPE, rwords = (dir == 0 ? recv_len_lo : recv_len_hi) * WPE;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(v + (dir == 0 ? send_lo : send_hi));
    uint32_t *dstv = reinterpret_cast<uint32_t *>(v + (dir == 0 ? recv_lo : recv_hi));
    {   // store: we are the peer's upper neighbour (slot 1) when it is our lower one, and vice versa
        LinkLine *dst = reinterpret_cast<LinkLine *>(L.win[peer] + L.lay.halo_off(parity, dir == 0 ? 1 : 0));
        const size_t nl = (swords + 1) / 2, per = (nl + nchunks - 1) / nchunks, a = (size_t)chunk * per, b = a + per < nl ? a + per : nl;
        for (size_t j = a + threadIdx.x; j < b; j += kThreads)
            ll_store(dst + j, src[2 * j], 2 * j + 1 < swords ? src[2 * j + 1] : 0u, seq);
        // handshake line (last line of the slot): keeps every exchange symmetric between neighbours even when one
        // direction carries no payload (cell rows travel upwards only), which the buffer-reuse argument relies on
        if (chunk == 0 && threadIdx.x == 0) ll_store(dst + (L.lay.halo_cap - 1), 0u, 0u, seq);
    }
    {   // poll our slot `dir` and unpack
        const LinkLine *in = reinterpret_cast<const LinkLine *>(L.win[L.rank] + L.lay.halo_off(parity, dir));
        const size_t nl = (rwords + 1) / 2, per = (nl + nchunks - 1) / nchunks, a = (size_t)chunk * per, b = a + per < nl ? a + per : nl;
        for (size_t j = a + threadIdx.x; j < b; j += kThreads) {
            uint32_t w0 = 0, w1 = 0;
            if (!ll_load(L, in + j, seq, w0, w1)) return;
            dstv[2 * j] = w0;
            if (2 * j + 1 < rwords) dstv[2 * j + 1] = w1;
        }
        if (chunk == 0 && threadIdx.x == 0) {
            uint32_t w0, w1;
            ll_load(L, in + (L.lay.halo_cap - 1), seq, w0, w1);
        }
    }
}

// All-reduce (sum or max) of `count` <= kLinkArMax doubles at scalars[0..count): every rank stores its values as lines
// into every window (its own included), then thread k polls the lines of all ranks for value k and reduces in rank
// order (identical bits on every rank).  With `pa` the block sums of up to two partial arrays are taken first (the
// Krylov kernels' finalize step and its all-reduce in ONE launch).
__global__ void __launch_bounds__(kThreads)
    k_link_allreduce(LinkDev L, double *scalars, int count, int is_max, const double *pa, const double *pb, int np,
                     int slotA, int slotB, unsigned int seq) {
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
            mine[0] = a;
            if (pb) mine[1] = b;
        }
        scalars += slotA;
    } else if (threadIdx.x < count) {
        mine[threadIdx.x] = scalars[threadIdx.x];
    }
    __syncthreads();
    const int parity = (int)(seq & 1u);
    for (int t = threadIdx.x; t < L.nranks * count; t += blockDim.x) {      // thread (r, k): value k into rank r's window
        const int r = t / count, k = t % count;
        LinkLine *dst = reinterpret_cast<LinkLine *>(L.win[r] + LinkLayout::kAr) + ((size_t)parity * kLinkMaxRanks + L.rank) * kLinkArMax;
        ll_store_double(dst + k, mine[k], seq);
    }
    if (threadIdx.x < count) {
        const LinkLine *src = reinterpret_cast<const LinkLine *>(L.win[L.rank] + LinkLayout::kAr) + (size_t)parity * kLinkMaxRanks * kLinkArMax;
        double acc = ll_load_double(L, src + threadIdx.x, seq);
        for (int r = 1; r < L.nranks; ++r) {
            const double v = ll_load_double(L, src + (size_t)r * kLinkArMax + threadIdx.x, seq);
            acc = is_max ? fmax(acc, v) : acc + v;
        }
        scalars[threadIdx.x] = acc;
    }
}

// All-gather of contiguous blocks into a replicated vector g[0..total): rank r contributes g[r*blk .. (r+1)*blk) (the
// last rank `tail` more).  Every rank stores its block as lines into all windows, then polls the lines of the whole
// vector in its own window and writes g (its own block included, so stale data outside it never survives).
__global__ void __launch_bounds__(kThreads)
    k_link_gather(LinkDev L, double *g, size_t blk, size_t tail, unsigned int seq) {
    const int parity = (int)(seq & 1u);
    const size_t myoff = (size_t)L.rank * blk, mylen = blk + (L.rank == L.nranks - 1 ? tail : 0);
    const size_t total = blk * L.nranks + tail;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < L.nranks; ++r) {
        LinkLine *dst = reinterpret_cast<LinkLine *>(L.win[r] + L.lay.gather_off(parity)) + myoff;
        for (size_t k = tid; k < mylen; k += nth) ll_store_double(dst + k, g[myoff + k], seq);
    }
    const LinkLine *src = reinterpret_cast<const LinkLine *>(L.win[L.rank] + L.lay.gather_off(parity));
    for (size_t k = tid; k < total; k += nth) g[k] = ll_load_double(L, src + k, seq);
}

// host-side state of the transport
struct Link {
    bool active = false;
    int rank = 0, nranks = 1, device = -1;
    char *local = nullptr;                 // this rank's window (cudaMalloc)
    char *win[kLinkMaxRanks] = {nullptr};  // mapped peers
    LinkLayout lay;
    unsigned int seq_halo = 0, seq_ar = 0, seq_gather = 0;   // per-kind collective counters (same on every rank); travel in every line
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
