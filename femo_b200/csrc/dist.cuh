// dist.cuh -- multi-GPU plumbing: one process per GPU, y-slab partition of lattice
// meshes with a one-cell ghost layer, NCCL point-to-point halo exchange of whole
// lattice rows (contiguous in the canonical numbering, so no packing kernels) and
// all-reduced scalars.  Replaces what the reference would get from PETSc
// VecScatter / ghostUpdate (femo/fea/utils_dolfinx.py:167,354-358) and MPI
// reductions (:236) -- which are only nominal there.
//
// NCCL is resolved with dlopen at femo_comm_init so that libfemo_b200.so loads
// on CPU-only boxes and shares the NCCL already loaded by torch when present.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace femo {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    void *handle = nullptr;
};

struct Comm {
    bool active = false;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    NcclApi api;
    long long halo_exchanges = 0, allreduces = 0;
};

static Comm g_comm;

#define FEMO_NCCL(call)                                                                                   \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess)                                                                           \
            return femo::set_err(FEMO_ECUDA, std::string(#call) + ": " +                                  \
                                                 (g_comm.api.GetErrorString ? g_comm.api.GetErrorString(r__) : "nccl error")); \
    } while (0)

static int nccl_load() {
    NcclApi &a = g_comm.api;
    if (a.handle) return FEMO_OK;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return set_err(FEMO_ENODEVICE, std::string("cannot load NCCL: ") + dlerror());
#define FEMO_SYM(field, name)                                                        \
    *(void **)(&a.field) = dlsym(a.handle, name);                                    \
    if (!a.field) return set_err(FEMO_ENODEVICE, std::string("NCCL symbol missing: ") + name);
    FEMO_SYM(GetUniqueId, "ncclGetUniqueId")
    FEMO_SYM(CommInitRank, "ncclCommInitRank")
    FEMO_SYM(CommDestroy, "ncclCommDestroy")
    FEMO_SYM(AllReduce, "ncclAllReduce")
    FEMO_SYM(AllGather, "ncclAllGather")
    FEMO_SYM(Broadcast, "ncclBroadcast")
    FEMO_SYM(Send, "ncclSend")
    FEMO_SYM(Recv, "ncclRecv")
    FEMO_SYM(GroupStart, "ncclGroupStart")
    FEMO_SYM(GroupEnd, "ncclGroupEnd")
    FEMO_SYM(GetErrorString, "ncclGetErrorString")
#undef FEMO_SYM
    return FEMO_OK;
}

// Partition gny cell rows over nranks (gny % nranks == 0).  Rank r owns node rows [a,b) with
// a = r*gny/nranks, b = a + gny/nranks (+ the top boundary row on the last rank) and cell rows [a,b);
// its local mesh holds cell rows [max(a-1,0), b): one ghost cell row below, one ghost node row above.
static inline SlabInfo make_slab(int gny, int rank, int nranks) {
    SlabInfo s;
    s.active = nranks > 1;
    s.rank = rank;
    s.nranks = nranks;
    s.gny = gny;
    const int rows = gny / nranks, a = rank * rows, b = a + rows;
    s.crow0 = (a > 0) ? a - 1 : 0;
    s.ncrows = b - s.crow0;
    s.own0 = a - s.crow0;
    s.own1 = s.own0 + rows + (rank == nranks - 1 ? 1 : 0);
    s.cown0 = a - s.crow0;
    s.cown1 = s.cown0 + rows;
    return s;
}

}  // namespace femo
