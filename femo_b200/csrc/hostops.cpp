// hostops.cpp -- multi-threaded bulk operations on HOST vectors for the API boundary.
//
// The femo callbacks exchange dense fp64 numpy vectors (femo/csdl_opt/state_model.py:75-200: results are assigned,
// jacvec products accumulated with +=).  At 16M dofs one single-threaded pass over such a vector costs as much as a
// multigrid cycle on the GPU; these run the copies / accumulations of the host side on all cores.  No device code.
#include <omp.h>

#include <cstdint>
#include <cstring>

#include "../../include/femo_b200.h"

namespace {
constexpr int64_t kChunk = 1 << 16;   // doubles per work item (512 KB)
}

extern "C" {

/* dst = alpha * src */
void femo_host_scaled_copy(double *dst, const double *src, int64_t n, double alpha) {
    const int64_t nchunks = (n + kChunk - 1) / kChunk;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t a = c * kChunk, b = a + kChunk < n ? a + kChunk : n;
        if (alpha == 1.0) {
            memcpy(dst + a, src + a, (size_t)(b - a) * sizeof(double));
        } else {
            for (int64_t i = a; i < b; ++i) dst[i] = alpha * src[i];
        }
    }
}

/* dst += alpha * src */
void femo_host_axpy(double *dst, const double *src, int64_t n, double alpha) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) dst[i] += alpha * src[i];
}

/* dst = value */
void femo_host_fill(double *dst, int64_t n, double value) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) dst[i] = value;
}

}  // extern "C"
