// stencil.cuh -- diagonal (DIA) storage of lattice operators for the multigrid V-cycle.
//
// On a lattice mesh every row of the scalar P1 Jacobian couples a node with the same <= 9 relative
// neighbours, so the level matrix is stored as D planes (one per column offset col - row) of fp32 values:
// no column indices, no row pointers, 4*D bytes per row instead of the 8*D + 4 of the fp32 CSR copy, and
// every load of the operator kernel is coalesced with no staging in shared memory.  The planes are a
// re-layout of the assembled (BC'd) CSR values, so the operator is the same matrix the reference would
// factorise (femo/fea/utils_dolfinx.py:476-512) rounded to fp32 -- used inside the preconditioner only.
//
// Row epilogues: the Chebyshev smoother steps of multigrid.cuh are fused into the operator, the
// Jacobi scaling 1/a_ii is taken from the diagonal plane (no dinv vector), the zero-guess degree-2
// pre-smoother is ONE kernel, and residual/direction vectors that nobody reads are not written.
//
// Included by engine.cu after SpmvEpi is defined.
#pragma once
#include "common.cuh"

namespace femo {

// CSR (fp64) -> DIA planes (fp32); entries absent from the CSR row are stored as 0.  `bad` is raised when an
// entry's offset is not in the table (the host checked the pattern, so this never fires on a sane layout).
// The same pass yields what the smoother set-up needs from the fp64 values: dinv_i = 1/a_ii and the per-CTA maximum of
// the Gershgorin bound sum_j |a_ij| / |a_ii| over the owned rows [o0,o1) (partials[blockIdx.x]).
template <int D>
__global__ void __launch_bounds__(kThreads)
    k_csr_to_dia(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ vals,
                 DiaMat A, float *__restrict__ planes, int *__restrict__ bad, double *__restrict__ dinv, int64_t o0,
                 int64_t o1, double *__restrict__ partials) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double g = 0.0;
    if (i < A.n) {
        float a[D];
#pragma unroll
        for (int s = 0; s < D; ++s) a[s] = 0.0f;
        double diag = 1.0, sabs = 0.0;
        for (int32_t t = rowptr[i]; t < rowptr[i + 1]; ++t) {
            const int d = col[t] - (int)i;
            const double v64 = vals[t];
            const float v = (float)v64;
            sabs += fabs(v64);
            if (d == 0) diag = v64;
            bool hit = false;
#pragma unroll
            for (int s = 0; s < D; ++s)
                if (d == A.off[s]) {
                    a[s] = v;
                    hit = true;
                }
            if (!hit) *bad = 1;
        }
#pragma unroll
        for (int s = 0; s < D; ++s) planes[(int64_t)s * A.np + i] = a[s];
        const double di = (diag != 0.0) ? 1.0 / diag : 1.0;
        dinv[i] = di;
        if (i >= o0 && i < o1) g = sabs * fabs(di);
    }
    __shared__ double sh[kThreads];
    sh[threadIdx.x] = g;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

enum DiaMode {
    DIA_PLAIN = 0,   // y = A x, or y = b - A x when b is given
    DIA_CHEB0 = 1,   // r = b - A x ; rout = r ; dout = c1 r / a_ii                     (first Chebyshev step, x untouched)
    DIA_CHEBK = 2,   // r = rin - A d ; [rout = r] ; dn = c1 d_i + c2 r / a_ii ; [dout = dn] ; xacc (+)= ... as SpmvEpi.xmode
    DIA_PRE2 = 3     // zero-guess degree-2 smoother: d0 = c0 b / a ; r = b - A d0 ; x = d0 + c1 d0 + c2 r / a_ii
};

struct DiaEpi {
    const double *b = nullptr, *rin = nullptr;
    double *rout = nullptr, *dout = nullptr, *xacc = nullptr;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    int xmode = 0;
};

__device__ __forceinline__ double dia_rcp(float d) { return d != 0.0f ? (double)(1.0f / d) : 1.0; }

template <int MODE, int D>
__global__ void __launch_bounds__(kThreads)
    k_dia_apply(DiaMat A, const double *__restrict__ x, double *__restrict__ y, DiaEpi E) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    float a[D];
#pragma unroll
    for (int s = 0; s < D; ++s) a[s] = __ldcs(A.v + (int64_t)s * A.np + i);
    float ad = 0.0f;             // a_ii (selected under unrolling: no dynamically indexed registers)
#pragma unroll
    for (int s = 0; s < D; ++s)
        if (s == A.sdiag) ad = a[s];
    double acc = 0.0;
    if (MODE == DIA_PRE2) {
        const float *dg = A.v + (int64_t)A.sdiag * A.np;
        double d0i = 0.0;
#pragma unroll
        for (int s = 0; s < D; ++s) {
            const int64_t j = i + A.off[s];
            double v = 0.0;
            if (j >= 0 && j < A.n) v = E.c0 * dia_rcp(s == A.sdiag ? ad : __ldg(dg + j)) * __ldg(E.b + j);
            if (s == A.sdiag) d0i = v;
            acc = fma((double)a[s], v, acc);
        }
        const double r = __ldg(E.b + i) - acc;
        y[i] = d0i + E.c1 * d0i + E.c2 * dia_rcp(ad) * r;
        return;
    }
    double xi = 0.0;
#pragma unroll
    for (int s = 0; s < D; ++s) {
        const int64_t j = i + A.off[s];
        const double v = (j >= 0 && j < A.n) ? __ldg(x + j) : 0.0;
        if (s == A.sdiag) xi = v;
        acc = fma((double)a[s], v, acc);
    }
    if (MODE == DIA_PLAIN) {
        y[i] = E.b ? E.b[i] - acc : acc;
    } else if (MODE == DIA_CHEB0) {
        const double r = E.b[i] - acc;
        E.rout[i] = r;
        E.dout[i] = E.c1 * dia_rcp(ad) * r;
    } else {
        const double r = E.rin[i] - acc;
        if (E.rout) E.rout[i] = r;
        const double dn = E.c1 * xi + E.c2 * dia_rcp(ad) * r;
        if (E.dout) E.dout[i] = dn;
        if (E.xmode == 0) E.xacc[i] += dn;
        else if (E.xmode == 1) E.xacc[i] += xi + dn;
        else E.xacc[i] = xi + dn;
    }
}

// Zero-guess degree-2 pre-smoother of the 7-point lattice stencil (offsets {-w-1,-w | -1,0,1 | w,w+1}) with the scaled
// right-hand side d0 = c0 b / a_jj staged in shared memory: a CTA of 256 consecutive rows needs d0 on three contiguous
// index ranges (the lattice rows below / of / above), each computed ONCE per CTA (one b load, one diagonal load, one
// reciprocal per entry) instead of 7 gathers + 7 reciprocals per row.
__global__ void __launch_bounds__(kThreads)
    k_dia_pre2(DiaMat A, const double *__restrict__ b, double *__restrict__ x, DiaEpi E) {
    __shared__ double sd[3][kThreads + 2];
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x;
    const int w = A.off[5];
    const float *dg = A.v + 3 * A.np;
    const int64_t start[3] = {i0 - w - 1, i0 - 1, i0 + w};
#pragma unroll
    for (int g = 0; g < 3; ++g)
        for (int t = threadIdx.x; t < kThreads + 2; t += kThreads) {
            const int64_t j = start[g] + t;
            sd[g][t] = (j >= 0 && j < A.n) ? E.c0 * dia_rcp(__ldg(dg + j)) * __ldg(b + j) : 0.0;
        }
    __syncthreads();
    const int64_t i = i0 + threadIdx.x;
    if (i >= A.n) return;
    const int t = threadIdx.x;
    float a[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) a[s] = __ldcs(A.v + (int64_t)s * A.np + i);
    double acc = (double)a[0] * sd[0][t];
    acc = fma((double)a[1], sd[0][t + 1], acc);
    acc = fma((double)a[2], sd[1][t], acc);
    const double d0i = sd[1][t + 1];
    acc = fma((double)a[3], d0i, acc);
    acc = fma((double)a[4], sd[1][t + 2], acc);
    acc = fma((double)a[5], sd[2][t], acc);
    acc = fma((double)a[6], sd[2][t + 1], acc);
    const double r = __ldg(b + i) - acc;
    x[i] = d0i + E.c1 * d0i + E.c2 * dia_rcp(a[3]) * r;
}

}  // namespace femo
