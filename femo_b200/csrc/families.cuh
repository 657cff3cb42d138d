// families.cuh -- per-cell / per-facet quadrature kernels of the form families.
//
// These replace the FFCx-generated tabulate_tensor kernels that the reference
// JIT-compiles for every `form(...)` (femo/fea/utils_dolfinx.py:173,179,185,
// 193-197).  One thread owns one cell (or exterior facet): it gathers vertex
// coordinates and coefficients, evaluates the element tensor in registers with
// the quadrature degrees of SURVEY.md Appendix A.3, and stores entry k of
// entity e at scratch[k*ne + e] (SoA planes => fully coalesced stores).  The
// sorted segmented reduction in engine.cu then sums the planes into the CSR
// values / vector entries in a fixed order (no atomics).
#pragma once
#include "common.cuh"

namespace femo {

// quadrature tables (filled by femo_problem_upload)
__constant__ double c_tri6[6][3];     // degree-4 rule: xi, eta, w   (sum w = 1/2)
__constant__ double c_gl5[5][2];      // 5-pt Gauss-Legendre on [0,1]: s, w
__constant__ double c_tri49[49][3];   // collapsed 7x7 Gauss rule, degree 12
// angle-addition table of the analytic u_ex on the current problem's uniform lattice (refreshed before every launch that
// reads it): [triangle type][point] = cos(2 pi dx), sin(2 pi dx), cos(pi dy), sin(pi dy).  With the point loop fully
// unrolled and one code path per triangle type every entry is a constant-bank operand of an FMA (no load instructions).
__constant__ double c_uex49[2][49][4];

struct TriArgs {
    const double *coords;    // (nverts,2) AoS
    const int32_t *cellsT;   // (3,ncells) SoA
    int64_t ncells;
    const int32_t *bf_cell, *bf_local;
    int64_t nfacets;
    const double *u, *f, *uex;
    double alpha, beta;
    double *out;             // scratch base of this block
    // uniform right-diagonal lattices: u_ex(x0 + dx_q, y0 + dy_q) by angle addition from 2 sincos per cell and a
    // per-point table tab[(type*49+q)*4] = cos(2 pi dx), sin(2 pi dx), cos(pi dy), sin(pi dy)  (nullptr: general path)
    const double *uex_tab;
};

struct Tri {
    int v[3];
    double X[3][2];
    double g[3][2];
    double a2;  // |det J| = 2*area
};

__device__ __forceinline__ void tri_load(const TriArgs &A, int64_t c, Tri &T) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T.v[a] = A.cellsT[a * A.ncells + c];
        const double2 xy = __ldg(reinterpret_cast<const double2 *>(A.coords) + T.v[a]);
        T.X[a][0] = xy.x;
        T.X[a][1] = xy.y;
    }
    const double e1x = T.X[1][0] - T.X[0][0], e1y = T.X[1][1] - T.X[0][1];
    const double e2x = T.X[2][0] - T.X[0][0], e2y = T.X[2][1] - T.X[0][1];
    const double det = e1x * e2y - e1y * e2x;
    const double id = 1.0 / det;
    T.g[1][0] = e2y * id;  T.g[1][1] = -e2x * id;
    T.g[2][0] = -e1y * id; T.g[2][1] = e1x * id;
    T.g[0][0] = -(T.g[1][0] + T.g[2][0]);
    T.g[0][1] = -(T.g[1][1] + T.g[2][1]);
    T.a2 = fabs(det);
}

__device__ __forceinline__ double uex_nlp(double x, double y) {
    // examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py:144-145
    const double pi = 3.141592653589793;
    return sin(2.0 * pi * x) * sin(pi * y);
}

// u_ex at quadrature point q of the degree-12 rule: table path on uniform lattices, direct evaluation otherwise
struct UexCell {
    double S0, C0, S1, C1;
    const double *tab;
};
__device__ __forceinline__ UexCell uex_cell(const TriArgs &A, const Tri &T) {
    UexCell U;
    U.tab = nullptr;
    if (A.uex_tab) {
        sincospi(2.0 * T.X[0][0], &U.S0, &U.C0);
        sincospi(T.X[0][1], &U.S1, &U.C1);
        U.tab = A.uex_tab + ((T.X[1][1] == T.X[0][1]) ? 0 : 49 * 4);   // lower [v0,v1,v3] or upper [v0,v2,v3] triangle
    }
    return U;
}
__device__ __forceinline__ double uex_q(const UexCell &U, int q, double x, double y) {
    if (U.tab) {
        const double *t = U.tab + 4 * q;
        return (U.S0 * t[0] + U.C0 * t[1]) * (U.S1 * t[2] + U.C1 * t[3]);
    }
    return uex_nlp(x, y);
}

enum CellOp { OP_RES = 0, OP_JAC = 1, OP_DRDM = 2, OP_OUT = 3, OP_OUT_DU = 4, OP_OUT_DM = 5,
              OP_OUT_BOTH = 6 /* functional (plane 3) and its state gradient (planes 0-2) in one quadrature pass */ };

// ---------------------------------------------------------------------------
// family 1: Poisson, P1 state, DG0 source  (examples/poisson_opt/run_poisson_opt.py)
//   R = int grad u . grad v - f v dx                      (:32-38)
//   J = int 1/2 (u-u_ex)^2 + alpha/2 f^2 dx, u_ex in P1   (:74-76,108)
// ---------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(kThreads) k_poisson_p1_cell(TriArgs A) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= A.ncells) return;
    const int64_t ne = A.ncells;
    Tri T;
    tri_load(A, c, T);
    if (OP == OP_RES) {
        double gu[2] = {0.0, 0.0};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double ua = A.u[T.v[a]];
            gu[0] += ua * T.g[a][0];
            gu[1] += ua * T.g[a][1];
        }
        const double f = A.f[c];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            A.out[a * ne + c] = 0.5 * T.a2 * (gu[0] * T.g[a][0] + gu[1] * T.g[a][1] - f * (1.0 / 3.0));
    } else if (OP == OP_JAC) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                A.out[(a * 3 + b) * ne + c] = 0.5 * T.a2 * (T.g[a][0] * T.g[b][0] + T.g[a][1] * T.g[b][1]);
    } else if (OP == OP_DRDM) {
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + c] = -0.5 * T.a2 * (1.0 / 3.0);
    } else if (OP == OP_OUT || OP == OP_OUT_DU) {
        double e[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) e[a] = A.u[T.v[a]] - A.uex[T.v[a]];
        // degree-2 rule: barycentric (2/3,1/6,1/6) and rotations, w = 1/6
        const double w = T.a2 * (1.0 / 6.0);
        double val = 0.0, ge[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double ph[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) ph[a] = (a == q) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double eq = e[0] * ph[0] + e[1] * ph[1] + e[2] * ph[2];
            val += w * 0.5 * eq * eq;
#pragma unroll
            for (int a = 0; a < 3; ++a) ge[a] += w * eq * ph[a];
        }
        if (OP == OP_OUT) {
            const double f = A.f[c];
            A.out[c] = val + 0.5 * T.a2 * 0.5 * A.alpha * f * f;
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) A.out[a * ne + c] = ge[a];
        }
    } else if (OP == OP_OUT_DM) {
        A.out[c] = 0.5 * T.a2 * A.alpha * A.f[c];
    }
}

// ---------------------------------------------------------------------------
// family 2: nonlinear Poisson + symmetric Nitsche
//   (examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py)
//   cells  (:88-95):  int grad u . grad v + u^3 v - f v dx       degree 4 -> 6 pts
//   facets (:97-116): -(grad u.n) v + (u_ex-u)(grad v.n) + beta/h (u-u_ex) v
//                                                                 degree 9 -> 5-pt Gauss
//   J (:140-142): int 1/2 (u-u_ex)^2 + alpha/2 f^2 dx, u_ex analytic  degree 12
// ---------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(kThreads) k_nlpoisson_p1_cell(TriArgs A) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if ((OP == OP_OUT || OP == OP_OUT_DU || OP == OP_OUT_BOTH) && A.uex_tab) {
        // lattice meshes alternate lower / upper triangles: let each half of the CTA take one type so that the
        // angle-addition table reads of the 49-point rule are warp-uniform (even cells: warps 0-3, odd cells: warps 4-7)
        const int t = threadIdx.x;
        c = blockIdx.x * (int64_t)blockDim.x + 2 * (t % (kThreads / 2)) + t / (kThreads / 2);
    }
    if (c >= A.ncells) return;
    const int64_t ne = A.ncells;
    Tri T;
    tri_load(A, c, T);
    if (OP == OP_DRDM) {
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + c] = -0.5 * T.a2 * (1.0 / 3.0);
        return;
    }
    if (OP == OP_OUT_DM) {
        A.out[c] = 0.5 * T.a2 * A.alpha * A.f[c];
        return;
    }
    double u[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) u[a] = A.u[T.v[a]];
    if (OP == OP_RES) {
        const double gu0 = u[0] * T.g[0][0] + u[1] * T.g[1][0] + u[2] * T.g[2][0];
        const double gu1 = u[0] * T.g[0][1] + u[1] * T.g[1][1] + u[2] * T.g[2][1];
        const double f = A.f[c];
        double R[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) R[a] = 0.5 * T.a2 * (gu0 * T.g[a][0] + gu1 * T.g[a][1]);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double ph[3] = {1.0 - c_tri6[q][0] - c_tri6[q][1], c_tri6[q][0], c_tri6[q][1]};
            const double uq = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2];
            const double s = c_tri6[q][2] * T.a2 * (uq * uq * uq - f);
#pragma unroll
            for (int a = 0; a < 3; ++a) R[a] += s * ph[a];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + c] = R[a];
    } else if (OP == OP_JAC) {
        double K[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) K[a][b] = 0.5 * T.a2 * (T.g[a][0] * T.g[b][0] + T.g[a][1] * T.g[b][1]);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double ph[3] = {1.0 - c_tri6[q][0] - c_tri6[q][1], c_tri6[q][0], c_tri6[q][1]};
            const double uq = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2];
            const double s = c_tri6[q][2] * T.a2 * 3.0 * uq * uq;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) K[a][b] += s * ph[a] * ph[b];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) A.out[(a * 3 + b) * ne + c] = K[a][b];
    } else if (A.uex_tab) {  // OP_OUT / OP_OUT_DU on a uniform lattice: u_ex by angle addition from the constant table
        double val = 0.0, ge[3] = {0.0, 0.0, 0.0};
        double S0, C0, S1, C1;
        sincospi(2.0 * T.X[0][0], &S0, &C0);
        sincospi(T.X[0][1], &S1, &C1);
        const bool upper = T.X[1][1] != T.X[0][1];          // lower [v0,v1,v3] or upper [v0,v2,v3] triangle (warp-uniform)
#define FEMO_OUT49(TY)                                                                                              \
        _Pragma("unroll") for (int q = 0; q < 49; ++q) {                                                            \
            const double ph1 = c_tri49[q][0], ph2 = c_tri49[q][1], ph0 = 1.0 - ph1 - ph2;                           \
            const double ex = (S0 * c_uex49[TY][q][0] + C0 * c_uex49[TY][q][1]) * (S1 * c_uex49[TY][q][2] + C1 * c_uex49[TY][q][3]); \
            const double eq = u[0] * ph0 + u[1] * ph1 + u[2] * ph2 - ex;                                            \
            const double w = c_tri49[q][2] * T.a2;                                                                  \
            val += w * 0.5 * eq * eq;                                                                               \
            ge[0] += w * eq * ph0; ge[1] += w * eq * ph1; ge[2] += w * eq * ph2;                                    \
        }
        if (upper) { FEMO_OUT49(1) } else { FEMO_OUT49(0) }
#undef FEMO_OUT49
        if (OP == OP_OUT || OP == OP_OUT_BOTH) {
            const double f = A.f[c];
            A.out[(OP == OP_OUT_BOTH ? 3 * ne : 0) + c] = val + 0.5 * T.a2 * 0.5 * A.alpha * f * f;
        }
        if (OP == OP_OUT_DU || OP == OP_OUT_BOTH) {
#pragma unroll
            for (int a = 0; a < 3; ++a) A.out[a * ne + c] = ge[a];
        }
    } else {  // OP_OUT / OP_OUT_DU : degree-12 rule, u_ex evaluated at the points
        double val = 0.0, ge[3] = {0.0, 0.0, 0.0};
        const UexCell U = uex_cell(A, T);
#pragma unroll 7
        for (int q = 0; q < 49; ++q) {
            const double ph[3] = {1.0 - c_tri49[q][0] - c_tri49[q][1], c_tri49[q][0], c_tri49[q][1]};
            const double x = ph[0] * T.X[0][0] + ph[1] * T.X[1][0] + ph[2] * T.X[2][0];
            const double y = ph[0] * T.X[0][1] + ph[1] * T.X[1][1] + ph[2] * T.X[2][1];
            const double eq = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2] - uex_q(U, q, x, y);
            const double w = c_tri49[q][2] * T.a2;
            val += w * 0.5 * eq * eq;
#pragma unroll
            for (int a = 0; a < 3; ++a) ge[a] += w * eq * ph[a];
        }
        if (OP == OP_OUT || OP == OP_OUT_BOTH) {
            const double f = A.f[c];
            A.out[(OP == OP_OUT_BOTH ? 3 * ne : 0) + c] = val + 0.5 * T.a2 * 0.5 * A.alpha * f * f;
        }
        if (OP == OP_OUT_DU || OP == OP_OUT_BOTH) {
#pragma unroll
            for (int a = 0; a < 3; ++a) A.out[a * ne + c] = ge[a];
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_nlpoisson_p1_facet(TriArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= A.nfacets) return;
    const int64_t ne = A.nfacets;
    const int64_t c = A.bf_cell[e];
    const int l = A.bf_local[e];
    Tri T;
    tri_load(A, c, T);
    const int la = (l == 0) ? 1 : 0, lb = (l == 2) ? 1 : 2;   // facet l is opposite vertex l
    const double tx = T.X[lb][0] - T.X[la][0], ty = T.X[lb][1] - T.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (T.X[la][0] - T.X[l][0]) + ny * (T.X[la][1] - T.X[l][1]) < 0.0) { nx = -nx; ny = -ny; }
    // UFL CellDiameter: longest vertex distance
    double h2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int b = (a + 1) % 3;
        const double dx = T.X[a][0] - T.X[b][0], dy = T.X[a][1] - T.X[b][1];
        h2 = fmax(h2, dx * dx + dy * dy);
    }
    const double bh = A.beta / sqrt(h2);
    double gn[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) gn[a] = T.g[a][0] * nx + T.g[a][1] * ny;
    if (OP == OP_RES) {
        double u[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) u[a] = A.u[T.v[a]];
        const double dudn = u[0] * gn[0] + u[1] * gn[1] + u[2] * gn[2];
        double R[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const double s = c_gl5[q][0], w = c_gl5[q][1] * len;
            double ph[3] = {0.0, 0.0, 0.0};
            ph[la] = 1.0 - s;
            ph[lb] = s;
            const double uq = u[la] * (1.0 - s) + u[lb] * s;
            const double ex = uex_nlp(T.X[la][0] + s * tx, T.X[la][1] + s * ty);
#pragma unroll
            for (int a = 0; a < 3; ++a) R[a] += w * (-dudn * ph[a] + (ex - uq) * gn[a] + bh * (uq - ex) * ph[a]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + e] = R[a];
    } else {  // OP_JAC
        double K[3][3] = {{0.0}};
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const double s = c_gl5[q][0], w = c_gl5[q][1] * len;
            double ph[3] = {0.0, 0.0, 0.0};
            ph[la] = 1.0 - s;
            ph[lb] = s;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) K[a][b] += w * (-ph[a] * gn[b] - gn[a] * ph[b] + bh * ph[a] * ph[b]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) A.out[(a * 3 + b) * ne + e] = K[a][b];
    }
}

}  // namespace femo
