// families7.cuh -- family 9: nonlinear Poisson with a P2 (quadratic Lagrange) state on triangles, the
// "P2" variant of config 2 named by BASELINE.json / SURVEY.md section 8d
// (examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py:88-116,140-145 with V = CG2):
//   cells :  int grad u . grad v + u^3 v - f v dx              degree 8 -> collapsed 5x5 Gauss (25 pts)
//   facets: -(grad u.n) v + (u_ex-u)(grad v.n) + beta/h (u-u_ex) v         6-pt Gauss
//   J     :  int 1/2 (u-u_ex)^2 + alpha/2 f^2 dx, u_ex analytic            degree-12 rule (49 pts)
// Local dofs: vertices 0..2, then the midpoint of the edge opposite vertex i as 3+i (basix).  Basis in
// barycentric coordinates: phi_i = l_i (2 l_i - 1), phi_{3+i} = 4 l_j l_k.
// One thread per cell / exterior facet, element tensors in registers, SoA scratch planes.
#pragma once
#include "common.cuh"
#include "families.cuh"

namespace femo {

__constant__ double c_tri25[25][3];   // collapsed 5x5 Gauss rule, degree 9: xi, eta, w
__constant__ double c_gl6[6][2];      // 6-pt Gauss-Legendre on [0,1]

struct P2Args {
    TriArgs T;                 // geometry, facets, u (P2 dofs), f, alpha, beta, out
    const int32_t *edgesT;     // (3,ncells) SoA: edge opposite local vertex i
    int64_t nverts;
};

// basis values and physical gradients at barycentric point (l0,l1,l2)
__device__ __forceinline__ void p2_basis(const Tri &T, const double l[3], double ph[6], double gp[6][2]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        ph[i] = l[i] * (2.0 * l[i] - 1.0);
        ph[3 + i] = 4.0 * l[j] * l[k];
        const double d = 4.0 * l[i] - 1.0;
        gp[i][0] = d * T.g[i][0];
        gp[i][1] = d * T.g[i][1];
        gp[3 + i][0] = 4.0 * (l[j] * T.g[k][0] + l[k] * T.g[j][0]);
        gp[3 + i][1] = 4.0 * (l[j] * T.g[k][1] + l[k] * T.g[j][1]);
    }
}
__device__ __forceinline__ void p2_values(const double l[3], double ph[6]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ph[i] = l[i] * (2.0 * l[i] - 1.0);
        ph[3 + i] = 4.0 * l[(i + 1) % 3] * l[(i + 2) % 3];
    }
}

__device__ __forceinline__ void p2_load_u(const P2Args &A, const Tri &T, int64_t c, double u[6]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        u[a] = A.T.u[T.v[a]];
        u[3 + a] = A.T.u[A.nverts + A.edgesT[a * A.T.ncells + c]];
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_nlpoisson_p2_cell(P2Args A) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= A.T.ncells) return;
    const int64_t ne = A.T.ncells;
    Tri T;
    tri_load(A.T, c, T);
    double *out = A.T.out;
    if (OP == OP_DRDM) {       // -int phi_a dx: 0 at the vertices, area/3 at the edge midpoints
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            out[a * ne + c] = 0.0;
            out[(3 + a) * ne + c] = -0.5 * T.a2 * (1.0 / 3.0);
        }
        return;
    }
    if (OP == OP_OUT_DM) {
        out[c] = 0.5 * T.a2 * A.T.alpha * A.T.f[c];
        return;
    }
    double u[6];
    p2_load_u(A, T, c, u);
    if (OP == OP_RES) {
        const double f = A.T.f[c];
        double R[6] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < 25; ++q) {
            const double l[3] = {1.0 - c_tri25[q][0] - c_tri25[q][1], c_tri25[q][0], c_tri25[q][1]};
            double ph[6], gp[6][2];
            p2_basis(T, l, ph, gp);
            double uq = 0.0, gx = 0.0, gy = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                uq += u[a] * ph[a];
                gx += u[a] * gp[a][0];
                gy += u[a] * gp[a][1];
            }
            const double w = c_tri25[q][2] * T.a2, s = uq * uq * uq - f;
#pragma unroll
            for (int a = 0; a < 6; ++a) R[a] += w * (gx * gp[a][0] + gy * gp[a][1] + s * ph[a]);
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) out[a * ne + c] = R[a];
    } else if (OP == OP_JAC) {
        double K[6][6];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) K[a][b] = 0.0;
        for (int q = 0; q < 25; ++q) {
            const double l[3] = {1.0 - c_tri25[q][0] - c_tri25[q][1], c_tri25[q][0], c_tri25[q][1]};
            double ph[6], gp[6][2];
            p2_basis(T, l, ph, gp);
            double uq = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) uq += u[a] * ph[a];
            const double w = c_tri25[q][2] * T.a2, s = 3.0 * uq * uq;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) K[a][b] += w * (gp[a][0] * gp[b][0] + gp[a][1] * gp[b][1] + s * ph[a] * ph[b]);
        }
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) out[(a * 6 + b) * ne + c] = K[a][b];
    } else {  // OP_OUT / OP_OUT_DU
        double val = 0.0, ge[6] = {0, 0, 0, 0, 0, 0};
        const UexCell U = uex_cell(A.T, T);
        for (int q = 0; q < 49; ++q) {
            const double l[3] = {1.0 - c_tri49[q][0] - c_tri49[q][1], c_tri49[q][0], c_tri49[q][1]};
            double ph[6];
            p2_values(l, ph);
            const double x = l[0] * T.X[0][0] + l[1] * T.X[1][0] + l[2] * T.X[2][0];
            const double y = l[0] * T.X[0][1] + l[1] * T.X[1][1] + l[2] * T.X[2][1];
            double eq = -uex_q(U, q, x, y);
#pragma unroll
            for (int a = 0; a < 6; ++a) eq += u[a] * ph[a];
            const double w = c_tri49[q][2] * T.a2;
            val += w * 0.5 * eq * eq;
#pragma unroll
            for (int a = 0; a < 6; ++a) ge[a] += w * eq * ph[a];
        }
        if (OP == OP_OUT) {
            const double f = A.T.f[c];
            out[c] = val + 0.5 * T.a2 * 0.5 * A.T.alpha * f * f;
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) out[a * ne + c] = ge[a];
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_nlpoisson_p2_facet(P2Args A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= A.T.nfacets) return;
    const int64_t ne = A.T.nfacets;
    const int64_t c = A.T.bf_cell[e];
    const int l = A.T.bf_local[e];
    Tri T;
    tri_load(A.T, c, T);
    const int la = (l == 0) ? 1 : 0, lb = (l == 2) ? 1 : 2;   // facet l is opposite vertex l
    const double tx = T.X[lb][0] - T.X[la][0], ty = T.X[lb][1] - T.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (T.X[la][0] - T.X[l][0]) + ny * (T.X[la][1] - T.X[l][1]) < 0.0) { nx = -nx; ny = -ny; }
    double h2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int b = (a + 1) % 3;
        const double dx = T.X[a][0] - T.X[b][0], dy = T.X[a][1] - T.X[b][1];
        h2 = fmax(h2, dx * dx + dy * dy);
    }
    const double bh = A.T.beta / sqrt(h2);
    double *out = A.T.out;
    double u[6];
    if (OP == OP_RES) p2_load_u(A, T, c, u);
    double R[6] = {0, 0, 0, 0, 0, 0};
    double K[6][6];
    if (OP == OP_JAC) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) K[a][b] = 0.0;
    }
    for (int q = 0; q < 6; ++q) {
        const double s = c_gl6[q][0], w = c_gl6[q][1] * len;
        double lam[3] = {0.0, 0.0, 0.0};
        lam[la] = 1.0 - s;
        lam[lb] = s;
        double ph[6], gp[6][2], gn[6];
        p2_basis(T, lam, ph, gp);
#pragma unroll
        for (int a = 0; a < 6; ++a) gn[a] = gp[a][0] * nx + gp[a][1] * ny;
        if (OP == OP_RES) {
            double uq = 0.0, dudn = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                uq += u[a] * ph[a];
                dudn += u[a] * gn[a];
            }
            const double ex = uex_nlp(T.X[la][0] + s * tx, T.X[la][1] + s * ty);
#pragma unroll
            for (int a = 0; a < 6; ++a) R[a] += w * (-dudn * ph[a] + (ex - uq) * gn[a] + bh * (uq - ex) * ph[a]);
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) K[a][b] += w * (-ph[a] * gn[b] - gn[a] * ph[b] + bh * ph[a] * ph[b]);
        }
    }
    if (OP == OP_RES) {
#pragma unroll
        for (int a = 0; a < 6; ++a) out[a * ne + e] = R[a];
    } else {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) out[(a * 6 + b) * ne + e] = K[a][b];
    }
}

// ---- p-multigrid transfers between the P2 level and the P1 level on the same mesh ---------------------
// prolongation: vertex values copied, edge-midpoint values = mean of the edge's two vertex values
template <bool ADD>
__global__ void __launch_bounds__(kThreads)
    k_p1_to_p2(const int32_t *__restrict__ edge_verts, int64_t nverts, int64_t nedges, const double *__restrict__ x1,
               double *__restrict__ x2, const uint8_t *__restrict__ mask2) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nverts + nedges) return;
    if (mask2 && mask2[i]) {
        if (!ADD) x2[i] = 0.0;
        return;
    }
    const double v = (i < nverts) ? x1[i] : 0.5 * (x1[edge_verts[2 * (i - nverts)]] + x1[edge_verts[2 * (i - nverts) + 1]]);
    if (ADD) x2[i] += v;
    else x2[i] = v;
}

// restriction = transpose, in gather form through the vertex -> incident edges lists (fixed order)
__global__ void __launch_bounds__(kThreads)
    k_p2_to_p1_restrict(const int32_t *__restrict__ vptr, const int32_t *__restrict__ vedge, int64_t nverts,
                        const double *__restrict__ r2, double *__restrict__ r1, const uint8_t *__restrict__ mask2,
                        const uint8_t *__restrict__ mask1) {
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= nverts) return;
    if (mask1 && mask1[v]) {
        r1[v] = 0.0;
        return;
    }
    double acc = (mask2 && mask2[v]) ? 0.0 : r2[v];
    for (int32_t k = vptr[v]; k < vptr[v + 1]; ++k) {
        const int64_t d = nverts + vedge[k];
        if (!(mask2 && mask2[d])) acc += 0.5 * r2[d];
    }
    r1[v] = acc;
}

}  // namespace femo
