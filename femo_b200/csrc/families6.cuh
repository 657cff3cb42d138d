// families6.cuh -- family 8: SIMP linear elasticity on trilinear hexahedra (vector CG1 state, DG0
// density), the 3-D extension of examples/beam_topo_opt/run_topo_opt_cantilever_beam.py:62-86 named by
// SURVEY.md section 8d (C4-3D):
//   R = int sigma(u):eps(v) dx - int_{ds(100)} f.v ds,  sigma = lam tr(eps) I + 2 mu eps,
//   E = rho^p, lam = E nu/((1+nu)(1-2nu)), mu = E/(2(1+nu));  outputs 0 avg density, 1 compliance.
// 2x2x2 Gauss points (degree 2 per direction, exact on parallelepipeds).
//
// Thread mapping: a CTA of 256 threads owns 32 consecutive cells; lane = cell, warp w = 0..7.
//   phase 1: thread (cell, q = w) computes the Jacobian at Gauss point q and stores the 8 physical basis
//            gradients + weight in shared memory ([q][a][d][lane]: conflict-free); for the residual-type
//            ops also the weighted stress at q
//   phase 2: thread (cell, a = w) forms row-node a's entries (JAC: 8 blocks of 3x3; RES/DRDM: 3 values)
// Every global store has lane = consecutive cell in one SoA plane => fully coalesced; the element tensor
// (24x24) never lives in one thread's registers.
#pragma once
#include "common.cuh"
#include "families.cuh"

namespace femo {

struct HexArgs {
    const double *coords;      // (nverts,3) AoS
    const int32_t *cellsT;     // (8,ncells)
    int64_t ncells;
    const int32_t *fb_cell, *fb_local;
    int64_t nfacets;
    const double *u, *rho;
    double nu, f[3], penal, volume;
    int out_id;
    double *out;
};

constexpr int kHexCells = 32;                                        // cells per CTA
constexpr int kHexSmem = (8 * 8 * 3 + 8 + 8 * 6) * kHexCells * 8;     // gradients + weights + stresses (bytes)

template <int OP>
__global__ void __launch_bounds__(kThreads) k_simp_hex_cell(HexArgs A) {
    extern __shared__ double hs[];
    double *G = hs;                              // [q][a][d][lane]
    double *W = hs + 8 * 8 * 3 * kHexCells;      // [q][lane]
    double *S = W + 8 * kHexCells;               // [q][6][lane]  weighted stress (xx,yy,zz,yz,xz,xy)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t c = blockIdx.x * (int64_t)kHexCells + lane;
    const int64_t ne = A.ncells;
    const bool valid = c < ne;
    constexpr bool kGrad = (OP == OP_RES || OP == OP_JAC || OP == OP_DRDM);
    constexpr bool kStress = (OP == OP_RES || OP == OP_DRDM);
    const double lam = A.nu / ((1.0 + A.nu) * (1.0 - 2.0 * A.nu)), mu = 1.0 / (2.0 * (1.0 + A.nu));
    int v[8];
    if (valid) {
        // ---- phase 1: Gauss point q = w -------------------------------------------------
        double X[8][3];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            v[a] = A.cellsT[a * ne + c];
            X[a][0] = __ldg(A.coords + 3 * (int64_t)v[a]);
            X[a][1] = __ldg(A.coords + 3 * (int64_t)v[a] + 1);
            X[a][2] = __ldg(A.coords + 3 * (int64_t)v[a] + 2);
        }
        const double g0 = 0.5 - 0.28867513459481287, g1 = 0.5 + 0.28867513459481287;
        const double t[3] = {(w & 1) ? g1 : g0, (w & 2) ? g1 : g0, (w & 4) ? g1 : g0};
        double dN[8][3];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const double l0 = (a & 1) ? t[0] : 1.0 - t[0], l1 = (a & 2) ? t[1] : 1.0 - t[1], l2 = (a & 4) ? t[2] : 1.0 - t[2];
            dN[a][0] = ((a & 1) ? 1.0 : -1.0) * l1 * l2;
            dN[a][1] = ((a & 2) ? 1.0 : -1.0) * l0 * l2;
            dN[a][2] = ((a & 4) ? 1.0 : -1.0) * l0 * l1;
        }
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};    // J[i][j] = d x_i / d xi_j
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) J[i][j] += X[a][i] * dN[a][j];
        double C[3][3];                                           // cofactors: inv(J) = C^T / det
        C[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        C[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        C[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        C[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
        C[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
        C[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
        C[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        C[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
        C[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double det = J[0][0] * C[0][0] + J[0][1] * C[0][1] + J[0][2] * C[0][2];
        const double wq = 0.125 * fabs(det);
        W[w * kHexCells + lane] = wq;
        if (kGrad) {
            const double id = 1.0 / det;
            double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // grad u at q
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                double g[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    // dN_a/dx_d = sum_j dN_a/dxi_j * inv(J)[j][d],  inv(J)[j][d] = C[d][j]/det
                    g[d] = (dN[a][0] * C[d][0] + dN[a][1] * C[d][1] + dN[a][2] * C[d][2]) * id;
                    G[((w * 8 + a) * 3 + d) * kHexCells + lane] = g[d];
                }
                if (kStress) {
                    const double ua[3] = {A.u[3 * (int64_t)v[a]], A.u[3 * (int64_t)v[a] + 1], A.u[3 * (int64_t)v[a] + 2]};
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int d = 0; d < 3; ++d) H[i][d] += ua[i] * g[d];
                }
            }
            if (kStress) {
                const double tr = H[0][0] + H[1][1] + H[2][2];
                double *s = S + (w * 6) * kHexCells + lane;
                s[0 * kHexCells] = wq * (lam * tr + 2.0 * mu * H[0][0]);
                s[1 * kHexCells] = wq * (lam * tr + 2.0 * mu * H[1][1]);
                s[2 * kHexCells] = wq * (lam * tr + 2.0 * mu * H[2][2]);
                s[3 * kHexCells] = wq * mu * (H[1][2] + H[2][1]);
                s[4 * kHexCells] = wq * mu * (H[0][2] + H[2][0]);
                s[5 * kHexCells] = wq * mu * (H[0][1] + H[1][0]);
            }
        }
    }
    __syncthreads();
    if (!valid) return;
    // ---- phase 2: row node a = w ------------------------------------------------------------
    const double rho = A.rho[c];
    if (OP == OP_OUT || OP == OP_OUT_DM) {
        if (w == 0) {
            double vol = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) vol += W[q * kHexCells + lane];
            A.out[c] = (OP == OP_OUT ? rho : 1.0) * vol / A.volume;
        }
        return;
    }
    if (OP == OP_JAC) {
        const double E = pow(rho, A.penal);
        double ga[8][3], wq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            wq[q] = E * W[q * kHexCells + lane];
#pragma unroll
            for (int d = 0; d < 3; ++d) ga[q][d] = G[((q * 8 + w) * 3 + d) * kHexCells + lane];
        }
#pragma unroll 1
        for (int b = 0; b < 8; ++b) {
            double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double gb[3] = {G[((q * 8 + b) * 3 + 0) * kHexCells + lane], G[((q * 8 + b) * 3 + 1) * kHexCells + lane],
                                      G[((q * 8 + b) * 3 + 2) * kHexCells + lane]};
                const double dotw = wq[q] * mu * (ga[q][0] * gb[0] + ga[q][1] * gb[1] + ga[q][2] * gb[2]);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        acc[i][j] += wq[q] * (lam * ga[q][i] * gb[j] + mu * ga[q][j] * gb[i]) + (i == j ? dotw : 0.0);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) A.out[(int64_t)((3 * w + i) * 24 + 3 * b + j) * ne + c] = acc[i][j];
        }
    } else {   // OP_RES: rho^p K u ; OP_DRDM: p rho^(p-1) K u   (r_a = sum_q sigma_q . grad N_a)
        const double E = (OP == OP_RES) ? pow(rho, A.penal) : A.penal * pow(rho, A.penal - 1.0);
        double r[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double g[3] = {G[((q * 8 + w) * 3 + 0) * kHexCells + lane], G[((q * 8 + w) * 3 + 1) * kHexCells + lane],
                                 G[((q * 8 + w) * 3 + 2) * kHexCells + lane]};
            const double *s = S + (q * 6) * kHexCells + lane;
            const double sxx = s[0], syy = s[kHexCells], szz = s[2 * kHexCells], syz = s[3 * kHexCells], sxz = s[4 * kHexCells],
                         sxy = s[5 * kHexCells];
            r[0] += sxx * g[0] + sxy * g[1] + sxz * g[2];
            r[1] += sxy * g[0] + syy * g[1] + syz * g[2];
            r[2] += sxz * g[0] + syz * g[1] + szz * g[2];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) A.out[(int64_t)(3 * w + i) * ne + c] = E * r[i];
    }
}

// local vertices of the six hexahedron facets, tensor order on the face (s fastest)
__device__ __constant__ int c_hex_facet[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {0, 2, 4, 6}, {1, 3, 5, 7}, {2, 3, 6, 7}, {4, 5, 6, 7}};

// traction facets: OP_RES -> -int f.v ds, OP_OUT -> int u.f ds, OP_OUT_DU -> int f.v ds   (2x2 Gauss on the face)
template <int OP>
__global__ void __launch_bounds__(kThreads) k_simp_hex_facet(HexArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= A.nfacets) return;
    const int64_t ne = A.nfacets;
    const int64_t c = A.fb_cell[e];
    const int l = A.fb_local[e];
    int v[4];
    double X[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = A.cellsT[c_hex_facet[l][k] * A.ncells + c];
#pragma unroll
        for (int d = 0; d < 3; ++d) X[k][d] = __ldg(A.coords + 3 * (int64_t)v[k] + d);
    }
    const double g0 = 0.5 - 0.28867513459481287, g1 = 0.5 + 0.28867513459481287;
    double m[4] = {0, 0, 0, 0};                   // int N_k ds
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double s = (q & 1) ? g1 : g0, t = (q & 2) ? g1 : g0;
        const double N[4] = {(1 - s) * (1 - t), s * (1 - t), (1 - s) * t, s * t};
        const double dNs[4] = {-(1 - t), (1 - t), -t, t}, dNt[4] = {-(1 - s), -s, (1 - s), s};
        double xs[3] = {0, 0, 0}, xt[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                xs[d] += X[k][d] * dNs[k];
                xt[d] += X[k][d] * dNt[k];
            }
        const double nx = xs[1] * xt[2] - xs[2] * xt[1], ny = xs[2] * xt[0] - xs[0] * xt[2], nz = xs[0] * xt[1] - xs[1] * xt[0];
        const double da = 0.25 * sqrt(nx * nx + ny * ny + nz * nz);
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] += da * N[k];
    }
    if (OP == OP_OUT) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            acc += m[k] * (A.f[0] * A.u[3 * (int64_t)v[k]] + A.f[1] * A.u[3 * (int64_t)v[k] + 1] + A.f[2] * A.u[3 * (int64_t)v[k] + 2]);
        A.out[e] = acc;
    } else {
        const double sg = (OP == OP_RES) ? -1.0 : 1.0;
        for (int a = 0; a < 8; ++a) {
            double ma = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) ma = (c_hex_facet[l][k] == a) ? m[k] : ma;
#pragma unroll
            for (int d = 0; d < 3; ++d) A.out[(int64_t)(3 * a + d) * ne + e] = sg * ma * A.f[d];
        }
    }
}

}  // namespace femo
