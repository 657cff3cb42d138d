// families5.cuh -- family 6: hyperelastic mesh motion of the motor example
// (examples/em_motor_opt/motor_pde.py:134-183 pdeResMM, :199-210 area_form).
//
//   F = I + grad(uhat), E = (F^T F - I)/2, S = det(F)^-3 (2E + tr(E)/3 I), P = F S
//   R = int P:grad(v) dx + sum over tagged one-sided facets of
//         -(P n).v + (dP[v] n).(uhat - g) + beta0/(det(F)^3 h_E) v.(uhat - g)
// State uhat and input g are vector P1 (2 interleaved components).  The residual already contains the
// Gateaux derivative dP[v]; it is obtained from duals seeded on the six element dofs, and dR/duhat, dR/dg
// from a second, outer layer of duals (Dual<6, Dual<6>>) -- exact second derivatives, no differencing.
#pragma once
#include "dual.cuh"
#include "families.cuh"

namespace femo {

struct MmArgs {
    const double *coords;
    const int32_t *cellsT;
    int64_t ncells;
    const int32_t *fb_cell, *fb_local;
    int64_t nfacets;
    const int32_t *tag;
    const double *uh, *g;
    double beta0;
    int out_id;
    double *out;
};

template <class T>
__device__ __forceinline__ void mm_F(const Tri &G, const T uh[3][2], T F[2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = uh[0][i] * G.g[0][j] + uh[1][i] * G.g[1][j] + uh[2][i] * G.g[2][j] + (i == j ? 1.0 : 0.0);
}

template <class T>
__device__ __forceinline__ void mm_P(const T F[2][2], T P[2][2], T &J) {
    T E[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) E[i][j] = 0.5 * (F[0][i] * F[0][j] + F[1][i] * F[1][j] - (i == j ? 1.0 : 0.0));
    const T trE = E[0][0] + E[1][1];
    J = F[0][0] * F[1][1] - F[0][1] * F[1][0];
    const T s = 1.0 / (J * J * J);
    T S[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            T c = 2.0 * E[i][j];
            if (i == j) c = c + (1.0 / 3.0) * trE;
            S[i][j] = s * c;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) P[i][j] = F[i][0] * S[0][j] + F[i][1] * S[1][j];
}

template <class T>
__device__ __forceinline__ void mm_cell_residual(const Tri &G, const T uh[3][2], T R[6]) {
    T F[2][2], P[2][2], J;
    mm_F(G, uh, F);
    mm_P(F, P, J);
    const double area = 0.5 * G.a2;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 2; ++i) R[2 * a + i] = area * (P[i][0] * G.g[a][0] + P[i][1] * G.g[a][1]);
}

template <class T>
__device__ void mm_facet_residual(const MmArgs &A, const Tri &G, int l, const T uh[3][2], const T gg[3][2], T R[6]) {
    typedef Dual<6, T> D6;
    D6 uu[3][2], F[2][2], P[2][2], J;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uu[a][c] = lift<6, T>(uh[a][c]);
            uu[a][c].d[2 * a + c] = T(1.0);
        }
    mm_F(G, uu, F);
    mm_P(F, P, J);
    const int la = (l == 0) ? 1 : 0, lb = (l == 2) ? 1 : 2;
    const double tx = G.X[lb][0] - G.X[la][0], ty = G.X[lb][1] - G.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
    double n[2] = {ty / len, -tx / len};
    if (n[0] * (G.X[la][0] - G.X[l][0]) + n[1] * (G.X[la][1] - G.X[l][1]) < 0.0) { n[0] = -n[0]; n[1] = -n[1]; }
    double h2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int b = (a + 1) % 3;
        const double dx = G.X[a][0] - G.X[b][0], dy = G.X[a][1] - G.X[b][1];
        h2 = fmax(h2, dx * dx + dy * dy);
    }
    T d[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) d[a][c] = uh[a][c] - gg[a][c];
    T wint[2], pen_a[2], pen_b[2], Pn[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        wint[c] = (0.5 * len) * (d[la][c] + d[lb][c]);                       // int (uhat - g) ds
        pen_a[c] = (len / 3.0) * d[la][c] + (len / 6.0) * d[lb][c];          // facet mass matrix
        pen_b[c] = (len / 6.0) * d[la][c] + (len / 3.0) * d[lb][c];
        Pn[c] = P[c][0].v * n[0] + P[c][1].v * n[1];
    }
    const T Jv = J.v;
    const T bh = (A.beta0 / sqrt(h2)) / (Jv * Jv * Jv);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int k = 2 * a + c;
            // (dP[phi_a e_c] n) . int (uhat - g) ds
            T r = (P[0][0].d[k] * n[0] + P[0][1].d[k] * n[1]) * wint[0] + (P[1][0].d[k] * n[0] + P[1][1].d[k] * n[1]) * wint[1];
            if (a == la) r = r - (0.5 * len) * Pn[c] + bh * pen_a[c];
            if (a == lb) r = r - (0.5 * len) * Pn[c] + bh * pen_b[c];
            R[k] = r;
        }
}

__device__ __forceinline__ bool mm_out_selected(int out_id, int tag) {
    // winding_id = [15], magnet_id = [3], steel_id = [1, 2]  (run_motor_opt.py:68-70)
    return out_id == 0 ? tag == 15 : (out_id == 1 ? tag == 3 : (tag == 1 || tag == 2));
}

template <int OP, int ENTITY>
__global__ void __launch_bounds__(64) k_motor_mm(MmArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t ne = ENTITY ? A.nfacets : A.ncells;
    if (e >= ne) return;
    const int64_t c = ENTITY ? A.fb_cell[e] : e;
    const int l = ENTITY ? A.fb_local[e] : 0;
    TriArgs TA;
    TA.coords = A.coords;
    TA.cellsT = A.cellsT;
    TA.ncells = A.ncells;
    Tri G;
    tri_load(TA, c, G);
    double uh[3][2], gg[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        uh[a][0] = A.uh[2 * G.v[a]];
        uh[a][1] = A.uh[2 * G.v[a] + 1];
        gg[a][0] = A.g ? A.g[2 * G.v[a]] : 0.0;
        gg[a][1] = A.g ? A.g[2 * G.v[a] + 1] : 0.0;
    }
    if (OP == OP_RES) {
        double R[6];
        if (ENTITY) mm_facet_residual<double>(A, G, l, uh, gg, R);
        else mm_cell_residual<double>(G, uh, R);
#pragma unroll
        for (int k = 0; k < 6; ++k) A.out[k * ne + e] = R[k];
    } else if (OP == OP_JAC || OP == OP_DRDM) {
        typedef Dual<6> D;
        D du[3][2], dg[3][2], R[6];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                du[a][i] = D(uh[a][i]);
                dg[a][i] = D(gg[a][i]);
                if (OP == OP_JAC) du[a][i].d[2 * a + i] = 1.0;
                else dg[a][i].d[2 * a + i] = 1.0;
            }
        if (ENTITY) mm_facet_residual<D>(A, G, l, du, dg, R);
        else mm_cell_residual<D>(G, du, R);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < 6; ++k) A.out[(r * 6 + k) * ne + e] = R[r].d[k];
    } else if (!ENTITY && (OP == OP_OUT || OP == OP_OUT_DU)) {
        const bool sel = mm_out_selected(A.out_id, A.tag[c]);
        if (OP == OP_OUT) {
            double F[2][2];
            mm_F(G, uh, F);
            A.out[e] = sel ? 0.5 * G.a2 * (F[0][0] * F[1][1] - F[0][1] * F[1][0]) : 0.0;
        } else {
            typedef Dual<6> D;
            D du[3][2], F[2][2];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    du[a][i] = D(uh[a][i]);
                    du[a][i].d[2 * a + i] = 1.0;
                }
            mm_F(G, du, F);
            const D J = F[0][0] * F[1][1] - F[0][1] * F[1][0];
#pragma unroll
            for (int k = 0; k < 6; ++k) A.out[k * ne + e] = sel ? 0.5 * G.a2 * J.d[k] : 0.0;
        }
    }
}

}  // namespace femo
