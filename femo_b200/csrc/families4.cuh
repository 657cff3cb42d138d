// families4.cuh -- family 7: nonlinear magnetostatics of the motor example on a moving mesh
// (examples/em_motor_opt/motor_pde.py: RelativePermeability :12-35, JS :46-87, pdeResEM :90-130,
// B_power_form :186-197; ALE kinematics gradx / J / F of femo/fea/utils_dolfinx.py:34-66).
//
// State u = A_z (scalar P1), input uhat = mesh displacement (vector P1, 2 comps interleaved).
// All integrands are cellwise constant except v*J, so the cell tensor is closed-form; the exterior
// facet (symmetric Nitsche with the Nanson-transformed normal) is integrated with 2-pt Gauss.
// The element residual is written ONCE as a template on the scalar type; dR/du (3x3), dR/duhat (3x6)
// and the functional gradients come from forward-mode dual numbers seeded on the element dofs, i.e.
// the exact Gateaux derivative UFL would form (utils_dolfinx.py:313-314), not a finite difference.
#pragma once
#include "dual.cuh"
#include "families.cuh"

namespace femo {

// parameter slots of the family (femo_problem::params)
enum EmParam { EM_MU0 = 0, EM_HC, EM_IQ, EM_ANGLE, EM_P, EM_S, EM_JS_SCALE, EM_BETA, EM_X1, EM_X2, EM_LIN = 10, EM_CUB = 12,
               EM_EXP = 16, EM_EXPO0 = 19, EM_EXPO1 = 20, EM_NPARAM = 21 };

struct EmArgs {
    const double *coords;
    const int32_t *cellsT;
    int64_t ncells;
    const int32_t *fb_cell, *fb_local;
    int64_t nfacets;
    const int32_t *tag;
    const double *u, *uh;
    double prm[EM_NPARAM];
    int out_id;
    double *out;
};

// mu_r of the laminated steel (subdomains 1, 2): linear / cubic bridge / exponential saturation
template <class T>
__device__ __forceinline__ T em_mur_steel(const EmArgs &A, const T &nb) {
    const double x = valof(nb);
    if (x < A.prm[EM_X1]) return A.prm[EM_LIN] * nb + A.prm[EM_LIN + 1];
    if (x < A.prm[EM_X2])
        return ((A.prm[EM_CUB] * nb + A.prm[EM_CUB + 1]) * nb + A.prm[EM_CUB + 2]) * nb + A.prm[EM_CUB + 3];
    return A.prm[EM_EXP] * dexp(A.prm[EM_EXP + 1] * nb + A.prm[EM_EXP + 2]) + 1.0;
}

template <class T>
struct EmKin {
    T gx[2];        // gradx(u) = grad(u) F^-1
    T gv[3][2];     // gradx(phi_a)
    T det;          // J = det F
    T Fi[2][2];     // F^-1
};

template <class T>
__device__ __forceinline__ void em_kinematics(const Tri &G, const T u[3], const T uh[3][2], EmKin<T> &K) {
    T gu[2] = {u[0] * G.g[0][0] + u[1] * G.g[1][0] + u[2] * G.g[2][0], u[0] * G.g[0][1] + u[1] * G.g[1][1] + u[2] * G.g[2][1]};
    T F[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = uh[0][i] * G.g[0][j] + uh[1][i] * G.g[1][j] + uh[2][i] * G.g[2][j] + (i == j ? 1.0 : 0.0);
    K.det = F[0][0] * F[1][1] - F[0][1] * F[1][0];
    const T id = 1.0 / K.det;
    K.Fi[0][0] = F[1][1] * id;  K.Fi[0][1] = -F[0][1] * id;
    K.Fi[1][0] = -F[1][0] * id; K.Fi[1][1] = F[0][0] * id;
#pragma unroll
    for (int j = 0; j < 2; ++j) K.gx[j] = gu[0] * K.Fi[0][j] + gu[1] * K.Fi[1][j];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j) K.gv[a][j] = G.g[a][0] * K.Fi[0][j] + G.g[a][1] * K.Fi[1][j];
}

// cell residual: A nu (gradx u . gradx v_a) J  -  js_scale (magnet + winding sources)
template <class T>
__device__ __forceinline__ void em_cell_residual(const EmArgs &A, const Tri &G, int tag, const T u[3], const T uh[3][2], T R[3]) {
    EmKin<T> K;
    em_kinematics(G, u, uh, K);
    const double area = 0.5 * G.a2, pi = 3.141592653589793;
    const T B2 = K.gx[0] * K.gx[0] + K.gx[1] * K.gx[1];
    T nu;
    if (tag == 1 || tag == 2) nu = 1.0 / (A.prm[EM_MU0] * em_mur_steel(A, dsqrt(B2 + 3e-16)));
    else nu = T(1.0 / (A.prm[EM_MU0] * ((tag >= 3 && tag <= 14) ? 1.05 : 1.0)));
    const int p = (int)A.prm[EM_P], s = (int)A.prm[EM_S];
    double Hx = 0.0, Hy = 0.0, cur = 0.0;
    if (tag >= 3 && tag < 3 + p) {
        const int i = tag - 3;
        const double fa = 2.0 * pi / p / 2.0 + i * (2.0 * pi / p) + A.prm[EM_ANGLE] * 2.0 / p;
        const double sg = (i & 1) ? -1.0 : 1.0;
        Hx = sg * A.prm[EM_HC] * cos(fa);
        Hy = sg * A.prm[EM_HC] * sin(fa);
    } else if (tag >= 15 && tag < 15 + s) {
        const int w = tag - 15, pole = w / 3, k = w % 3;
        const double ang = A.prm[EM_ANGLE] + (k == 0 ? -2.0 * pi / 3.0 : (k == 2 ? 2.0 * pi / 3.0 : 0.0));
        const double amp = A.prm[EM_IQ] * sin(ang) + 3e-16;                     // JB, JA, JC (+ DOLFIN_EPS)
        const double sg = ((pole + (k == 1 ? 0 : 1)) & 1) ? -1.0 : 1.0;
        cur = amp * sg;
    }
    const double js = A.prm[EM_JS_SCALE];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T r = area * nu * K.det * (K.gx[0] * K.gv[a][0] + K.gx[1] * K.gv[a][1]);
        if (Hx != 0.0 || Hy != 0.0) r = r - (js * area) * K.det * (Hx * K.gv[a][1] - Hy * K.gv[a][0]);
        if (cur != 0.0) r = r - (js * cur * area / 3.0) * K.det;
        R[a] = r;
    }
}

// exterior facet (one-sided): 2 * [ coeff (-(gradx u.nN) v - (gradx v.nN) u) + beta/h coeff |nN| v u ] ds, g = 0,
// nN = J F^-T n (Nanson), coeff = 1/(mu0 mu_r,steel(|B|)) for both boundary components
template <class T>
__device__ __forceinline__ void em_facet_residual(const EmArgs &A, const Tri &G, int l, const T u[3], const T uh[3][2], T R[3]) {
    EmKin<T> K;
    em_kinematics(G, u, uh, K);
    const int la = (l == 0) ? 1 : 0, lb = (l == 2) ? 1 : 2;
    const double tx = G.X[lb][0] - G.X[la][0], ty = G.X[lb][1] - G.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (G.X[la][0] - G.X[l][0]) + ny * (G.X[la][1] - G.X[l][1]) < 0.0) { nx = -nx; ny = -ny; }
    double h2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int b = (a + 1) % 3;
        const double dx = G.X[a][0] - G.X[b][0], dy = G.X[a][1] - G.X[b][1];
        h2 = fmax(h2, dx * dx + dy * dy);
    }
    const double bh = A.prm[EM_BETA] / sqrt(h2);
    const T nN[2] = {K.det * (K.Fi[0][0] * nx + K.Fi[1][0] * ny), K.det * (K.Fi[0][1] * nx + K.Fi[1][1] * ny)};
    const T nrmN = dsqrt(nN[0] * nN[0] + nN[1] * nN[1]);
    const T B2 = K.gx[0] * K.gx[0] + K.gx[1] * K.gx[1];
    const T coeff = 1.0 / (A.prm[EM_MU0] * em_mur_steel(A, dsqrt(B2 + 3e-16)));
    const T gxn = K.gx[0] * nN[0] + K.gx[1] * nN[1];
    T gvn[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) gvn[a] = K.gv[a][0] * nN[0] + K.gv[a][1] * nN[1];
#pragma unroll
    for (int a = 0; a < 3; ++a) R[a] = T(0.0);
    const double gs[2] = {0.5 - 0.28867513459481287, 0.5 + 0.28867513459481287};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        double ph[3] = {0.0, 0.0, 0.0};
        ph[la] = 1.0 - gs[q];
        ph[lb] = gs[q];
        const T uq = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2];
        const double w = 0.5 * len * 2.0;                                        // Gauss weight 1/2, both components
#pragma unroll
        for (int a = 0; a < 3; ++a)
            R[a] = R[a] + w * (coeff * ((-1.0 * ph[a]) * gxn - gvn[a] * uq) + (bh * ph[a]) * coeff * nrmN * uq);
    }
}

// int |B|^n J dx over the steel subdomains (B_power_form)
template <class T>
__device__ __forceinline__ T em_cell_output(const EmArgs &A, const Tri &G, int tag, int k, const T u[3], const T uh[3][2]) {
    if (tag != 1 && tag != 2) return T(0.0);
    EmKin<T> K;
    em_kinematics(G, u, uh, K);
    const T Bm = dsqrt(K.gx[0] * K.gx[0] + K.gx[1] * K.gx[1]);
    return (0.5 * G.a2) * dpow(Bm, A.prm[k == 0 ? EM_EXPO0 : EM_EXPO1]) * K.det;
}

__device__ __forceinline__ void em_load(const EmArgs &A, int64_t c, Tri &G, double u[3], double uh[3][2]) {
    TriArgs TA;
    TA.coords = A.coords;
    TA.cellsT = A.cellsT;
    TA.ncells = A.ncells;
    tri_load(TA, c, G);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        u[a] = A.u[G.v[a]];
        uh[a][0] = A.uh[2 * G.v[a]];
        uh[a][1] = A.uh[2 * G.v[a] + 1];
    }
}

// shared driver: ENTITY 0 = cells, 1 = exterior facets
template <int OP, int ENTITY>
__global__ void __launch_bounds__(128) k_motor_em(EmArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t ne = ENTITY ? A.nfacets : A.ncells;
    if (e >= ne) return;
    const int64_t c = ENTITY ? A.fb_cell[e] : e;
    const int l = ENTITY ? A.fb_local[e] : 0;
    const int tag = A.tag[c];
    Tri G;
    double u[3], uh[3][2];
    em_load(A, c, G, u, uh);
    if (OP == OP_RES) {
        double R[3];
        if (ENTITY) em_facet_residual<double>(A, G, l, u, uh, R);
        else em_cell_residual<double>(A, G, tag, u, uh, R);
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + e] = R[a];
    } else if (OP == OP_JAC) {
        Dual<3> du[3], duh[3][2], R[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            du[a] = Dual<3>(u[a]);
            du[a].d[a] = 1.0;
            duh[a][0] = Dual<3>(uh[a][0]);
            duh[a][1] = Dual<3>(uh[a][1]);
        }
        if (ENTITY) em_facet_residual<Dual<3>>(A, G, l, du, duh, R);
        else em_cell_residual<Dual<3>>(A, G, tag, du, duh, R);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) A.out[(a * 3 + b) * ne + e] = R[a].d[b];
    } else if (OP == OP_DRDM) {
        Dual<6> du[3], duh[3][2], R[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            du[a] = Dual<6>(u[a]);
            duh[a][0] = Dual<6>(uh[a][0]);
            duh[a][0].d[2 * a] = 1.0;
            duh[a][1] = Dual<6>(uh[a][1]);
            duh[a][1].d[2 * a + 1] = 1.0;
        }
        if (ENTITY) em_facet_residual<Dual<6>>(A, G, l, du, duh, R);
        else em_cell_residual<Dual<6>>(A, G, tag, du, duh, R);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) A.out[(a * 6 + b) * ne + e] = R[a].d[b];
    } else if (!ENTITY && OP == OP_OUT) {
        A.out[e] = em_cell_output<double>(A, G, tag, A.out_id, u, uh);
    } else if (!ENTITY && OP == OP_OUT_DU) {
        Dual<3> du[3], duh[3][2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            du[a] = Dual<3>(u[a]);
            du[a].d[a] = 1.0;
            duh[a][0] = Dual<3>(uh[a][0]);
            duh[a][1] = Dual<3>(uh[a][1]);
        }
        const Dual<3> J = em_cell_output<Dual<3>>(A, G, tag, A.out_id, du, duh);
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + e] = J.d[a];
    } else if (!ENTITY && OP == OP_OUT_DM) {
        Dual<6> du[3], duh[3][2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            du[a] = Dual<6>(u[a]);
            duh[a][0] = Dual<6>(uh[a][0]);
            duh[a][0].d[2 * a] = 1.0;
            duh[a][1] = Dual<6>(uh[a][1]);
            duh[a][1].d[2 * a + 1] = 1.0;
        }
        const Dual<6> J = em_cell_output<Dual<6>>(A, G, tag, A.out_id, du, duh);
#pragma unroll
        for (int b = 0; b < 6; ++b) A.out[b * ne + e] = J.d[b];
    }
}

}  // namespace femo
