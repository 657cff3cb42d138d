// families8.cuh -- family 10: Reissner-Mindlin plate, the flat-mid-surface case of the reference's shell examples
// (examples/test_shell_m3l/shell_pde.py:219-311: ShellElement "CG2CG1", ElasticModel.weakFormResidual with
// penalty=True, linear_problem=True; outputs compliance :281-282, mass :287-288, elastic_energy :290-293).  The forms
// of the reference live in the un-vendored package shell_analysis_fenicsx; the published Reissner-Mindlin
// formulation is restated in oracle/rm_plate.py (PARITY UNPINNED, see there) and mirrored here term by term:
//
//   U = 1/2 int D(t) [(1-nu) kappa:kappa + nu tr(kappa)^2] dx  +  1/2 int ks G t |grad w - theta|^2 dx_reduced
//     + 1/2 pen int_{Gamma_c} (w^2 + theta.theta) ds ,      R = dU - int f v dx ,   D = E t^3 / (12 (1 - nu^2))
//
// w in CG2, theta in CG1^2, thickness t and load f in CG1.  Local dofs (12): w at vertices 0..2, w at the midpoints of
// the edges opposite vertices 0..2, (theta_x, theta_y) per vertex; global [w vertices | w edges | theta interleaved].
// Quadrature: bending weight and load 6-point degree-4 rule, shear 3-point degree-2 rule (reduced), penalty 5-pt Gauss.
// The element operator is applied as y = K(t) x (or dK/dt_a x) without forming K; the Jacobian columns are K e_b.
// One thread per cell / clamped facet, SoA scratch planes, sorted segmented reduction afterwards (engine.cu).
#pragma once
#include "common.cuh"
#include "families.cuh"

namespace femo {

struct RmArgs {
    TriArgs T;                  // geometry, facets, u = state, out
    const int32_t *edgesT;      // (3,ncells) SoA: edge opposite local vertex i
    int64_t nverts, nedges;
    const double *t, *f;        // thickness, load (CG1)
    double E, nu, pen, rho;
    int out_id, slot;
};

__constant__ double c_tri3[3][3];     // degree-2 rule: (1/6,1/6), (1/6,2/3), (2/3,1/6), weights 1/6

__device__ __forceinline__ void rm_load(const RmArgs &A, const Tri &T, int64_t c, const double *v, double x[12]) {
    const int64_t th = A.nverts + A.nedges;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        x[a] = v[T.v[a]];
        x[3 + a] = v[A.nverts + A.edgesT[a * A.T.ncells + c]];
        x[6 + 2 * a] = v[th + 2 * (int64_t)T.v[a]];
        x[7 + 2 * a] = v[th + 2 * (int64_t)T.v[a] + 1];
    }
}

// y = K(t) x  (a_t < 0)  or  y = dK/dt_{a_t} x
__device__ __forceinline__ void rm_apply(const Tri &T, const double te[3], double E, double nu, const double x[12], double y[12],
                                         int a_t) {
    const double Db = E / (12.0 * (1.0 - nu * nu)), Gs = (5.0 / 6.0) * E / (2.0 * (1.0 + nu));
#pragma unroll
    for (int i = 0; i < 12; ++i) y[i] = 0.0;
    // bending: kappa is constant on the cell, only the scalar weight int D(t) dx is integrated
    double Wb = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double l[3] = {1.0 - c_tri6[q][0] - c_tri6[q][1], c_tri6[q][0], c_tri6[q][1]};
        const double tq = te[0] * l[0] + te[1] * l[1] + te[2] * l[2];
        const double wq = c_tri6[q][2] * T.a2;
        Wb += a_t < 0 ? wq * Db * tq * tq * tq : wq * Db * 3.0 * tq * tq * l[a_t];
    }
    double kxx = 0.0, kyy = 0.0, kxy = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        kxx += T.g[a][0] * x[6 + 2 * a];
        kyy += T.g[a][1] * x[7 + 2 * a];
        kxy += T.g[a][1] * x[6 + 2 * a] + T.g[a][0] * x[7 + 2 * a];
    }
    const double mxx = kxx + nu * kyy, myy = nu * kxx + kyy, mxy = 0.5 * (1.0 - nu) * kxy;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        y[6 + 2 * a] += Wb * (T.g[a][0] * mxx + T.g[a][1] * mxy);
        y[7 + 2 * a] += Wb * (T.g[a][1] * myy + T.g[a][0] * mxy);
    }
    // transverse shear, reduced rule
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double l[3] = {1.0 - c_tri3[q][0] - c_tri3[q][1], c_tri3[q][0], c_tri3[q][1]};
        double gp[6][2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            const double d = 4.0 * l[i] - 1.0;
            gp[i][0] = d * T.g[i][0];
            gp[i][1] = d * T.g[i][1];
            gp[3 + i][0] = 4.0 * (l[j] * T.g[k][0] + l[k] * T.g[j][0]);
            gp[3 + i][1] = 4.0 * (l[j] * T.g[k][1] + l[k] * T.g[j][1]);
        }
        double gx = 0.0, gy = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            gx += gp[i][0] * x[i];
            gy += gp[i][1] * x[i];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            gx -= l[a] * x[6 + 2 * a];
            gy -= l[a] * x[7 + 2 * a];
        }
        const double tq = te[0] * l[0] + te[1] * l[1] + te[2] * l[2];
        const double s = c_tri3[q][2] * T.a2 * Gs * (a_t < 0 ? tq : l[a_t]);
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] += s * (gp[i][0] * gx + gp[i][1] * gy);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            y[6 + 2 * a] -= s * l[a] * gx;
            y[7 + 2 * a] -= s * l[a] * gy;
        }
    }
}

// L[i][b] = int phi^w_i phi^f_b dx (degree-4 rule)
__device__ __forceinline__ void rm_load_matrix(const Tri &T, double L[6][3]) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int b = 0; b < 3; ++b) L[i][b] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double l[3] = {1.0 - c_tri6[q][0] - c_tri6[q][1], c_tri6[q][0], c_tri6[q][1]};
        double ph[6];
        p2_values(l, ph);
        const double wq = c_tri6[q][2] * T.a2;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int b = 0; b < 3; ++b) L[i][b] += wq * ph[i] * l[b];
    }
}

template <int OP>
__global__ void __launch_bounds__(128) k_rm_plate_cell(RmArgs A) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= A.T.ncells) return;
    const int64_t ne = A.T.ncells;
    Tri T;
    tri_load(A.T, c, T);
    double *out = A.T.out;
    double te[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) te[a] = A.t[T.v[a]];
    if (OP == OP_JAC) {
        for (int b = 0; b < 12; ++b) {
            double x[12], y[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = (i == b) ? 1.0 : 0.0;
            rm_apply(T, te, A.E, A.nu, x, y, -1);
#pragma unroll
            for (int a = 0; a < 12; ++a) out[(a * 12 + b) * ne + c] = y[a];
        }
        return;
    }
    if (OP == OP_DRDM && A.slot == 1) {       // load: -int phi^w_i phi^f_b dx
        double L[6][3];
        rm_load_matrix(T, L);
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int b = 0; b < 3; ++b) out[(i * 3 + b) * ne + c] = i < 6 ? -L[i][b] : 0.0;
        return;
    }
    if ((OP == OP_OUT || OP == OP_OUT_DM) && A.out_id == 1) {       // mass rho int t dx
        if (OP == OP_OUT) out[c] = A.rho * T.a2 / 6.0 * (te[0] + te[1] + te[2]);
        else
#pragma unroll
            for (int a = 0; a < 3; ++a) out[a * ne + c] = A.rho * T.a2 / 6.0;
        return;
    }
    double u[12];
    rm_load(A, T, c, A.T.u, u);
    if (OP == OP_RES) {
        double y[12], L[6][3];
        rm_apply(T, te, A.E, A.nu, u, y, -1);
        rm_load_matrix(T, L);
        double fe[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) fe[b] = A.f[T.v[b]];
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] -= L[i][0] * fe[0] + L[i][1] * fe[1] + L[i][2] * fe[2];
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i * ne + c] = y[i];
    } else if (OP == OP_DRDM) {               // thickness: d(K(t) u)/dt_a
        for (int a = 0; a < 3; ++a) {
            double y[12];
            rm_apply(T, te, A.E, A.nu, u, y, a);
#pragma unroll
            for (int i = 0; i < 12; ++i) out[(i * 3 + a) * ne + c] = y[i];
        }
    } else if (OP == OP_OUT || OP == OP_OUT_DU) {
        if (A.out_id == 0) {                  // compliance 1/2 int w^2 dx
            double val = 0.0, ge[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const double l[3] = {1.0 - c_tri6[q][0] - c_tri6[q][1], c_tri6[q][0], c_tri6[q][1]};
                double ph[6];
                p2_values(l, ph);
                double wq = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) wq += u[i] * ph[i];
                const double w = c_tri6[q][2] * T.a2;
                val += 0.5 * w * wq * wq;
#pragma unroll
                for (int i = 0; i < 6; ++i) ge[i] += w * wq * ph[i];
            }
            if (OP == OP_OUT) out[c] = val;
            else
#pragma unroll
                for (int i = 0; i < 12; ++i) out[i * ne + c] = i < 6 ? ge[i] : 0.0;
        } else {                              // elastic energy 1/2 u K u
            double y[12];
            rm_apply(T, te, A.E, A.nu, u, y, -1);
            if (OP == OP_OUT) {
                double e = 0.0;
#pragma unroll
                for (int i = 0; i < 12; ++i) e += u[i] * y[i];
                out[c] = 0.5 * e;
            } else
#pragma unroll
                for (int i = 0; i < 12; ++i) out[i * ne + c] = y[i];
        }
    } else if (OP == OP_OUT_DM) {             // d(elastic energy)/dt_a = 1/2 u dK/dt_a u   (compliance: 0)
        for (int a = 0; a < 3; ++a) {
            double e = 0.0;
            if (A.out_id == 2) {
                double y[12];
                rm_apply(T, te, A.E, A.nu, u, y, a);
#pragma unroll
                for (int i = 0; i < 12; ++i) e += u[i] * y[i];
            }
            out[a * ne + c] = 0.5 * e;
        }
    }
}

// penalty clamp on the facets of block 2: y = K_f x with K_f = pen int N^T N ds, N = (w; theta_x; theta_y)
__device__ __forceinline__ void rm_facet_apply(const Tri &T, int l_, double pen, const double x[12], double y[12]) {
    const int la = (l_ == 0) ? 1 : 0, lb = (l_ == 2) ? 1 : 2;
    const double tx = T.X[lb][0] - T.X[la][0], ty = T.X[lb][1] - T.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
#pragma unroll
    for (int i = 0; i < 12; ++i) y[i] = 0.0;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double s = c_gl5[q][0], wq = c_gl5[q][1] * len * pen;
        double l[3] = {0.0, 0.0, 0.0};
        l[la] = 1.0 - s;
        l[lb] = s;
        double ph[6];
        p2_values(l, ph);
        double wv = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) wv += ph[i] * x[i];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            t0 += l[a] * x[6 + 2 * a];
            t1 += l[a] * x[7 + 2 * a];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] += wq * ph[i] * wv;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            y[6 + 2 * a] += wq * l[a] * t0;
            y[7 + 2 * a] += wq * l[a] * t1;
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(128) k_rm_plate_facet(RmArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= A.T.nfacets) return;
    const int64_t ne = A.T.nfacets;
    const int64_t c = A.T.bf_cell[e];
    const int l = A.T.bf_local[e];
    Tri T;
    tri_load(A.T, c, T);
    double *out = A.T.out;
    if (OP == OP_RES) {
        double u[12], y[12];
        rm_load(A, T, c, A.T.u, u);
        rm_facet_apply(T, l, A.pen, u, y);
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i * ne + e] = y[i];
    } else {
        for (int b = 0; b < 12; ++b) {
            double x[12], y[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = (i == b) ? 1.0 : 0.0;
            rm_facet_apply(T, l, A.pen, x, y);
#pragma unroll
            for (int a = 0; a < 12; ++a) out[(a * 12 + b) * ne + e] = y[a];
        }
    }
}

}  // namespace femo
