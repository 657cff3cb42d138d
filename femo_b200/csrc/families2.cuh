// families2.cuh -- element kernels of family 3 (Euler-Bernoulli beam, Hermite-3 on
// an interval) and family 4 (SIMP linear elasticity, Q1 quadrilaterals, vector
// state).  Same conventions as families.cuh: one thread per cell / tagged facet,
// element tensors in registers, entry k of entity e stored at out[k*ne + e].
#pragma once
#include "common.cuh"
#include "families.cuh"

namespace femo {

// ---------------------------------------------------------------------------
// family 3: examples/beam_thickness_opt/run_thickness_opt_cantilever_beam.py
//   R = int v'' (E b t^3/12) u'' dx - f v|_{ds(100)}             (:64-79,127-131)
//   outputs: 0 compliance f u|_{ds(100)} (:84-85), 1 volume int t b L dx (:81-82)
// Hermite dofs per vertex: (value, REFERENCE derivative) -- identity push-forward of
// basix 0.5 [upstream]; u'' = (1/h^2) d2/dxi2, so Ke = EI/h^3 * int H''H'' dxi
// (degree 2 -> 2-pt Gauss, exact: the classical beam matrix with unit length).
// ---------------------------------------------------------------------------
struct BeamArgs {
    const double *coords;      // (nverts)
    const int32_t *cellsT;     // (2,ncells)
    int64_t ncells;
    const int32_t *fb_cell, *fb_local;
    int64_t nfacets;
    const double *u, *t;
    double E, width, L, f;
    int out_id;
    double *out;
};

__device__ __constant__ double c_beam_k[4][4] = {
    {12.0, 6.0, -12.0, 6.0}, {6.0, 4.0, -6.0, 2.0}, {-12.0, -6.0, 12.0, -6.0}, {6.0, 2.0, -6.0, 4.0}};

template <int OP>
__global__ void __launch_bounds__(kThreads) k_beam_cell(BeamArgs A) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= A.ncells) return;
    const int64_t ne = A.ncells;
    const int v0 = A.cellsT[c], v1 = A.cellsT[ne + c];
    const double h = A.coords[v1] - A.coords[v0];
    const double t = A.t[c];
    if (OP == OP_JAC) {
        const double s = A.E * A.width * t * t * t / 12.0 / (h * h * h);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) A.out[(a * 4 + b) * ne + c] = s * c_beam_k[a][b];
    } else if (OP == OP_RES || OP == OP_DRDM) {
        const double ue[4] = {A.u[2 * v0], A.u[2 * v0 + 1], A.u[2 * v1], A.u[2 * v1 + 1]};
        const double s = (OP == OP_RES ? A.E * A.width * t * t * t / 12.0 : A.E * A.width * 3.0 * t * t / 12.0) / (h * h * h);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) acc += c_beam_k[a][b] * ue[b];
            A.out[a * ne + c] = s * acc;
        }
    } else if (OP == OP_OUT) {          // volume
        A.out[c] = t * A.width * A.L * h;
    } else if (OP == OP_OUT_DM) {       // d volume / d t
        A.out[c] = A.width * A.L * h;
    }
}

// tagged point facets: OP_RES -> -f H_a(xi), OP_OUT -> f u(x_facet), OP_OUT_DU -> f H_a(xi)
template <int OP>
__global__ void __launch_bounds__(kThreads) k_beam_facet(BeamArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= A.nfacets) return;
    const int64_t ne = A.nfacets;
    const int64_t c = A.fb_cell[e];
    const int l = A.fb_local[e];                 // facet k of an interval is its vertex k: xi = k
    if (OP == OP_OUT) {
        const int v = A.cellsT[l * A.ncells + c];
        A.out[e] = A.f * A.u[2 * v];
    } else {
        const double s = (OP == OP_RES) ? -A.f : A.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) A.out[a * ne + e] = (a == 2 * l) ? s : 0.0;
    }
}

// ---------------------------------------------------------------------------
// family 4: examples/beam_topo_opt/run_topo_opt_cantilever_beam.py
//   R = int sigma(u):eps(v) dx - int_{ds(100)} f.v ds                 (:62-77)
//   E = rho^p (SIMP, p = 3), nu = 0.3, lambda = E nu/((1+nu)(1-2nu)), mu = E/(2(1+nu))
//   cells: degree 2 -> 2x2 Gauss; traction facets: f.v is linear -> closed form len/2 per vertex
//   outputs: 0 avg_density int rho/|Omega| dx (:79-83), 1 compliance int_{ds(100)} u.f ds (:85-86)
// Local dof = 2*a + comp, Q1 vertices in tensor order.
// ---------------------------------------------------------------------------
struct QuadArgs {
    const double *coords;      // (nverts,2)
    const int32_t *cellsT;     // (4,ncells)
    int64_t ncells;
    const int32_t *fb_cell, *fb_local;
    int64_t nfacets;
    const double *u, *rho;
    double nu, fx, fy, penal, volume;
    int out_id;
    double *out;
};

struct QuadGeom {
    int v[4];
    double X[4][2];
    double G[4][4][2];   // physical gradients at the 4 Gauss points
    double w[4];         // weight * |det J|
};

__device__ __forceinline__ void quad_load(const QuadArgs &A, int64_t c, QuadGeom &Q, bool need_grads) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        Q.v[a] = A.cellsT[a * A.ncells + c];
        const double2 xy = __ldg(reinterpret_cast<const double2 *>(A.coords) + Q.v[a]);
        Q.X[a][0] = xy.x;
        Q.X[a][1] = xy.y;
    }
    const double g0 = 0.5 - 0.28867513459481287, g1 = 0.5 + 0.28867513459481287;   // (1 -/+ 1/sqrt 3)/2
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double xi = (q >> 1) ? g1 : g0, et = (q & 1) ? g1 : g0;   // same point order as the oracle's square rule
        const double dNx[4] = {-(1.0 - et), (1.0 - et), -et, et};
        const double dNy[4] = {-(1.0 - xi), -xi, (1.0 - xi), xi};
        double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            J00 += Q.X[a][0] * dNx[a];
            J01 += Q.X[a][0] * dNy[a];
            J10 += Q.X[a][1] * dNx[a];
            J11 += Q.X[a][1] * dNy[a];
        }
        const double det = J00 * J11 - J01 * J10;
        Q.w[q] = 0.25 * fabs(det);
        if (need_grads) {
            const double id = 1.0 / det;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                Q.G[q][a][0] = (J11 * dNx[a] - J10 * dNy[a]) * id;
                Q.G[q][a][1] = (-J01 * dNx[a] + J00 * dNy[a]) * id;
            }
        }
    }
}

// unit-modulus stiffness entry K[(a,ci),(b,cj)] summed over the 2x2 Gauss points
__device__ __forceinline__ double quad_khat(const QuadGeom &Q, double lam, double mu, int a, int ci, int b, int cj) {
    double k = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double t = lam * Q.G[q][a][ci] * Q.G[q][b][cj] + mu * Q.G[q][a][cj] * Q.G[q][b][ci];
        if (ci == cj) t += mu * (Q.G[q][a][0] * Q.G[q][b][0] + Q.G[q][a][1] * Q.G[q][b][1]);
        k += Q.w[q] * t;
    }
    return k;
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_simp_q1_cell(QuadArgs A) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= A.ncells) return;
    const int64_t ne = A.ncells;
    QuadGeom Q;
    quad_load(A, c, Q, OP == OP_RES || OP == OP_JAC || OP == OP_DRDM);
    const double rho = A.rho[c];
    if (OP == OP_OUT || OP == OP_OUT_DM) {
        const double area = Q.w[0] + Q.w[1] + Q.w[2] + Q.w[3];
        A.out[c] = (OP == OP_OUT ? rho : 1.0) * area / A.volume;
        return;
    }
    const double lam = A.nu / ((1.0 + A.nu) * (1.0 - 2.0 * A.nu)), mu = 1.0 / (2.0 * (1.0 + A.nu));
    if (OP == OP_JAC) {
        const double E = pow(rho, A.penal);
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int s = 0; s < 8; ++s)
                A.out[(r * 8 + s) * ne + c] = E * quad_khat(Q, lam, mu, r >> 1, r & 1, s >> 1, s & 1);
    } else {  // OP_RES: rho^p K u ; OP_DRDM: p rho^(p-1) K u
        double ue[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            ue[2 * a] = A.u[2 * Q.v[a]];
            ue[2 * a + 1] = A.u[2 * Q.v[a] + 1];
        }
        const double E = (OP == OP_RES) ? pow(rho, A.penal) : A.penal * pow(rho, A.penal - 1.0);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int s = 0; s < 8; ++s) acc += quad_khat(Q, lam, mu, r >> 1, r & 1, s >> 1, s & 1) * ue[s];
            A.out[r * ne + c] = E * acc;
        }
    }
}

// traction facets: OP_RES -> -int f.v ds, OP_OUT -> int u.f ds, OP_OUT_DU -> int f.v ds
template <int OP>
__global__ void __launch_bounds__(kThreads) k_simp_q1_facet(QuadArgs A) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= A.nfacets) return;
    const int64_t ne = A.nfacets;
    const int64_t c = A.fb_cell[e];
    const int l = A.fb_local[e];
    // basix quadrilateral facets: (v0,v1) (v0,v2) (v1,v3) (v2,v3)
    const int la = (l == 0 || l == 1) ? 0 : (l == 2 ? 1 : 2);
    const int lb = (l == 0) ? 1 : (l == 1 ? 2 : 3);
    const int va = A.cellsT[la * A.ncells + c], vb = A.cellsT[lb * A.ncells + c];
    const double2 pa = __ldg(reinterpret_cast<const double2 *>(A.coords) + va);
    const double2 pb = __ldg(reinterpret_cast<const double2 *>(A.coords) + vb);
    const double half = 0.5 * sqrt((pb.x - pa.x) * (pb.x - pa.x) + (pb.y - pa.y) * (pb.y - pa.y));
    if (OP == OP_OUT) {
        A.out[e] = half * (A.fx * (A.u[2 * va] + A.u[2 * vb]) + A.fy * (A.u[2 * va + 1] + A.u[2 * vb + 1]));
    } else {
        const double s = (OP == OP_RES) ? -half : half;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double on = (a == la || a == lb) ? s : 0.0;
            A.out[(2 * a) * ne + e] = on * A.fx;
            A.out[(2 * a + 1) * ne + e] = on * A.fy;
        }
    }
}

}  // namespace femo
