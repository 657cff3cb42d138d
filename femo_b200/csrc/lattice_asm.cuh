// lattice_asm.cuh -- node-centric Jacobian assembly on right-diagonal triangle lattices.
//
// The general path (families.cuh + the sorted segmented reduction of engine.cu) writes every element matrix to
// scratch planes and gathers them through an index map: 8*9 bytes per cell out, the same back in, plus 4 bytes of map
// per contribution -- 4.8 GB for the 16M-dof benchmark Jacobian (1.45 ms + 0.62 ms).  On a lattice the cells around a
// node and the node's local index in each of them are index arithmetic, so one thread per node evaluates the ROWS it
// owns of its <= 6 incident element matrices (and of the Nitsche facet matrices of boundary triangles) and writes the
// <= 7 CSR values of its row directly: no scratch, no map, u and f read once through the cache.  The total flop count
// equals the cell-centric kernel's (every element row is computed exactly once).  Values agree with the general path to
// round-off (different, but fixed, summation order => still bit-reproducible run to run); both are tested against the
// oracle at 1e-12.  Replaces assemble_matrix(form(J)) of the reference (femo/fea/utils_dolfinx.py:181-202) for
// examples/nonlinear_poisson_opt/run_nonlinear_poisson_opt.py:88-116.
#pragma once
#include "families.cuh"

namespace femo {

struct LatJacArgs {
    TriArgs T;
    int nx, ny;                 // local lattice (cells)
    int ext_bottom, ext_top;    // whether local rows 0 / ny are true domain boundaries (slabs: only on the end ranks)
    const int32_t *rowptr, *col;
    const uint8_t *bcflag;
    const double *bc_diag;
    double *out, *out_bc;
};

// row a of the cell matrix  int grad(phi_a).grad(phi_b) + 3 u^2 phi_a phi_b dx  (degree-4 rule, as k_nlpoisson_p1_cell<OP_JAC>)
__device__ __forceinline__ void nlp_cell_jac_row(const Tri &T, const double u[3], int a, double row[3]) {
#pragma unroll
    for (int b = 0; b < 3; ++b) row[b] = 0.5 * T.a2 * (T.g[a][0] * T.g[b][0] + T.g[a][1] * T.g[b][1]);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double ph[3] = {1.0 - c_tri6[q][0] - c_tri6[q][1], c_tri6[q][0], c_tri6[q][1]};
        const double uq = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2];
        const double s = c_tri6[q][2] * T.a2 * 3.0 * uq * uq * ph[a];
#pragma unroll
        for (int b = 0; b < 3; ++b) row[b] += s * ph[b];
    }
}

// row a of the symmetric-Nitsche matrix of exterior facet l (opposite local vertex l), as k_nlpoisson_p1_facet<OP_JAC>
__device__ __noinline__ void nlp_facet_jac_row(const Tri &T, int l, double beta, int a, double row[3]) {
    const int la = (l == 0) ? 1 : 0, lb = (l == 2) ? 1 : 2;
    const double tx = T.X[lb][0] - T.X[la][0], ty = T.X[lb][1] - T.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (T.X[la][0] - T.X[l][0]) + ny * (T.X[la][1] - T.X[l][1]) < 0.0) { nx = -nx; ny = -ny; }
    double h2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d = (c + 1) % 3;
        const double dx = T.X[c][0] - T.X[d][0], dy = T.X[c][1] - T.X[d][1];
        h2 = fmax(h2, dx * dx + dy * dy);
    }
    const double bh = beta / sqrt(h2);
    double gn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gn[c] = T.g[c][0] * nx + T.g[c][1] * ny;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double s = c_gl5[q][0], w = c_gl5[q][1] * len;
        double ph[3] = {0.0, 0.0, 0.0};
        ph[la] = 1.0 - s;
        ph[lb] = s;
#pragma unroll
        for (int b = 0; b < 3; ++b) row[b] += w * (-ph[a] * gn[b] - gn[a] * ph[b] + bh * ph[a] * ph[b]);
    }
}

// Uniform lattice: all lower triangles are congruent and so are all upper ones, so the P1 gradients and |det J| are
// two constant sets (computed on the host from the lattice spacing, passed as kernel arguments) -- no connectivity or
// coordinate loads and no divisions in the interior.  The 7 state values of the node's stencil are loaded once and
// every incident triangle picks its three by compile-time slot; element rows land in compile-time slots as well.
// Boundary triangles add their Nitsche facet rows through the general geometry path (tri_load), O(boundary) work.
struct LatGeom {
    double gl[3][2], gu[3][2];   // gradients of the lower [v0,v1,v3] / upper [v0,v2,v3] triangle
    double a2;                   // |det J| = hx * hy
};

// Exact integrals of barycentric monomials on a triangle: int l0^m0 l1^m1 l2^m2 dx = |det J| m0! m1! m2! / (m0+m1+m2+2)!.
// The degree-4 quadrature rule of the cell kernels integrates the quartic integrands u^2 phi_a phi_b and u^3 phi_a
// exactly, so the closed forms below give the same element rows up to round-off at a third of the arithmetic.
// kQuartJac[a][b][p]: 720 / |det J| * (coefficient of the p-th product of {u0u0, u1u1, u2u2, u0u1, u0u2, u1u2} in
// int u^2 phi_a phi_b dx); kQuartRes[a][m]: the same for the m-th cubic monomial u_i u_j u_k (i <= j <= k, lexicographic)
// in int u^3 phi_a dx.  Indexed with compile-time constants only (fully unrolled loops): folded into immediates.
__device__ constexpr double kQuartJac[3][3][6] = {{{24.0, 4.0, 4.0, 12.0, 12.0, 4.0}, {6.0, 6.0, 2.0, 8.0, 4.0, 4.0}, {6.0, 2.0, 6.0, 4.0, 8.0, 4.0}}, {{6.0, 6.0, 2.0, 8.0, 4.0, 4.0}, {4.0, 24.0, 4.0, 12.0, 4.0, 12.0}, {2.0, 6.0, 6.0, 4.0, 4.0, 8.0}}, {{6.0, 2.0, 6.0, 4.0, 8.0, 4.0}, {2.0, 6.0, 6.0, 4.0, 4.0, 8.0}, {4.0, 4.0, 24.0, 4.0, 12.0, 12.0}}};
__device__ constexpr double kQuartRes[3][10] = {{24.0, 18.0, 18.0, 12.0, 12.0, 12.0, 6.0, 6.0, 6.0, 6.0}, {6.0, 12.0, 6.0, 18.0, 12.0, 6.0, 24.0, 18.0, 12.0, 6.0}, {6.0, 6.0, 12.0, 6.0, 12.0, 18.0, 6.0, 12.0, 18.0, 24.0}};
template <int a>
__device__ __forceinline__ void nlp_cell_jac_row_const(const double g[3][2], double a2, const double u[3], double row[3]) {
    const double u2[6] = {u[0] * u[0], u[1] * u[1], u[2] * u[2], u[0] * u[1], u[0] * u[2], u[1] * u[2]};
    const double c = 3.0 * a2 / 720.0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        double m = 0.0;
#pragma unroll
        for (int p = 0; p < 6; ++p) m = fma(kQuartJac[a][b][p], u2[p], m);
        row[b] = 0.5 * a2 * (g[a][0] * g[b][0] + g[a][1] * g[b][1]) + c * m;
    }
}

// the six triangles around a lattice node: cell offset (dx, dy), upper?, local index of the node in the triangle, and
// the stencil slots {-w-1, -w, -1, 0, 1, w, w+1} of the triangle's three vertices
__device__ constexpr int kNodeTri[6][7] = {{0, 0, 0, 0, 3, 4, 6}, {0, 0, 1, 0, 3, 5, 6}, {-1, 0, 0, 1, 2, 3, 5},
                                           {0, -1, 1, 1, 1, 3, 4}, {-1, -1, 0, 2, 0, 1, 3}, {-1, -1, 1, 2, 0, 2, 3}};

// optional extra outputs of the node-centric Jacobian: the row directly in the layouts the multigrid-preconditioned
// Krylov solve consumes (stencil.cuh) -- 7 fp32 planes + the fp32 Jacobi scaling plane for the V-cycle, 7 fp64 planes
// for the recurrence, 1/a_ii in fp64 and the per-CTA Gershgorin maxima -- so no CSR -> DIA conversion pass is needed
struct LatDiaOut {
    float *planes32 = nullptr;
    double *planes64 = nullptr, *dinv = nullptr, *gpart = nullptr;
    int64_t np = 0, o0 = 0, o1 = 0;
    int use_bc = 0;             // the level matrix is the BC'd copy (rows / columns of Dirichlet dofs replaced)
    int gpart_off = 0;          // first slot of this launch's per-CTA maxima (the perimeter launch writes behind the interior's)
};

// boundary triangles (O(perimeter) of them): the Nitsche facet rows with the cell's own geometry, kept out of line so
// that the interior path of the node kernels stays a few hundred instructions
__device__ __noinline__ void node_facet_jac(const TriArgs &A, int64_t cell, int first, int second, int a, double row[3]) {
    Tri Tg;
    tri_load(A, cell, Tg);
    if (first) nlp_facet_jac_row(Tg, 2, A.beta, a, row);
    if (second) nlp_facet_jac_row(Tg, 0, A.beta, a, row);
}

// The node kernels run as two launches.  BND = false: the nodes at least two lattice rows / columns away from the boundary
// -- all six incident triangles exist and none of them owns a boundary facet (the symmetric Nitsche terms involve the
// normal derivative of the test function, so the vertex OPPOSITE a boundary facet, one ring inside, gets facet rows
// too) -- so the instantiation holds no range checks, no facet code, no calls and no stack frame (round 2 ncu / ptxas:
// the single kernel needed 80 - 88 registers and a 400-byte frame because of the out-of-line Nitsche path, 2 CTAs per
// SM, 1.18 / 0.87 ms for 1.3 / 0.4 GB of traffic).  BND = true: the two outer rings (all nodes on lattices narrower
// than 4 cells) with the complete general code.  Thread t of the band launch -> node (i, j):
__host__ __device__ __forceinline__ bool lattice_small(int nx, int ny) { return nx < 4 || ny < 4; }
__device__ __forceinline__ bool lattice_band_node(int t, int nx, int ny, int &i, int &j) {
    const int w = nx + 1, h = ny + 1;
    if (lattice_small(nx, ny)) {
        if (t >= w * h) return false;
        j = t / w;
        i = t - j * w;
        return true;
    }
    if (t < 4 * w) {                        // rows 0, 1, ny - 1, ny
        const int k = t / w;
        i = t - k * w;
        j = k < 2 ? k : ny - 3 + k;
        return true;
    }
    t -= 4 * w;
    const int hh = h - 4;                   // columns 0, 1, nx - 1, nx of the rows 2 .. ny - 2
    if (t < 4 * hh) {
        const int k = t / hh;
        j = 2 + (t - k * hh);
        i = k < 2 ? k : nx - 3 + k;
        return true;
    }
    return false;
}
static inline int lattice_band_count(int nx, int ny) {
    return lattice_small(nx, ny) ? (nx + 1) * (ny + 1) : 4 * (nx + 1) + 4 * (ny - 3);
}
// node of thread (blockIdx, threadIdx) in the interior / band launch; false = idle thread
template <bool BND>
__device__ __forceinline__ bool lattice_node_of_thread(int nx, int ny, int &i, int &j, int &r) {
    const int w = nx + 1, t = blockIdx.x * kThreads + threadIdx.x;
    if (BND) {
        if (!lattice_band_node(t, nx, ny, i, j)) return false;
        r = j * w + i;
        return true;
    }
    if (t >= w * (ny + 1) || lattice_small(nx, ny)) return false;
    r = t;
    j = t / w;
    i = t - j * w;
    return i >= 2 && j >= 2 && i <= nx - 2 && j <= ny - 2;
}

// contribution of incident triangle T (compile-time: local index and slots fold into the closed-form coefficients)
template <int T, bool BND>
__device__ __forceinline__ void node_jac_tri(const LatJacArgs &A, const LatGeom &G, int i, int j, const double u7[7], double v[7]) {
    constexpr int up = kNodeTri[T][2], a = kNodeTri[T][3], s0 = kNodeTri[T][4], s1 = kNodeTri[T][5], s2 = kNodeTri[T][6];
    const int ci = i + kNodeTri[T][0], cj = j + kNodeTri[T][1];
    if (BND && (ci < 0 || cj < 0 || ci >= A.nx || cj >= A.ny)) return;
    const double u[3] = {u7[s0], u7[s1], u7[s2]};
    double row[3];
    nlp_cell_jac_row_const<a>(up ? G.gu : G.gl, G.a2, u, row);
    if (BND) {
        const bool fb = !up && cj == 0 && A.ext_bottom, fr = !up && ci == A.nx - 1;
        const bool fl = up && ci == 0, ft = up && cj == A.ny - 1 && A.ext_top;
        if (fb || fr || fl || ft) node_facet_jac(A.T, 2 * (cj * A.nx + ci) + up, fb || fl, fr || ft, a, row);
    }
    v[s0] += row[0];
    v[s1] += row[1];
    v[s2] += row[2];
}

template <bool BND>
__global__ void __launch_bounds__(kThreads) k_nlpoisson_p1_node_jac(LatJacArgs A, LatGeom G, LatDiaOut O) {
    const int w = A.nx + 1;
    int i, j, r;                                                  // int32 dofs everywhere (FEMO_ELIMIT otherwise)
    __shared__ double sh_g[kThreads];
    double gersh = 0.0;
    if (lattice_node_of_thread<BND>(A.nx, A.ny, i, j, r)) {
    // stencil slots {-w-1, -w, -1, 0, 1, w, w+1} and which of them exist on the local lattice
    const bool present[7] = {!BND || (i > 0 && j > 0), !BND || j > 0, !BND || i > 0, true, !BND || i < A.nx, !BND || j < A.ny,
                             !BND || (i < A.nx && j < A.ny)};
    const int off[7] = {-w - 1, -w, -1, 0, 1, w, w + 1};
    double u7[7], v[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        u7[s] = present[s] ? __ldg(A.T.u + (r + off[s])) : 0.0;
        v[s] = 0.0;
    }
    node_jac_tri<0, BND>(A, G, i, j, u7, v);
    node_jac_tri<1, BND>(A, G, i, j, u7, v);
    node_jac_tri<2, BND>(A, G, i, j, u7, v);
    node_jac_tri<3, BND>(A, G, i, j, u7, v);
    node_jac_tri<4, BND>(A, G, i, j, u7, v);
    node_jac_tri<5, BND>(A, G, i, j, u7, v);
    // CSR positions: the row holds the present slots in ascending column order
    int32_t pos = A.rowptr[r];
    double sabs = 0.0, diag = 1.0;
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        double m = 0.0;                     // entry of the level matrix (plain or BC'd) in slot s
        if (present[s]) {
            double vb = v[s];
            if (A.out_bc || O.use_bc) {
                const uint8_t fl = A.bcflag ? A.bcflag[pos] : 0;
                vb = (fl == 0) ? v[s] : (fl == 1 ? 0.0 : A.bc_diag[A.col[pos]]);
            }
            if (A.out) A.out[pos] = v[s];
            if (A.out_bc) A.out_bc[pos] = vb;
            m = O.use_bc ? vb : v[s];
            ++pos;
        }
        if (O.planes32) {
            O.planes32[(int64_t)s * O.np + r] = (float)m;
            if (O.planes64) O.planes64[(int64_t)s * O.np + r] = m;
            sabs += fabs(m);
            if (s == 3) diag = m;
        }
    }
    if (O.planes32) {
        const double di = (diag != 0.0) ? 1.0 / diag : 1.0;
        O.planes32[(int64_t)7 * O.np + r] = (float)di;
        O.dinv[r] = di;
        if (r >= O.o0 && r < O.o1) gersh = sabs * fabs(di);
    }
    }
    if (O.gpart) {                          // per-CTA maximum of the Gershgorin bound (uniform branch)
        sh_g[threadIdx.x] = gersh;
        __syncthreads();
        for (int o = kThreads / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh_g[threadIdx.x] = fmax(sh_g[threadIdx.x], sh_g[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) O.gpart[O.gpart_off + blockIdx.x] = sh_g[0];
    }
}

// ---- residual rows, same node-centric scheme --------------------------------------------------------------------
// entry a of the cell vector  int grad(u).grad(phi_a) + (u^3 - f) phi_a dx  (as k_nlpoisson_p1_cell<OP_RES>)
template <int a>
__device__ __forceinline__ double nlp_cell_res_entry(const double g[3][2], double a2, const double u[3], double f) {
    const double gu0 = u[0] * g[0][0] + u[1] * g[1][0] + u[2] * g[2][0];
    const double gu1 = u[0] * g[0][1] + u[1] * g[1][1] + u[2] * g[2][1];
    const double u00 = u[0] * u[0], u11 = u[1] * u[1], u22 = u[2] * u[2], u01 = u[0] * u[1];
    // cubic monomials u_i u_j u_k, i <= j <= k, lexicographic
    const double cub[10] = {u00 * u[0], u00 * u[1], u00 * u[2], u01 * u[1], u01 * u[2], u[0] * u22,
                            u11 * u[1], u11 * u[2], u[1] * u22, u22 * u[2]};
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < 10; ++k) m = fma(kQuartRes[a][k], cub[k], m);
    return 0.5 * a2 * (gu0 * g[a][0] + gu1 * g[a][1]) + a2 * (m * (1.0 / 720.0) - f * (1.0 / 6.0));
}

// entry a of the Nitsche vector of exterior facet l (as k_nlpoisson_p1_facet<OP_RES>)
__device__ __noinline__ double nlp_facet_res_entry(const Tri &T, int l, double beta, const double u[3], int a) {
    const int la = (l == 0) ? 1 : 0, lb = (l == 2) ? 1 : 2;
    const double tx = T.X[lb][0] - T.X[la][0], ty = T.X[lb][1] - T.X[la][1];
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (T.X[la][0] - T.X[l][0]) + ny * (T.X[la][1] - T.X[l][1]) < 0.0) { nx = -nx; ny = -ny; }
    double h2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d = (c + 1) % 3;
        const double dx = T.X[c][0] - T.X[d][0], dy = T.X[c][1] - T.X[d][1];
        h2 = fmax(h2, dx * dx + dy * dy);
    }
    const double bh = beta / sqrt(h2);
    double gn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gn[c] = T.g[c][0] * nx + T.g[c][1] * ny;
    const double dudn = u[0] * gn[0] + u[1] * gn[1] + u[2] * gn[2];
    double R = 0.0;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double s = c_gl5[q][0], w = c_gl5[q][1] * len;
        double ph[3] = {0.0, 0.0, 0.0};
        ph[la] = 1.0 - s;
        ph[lb] = s;
        const double uq = u[la] * (1.0 - s) + u[lb] * s;
        const double ex = uex_nlp(T.X[la][0] + s * tx, T.X[la][1] + s * ty);
        R += w * (-dudn * ph[a] + (ex - uq) * gn[a] + bh * (uq - ex) * ph[a]);
    }
    return R;
}

__device__ __noinline__ double node_facet_res(const TriArgs &A, int64_t cell, int first, int second, const double u[3], int a) {
    Tri Tg;
    tri_load(A, cell, Tg);
    double R = 0.0;
    if (first) R += nlp_facet_res_entry(Tg, 2, A.beta, u, a);
    if (second) R += nlp_facet_res_entry(Tg, 0, A.beta, u, a);
    return R;
}

template <int T, bool BND>
__device__ __forceinline__ double node_res_tri(const LatJacArgs &A, const LatGeom &G, int i, int j, const double u7[7]) {
    constexpr int up = kNodeTri[T][2], a = kNodeTri[T][3];
    const int ci = i + kNodeTri[T][0], cj = j + kNodeTri[T][1];
    if (BND && (ci < 0 || cj < 0 || ci >= A.nx || cj >= A.ny)) return 0.0;
    const int c = 2 * (cj * A.nx + ci) + up;
    const double u[3] = {u7[kNodeTri[T][4]], u7[kNodeTri[T][5]], u7[kNodeTri[T][6]]};
    double R = nlp_cell_res_entry<a>(up ? G.gu : G.gl, G.a2, u, __ldg(A.T.f + c));
    if (BND) {
        const bool fb = !up && cj == 0 && A.ext_bottom, fr = !up && ci == A.nx - 1;
        const bool fl = up && ci == 0, ft = up && cj == A.ny - 1 && A.ext_top;
        if (fb || fr || fl || ft) R += node_facet_res(A.T, c, fb || fl, fr || ft, u, a);
    }
    return R;
}

template <bool BND>
__global__ void __launch_bounds__(kThreads) k_nlpoisson_p1_node_res(LatJacArgs A, LatGeom G, double *__restrict__ out) {
    const int w = A.nx + 1;
    int i, j, r;
    if (!lattice_node_of_thread<BND>(A.nx, A.ny, i, j, r)) return;
    const bool present[7] = {!BND || (i > 0 && j > 0), !BND || j > 0, !BND || i > 0, true, !BND || i < A.nx, !BND || j < A.ny,
                             !BND || (i < A.nx && j < A.ny)};
    const int off[7] = {-w - 1, -w, -1, 0, 1, w, w + 1};
    double u7[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) u7[s] = present[s] ? __ldg(A.T.u + (r + off[s])) : 0.0;
    double R = 0.0;
    R += node_res_tri<0, BND>(A, G, i, j, u7);
    R += node_res_tri<1, BND>(A, G, i, j, u7);
    R += node_res_tri<2, BND>(A, G, i, j, u7);
    R += node_res_tri<3, BND>(A, G, i, j, u7);
    R += node_res_tri<4, BND>(A, G, i, j, u7);
    R += node_res_tri<5, BND>(A, G, i, j, u7);
    out[r] = R;
}

}  // namespace femo
