// cpu_path.cpp -- CPU restatement of the benched hot path in C++/OpenMP (TEST / BASELINE INFRASTRUCTURE ONLY).
//
// One state+adjoint solve of the reference's examples/nonlinear_poisson_opt (run_nonlinear_poisson_opt.py:86-116
// residual with symmetric Nitsche terms, :140-145 output, :220 SNES) on the n x n "right"-diagonal unit square,
// restating on host cores what the reference delegates to dolfinx/PETSc (femo/fea/utils_dolfinx.py:175-202 assembly,
// :336-416 SNES, :476-512 linear solves, femo/csdl_opt/state_model.py:87-200 callback chain) with the SAME algorithm
// the CUDA engine uses in place of MUMPS: full-multigrid start + CG preconditioned by a geometric-multigrid V-cycle
// (rediscretised coarse operators, Chebyshev(2)-Jacobi smoothing with a Gershgorin bound, fp32 operator planes inside
// the V-cycle, dense coarsest solve), inexact Newton with PETSc's SNES stopping tests.
//
// It follows oracle/families.py (NonlinearPoissonP1) and oracle/solvers.py (StatePath.solve_snes / total_derivative)
// term by term and is pinned against them in tests/test_cpu_path.py (residual, Jacobian, functional, gradients at
// 1e-12; state / adjoint / dJ/df at 1e-8 against SuperLU).  bench.py times it on all host cores as `cpu_baseline`
// and as the `--impl reference` arm; tests use it as a second, fast oracle for the GPU path at n >= 1024.
// Nothing under femo_b200/ links or loads this file.
//
// Build: make -C oracle  (g++ -O3 -march=x86-64-v3 -fopenmp -shared -fPIC)
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// phase timers (FEMO_CPU_PROFILE=1 prints them)
struct Timers {
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
Timers g_tm;
struct Tick {
    int k;
    double t0;
    explicit Tick(int k_) : k(k_), t0(omp_get_wtime()) {}
    ~Tick() { g_tm.t[k] += omp_get_wtime() - t0; }
};

constexpr double kPi = 3.14159265358979323846;
constexpr double kAlpha = 6e-7, kBeta = 10.0;   // run_nonlinear_poisson_opt.py:79,196

// ---- quadrature (oracle/quadrature.py) ------------------------------------------------------------------------
struct Rules {
    double t6[6][3];     // degree-4 triangle rule (6 points): xi, eta, weight (weights sum to 1/2)
    double g5[5][2];     // 5-point Gauss-Legendre on [0,1]
    double t49[49][3];   // collapsed 7x7 Gauss rule (degree 12)
};

void gauss01(int m, double *x, double *w) {   // Newton on P_m, as numpy.polynomial.legendre.leggauss mapped to [0,1]
    for (int i = 0; i < m; ++i) {
        double z = std::cos(kPi * (i + 0.75) / (m + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= m; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
            }
            pp = m * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        x[m - 1 - i] = 0.5 * (z + 1.0);
        w[m - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
    }
}

const Rules &rules() {
    static Rules R;
    static bool init = false;
    if (!init) {
        const double s10 = std::sqrt(10.0), t = std::sqrt(38.0 - 44.0 * std::sqrt(2.0 / 5.0));
        const double a[2] = {(8.0 - s10 + t) / 18.0, (8.0 - s10 - t) / 18.0};
        const double sw = std::sqrt(213125.0 - 53320.0 * s10);
        const double wg[2] = {(620.0 + sw) / 3720.0, (620.0 - sw) / 3720.0};
        for (int k = 0; k < 2; ++k) {
            const double b = 1.0 - 2.0 * a[k];
            const double p[3][2] = {{a[k], a[k]}, {a[k], b}, {b, a[k]}};
            for (int j = 0; j < 3; ++j) {
                R.t6[3 * k + j][0] = p[j][0];
                R.t6[3 * k + j][1] = p[j][1];
                R.t6[3 * k + j][2] = 0.5 * wg[k];
            }
        }
        double x[7], w[7];
        gauss01(5, x, w);
        for (int i = 0; i < 5; ++i) { R.g5[i][0] = x[i]; R.g5[i][1] = w[i]; }
        gauss01(7, x, w);
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 7; ++j) {
                R.t49[7 * i + j][0] = x[i];
                R.t49[7 * i + j][1] = x[j] * (1.0 - x[i]);
                R.t49[7 * i + j][2] = w[i] * w[j] * (1.0 - x[i]);
            }
        init = true;
    }
    return R;
}

// ---- lattice ---------------------------------------------------------------------------------------------------
// nodes (nx+1) x (ny+1), node (i,j) -> j*(nx+1)+i; cell (i,j) holds triangles 2*(j*nx+i)+{0,1} with vertices
// lower [v0,v1,v3], upper [v0,v2,v3] (v0=(i,j), v1=(i+1,j), v2=(i,j+1), v3=(i+1,j+1)): oracle/mesh.py unit_square_tri
struct Lat {
    int nx = 0, ny = 0;
    int64_t N() const { return (int64_t)(nx + 1) * (ny + 1); }
    int64_t M() const { return 2 * (int64_t)nx * ny; }
    double X(int i) const { return (double)i / nx; }     // same formula as lo + (hi-lo)*i/n with lo=0, hi=1
    double Y(int j) const { return (double)j / ny; }
};

struct Tri {
    int64_t v[3];
    double x[3][2];
    double G[3][2];   // physical gradients of the P1 basis
    double detJ;      // |det J| = 2 area
    double h;         // cell diameter
};

inline void tri_of(const Lat &L, int i, int j, int upper, Tri &T) {
    const int w = L.nx + 1;
    const int64_t v0 = (int64_t)j * w + i, v1 = v0 + 1, v2 = v0 + w, v3 = v2 + 1;
    const double x0 = L.X(i), x1 = L.X(i + 1), y0 = L.Y(j), y1 = L.Y(j + 1);
    if (!upper) {
        T.v[0] = v0; T.v[1] = v1; T.v[2] = v3;
        T.x[0][0] = x0; T.x[0][1] = y0; T.x[1][0] = x1; T.x[1][1] = y0; T.x[2][0] = x1; T.x[2][1] = y1;
    } else {
        T.v[0] = v0; T.v[1] = v2; T.v[2] = v3;
        T.x[0][0] = x0; T.x[0][1] = y0; T.x[1][0] = x0; T.x[1][1] = y1; T.x[2][0] = x1; T.x[2][1] = y1;
    }
    const double J00 = T.x[1][0] - T.x[0][0], J01 = T.x[2][0] - T.x[0][0];
    const double J10 = T.x[1][1] - T.x[0][1], J11 = T.x[2][1] - T.x[0][1];
    const double det = J00 * J11 - J01 * J10;
    T.detJ = std::fabs(det);
    const double i00 = J11 / det, i01 = -J01 / det, i10 = -J10 / det, i11 = J00 / det;
    const double gr[3][2] = {{-1.0, -1.0}, {1.0, 0.0}, {0.0, 1.0}};
    for (int a = 0; a < 3; ++a) {   // G_a = Jinv^T gref_a
        T.G[a][0] = i00 * gr[a][0] + i10 * gr[a][1];
        T.G[a][1] = i01 * gr[a][0] + i11 * gr[a][1];
    }
    double h = 0.0;
    for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b)
            h = std::max(h, std::hypot(T.x[a][0] - T.x[b][0], T.x[a][1] - T.x[b][1]));
    T.h = h;
}

// exterior facets of a cell's triangles: local facet fl (opposite local vertex fl), vertices lf = local_facets[fl]
struct Facet {
    int upper, fl;
};
inline int facets_of(const Lat &L, int i, int j, Facet out[4]) {
    int n = 0;
    if (j == 0) out[n++] = {0, 2};            // bottom: lower triangle, edge v0-v1
    if (i == L.nx - 1) out[n++] = {0, 0};     // right: lower triangle, edge v1-v3
    if (i == 0) out[n++] = {1, 2};            // left: upper triangle, edge v0-v2
    if (j == L.ny - 1) out[n++] = {1, 0};     // top: upper triangle, edge v2-v3
    return n;
}
const int kLocalFacets[3][2] = {{1, 2}, {0, 2}, {0, 1}};

struct FacetGeom {
    double len, n[2], P[2], Q[2], gn[3];
    int a0, a1;
};
inline void facet_geom(const Tri &T, int fl, FacetGeom &F) {
    F.a0 = kLocalFacets[fl][0];
    F.a1 = kLocalFacets[fl][1];
    for (int d = 0; d < 2; ++d) { F.P[d] = T.x[F.a0][d]; F.Q[d] = T.x[F.a1][d]; }
    const double tx = F.Q[0] - F.P[0], ty = F.Q[1] - F.P[1];
    F.len = std::hypot(tx, ty);
    double nx = ty / F.len, ny = -tx / F.len;
    const double s = nx * (F.P[0] - T.x[fl][0]) + ny * (F.P[1] - T.x[fl][1]);
    const double sg = (s > 0) - (s < 0);
    F.n[0] = nx * sg;
    F.n[1] = ny * sg;
    for (int a = 0; a < 3; ++a) F.gn[a] = T.G[a][0] * F.n[0] + T.G[a][1] * F.n[1];
}

inline double u_exact(double x, double y) { return std::sin(2.0 * kPi * x) * std::sin(kPi * y); }

// ---- element tensors (oracle/families.py NonlinearPoissonP1) -----------------------------------------------
inline void cell_residual(const Tri &T, const double ue[3], double f, double Re[3]) {
    const Rules &Q = rules();
    double gu[2] = {0, 0};
    for (int a = 0; a < 3; ++a) { gu[0] += ue[a] * T.G[a][0]; gu[1] += ue[a] * T.G[a][1]; }
    double kg[3];
    for (int a = 0; a < 3; ++a) kg[a] = gu[0] * T.G[a][0] + gu[1] * T.G[a][1];
    Re[0] = Re[1] = Re[2] = 0.0;
    for (int q = 0; q < 6; ++q) {
        const double ph[3] = {1.0 - Q.t6[q][0] - Q.t6[q][1], Q.t6[q][0], Q.t6[q][1]};
        const double uq = ue[0] * ph[0] + ue[1] * ph[1] + ue[2] * ph[2];
        const double wq = Q.t6[q][2] * T.detJ, c = uq * uq * uq - f;
        for (int a = 0; a < 3; ++a) Re[a] += wq * (kg[a] + c * ph[a]);
    }
}

inline void cell_jacobian(const Tri &T, const double ue[3], double Ae[3][3]) {
    const Rules &Q = rules();
    double K[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            K[a][b] = T.G[a][0] * T.G[b][0] + T.G[a][1] * T.G[b][1];
            Ae[a][b] = 0.0;
        }
    for (int q = 0; q < 6; ++q) {
        const double ph[3] = {1.0 - Q.t6[q][0] - Q.t6[q][1], Q.t6[q][0], Q.t6[q][1]};
        const double uq = ue[0] * ph[0] + ue[1] * ph[1] + ue[2] * ph[2];
        const double wq = Q.t6[q][2] * T.detJ, c = 3.0 * uq * uq;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Ae[a][b] += wq * (K[a][b] + c * ph[a] * ph[b]);
    }
}

inline void facet_residual(const Tri &T, int fl, const double ue[3], double Rf[3]) {
    const Rules &Q = rules();
    FacetGeom F;
    facet_geom(T, fl, F);
    double dudn = 0.0;
    for (int a = 0; a < 3; ++a) dudn += ue[a] * F.gn[a];
    const double bh = kBeta / T.h;
    Rf[0] = Rf[1] = Rf[2] = 0.0;
    for (int q = 0; q < 5; ++q) {
        const double s = Q.g5[q][0], wq = Q.g5[q][1] * F.len;
        double ph[3] = {0, 0, 0};
        ph[F.a0] = 1.0 - s;
        ph[F.a1] = s;
        const double uq = ue[0] * ph[0] + ue[1] * ph[1] + ue[2] * ph[2];
        const double ex = u_exact(F.P[0] + s * (F.Q[0] - F.P[0]), F.P[1] + s * (F.Q[1] - F.P[1]));
        for (int a = 0; a < 3; ++a) Rf[a] += wq * (-dudn * ph[a] + (ex - uq) * F.gn[a] + bh * (uq - ex) * ph[a]);
    }
}

inline void facet_jacobian(const Tri &T, int fl, double Af[3][3]) {
    const Rules &Q = rules();
    FacetGeom F;
    facet_geom(T, fl, F);
    const double bh = kBeta / T.h;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Af[a][b] = 0.0;
    for (int q = 0; q < 5; ++q) {
        const double s = Q.g5[q][0], wq = Q.g5[q][1] * F.len;
        double ph[3] = {0, 0, 0};
        ph[F.a0] = 1.0 - s;
        ph[F.a1] = s;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Af[a][b] += wq * (-ph[a] * F.gn[b] - F.gn[a] * ph[b] + bh * ph[a] * ph[b]);
    }
}

// ---- DIA operator: 7 planes, offsets {-w-1,-w,-1,0,1,w,w+1}, plane s at [s*N + i] ------------------------------
inline int dia_slot(int64_t d, int w) {
    if (d == 0) return 3;
    if (d == 1) return 4;
    if (d == -1) return 2;
    if (d == w) return 5;
    if (d == w + 1) return 6;
    if (d == -w) return 1;
    return 0;   // -w-1
}

// Cell rows are processed in two colours (even / odd j): cells of row j touch node rows j and j+1 only,
// so rows of one colour never write the same entry.  The summation order differs from the numpy oracle
// (round-off level only; tests compare at 1e-12 relative to the largest entry).
template <class FN>
void for_cell_rows(const Lat &L, FN fn) {
    for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
        for (int j = colour; j < L.ny; j += 2)
            for (int i = 0; i < L.nx; ++i) fn(i, j);
    }
}

void assemble_residual(const Lat &L, const double *u, const double *f, double *R) {
    const int64_t N = L.N();
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < N; ++k) R[k] = 0.0;
    for_cell_rows(L, [&](int i, int j) {
        Tri T;
        Facet fc[4];
        const int nf = facets_of(L, i, j, fc);
        for (int up = 0; up < 2; ++up) {
            tri_of(L, i, j, up, T);
            const double ue[3] = {u[T.v[0]], u[T.v[1]], u[T.v[2]]};
            double Re[3];
            cell_residual(T, ue, f[2 * ((int64_t)j * L.nx + i) + up], Re);
            for (int a = 0; a < 3; ++a) R[T.v[a]] += Re[a];
            for (int k = 0; k < nf; ++k)
                if (fc[k].upper == up) {
                    facet_residual(T, fc[k].fl, ue, Re);
                    for (int a = 0; a < 3; ++a) R[T.v[a]] += Re[a];
                }
        }
    });
}

void assemble_jacobian(const Lat &L, const double *u, double *A /* 7*N */) {
    const int64_t N = L.N();
    const int w = L.nx + 1;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < 7 * N; ++k) A[k] = 0.0;
    for_cell_rows(L, [&](int i, int j) {
        Tri T;
        Facet fc[4];
        const int nf = facets_of(L, i, j, fc);
        for (int up = 0; up < 2; ++up) {
            tri_of(L, i, j, up, T);
            const double ue[3] = {u[T.v[0]], u[T.v[1]], u[T.v[2]]};
            double Ae[3][3];
            cell_jacobian(T, ue, Ae);
            for (int k = 0; k < nf; ++k)
                if (fc[k].upper == up) {
                    double Af[3][3];
                    facet_jacobian(T, fc[k].fl, Af);
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) Ae[a][b] += Af[a][b];
                }
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) A[(int64_t)dia_slot(T.v[b] - T.v[a], w) * N + T.v[a]] += Ae[a][b];
        }
    });
}

// functional J = int 1/2 (u - u_ex)^2 + alpha/2 f^2 dx, dJ/du, dJ/df.  u_ex at the 49 points of a cell comes from
// one sincos pair per cell corner and an angle-addition table of the point offsets (uniform lattice), the same
// evaluation strategy as the CUDA kernel; values agree with the direct formula to round-off.
void assemble_output(const Lat &L, const double *u, const double *f, double *Jout, double *dJdu, double *dJdf) {
    const Rules &Q = rules();
    const int64_t N = L.N();
    const double hx = 1.0 / L.nx, hy = 1.0 / L.ny;
    // offsets of the 49 points from the cell's v0 in both triangle types
    double tab[2][49][4];
    for (int t = 0; t < 2; ++t)
        for (int q = 0; q < 49; ++q) {
            const double xi = Q.t49[q][0], et = Q.t49[q][1];
            const double dx = t == 0 ? hx * (xi + et) : hx * et, dy = t == 0 ? hy * et : hy * (xi + et);
            tab[t][q][0] = std::cos(2.0 * kPi * dx); tab[t][q][1] = std::sin(2.0 * kPi * dx);
            tab[t][q][2] = std::cos(kPi * dy); tab[t][q][3] = std::sin(kPi * dy);
        }
    if (dJdu) {
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < N; ++k) dJdu[k] = 0.0;
    }
    double Jsum = 0.0;
    for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static) reduction(+ : Jsum)
        for (int j = colour; j < L.ny; j += 2) {
            const double sy = std::sin(kPi * L.Y(j)), cy = std::cos(kPi * L.Y(j));
            for (int i = 0; i < L.nx; ++i) {
                const double sx = std::sin(2.0 * kPi * L.X(i)), cx = std::cos(2.0 * kPi * L.X(i));
                Tri T;
                for (int up = 0; up < 2; ++up) {
                    tri_of(L, i, j, up, T);
                    const double ue[3] = {u[T.v[0]], u[T.v[1]], u[T.v[2]]};
                    const int64_t c = 2 * ((int64_t)j * L.nx + i) + up;
                    double val = 0.0, ge[3] = {0, 0, 0};
                    for (int q = 0; q < 49; ++q) {
                        const double xi = Q.t49[q][0], et = Q.t49[q][1];
                        const double ph[3] = {1.0 - xi - et, xi, et};
                        const double *o = tab[up][q];
                        const double ex = (sx * o[0] + cx * o[1]) * (sy * o[2] + cy * o[3]);
                        const double eq = ue[0] * ph[0] + ue[1] * ph[1] + ue[2] * ph[2] - ex;
                        const double wq = Q.t49[q][2] * T.detJ;
                        val += wq * 0.5 * eq * eq;
                        for (int a = 0; a < 3; ++a) ge[a] += wq * eq * ph[a];
                    }
                    val += 0.5 * T.detJ * 0.5 * kAlpha * f[c] * f[c];
                    Jsum += val;
                    if (dJdu)
                        for (int a = 0; a < 3; ++a) dJdu[T.v[a]] += ge[a];
                    if (dJdf) dJdf[c] = 0.5 * T.detJ * kAlpha * f[c];
                }
            }
        }
    }
    if (Jout) *Jout = Jsum;
}

// ---- vector kernels ----------------------------------------------------------------------------------------------
double dot(const double *a, const double *b, int64_t n) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// y = A x (or b - A x) for a DIA operator; VT = double (Krylov recurrence) or float (V-cycle copy)
template <class VT>
void dia_apply(const VT *A, int64_t N, int w, const double *x, double *y, const double *b) {
    const int64_t off[7] = {-(int64_t)w - 1, -(int64_t)w, -1, 0, 1, w, (int64_t)w + 1};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double acc = 0.0;
        for (int s = 0; s < 7; ++s) {
            const int64_t j = i + off[s];
            if (j >= 0 && j < N) acc += (double)A[(int64_t)s * N + i] * x[j];
        }
        y[i] = b ? b[i] - acc : acc;
    }
}

// ---- multigrid (mirrors femo_b200/csrc/multigrid.cuh) -----------------------------------------------------
struct Level {
    Lat L;
    std::vector<double> A;      // fp64 planes (level 0: borrowed through Ap)
    const double *Ap = nullptr;
    std::vector<float> A32;
    std::vector<double> dinv, x, b, r, d, q, u, fb, fx, dense;
    double lmax = 2.0;
};

inline double hat_p1(double dx, double dy) {
    const double v = 1.0 - std::max(std::max(dx, dy), 0.0) + std::min(std::min(dx, dy), 0.0);
    return v > 0.0 ? v : 0.0;
}

// dst (+)= P src: evaluate the coarse (src) P1 interpolant at the nodes of dst's lattice
void interp(const Lat &s, const Lat &d, const double *src, double *dst, bool add) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j <= d.ny; ++j)
        for (int i = 0; i <= d.nx; ++i) {
            const double X = (double)i * ((double)s.nx / (double)d.nx), Y = (double)j * ((double)s.ny / (double)d.ny);
            const int I = std::min((int)X, s.nx - 1), J = std::min((int)Y, s.ny - 1);
            double acc = 0.0;
            for (int b = 0; b < 2; ++b)
                for (int a = 0; a < 2; ++a) {
                    const double wgt = hat_p1(X - (double)(I + a), Y - (double)(J + b));
                    if (wgt > 0.0) acc += wgt * src[(int64_t)(J + b) * (s.nx + 1) + (I + a)];
                }
            double &o = dst[(int64_t)j * (d.nx + 1) + i];
            o = add ? o + acc : acc;
        }
}

// rc = P^T rf with the weights of interp(coarse -> fine)
void restrict_(const Lat &f, const Lat &c, const double *rf, double *rc) {
    const double sx = (double)c.nx / (double)f.nx, sy = (double)c.ny / (double)f.ny;
#pragma omp parallel for schedule(static)
    for (int J = 0; J <= c.ny; ++J)
        for (int I = 0; I <= c.nx; ++I) {
            const int ilo = std::max(0, (int)std::floor((double)(I - 1) / sx)), ihi = std::min(f.nx, (int)std::ceil((double)(I + 1) / sx));
            const int jlo = std::max(0, (int)std::floor((double)(J - 1) / sy)), jhi = std::min(f.ny, (int)std::ceil((double)(J + 1) / sy));
            double acc = 0.0;
            for (int j = jlo; j <= jhi; ++j)
                for (int i = ilo; i <= ihi; ++i) {
                    const double X = (double)i * sx, Y = (double)j * sy;
                    const int I0 = std::min((int)X, c.nx - 1), J0 = std::min((int)Y, c.ny - 1);
                    if (I < I0 || I > I0 + 1 || J < J0 || J > J0 + 1) continue;
                    const double wgt = hat_p1(X - (double)I, Y - (double)J);
                    if (wgt > 0.0) acc += wgt * rf[(int64_t)j * (f.nx + 1) + i];
                }
            rc[(int64_t)J * (c.nx + 1) + I] = acc;
        }
}

struct Mg {
    std::vector<Level> lv;
    int degree = 2;
    double ratio = 4.0;
    long vcycles = 0;
    std::vector<std::vector<double>> kr;   // Krylov work vectors
};

constexpr int kCoarsest = 8;

void mg_build(Mg &M, int n) {
    M.lv.clear();
    int nx = n, ny = n;
    for (;;) {
        Level l;
        l.L.nx = nx;
        l.L.ny = ny;
        M.lv.push_back(std::move(l));
        if (nx <= kCoarsest && ny <= kCoarsest) break;
        if (nx > kCoarsest) nx = (nx + 1) / 2;
        if (ny > kCoarsest) ny = (ny + 1) / 2;
    }
    for (size_t k = 0; k < M.lv.size(); ++k) {
        Level &l = M.lv[k];
        const int64_t N = l.L.N();
        if (k > 0) l.A.assign(7 * N, 0.0);
        l.A32.assign(7 * N, 0.0f);
        for (auto *v : {&l.dinv, &l.x, &l.b, &l.r, &l.d, &l.q, &l.u, &l.fb, &l.fx}) v->assign(N, 0.0);
    }
    Level &c = M.lv.back();
    c.dense.assign((size_t)c.L.N() * c.L.N(), 0.0);
}

void dense_inverse(const double *A, int64_t N, int w, double *Inv) {   // Gauss-Jordan with partial pivoting
    const int n = (int)N;
    std::vector<double> D((size_t)n * n, 0.0);
    const int64_t off[7] = {-(int64_t)w - 1, -(int64_t)w, -1, 0, 1, w, (int64_t)w + 1};
    for (int i = 0; i < n; ++i)
        for (int s = 0; s < 7; ++s) {
            const int64_t j = i + off[s];
            if (j >= 0 && j < n) D[(size_t)i * n + j] = A[(int64_t)s * N + i];
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Inv[(size_t)i * n + j] = i == j ? 1.0 : 0.0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(D[(size_t)i * n + k]) > std::fabs(D[(size_t)p * n + k])) p = i;
        if (p != k)
            for (int j = 0; j < n; ++j) {
                std::swap(D[(size_t)k * n + j], D[(size_t)p * n + j]);
                std::swap(Inv[(size_t)k * n + j], Inv[(size_t)p * n + j]);
            }
        const double pv = 1.0 / D[(size_t)k * n + k];
        for (int j = 0; j < n; ++j) { D[(size_t)k * n + j] *= pv; Inv[(size_t)k * n + j] *= pv; }
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            const double fct = D[(size_t)i * n + k];
            if (fct == 0.0) continue;
            for (int j = 0; j < n; ++j) {
                D[(size_t)i * n + j] -= fct * D[(size_t)k * n + j];
                Inv[(size_t)i * n + j] -= fct * Inv[(size_t)k * n + j];
            }
        }
    }
}

// hierarchy for the fine operator A0 at the fine state u0: coarse levels rediscretised at the interpolated state
void mg_setup(Mg &M, const double *A0, const double *u0) {
    for (size_t k = 0; k < M.lv.size(); ++k) {
        Level &l = M.lv[k];
        const int64_t N = l.L.N();
        const int w = l.L.nx + 1;
        if (k == 0) {
            l.Ap = A0;
        } else {
            const Level &F = M.lv[k - 1];
            interp(F.L, l.L, k == 1 ? u0 : F.u.data(), l.u.data(), false);
            assemble_jacobian(l.L, l.u.data(), l.A.data());
            l.Ap = l.A.data();
        }
        const double *A = l.Ap;
        if (k + 1 == M.lv.size()) {
            dense_inverse(A, N, w, l.dense.data());
            break;
        }
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < 7 * N; ++t) l.A32[t] = (float)A[t];
        double mx = 0.0;
#pragma omp parallel for schedule(static) reduction(max : mx)
        for (int64_t i = 0; i < N; ++i) {
            double s = 0.0;
            for (int p = 0; p < 7; ++p) s += std::fabs(A[(int64_t)p * N + i]);
            const double d = A[3 * N + i], di = d != 0.0 ? 1.0 / d : 1.0;
            l.dinv[i] = di;
            mx = std::max(mx, s * std::fabs(di));
        }
        l.lmax = mx;
    }
}

// Chebyshev(deg)-Jacobi smoother on level l (fp32 planes), x ~ A^-1 b
void smooth(Mg &M, Level &l, const double *b, double *x, bool zero_guess) {
    const int64_t N = l.L.N();
    const int w = l.L.nx + 1;
    const double lmax = l.lmax, lmin = lmax / M.ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    double *d = l.d.data(), *r = l.r.data(), *q = l.q.data();
    const double *dinv = l.dinv.data();
    if (zero_guess) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            r[i] = b[i];
            d[i] = dinv[i] * b[i] / theta;
            x[i] = d[i];
        }
    } else {
        dia_apply(l.A32.data(), N, w, x, r, b);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            d[i] = dinv[i] * r[i] / theta;
            x[i] += d[i];
        }
    }
    for (int k = 2; k <= M.degree; ++k) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        const double c1 = rho_new * rho, c2 = 2.0 * rho_new / delta;
        dia_apply(l.A32.data(), N, w, d, q, nullptr);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            r[i] -= q[i];
            d[i] = c1 * d[i] + c2 * dinv[i] * r[i];
            x[i] += d[i];
        }
        rho = rho_new;
    }
}

void vcycle(Mg &M, size_t k, const double *b, double *x) {
    Level &l = M.lv[k];
    const int64_t N = l.L.N();
    if (k + 1 == M.lv.size()) {
        const int n = (int)N;
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int j = 0; j < n; ++j) acc += l.dense[(size_t)i * n + j] * b[j];
            x[i] = acc;
        }
        return;
    }
    if (k == 0) M.vcycles++;
    Level &c = M.lv[k + 1];
    smooth(M, l, b, x, true);
    std::vector<double> &res = l.fx;     // scratch of this level outside the FMG phase is l.q; keep r for the smoother
    (void)res;
    dia_apply(l.A32.data(), N, l.L.nx + 1, x, l.q.data(), b);   // q = b - A x
    restrict_(l.L, c.L, l.q.data(), c.b.data());
    vcycle(M, k + 1, c.b.data(), c.x.data());
    interp(c.L, l.L, c.x.data(), x, true);
    smooth(M, l, b, x, false);
}

// full-multigrid start: x ~ A^-1 b to discretisation accuracy
void fmg(Mg &M, const double *b, double *x, double *r0, double *e0) {
    const size_t nl = M.lv.size();
    for (size_t k = 0; k + 1 < nl; ++k)
        restrict_(M.lv[k].L, M.lv[k + 1].L, k == 0 ? b : M.lv[k].fb.data(), M.lv[k + 1].fb.data());
    for (size_t kk = nl; kk-- > 0;) {
        Level &l = M.lv[kk];
        const int64_t N = l.L.N();
        const double *bl = kk == 0 ? b : l.fb.data();
        double *xl = kk == 0 ? x : l.fx.data();
        if (kk + 1 == nl) {
            const int n = (int)N;
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc += l.dense[(size_t)i * n + j] * bl[j];
                xl[i] = acc;
            }
            continue;
        }
        Level &c = M.lv[kk + 1];
        interp(c.L, l.L, c.fx.data(), xl, false);
        double *rl = kk == 0 ? r0 : l.b.data(), *el = kk == 0 ? e0 : l.x.data();
        dia_apply(l.Ap, N, l.L.nx + 1, xl, rl, bl);       // fp64 residual of the prolonged iterate
        vcycle(M, kk, rl, el);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) xl[i] += el[i];
    }
}

struct KrylovInfo {
    int iterations = 0, converged = 0;
    double rnorm = 0, bnorm = 0;
};

// GMG-preconditioned CG on the fp64 operator A (symmetric); x0 replaced by the full-multigrid iterate
KrylovInfo pcg(Mg &M, const double *A, const double *u_state, const double *b, double *x, double rtol, double atol, int max_it) {
    Level &l0 = M.lv[0];
    const int64_t N = l0.L.N();
    const int w = l0.L.nx + 1;
    { Tick t(4); mg_setup(M, A, u_state); }
    M.kr.resize(4);
    for (auto &v : M.kr) v.resize(N);
    std::vector<double> &r = M.kr[0], &z = M.kr[1], &p = M.kr[2], &q = M.kr[3];
    KrylovInfo ki;
    ki.bnorm = std::sqrt(dot(b, b, N));
    { Tick t(6); fmg(M, b, x, r.data(), z.data()); }
    dia_apply(A, N, w, x, r.data(), b);
    double rnorm = std::sqrt(dot(r.data(), r.data(), N));
    const double tol = std::max(rtol * ki.bnorm, atol);
    double rz = 0.0;
    int it = 0;
    bool conv = rnorm <= tol;
    while (!conv && it < max_it) {
        { Tick t(5); vcycle(M, 0, r.data(), z.data()); }
        const double rz_new = dot(r.data(), z.data(), N);
        const double beta = it == 0 ? 0.0 : rz_new / rz;
        rz = rz_new;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) p[i] = z[i] + beta * p[i];
        dia_apply(A, N, w, p.data(), q.data(), nullptr);
        const double alpha = rz / dot(p.data(), q.data(), N);
        double rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rr)
        for (int64_t i = 0; i < N; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * q[i];
            rr += r[i] * r[i];
        }
        ++it;
        rnorm = std::sqrt(rr);
        if (!(rnorm == rnorm)) break;
        conv = rnorm <= tol;
    }
    ki.iterations = it;
    ki.converged = conv;
    ki.rnorm = rnorm;
    return ki;
}

// workspace that persists across steps of the same size, as the CUDA engine's arenas do
struct Workspace {
    int n = -1;
    Mg M;
    std::vector<double> A, b, y, dJdu;
};
Workspace g_ws;

}  // namespace

extern "C" {

int femo_cpu_version(void) { return 100; }

// element-level pieces, exposed so tests can pin them against the numpy oracle
void femo_cpu_nlp_residual(int n, const double *u, const double *f, double *R) {
    Lat L{n, n};
    assemble_residual(L, u, f, R);
}
void femo_cpu_nlp_jacobian_dia(int n, const double *u, double *planes /* 7*(n+1)^2 */) {
    Lat L{n, n};
    assemble_jacobian(L, u, planes);
}
void femo_cpu_nlp_output(int n, const double *u, const double *f, double *J, double *dJdu, double *dJdf) {
    Lat L{n, n};
    assemble_output(L, u, f, J, dJdu, dJdf);
}

// One state+adjoint solve (what one optimiser gradient evaluation triggers):
// SNES from u = 0, J, dJ/du, dJ/df, adjoint solve, dJ/df_total = dJ/df - dRdf^T lambda.
// info: [0] newton its, [1] krylov its (state), [2] krylov its (adjoint), [3] converged reason (1 abs, 2 rel, 3 stol),
//       [4] V-cycles, [5] threads used.   Returns 0 on success, 1 when SNES / a Krylov solve did not converge.
int femo_cpu_nlp_step(int n, const double *f, double *u, double *lam, double *grad, double *Jout, double krylov_rtol,
                      int nthreads, int64_t *info, double *fnorms /* [2]: ||F(u0)||, ||F(u)|| */) {
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Lat L{n, n};
    const int64_t N = L.N(), Mc = L.M();
    const double snes_atol = 1e-13, snes_rtol = 1e-13, snes_stol = 1e-8;
    if (g_ws.n != n) {
        g_ws = Workspace();
        mg_build(g_ws.M, n);
        g_ws.A.assign(7 * N, 0.0);
        g_ws.b.assign(N, 0.0);
        g_ws.y.assign(N, 0.0);
        g_ws.dJdu.assign(N, 0.0);
        g_ws.n = n;
    }
    Mg &M = g_ws.M;
    M.vcycles = 0;
    std::vector<double> &A = g_ws.A, &b = g_ws.b, &y = g_ws.y;
    g_tm = Timers();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) u[i] = 0.0;
    { Tick t(0); assemble_residual(L, u, f, b.data()); }
    double f0 = std::sqrt(dot(b.data(), b.data(), N)), fn = f0;
    int it = 0, kit = 0, reason = fn < snes_atol ? 1 : 0, bad = 0;
    while (!reason && it < 100) {
        { Tick t(1); assemble_jacobian(L, u, A.data()); }
        KrylovInfo ki;
        { Tick t(2); ki = pcg(M, A.data(), u, b.data(), y.data(), krylov_rtol, 0.1 * snes_atol, 100000); }
        if (!ki.converged) bad = 1;
        kit += ki.iterations;
        double sy = 0.0, sx = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sy, sx)
        for (int64_t i = 0; i < N; ++i) {
            u[i] -= y[i];
            sy += y[i] * y[i];
            sx += u[i] * u[i];
        }
        { Tick t(0); assemble_residual(L, u, f, b.data()); }
        fn = std::sqrt(dot(b.data(), b.data(), N));
        ++it;
        if (fn < snes_atol) reason = 1;
        else if (fn <= snes_rtol * f0) reason = 2;
        else if (std::sqrt(sy) < snes_stol * std::sqrt(sx)) reason = 3;
    }
    // linearisation at the converged state, functional and partials
    { Tick t(1); assemble_jacobian(L, u, A.data()); }
    std::vector<double> &dJdu = g_ws.dJdu;
    { Tick t(3); assemble_output(L, u, f, Jout, dJdu.data(), grad); }
    // adjoint: A^T lam = dJ/du (A symmetric)
    KrylovInfo ka;
    { Tick t(2); ka = pcg(M, A.data(), u, dJdu.data(), lam, krylov_rtol, 0.0, 100000); }
    if (!ka.converged) bad = 1;
    // dR/df = -detJ/6 per (vertex of cell, cell) (1-point rule): grad -= dRdf^T lam
    const double detJ = 1.0 / ((double)n * n);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            Tri T;
            for (int up = 0; up < 2; ++up) {
                tri_of(L, i, j, up, T);
                const int64_t c = 2 * ((int64_t)j * n + i) + up;
                // -(w detJ phi_a) with the degree-1 rule: one point (1/3,1/3), weight 1/2
                const double de = -(0.5 * T.detJ) * (1.0 / 3.0);
                grad[c] -= de * (lam[T.v[0]] + lam[T.v[1]] + lam[T.v[2]]);
            }
        }
    (void)detJ; (void)Mc;
    if (info) {
        info[0] = it; info[1] = kit; info[2] = ka.iterations; info[3] = reason; info[4] = M.vcycles;
        info[5] = omp_get_max_threads();
    }
    if (fnorms) { fnorms[0] = f0; fnorms[1] = fn; }
    if (getenv("FEMO_CPU_PROFILE"))
        fprintf(stderr, "cpu_path n=%d: residual %.2f s, jacobian %.2f s, pcg (incl. setup) %.2f s [setup %.2f, vcycle %.2f, fmg %.2f], output %.2f s\n",
                n, g_tm.t[0], g_tm.t[1], g_tm.t[2], g_tm.t[4], g_tm.t[5], g_tm.t[6], g_tm.t[3]);
    return (reason && !bad) ? 0 : 1;
}

}  // extern "C"
