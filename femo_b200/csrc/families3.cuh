// families3.cuh -- family 5: L2 projection onto CG1 / DG0 (`project`,
// femo/fea/utils_dolfinx.py:549-583): residual int (u - g) w dx, Jacobian = mass matrix.
// Sources g: analytic u_ex / f_ex of examples/nonlinear_poisson_opt (:144-145,167),
// a DG0 function raised to a power (examples/beam_topo_opt:264-268) or a CG1 function.
#pragma once
#include "families.cuh"

namespace femo {

struct MassArgs {
    const double *coords;
    const int32_t *cellsT;
    int64_t ncells;
    const double *u, *src;
    int target;   // 0 CG1, 1 DG0
    int source;   // 0 u_ex, 1 f_ex, 2 dg^power, 3 cg1 function
    double power;
    double *out;
};

template <int OP>
__global__ void __launch_bounds__(kThreads) k_mass_cell(MassArgs A) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= A.ncells) return;
    const int64_t ne = A.ncells;
    TriArgs TA;
    TA.coords = A.coords;
    TA.cellsT = A.cellsT;
    TA.ncells = A.ncells;
    Tri T;
    tri_load(TA, c, T);
    if (OP == OP_JAC) {
        if (A.target == 1) {
            A.out[c] = 0.5 * T.a2;
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) A.out[(a * 3 + b) * ne + c] = T.a2 * ((a == b) ? (1.0 / 12.0) : (1.0 / 24.0));
        }
        return;
    }
    if (OP != OP_RES) return;
    double u[3] = {0.0, 0.0, 0.0}, s[3] = {0.0, 0.0, 0.0};
    if (A.target == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) u[a] = A.u[T.v[a]];
    } else {
        u[0] = A.u[c];
    }
    double sconst = 0.0;
    if (A.source == 2) sconst = pow(A.src[c], A.power);
    if (A.source == 3) {
#pragma unroll
        for (int a = 0; a < 3; ++a) s[a] = A.src[T.v[a]];
    }
    const double pi = 3.141592653589793;
    double R[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < 49; ++q) {
        const double ph[3] = {1.0 - c_tri49[q][0] - c_tri49[q][1], c_tri49[q][0], c_tri49[q][1]};
        double g;
        if (A.source <= 1) {
            const double x = ph[0] * T.X[0][0] + ph[1] * T.X[1][0] + ph[2] * T.X[2][0];
            const double y = ph[0] * T.X[0][1] + ph[1] * T.X[1][1] + ph[2] * T.X[2][1];
            const double ue = uex_nlp(x, y);
            g = (A.source == 0) ? ue : 5.0 * pi * pi * ue + ue * ue * ue;
        } else if (A.source == 2) {
            g = sconst;
        } else {
            g = s[0] * ph[0] + s[1] * ph[1] + s[2] * ph[2];
        }
        const double w = c_tri49[q][2] * T.a2;
        if (A.target == 0) {
            const double d = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2] - g;
#pragma unroll
            for (int a = 0; a < 3; ++a) R[a] += w * d * ph[a];
        } else {
            R[0] += w * (u[0] - g);
        }
    }
    if (A.target == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) A.out[a * ne + c] = R[a];
    } else {
        A.out[c] = R[0];
    }
}

}  // namespace femo
