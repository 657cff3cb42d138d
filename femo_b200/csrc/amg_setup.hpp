// amg_setup.hpp -- host side of the algebraic multigrid preconditioner for meshes without a lattice hierarchy
// (unstructured Gmsh meshes, the motor annulus).  The reference factorises every
// Jacobian with MUMPS (utils_dolfinx.py:405-408,476-512); here non-lattice problems get smoothed aggregation
// (Vanek, Mandel, Brezina 1996: strength graph -> greedy aggregates -> damped-Jacobi smoothed piecewise-constant
// prolongator -> Galerkin coarse operator P^T A P) with everything that depends only on the PATTERN decided once on the
// host, and every number recomputed on the device whenever the Jacobian changes (amg.cuh).  The numeric phase is three
// sorted segmented reductions per level over index lists built here:
//     P[t]  = sum_{k in pp}  (delta_{row,col(k)} - omega * A[k] / a_row,row)        (entries of A that fall in aggregate J)
//     AP[t] = sum_{(k,l) in ap} A[k] * P[l]
//     Ac[t] = sum_{(l,m) in ac} P[l] * AP[m]
// with ascending sources inside each segment: fixed summation order, no atomics, bit-reproducible (same contract as
// the assembly's gather maps, layout.hpp).
#pragma once
#include <cstdint>
#include <vector>

namespace femo {

struct AmgOpts {
    double theta = 0.08;       // strength threshold |a_ij| >= theta sqrt(a_ii a_jj) on level 0 ...
    double theta_decay = 0.5;  // ... multiplied by this per coarser level
    int max_levels = 10;
    int coarse_size = 200;     // stop coarsening at or below this many dofs (one-CTA dense inverse: keep it small; limit 512)
    int block = 1;             // dofs per node (aggregation is node based, one constant vector per component)
    double omega_scale = 1.0;  // prolongator smoothing P = (I - omega_scale * 4/(3 lmax) D^-1 A) T; 0 = plain aggregation
};

struct AmgLevelHost {
    int64_t n = 0, nnz = 0;
    std::vector<int32_t> rowptr, col, rb;     // operator pattern of this level + SpMV row blocks
    std::vector<double> vals, dinv;           // host numeric copy (empty when the hierarchy was built from the pattern only)
    double lmax = 2.0, omega = 2.0 / 3.0;
    // transfer to the next level (all empty on the coarsest level)
    int64_t nc = 0, nnzP = 0, nnzAP = 0, nnzC = 0;
    std::vector<int32_t> agg;                               // dof -> coarse dof, -1 for isolated (Dirichlet) rows
    std::vector<int32_t> p_rowptr, p_col, p_row, p_rb;      // P (n x nc), CSR; p_row[t] = row of entry t
    std::vector<int32_t> pp_ptr, pp_src;                    // P entry t sums the A entries pp_src[pp_ptr[t] .. pp_ptr[t+1])
    std::vector<int32_t> r_rowptr, r_col, r_perm, r_rb;     // R = P^T, CSR; value k = P value r_perm[k]
    std::vector<int32_t> ap_rowptr, ap_col;                 // pattern of A P
    std::vector<int32_t> ap_ptr, ap_ia, ap_ib;              // AP entry t = sum A[ap_ia] * P[ap_ib]
    std::vector<int32_t> ac_ptr, ac_ia, ac_ib;              // coarse entry t = sum P[ac_ia] * AP[ac_ib]
    std::vector<double> p_vals, ap_vals;                    // host numeric
};

struct AmgHier {
    AmgOpts opts;
    bool with_values = false;
    std::vector<AmgLevelHost> lv;
};

// Builds the whole hierarchy.  vals may be null (pattern-only aggregation: every connection strong); isolated[i] != 0
// marks rows the coarse correction must not touch (Dirichlet rows of the BC'd Jacobian), may be null.
void amg_build(const int32_t *rowptr, const int32_t *col, int64_t n, const double *vals, const uint8_t *isolated,
               const AmgOpts &opts, AmgHier &out);
// Numeric phase on the host with the same formulas and summation order as the device kernels: fills dinv / lmax / omega /
// p_vals / ap_vals of level l and the values of level l+1.  Used while building (strength on coarse levels) and by the tests.
void amg_numeric_host(AmgHier &h, int level);

}  // namespace femo
