// amg_setup.cpp -- pattern phase of the smoothed-aggregation hierarchy (see amg_setup.hpp).
#include "amg_setup.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "layout.hpp"

namespace femo {
namespace {

constexpr int32_t kUnset = -2;

struct Triple {
    int32_t j, a, b;
    bool operator<(const Triple &o) const { return j != o.j ? j < o.j : (a != o.a ? a < o.a : b < o.b); }
};

// dof-level isolation: rows the coarse correction leaves alone (Dirichlet rows; rows without any off-diagonal coupling)
void find_isolated(const AmgLevelHost &L, bool with_values, const uint8_t *given, std::vector<uint8_t> &isol) {
    isol.assign(L.n, 0);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < L.n; ++r) {
        if (given && given[r]) {
            isol[r] = 1;
            continue;
        }
        if (!with_values) continue;
        double d = 0.0, off = 0.0;
        for (int32_t t = L.rowptr[r]; t < L.rowptr[r + 1]; ++t) {
            const double v = std::fabs(L.vals[t]);
            if (L.col[t] == r) d = v;
            else off = std::max(off, v);
        }
        if (off <= 1e-14 * d) isol[r] = 1;
    }
}

// Greedy aggregation (Vanek et al.) on the node graph; natural node order => deterministic.
void aggregate(const AmgLevelHost &L, int bs, double theta, bool with_values, const std::vector<uint8_t> &isol,
               std::vector<int32_t> &agg_node, int64_t &nagg) {
    const int64_t nn = L.n / bs;
    std::vector<double> dn(nn, 0.0);
    std::vector<uint8_t> iso(nn, 1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
        double d = 0.0;
        uint8_t all = 1;
        for (int a = 0; a < bs; ++a) {
            const int64_t r = i * bs + a;
            if (!isol[r]) all = 0;
            for (int32_t t = L.rowptr[r]; t < L.rowptr[r + 1]; ++t)
                if (L.col[t] == r) d += with_values ? std::fabs(L.vals[t]) : 1.0;
        }
        dn[i] = d;
        iso[i] = all;
    }
    // strong-neighbour lists: same-component coupling s_ij = sum_a |A[(i,a),(j,a)]| >= theta sqrt(d_i d_j)
    std::vector<int64_t> sptr(nn + 1, 0);
    std::vector<int32_t> sidx;
    std::vector<double> sw;
    {
        std::vector<int32_t> pos(nn, -1), loc;
        std::vector<double> w;
        for (int64_t i = 0; i < nn; ++i) {
            loc.clear();
            w.clear();
            if (!iso[i]) {
                for (int a = 0; a < bs; ++a) {
                    const int64_t r = i * bs + a;
                    for (int32_t t = L.rowptr[r]; t < L.rowptr[r + 1]; ++t) {
                        const int32_t c = L.col[t];
                        if (c % bs != a) continue;
                        const int32_t j = c / bs;
                        if (j == i || iso[j]) continue;
                        const double v = with_values ? std::fabs(L.vals[t]) : 1.0;
                        if (pos[j] < 0) {
                            pos[j] = (int32_t)loc.size();
                            loc.push_back(j);
                            w.push_back(v);
                        } else w[pos[j]] += v;
                    }
                }
                for (size_t k = 0; k < loc.size(); ++k) {
                    const int32_t j = loc[k];
                    pos[j] = -1;
                    const bool strong = with_values ? (w[k] > 0.0 && w[k] >= theta * std::sqrt(dn[i] * dn[j])) : true;
                    if (strong) {
                        sidx.push_back(j);
                        sw.push_back(w[k]);
                    }
                }
            }
            sptr[i + 1] = (int64_t)sidx.size();
        }
    }
    agg_node.assign(nn, kUnset);
    nagg = 0;
    for (int64_t i = 0; i < nn; ++i)
        if (iso[i]) agg_node[i] = -1;
    // pass 1: a node whose strong neighbourhood is still free seeds an aggregate
    for (int64_t i = 0; i < nn; ++i) {
        if (agg_node[i] != kUnset || sptr[i + 1] == sptr[i]) continue;
        bool free_nb = true;
        for (int64_t k = sptr[i]; k < sptr[i + 1] && free_nb; ++k) free_nb = agg_node[sidx[k]] == kUnset;
        if (!free_nb) continue;
        const int32_t id = (int32_t)nagg++;
        agg_node[i] = id;
        for (int64_t k = sptr[i]; k < sptr[i + 1]; ++k) agg_node[sidx[k]] = id;
    }
    // pass 2: leftovers join the pass-1 aggregate they are most strongly tied to
    {
        std::vector<int32_t> a1(agg_node);
        for (int64_t i = 0; i < nn; ++i) {
            if (a1[i] != kUnset) continue;
            double best = -1.0;
            int32_t pick = kUnset;
            for (int64_t k = sptr[i]; k < sptr[i + 1]; ++k) {
                const int32_t aj = a1[sidx[k]];
                if (aj >= 0 && sw[k] > best) {
                    best = sw[k];
                    pick = aj;
                }
            }
            if (pick != kUnset) agg_node[i] = pick;
        }
    }
    // pass 3: what is left forms aggregates with its free strong neighbours (or stays alone)
    for (int64_t i = 0; i < nn; ++i) {
        if (agg_node[i] != kUnset) continue;
        const int32_t id = (int32_t)nagg++;
        agg_node[i] = id;
        for (int64_t k = sptr[i]; k < sptr[i + 1]; ++k)
            if (agg_node[sidx[k]] == kUnset) agg_node[sidx[k]] = id;
    }
}

// CSR transpose bookkeeping of P: R row I lists the fine rows holding an entry in column I, ascending
void transpose_p(AmgLevelHost &L) {
    L.r_rowptr.assign(L.nc + 1, 0);
    for (int64_t t = 0; t < L.nnzP; ++t) L.r_rowptr[L.p_col[t] + 1]++;
    for (int64_t I = 0; I < L.nc; ++I) L.r_rowptr[I + 1] += L.r_rowptr[I];
    L.r_col.resize(L.nnzP);
    L.r_perm.resize(L.nnzP);
    std::vector<int32_t> fill(L.r_rowptr.begin(), L.r_rowptr.end() - 1);
    for (int64_t t = 0; t < L.nnzP; ++t) {
        const int32_t k = fill[L.p_col[t]]++;
        L.r_col[k] = L.p_row[t];
        L.r_perm[k] = (int32_t)t;
    }
    build_rowblocks(L.r_rowptr, L.nc, L.r_rb);
}

// rows [0,nrows): gen(row, out) appends the (destination column, source a, source b) triples of that row.
// Produces the CSR pattern (rowptr, col) of the destinations and the pair lists grouped by destination entry.
template <class Gen>
void build_pairs(int64_t nrows, Gen gen, std::vector<int32_t> &rowptr, std::vector<int32_t> &col, std::vector<int32_t> &ptr,
                 std::vector<int32_t> &ia, std::vector<int32_t> &ib) {
    std::vector<int64_t> cnt_e(nrows + 1, 0), cnt_p(nrows + 1, 0);
#pragma omp parallel
    {
        std::vector<Triple> tr;
#pragma omp for schedule(dynamic, 1024)
        for (int64_t r = 0; r < nrows; ++r) {
            tr.clear();
            gen(r, tr);
            std::sort(tr.begin(), tr.end());
            int64_t ne = 0;
            for (size_t k = 0; k < tr.size(); ++k)
                if (k == 0 || tr[k].j != tr[k - 1].j) ++ne;
            cnt_e[r + 1] = ne;
            cnt_p[r + 1] = (int64_t)tr.size();
        }
    }
    for (int64_t r = 0; r < nrows; ++r) {
        cnt_e[r + 1] += cnt_e[r];
        cnt_p[r + 1] += cnt_p[r];
    }
    if (cnt_p[nrows] >= (int64_t)1 << 31) throw LayoutError{-4, "AMG: Galerkin pair list exceeds 2^31 entries"};
    rowptr.resize(nrows + 1);
    for (int64_t r = 0; r <= nrows; ++r) rowptr[r] = (int32_t)cnt_e[r];
    col.resize(cnt_e[nrows]);
    ptr.resize(cnt_e[nrows] + 1);
    ia.resize(cnt_p[nrows]);
    ib.resize(cnt_p[nrows]);
    ptr[cnt_e[nrows]] = (int32_t)cnt_p[nrows];
#pragma omp parallel
    {
        std::vector<Triple> tr;
#pragma omp for schedule(dynamic, 1024)
        for (int64_t r = 0; r < nrows; ++r) {
            tr.clear();
            gen(r, tr);
            std::sort(tr.begin(), tr.end());
            int64_t e = cnt_e[r] - 1, q = cnt_p[r];
            for (size_t k = 0; k < tr.size(); ++k, ++q) {
                if (k == 0 || tr[k].j != tr[k - 1].j) {
                    ++e;
                    col[e] = tr[k].j;
                    ptr[e] = (int32_t)q;
                }
                ia[q] = tr[k].a;
                ib[q] = tr[k].b;
            }
        }
    }
}

// pattern of P, its source lists, R, the A P and P^T A P pair lists, and the next level's pattern
void build_transfer(AmgLevelHost &L, AmgLevelHost &C) {
    const int64_t n = L.n;
    // P: row r (not isolated) has one entry per distinct aggregate among its columns
    {
        std::vector<int32_t> dummy_b;
        auto gen = [&](int64_t r, std::vector<Triple> &tr) {
            if (L.agg[r] < 0) return;
            for (int32_t t = L.rowptr[r]; t < L.rowptr[r + 1]; ++t) {
                const int32_t J = L.agg[L.col[t]];
                if (J >= 0) tr.push_back(Triple{J, t, 0});
            }
        };
        build_pairs(n, gen, L.p_rowptr, L.p_col, L.pp_ptr, L.pp_src, dummy_b);
    }
    L.nnzP = (int64_t)L.p_col.size();
    L.p_row.resize(L.nnzP);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r)
        for (int32_t t = L.p_rowptr[r]; t < L.p_rowptr[r + 1]; ++t) L.p_row[t] = (int32_t)r;
    build_rowblocks(L.p_rowptr, n, L.p_rb);
    transpose_p(L);
    // A P
    {
        auto gen = [&](int64_t r, std::vector<Triple> &tr) {
            for (int32_t k = L.rowptr[r]; k < L.rowptr[r + 1]; ++k) {
                const int32_t c = L.col[k];
                for (int32_t l = L.p_rowptr[c]; l < L.p_rowptr[c + 1]; ++l) tr.push_back(Triple{L.p_col[l], k, l});
            }
        };
        build_pairs(n, gen, L.ap_rowptr, L.ap_col, L.ap_ptr, L.ap_ia, L.ap_ib);
    }
    L.nnzAP = (int64_t)L.ap_col.size();
    // P^T (A P)
    {
        auto gen = [&](int64_t I, std::vector<Triple> &tr) {
            for (int32_t k = L.r_rowptr[I]; k < L.r_rowptr[I + 1]; ++k) {
                const int32_t r = L.r_col[k], l = L.r_perm[k];
                for (int32_t m = L.ap_rowptr[r]; m < L.ap_rowptr[r + 1]; ++m) tr.push_back(Triple{L.ap_col[m], l, m});
            }
        };
        build_pairs(L.nc, gen, C.rowptr, C.col, L.ac_ptr, L.ac_ia, L.ac_ib);
    }
    C.n = L.nc;
    C.nnz = (int64_t)C.col.size();
    L.nnzC = C.nnz;
    build_rowblocks(C.rowptr, C.n, C.rb);
}

}  // namespace

void amg_numeric_host(AmgHier &h, int level) {
    AmgLevelHost &L = h.lv[level];
    const int64_t n = L.n;
    L.dinv.resize(n);
    double lmax = 0.0;
#pragma omp parallel for schedule(static) reduction(max : lmax)
    for (int64_t i = 0; i < n; ++i) {
        double d = 1.0, s = 0.0;
        for (int32_t t = L.rowptr[i]; t < L.rowptr[i + 1]; ++t) {
            const double v = L.vals[t];
            s += std::fabs(v);
            if (L.col[t] == i) d = v;
        }
        const double di = (d != 0.0) ? 1.0 / d : 1.0;
        L.dinv[i] = di;
        lmax = std::max(lmax, s * std::fabs(di));
    }
    L.lmax = lmax;
    L.omega = h.opts.omega_scale * (4.0 / (3.0 * lmax));
    if (level + 1 >= (int)h.lv.size()) return;
    AmgLevelHost &C = h.lv[level + 1];
    L.p_vals.resize(L.nnzP);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < L.nnzP; ++t) {
        const int32_t r = L.p_row[t];
        const double w = L.omega * L.dinv[r];
        double acc = 0.0;
        for (int32_t s = L.pp_ptr[t]; s < L.pp_ptr[t + 1]; ++s) {
            const int32_t k = L.pp_src[s];
            acc += (L.col[k] == r ? 1.0 : 0.0) - w * L.vals[k];
        }
        L.p_vals[t] = acc;
    }
    L.ap_vals.resize(L.nnzAP);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < L.nnzAP; ++t) {
        double acc = 0.0;
        for (int32_t s = L.ap_ptr[t]; s < L.ap_ptr[t + 1]; ++s) acc += L.vals[L.ap_ia[s]] * L.p_vals[L.ap_ib[s]];
        L.ap_vals[t] = acc;
    }
    C.vals.resize(C.nnz);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < C.nnz; ++t) {
        double acc = 0.0;
        for (int32_t s = L.ac_ptr[t]; s < L.ac_ptr[t + 1]; ++s) acc += L.p_vals[L.ac_ia[s]] * L.ap_vals[L.ac_ib[s]];
        C.vals[t] = acc;
    }
}

void amg_build(const int32_t *rowptr, const int32_t *col, int64_t n, const double *vals, const uint8_t *isolated,
               const AmgOpts &opts, AmgHier &out) {
    out = AmgHier();
    out.opts = opts;
    out.with_values = vals != nullptr;
    int bs = opts.block > 0 ? opts.block : 1;
    if (n % bs) bs = 1;
    out.lv.emplace_back();
    {
        AmgLevelHost &L = out.lv[0];
        L.n = n;
        L.nnz = rowptr[n];
        L.rowptr.assign(rowptr, rowptr + n + 1);
        L.col.assign(col, col + L.nnz);
        if (vals) L.vals.assign(vals, vals + L.nnz);
        build_rowblocks(L.rowptr, n, L.rb);
    }
    double theta = opts.theta;
    for (int l = 0; l + 1 < opts.max_levels; ++l) {
        if (out.lv[l].n <= opts.coarse_size) break;
        std::vector<uint8_t> isol;
        find_isolated(out.lv[l], out.with_values, l == 0 ? isolated : nullptr, isol);
        std::vector<int32_t> agg_node;
        int64_t nagg = 0;
        aggregate(out.lv[l], bs, theta, out.with_values, isol, agg_node, nagg);
        const int64_t nc = nagg * bs;
        if (nc == 0 || nc * 10 > out.lv[l].n * 8) break;      // coarsening stalled: this level is the coarsest
        out.lv.emplace_back();
        AmgLevelHost &L = out.lv[l], &C = out.lv[l + 1];
        L.nc = nc;
        L.agg.resize(L.n);
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < L.n; ++r) L.agg[r] = isol[r] ? -1 : agg_node[r / bs] * bs + (int32_t)(r % bs);
        build_transfer(L, C);
        if (out.with_values) amg_numeric_host(out, l);
        theta *= opts.theta_decay;
    }
    if (out.with_values) amg_numeric_host(out, (int)out.lv.size() - 1);
}

}  // namespace femo
