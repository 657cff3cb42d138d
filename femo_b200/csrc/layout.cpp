// layout.cpp -- see layout.hpp.
#include "layout.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>

#include "../../include/femo_b200.h"

namespace femo {

static const int64_t kInt32Max = 2147483647LL;

// ---------------------------------------------------------------------------
// meshes
// ---------------------------------------------------------------------------
static void lattice_coords(int nx, int ny, const double lo[2], const double hi[2], Mesh &m) {
    m.gdim = 2;
    m.n[0] = nx;
    m.n[1] = ny;
    m.lo[0] = lo[0]; m.lo[1] = lo[1];
    m.hi[0] = hi[0]; m.hi[1] = hi[1];
    m.nverts = (int64_t)(nx + 1) * (ny + 1);
    m.coords.resize(m.nverts * 2);
    for (int iy = 0; iy <= ny; ++iy)
        for (int ix = 0; ix <= nx; ++ix) {
            int64_t v = (int64_t)iy * (nx + 1) + ix;
            // same expression as the oracle (lo + (hi-lo)*i/n) so coordinates agree bit for bit
            m.coords[2 * v] = lo[0] + (hi[0] - lo[0]) * (double)ix / (double)nx;
            m.coords[2 * v + 1] = lo[1] + (hi[1] - lo[1]) * (double)iy / (double)ny;
        }
}

void make_unit_square_tri(int nx, int ny, const double lo[2], const double hi[2], Mesh &m, bool ext_bottom,
                          bool ext_top) {
    m = Mesh();
    m.kind = MESH_TRI;
    m.nvpc = 3;
    lattice_coords(nx, ny, lo, hi, m);
    m.ncells = 2LL * nx * ny;
    m.cells.resize(m.ncells * 3);
    for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
            int32_t v0 = iy * (nx + 1) + ix, v1 = v0 + 1, v2 = v0 + nx + 1, v3 = v2 + 1;
            int64_t c = 2LL * ((int64_t)iy * nx + ix);
            int32_t *a = &m.cells[c * 3];
            a[0] = v0; a[1] = v1; a[2] = v3;   // "right" diagonal v0-v3
            a[3] = v0; a[4] = v2; a[5] = v3;
            // exterior facets; local facet i is opposite local vertex i
            if (ix == nx - 1) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(0); }      // v1-v3
            if (iy == 0 && ext_bottom) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(2); }      // v0-v1
            if (iy == ny - 1 && ext_top) { m.bf_cell.push_back((int32_t)c + 1); m.bf_local.push_back(0); }  // v2-v3
            if (ix == 0)      { m.bf_cell.push_back((int32_t)c + 1); m.bf_local.push_back(2); }  // v0-v2
        }
}

void make_unit_square_tri_slab(int nx, int gny, int row0, int nrows, const double lo[2], const double hi[2], Mesh &m) {
    const double dy = hi[1] - lo[1];
    double l2[2] = {lo[0], lo[1] + dy * (double)row0 / (double)gny};
    double h2[2] = {hi[0], lo[1] + dy * (double)(row0 + nrows) / (double)gny};
    make_unit_square_tri(nx, nrows, l2, h2, m, row0 == 0, row0 + nrows == gny);
    for (int jl = 0; jl <= nrows; ++jl)
        for (int ix = 0; ix <= nx; ++ix)
            m.coords[2 * ((int64_t)jl * (nx + 1) + ix) + 1] = lo[1] + dy * (double)(row0 + jl) / (double)gny;
    // the slab remembers the extent of the whole domain (functionals normalise by |Omega|)
    m.lo[1] = lo[1];
    m.hi[1] = hi[1];
}

void make_rectangle_quad(int nx, int ny, const double lo[2], const double hi[2], Mesh &m) {
    m = Mesh();
    m.kind = MESH_QUAD;
    m.nvpc = 4;
    lattice_coords(nx, ny, lo, hi, m);
    m.ncells = (int64_t)nx * ny;
    m.cells.resize(m.ncells * 4);
    for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
            int32_t v0 = iy * (nx + 1) + ix;
            int64_t c = (int64_t)iy * nx + ix;
            int32_t *a = &m.cells[c * 4];
            a[0] = v0; a[1] = v0 + 1; a[2] = v0 + nx + 1; a[3] = v0 + nx + 2;
            // basix quadrilateral facets: (v0,v1) (v0,v2) (v1,v3) (v2,v3)
            if (iy == 0)      { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(0); }
            if (ix == 0)      { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(1); }
            if (ix == nx - 1) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(2); }
            if (iy == ny - 1) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(3); }
        }
}

void make_box_hex(int nx, int ny, int nz, const double lo[3], const double hi[3], Mesh &m) {
    make_box_hex_slab(nx, ny, nz, 0, nz, lo, hi, m);
}

void make_box_hex_slab(int nx, int ny, int gnz, int k0, int nz, const double lo[3], const double hi[3], Mesh &m) {
    m = Mesh();
    m.kind = MESH_HEX;
    m.nvpc = 8;
    m.gdim = 3;
    m.n[0] = nx; m.n[1] = ny; m.n[2] = nz;
    for (int d = 0; d < 3; ++d) { m.lo[d] = lo[d]; m.hi[d] = hi[d]; }
    const int64_t sx = nx + 1, sy = ny + 1;
    m.nverts = sx * sy * (nz + 1);
    m.ncells = (int64_t)nx * ny * nz;
    m.coords.resize(m.nverts * 3);
    for (int iz = 0; iz <= nz; ++iz)
        for (int iy = 0; iy <= ny; ++iy)
            for (int ix = 0; ix <= nx; ++ix) {
                const int64_t v = ((int64_t)iz * sy + iy) * sx + ix;
                m.coords[3 * v] = lo[0] + (hi[0] - lo[0]) * (double)ix / (double)nx;
                m.coords[3 * v + 1] = lo[1] + (hi[1] - lo[1]) * (double)iy / (double)ny;
                m.coords[3 * v + 2] = lo[2] + (hi[2] - lo[2]) * (double)(k0 + iz) / (double)gnz;
            }
    m.cells.resize(m.ncells * 8);
    for (int iz = 0; iz < nz; ++iz)
        for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix) {
                const int64_t c = ((int64_t)iz * ny + iy) * nx + ix;
                const int64_t v0 = ((int64_t)iz * sy + iy) * sx + ix;
                int32_t *a = &m.cells[c * 8];
                for (int k = 0; k < 8; ++k)
                    a[k] = (int32_t)(v0 + (k & 1) + ((k >> 1) & 1) * sx + ((k >> 2) & 1) * sx * sy);
                if (iz == 0 && k0 == 0) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(0); }
                if (iy == 0)      { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(1); }
                if (ix == 0)      { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(2); }
                if (ix == nx - 1) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(3); }
                if (iy == ny - 1) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(4); }
                if (iz == nz - 1 && k0 + nz == gnz) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(5); }
            }
}

void make_from_arrays(int kind, int gdim, int64_t nverts, const double *coords, int64_t ncells, const int32_t *cells, Mesh &m) {
    static const int tri_f[3][4] = {{1, 2, -1, -1}, {0, 2, -1, -1}, {0, 1, -1, -1}};
    static const int quad_f[4][4] = {{0, 1, -1, -1}, {0, 2, -1, -1}, {1, 3, -1, -1}, {2, 3, -1, -1}};
    static const int hex_f[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {0, 2, 4, 6}, {1, 3, 5, 7}, {2, 3, 6, 7}, {4, 5, 6, 7}};
    static const int int_f[2][4] = {{0, -1, -1, -1}, {1, -1, -1, -1}};
    m = Mesh();
    m.kind = kind;
    m.gdim = gdim;
    m.lattice = false;
    const int (*ft)[4] = nullptr;
    int nf = 0;
    switch (kind) {
        case MESH_INTERVAL: m.nvpc = 2; nf = 2; ft = int_f; break;
        case MESH_TRI: m.nvpc = 3; nf = 3; ft = tri_f; break;
        case MESH_QUAD: m.nvpc = 4; nf = 4; ft = quad_f; break;
        case MESH_HEX: m.nvpc = 8; nf = 6; ft = hex_f; break;
        default: throw LayoutError{FEMO_EINVAL, "unknown mesh kind"};
    }
    if (nverts < 1 || ncells < 1 || gdim < 1 || gdim > 3) throw LayoutError{FEMO_EINVAL, "empty mesh"};
    m.nverts = nverts;
    m.ncells = ncells;
    m.coords.assign(coords, coords + nverts * gdim);
    m.cells.assign(cells, cells + ncells * m.nvpc);
    for (int d = 0; d < gdim; ++d) { m.lo[d] = 1e300; m.hi[d] = -1e300; }
    for (int64_t v = 0; v < nverts; ++v)
        for (int d = 0; d < gdim; ++d) {
            m.lo[d] = std::min(m.lo[d], coords[v * gdim + d]);
            m.hi[d] = std::max(m.hi[d], coords[v * gdim + d]);
        }
    for (int64_t k = 0; k < ncells * m.nvpc; ++k)
        if (cells[k] < 0 || cells[k] >= nverts) throw LayoutError{FEMO_EINVAL, "cell vertex index out of range"};
    struct Fk { int32_t v[4]; int64_t owner; };
    std::vector<Fk> f((size_t)ncells * nf);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncells; ++c)
        for (int l = 0; l < nf; ++l) {
            Fk &e = f[c * nf + l];
            for (int k = 0; k < 4; ++k) e.v[k] = ft[l][k] < 0 ? -1 : cells[c * m.nvpc + ft[l][k]];
            std::sort(e.v, e.v + 4);
            e.owner = c * nf + l;
        }
    auto less = [](const Fk &a, const Fk &b) {
        for (int k = 0; k < 4; ++k)
            if (a.v[k] != b.v[k]) return a.v[k] < b.v[k];
        return a.owner < b.owner;
    };
    auto same = [](const Fk &a, const Fk &b) { return a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2] && a.v[3] == b.v[3]; };
    std::sort(f.begin(), f.end(), less);
    std::vector<int64_t> ext;
    for (size_t i = 0; i < f.size();) {
        size_t j = i + 1;
        while (j < f.size() && same(f[i], f[j])) ++j;
        if (j - i == 1) ext.push_back(f[i].owner);
        i = j;
    }
    std::sort(ext.begin(), ext.end());
    m.bf_cell.reserve(ext.size());
    m.bf_local.reserve(ext.size());
    for (int64_t o : ext) {
        m.bf_cell.push_back((int32_t)(o / nf));
        m.bf_local.push_back((int32_t)(o % nf));
    }
}

void Mesh::build_edges() {
    if (kind != MESH_TRI) throw LayoutError{FEMO_EINVAL, "edges are built for triangle meshes"};
    if (!cell_edges.empty()) return;
    std::vector<int64_t> keys((size_t)ncells * 3);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncells; ++c)
        for (int i = 0; i < 3; ++i) {
            const int64_t a = cells[c * 3 + (i + 1) % 3], b = cells[c * 3 + (i + 2) % 3];
            keys[c * 3 + i] = (std::min(a, b) << 32) | std::max(a, b);
        }
    std::vector<int64_t> uniq(keys);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    nedges = (int64_t)uniq.size();
    if (nverts + nedges > kInt32Max) throw LayoutError{FEMO_ELIMIT, "P2 space exceeds int32 dofs"};
    edge_verts.resize(nedges * 2);
    for (int64_t e = 0; e < nedges; ++e) {
        edge_verts[2 * e] = (int32_t)(uniq[e] >> 32);
        edge_verts[2 * e + 1] = (int32_t)(uniq[e] & 0xffffffffLL);
    }
    cell_edges.resize((size_t)ncells * 3);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ncells * 3; ++k)
        cell_edges[k] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), keys[k]) - uniq.begin());
}

void make_annulus_tri(int nr, int nth, double r0, double r1, Mesh &m) {
    m = Mesh();
    m.kind = MESH_TRI;
    m.nvpc = 3;
    m.gdim = 2;
    m.lattice = false;
    m.n[0] = nr;
    m.n[1] = nth;
    m.lo[0] = r0; m.hi[0] = r1;
    m.lo[1] = 0.0; m.hi[1] = 2.0 * 3.141592653589793;
    m.nverts = (int64_t)(nr + 1) * nth;
    m.ncells = 2LL * nr * nth;
    m.coords.resize(m.nverts * 2);
    const double pi = 3.141592653589793;
    for (int ir = 0; ir <= nr; ++ir)
        for (int it = 0; it < nth; ++it) {
            const double r = r0 + (r1 - r0) * (double)ir / (double)nr;
            const double th = 2.0 * pi * (double)it / (double)nth;
            const int64_t v = (int64_t)ir * nth + it;
            m.coords[2 * v] = r * std::cos(th);
            m.coords[2 * v + 1] = r * std::sin(th);
        }
    m.cells.resize(m.ncells * 3);
    for (int ir = 0; ir < nr; ++ir)
        for (int it = 0; it < nth; ++it) {
            const int tp = (it + 1) % nth;
            const int32_t v0 = ir * nth + it, v1 = (ir + 1) * nth + it, v2 = ir * nth + tp, v3 = (ir + 1) * nth + tp;
            const int64_t c = 2LL * ((int64_t)ir * nth + it);
            int32_t *a = &m.cells[c * 3];
            a[0] = v0; a[1] = v1; a[2] = v3;
            a[3] = v0; a[4] = v2; a[5] = v3;
            if (ir == nr - 1) { m.bf_cell.push_back((int32_t)c); m.bf_local.push_back(0); }      // outer circle: v1-v3
            if (ir == 0)      { m.bf_cell.push_back((int32_t)c + 1); m.bf_local.push_back(2); }  // inner circle: v0-v2
        }
}

void make_interval(int n, double x0, double x1, Mesh &m) {
    m = Mesh();
    m.kind = MESH_INTERVAL;
    m.nvpc = 2;
    m.gdim = 1;
    m.n[0] = n;
    m.lo[0] = x0;
    m.hi[0] = x1;
    m.nverts = n + 1;
    m.ncells = n;
    m.coords.resize(n + 1);
    for (int i = 0; i <= n; ++i) m.coords[i] = x0 + (x1 - x0) * (double)i / (double)n;
    m.cells.resize(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { m.cells[2 * i] = i; m.cells[2 * i + 1] = i + 1; }
    // facet k of an interval is its vertex k
    m.bf_cell.push_back(0); m.bf_local.push_back(0);
    m.bf_cell.push_back(n - 1); m.bf_local.push_back(1);
}

void Space::init(const Mesh &m, int element_, int block_) {
    element = element_;
    block = block_;
    if (element == EL_DG0) {
        ndofs = m.ncells * block;
        ndpc = block;
    } else if (element == EL_P2) {
        if (m.kind != MESH_TRI) throw LayoutError{FEMO_EINVAL, "P2 spaces need a triangle mesh"};
        if (m.cell_edges.empty()) throw LayoutError{FEMO_ESTATE, "P2 space: mesh edges were not built"};
        block = 1;
        ndofs = m.nverts + m.nedges;
        ndpc = 6;
    } else if (element == EL_RMP) {
        if (m.kind != MESH_TRI) throw LayoutError{FEMO_EINVAL, "Reissner-Mindlin plate spaces need a triangle mesh"};
        if (m.cell_edges.empty()) throw LayoutError{FEMO_ESTATE, "RM plate space: mesh edges were not built"};
        block = 1;
        ndofs = 3 * m.nverts + m.nedges;
        ndpc = 12;
    } else {
        if (element == EL_HERMITE3) block = 2;  // (value, reference derivative) per vertex
        ndofs = m.nverts * block;
        ndpc = m.nvpc * block;
    }
}

// ---------------------------------------------------------------------------
// incidences: for every dof of a space, the (block, local index, entity) triples
// that touch it, in ascending (block, local, entity) order.
// ---------------------------------------------------------------------------
struct Incidence {
    std::vector<int64_t> ptr;   // ndofs+1
    std::vector<int32_t> ent;   // entity within its block
    std::vector<int16_t> blk;   // block id
    std::vector<int16_t> loc;   // local dof index
};

static void build_incidence(const Mesh &m, const Space &s, const std::vector<IntegralBlock> &blocks, Incidence &inc) {
    const int nd = s.ndpc;
    inc.ptr.assign(s.ndofs + 1, 0);
    std::vector<int32_t> d(nd);
    for (size_t b = 0; b < blocks.size(); ++b)
        for (int64_t e = 0; e < blocks[b].ne; ++e) {
            int64_t cell = blocks[b].ent_cell ? blocks[b].ent_cell[e] : e;
            s.cell_dofs(m, cell, d.data());
            for (int a = 0; a < nd; ++a) inc.ptr[d[a] + 1]++;
        }
    for (int64_t i = 0; i < s.ndofs; ++i) inc.ptr[i + 1] += inc.ptr[i];
    int64_t tot = inc.ptr[s.ndofs];
    inc.ent.resize(tot);
    inc.blk.resize(tot);
    inc.loc.resize(tot);
    std::vector<int64_t> pos(inc.ptr.begin(), inc.ptr.end() - 1);
    // (block, local, entity) loop order => every dof's list is sorted the same way
    for (size_t b = 0; b < blocks.size(); ++b)
        for (int a = 0; a < nd; ++a)
            for (int64_t e = 0; e < blocks[b].ne; ++e) {
                int64_t cell = blocks[b].ent_cell ? blocks[b].ent_cell[e] : e;
                const int32_t dof = s.cell_dof(m, cell, a);
                int64_t q = pos[dof]++;
                inc.ent[q] = (int32_t)e;
                inc.blk[q] = (int16_t)b;
                inc.loc[q] = (int16_t)a;
            }
}

void build_vecmap(const Mesh &m, const Space &s, const std::vector<IntegralBlock> &blocks, VecMap &out) {
    Incidence inc;
    build_incidence(m, s, blocks, inc);
    std::vector<int64_t> off(blocks.size() + 1, 0);
    for (size_t b = 0; b < blocks.size(); ++b) off[b + 1] = off[b] + blocks[b].ne * s.ndpc;
    out.n = s.ndofs;
    out.scratch_len = off[blocks.size()];
    out.ncontrib = inc.ptr[s.ndofs];
    if (out.scratch_len > kInt32Max || out.ncontrib > kInt32Max) throw LayoutError{FEMO_ELIMIT, "vector map exceeds int32"};
    out.ptr.resize(s.ndofs + 1);
    out.src.resize(out.ncontrib);
    for (int64_t i = 0; i <= s.ndofs; ++i) out.ptr[i] = (int32_t)inc.ptr[i];
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < out.ncontrib; ++q) {
        int b = inc.blk[q];
        out.src[q] = (int32_t)(off[b] + (int64_t)inc.loc[q] * blocks[b].ne + inc.ent[q]);
    }
}

void build_pattern(const Mesh &m, const Space &rs, const Space &cs, const std::vector<IntegralBlock> &blocks,
                   Pattern &out) {
    Incidence inc;
    build_incidence(m, rs, blocks, inc);
    const int nr = rs.ndpc, nc = cs.ndpc;
    const int64_t N = rs.ndofs;
    std::vector<int64_t> off(blocks.size() + 1, 0);
    for (size_t b = 0; b < blocks.size(); ++b) off[b + 1] = off[b] + blocks[b].ne * (int64_t)nr * nc;
    out = Pattern();
    out.nrows = N;
    out.ncols = cs.ndofs;
    out.scratch_len = off[blocks.size()];
    out.ncontrib = inc.ptr[N] * nc;
    if (out.scratch_len > kInt32Max || out.ncontrib > kInt32Max)
        throw LayoutError{FEMO_ELIMIT, "matrix gather map exceeds int32 (use a smaller mesh per GPU)"};

    // pass 1: nnz per row
    std::vector<int64_t> rp(N + 1, 0);
#pragma omp parallel
    {
        std::vector<int32_t> cols, d(nc);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            cols.clear();
            for (int64_t q = inc.ptr[i]; q < inc.ptr[i + 1]; ++q) {
                const IntegralBlock &B = blocks[inc.blk[q]];
                int64_t cell = B.ent_cell ? B.ent_cell[inc.ent[q]] : inc.ent[q];
                cs.cell_dofs(m, cell, d.data());
                cols.insert(cols.end(), d.begin(), d.end());
            }
            std::sort(cols.begin(), cols.end());
            rp[i + 1] = std::unique(cols.begin(), cols.end()) - cols.begin();
        }
    }
    for (int64_t i = 0; i < N; ++i) rp[i + 1] += rp[i];
    out.nnz = rp[N];
    if (out.nnz > kInt32Max) throw LayoutError{FEMO_ELIMIT, "nnz exceeds int32"};
    out.rowptr.resize(N + 1);
    for (int64_t i = 0; i <= N; ++i) out.rowptr[i] = (int32_t)rp[i];
    out.col.resize(out.nnz);
    out.gptr.resize(out.nnz + 1);
    out.gsrc.resize(out.ncontrib);
    out.gptr[out.nnz] = (int32_t)out.ncontrib;

    // pass 2: columns + gather map; row i's contributions start at inc.ptr[i]*nc
#pragma omp parallel
    {
        std::vector<int64_t> pairs;  // (col << 32) | src
        std::vector<int32_t> d(nc);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            pairs.clear();
            for (int64_t q = inc.ptr[i]; q < inc.ptr[i + 1]; ++q) {
                int b = inc.blk[q];
                const IntegralBlock &B = blocks[b];
                int64_t e = inc.ent[q];
                int64_t cell = B.ent_cell ? B.ent_cell[e] : e;
                cs.cell_dofs(m, cell, d.data());
                for (int c = 0; c < nc; ++c) {
                    int64_t src = off[b] + ((int64_t)inc.loc[q] * nc + c) * B.ne + e;
                    pairs.push_back(((int64_t)d[c] << 32) | src);
                }
            }
            std::sort(pairs.begin(), pairs.end());
            int64_t t = rp[i] - 1, g = inc.ptr[i] * nc;
            int32_t last = -1;
            for (size_t k = 0; k < pairs.size(); ++k, ++g) {
                int32_t cj = (int32_t)(pairs[k] >> 32);
                if (cj != last) {
                    ++t;
                    out.col[t] = cj;
                    out.gptr[t] = (int32_t)g;
                    last = cj;
                }
                out.gsrc[g] = (int32_t)(pairs[k] & 0xffffffffLL);
            }
        }
    }

    // transpose bookkeeping
    out.square_symmetric = (&rs == &cs) || (rs.element == cs.element && rs.block == cs.block && rs.ndofs == cs.ndofs);
    out.t_perm.resize(out.nnz);
    if (out.square_symmetric) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i)
            for (int64_t t = rp[i]; t < rp[i + 1]; ++t) {
                int32_t j = out.col[t];
                const int32_t *b = &out.col[rp[j]], *e = &out.col[rp[j + 1]];
                const int32_t *f = std::lower_bound(b, e, (int32_t)i);
                out.t_perm[t] = (int32_t)(f - out.col.data());  // entry (j,i)
            }
    } else {
        out.t_rowptr.assign(out.ncols + 1, 0);
        for (int64_t t = 0; t < out.nnz; ++t) out.t_rowptr[out.col[t] + 1]++;
        for (int64_t j = 0; j < out.ncols; ++j) out.t_rowptr[j + 1] += out.t_rowptr[j];
        out.t_col.resize(out.nnz);
        std::vector<int32_t> pos(out.t_rowptr.begin(), out.t_rowptr.end() - 1);
        for (int64_t i = 0; i < N; ++i)
            for (int64_t t = rp[i]; t < rp[i + 1]; ++t) {
                int32_t q = pos[out.col[t]]++;
                out.t_col[q] = (int32_t)i;
                out.t_perm[q] = (int32_t)t;
            }
    }
    build_rowblocks(out.rowptr, out.nrows, out.rb);
    if (!out.square_symmetric) build_rowblocks(out.t_rowptr, out.ncols, out.t_rb);
}

void build_bsr3(Pattern &P) {
    P.b_rowptr.clear(); P.b_col.clear(); P.b_perm.clear(); P.b_rb.clear();
    if (P.nrows != P.ncols || P.nrows % 3 != 0 || P.nnz % 9 != 0) return;
    const int64_t nb = P.nrows / 3;
    std::vector<int32_t> brp(nb + 1, 0), bcol, bperm;
    bcol.reserve(P.nnz / 9);
    bperm.reserve(P.nnz);
    for (int64_t I = 0; I < nb; ++I) {
        const int32_t s0 = P.rowptr[3 * I], s1 = P.rowptr[3 * I + 1], s2 = P.rowptr[3 * I + 2], e2 = P.rowptr[3 * I + 3];
        const int32_t len = s1 - s0;
        if (len % 3 != 0 || s2 - s1 != len || e2 - s2 != len) return;          // rows of a triple must match
        for (int32_t k = 0; k < len; k += 3) {
            const int32_t c = P.col[s0 + k];
            if (c % 3 != 0) return;
            for (int r = 0; r < 3; ++r) {
                const int32_t base = P.rowptr[3 * I + r] + k;
                for (int j = 0; j < 3; ++j) {
                    if (P.col[base + j] != c + j) return;                       // incomplete block: keep CSR only
                    bperm.push_back(base + j);
                }
            }
            bcol.push_back(c / 3);
        }
        if (len / 3 > kBsrBlocks) return;                                      // a block row must fit one CTA
        brp[I + 1] = (int32_t)bcol.size();
    }
    P.b_rowptr = std::move(brp);
    P.b_col = std::move(bcol);
    P.b_perm = std::move(bperm);
    P.b_rb.push_back(0);
    int64_t r = 0;
    while (r < nb) {
        int64_t e = r + 1;
        while (e < nb && e - r < kBsrRows && P.b_rowptr[e + 1] - P.b_rowptr[r] <= kBsrBlocks) ++e;
        P.b_rb.push_back((int32_t)e);
        r = e;
    }
}

void build_rowblocks(const std::vector<int32_t> &rowptr, int64_t nrows, std::vector<int32_t> &rb) {
    rb.clear();
    rb.push_back(0);
    int64_t r = 0;
    while (r < nrows) {
        int64_t e = r + 1;  // a block always holds at least one row (longer rows are strided in-kernel)
        while (e < nrows && e - r < kSpmvRows && rowptr[e + 1] - rowptr[r] <= kSpmvCap) ++e;
        rb.push_back((int32_t)e);
        r = e;
    }
}

// ---------------------------------------------------------------------------
void gauss_legendre_01(int m, std::vector<double> &x, std::vector<double> &w) {
    x.resize(m);
    w.resize(m);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < m; ++i) {
        double z = std::cos(pi * (i + 0.75) / (m + 0.5));  // Chebyshev guess on [-1,1]
        double pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < m; ++j) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = m * (z * p1 - p2) / (z * z - 1.0);
            double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        // descending z -> store ascending on [0,1]
        x[m - 1 - i] = 0.5 * (z + 1.0);
        w[m - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);  // = 0.5 * 2/((1-z^2) pp^2)
    }
}

}  // namespace femo
