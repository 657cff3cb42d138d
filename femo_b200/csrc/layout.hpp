// layout.hpp -- host-side mesh, dofmap, sparsity pattern and segmented-reduction
// map builders (pure C++, no CUDA; exercised by the CPU test-suite).
//
// Replaces what the reference obtains from dolfinx mesh creation
// (femo/fea/utils_dolfinx.py:136-153), FunctionSpace dofmaps and
// create_matrix / sparsity patterns (utils_dolfinx.py:390 and implicitly in
// every assemble_matrix(form) at :185,195,575).  Numbering is the canonical
// lattice order described in DESIGN.md; every integer array built here is
// compared with `==` against the oracle in tests/test_layout.py.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace femo {

enum MeshKind { MESH_INTERVAL = 1, MESH_TRI = 2, MESH_QUAD = 3, MESH_HEX = 4 };
enum Element { EL_DG0 = 0, EL_VERTEX = 1 /* P1 / Q1: one node per vertex */, EL_HERMITE3 = 2,
               EL_P2 = 3 /* triangles: vertex nodes then edge-midpoint nodes */,
               EL_RMP = 4 /* Reissner-Mindlin plate, mixed P2 (deflection) x P1^2 (rotations) on triangles:
                             [w vertices | w edge midpoints | (theta_x, theta_y) per vertex], 12 dofs per cell */ };

struct Mesh {
    int kind = 0, gdim = 0, nvpc = 0;
    int64_t ncells = 0, nverts = 0;
    int n[3] = {0, 0, 0};
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    std::vector<double> coords;              // nverts*gdim, AoS
    std::vector<int32_t> cells;              // ncells*nvpc, AoS
    std::vector<int32_t> bf_cell, bf_local;  // exterior facets, sorted by (cell, local facet)
    std::vector<int32_t> cell_tag;           // optional subdomain id per cell (meshtags of dimension tdim)
    bool lattice = true;                     // (n[0]+1) x (n[1]+1) vertex lattice in row-major order
    // edges of a triangle mesh (built on demand for P2 spaces): edge e joins edge_verts[2e] < edge_verts[2e+1];
    // edges are numbered in lexicographic order of that pair; local edge i of a cell is opposite vertex i (basix)
    int64_t nedges = 0;
    std::vector<int32_t> edge_verts;         // nedges*2
    std::vector<int32_t> cell_edges;         // ncells*3
    void build_edges();
};

// ext_bottom / ext_top: whether the y = lo / y = hi edge is a true domain boundary (false for the
// interior interfaces of a y-slab of a partitioned mesh: no exterior facets are generated there)
void make_unit_square_tri(int nx, int ny, const double lo[2], const double hi[2], Mesh &m, bool ext_bottom = true,
                          bool ext_top = true);
// y-slab of the (nx x gny)-cell lattice on [lo,hi]: cell rows [row0, row0+nrows); vertex coordinates are
// computed with the GLOBAL formula lo + (hi-lo)*j/gny so they agree bit for bit with the unpartitioned mesh
void make_unit_square_tri_slab(int nx, int gny, int row0, int nrows, const double lo[2], const double hi[2], Mesh &m);
void make_rectangle_quad(int nx, int ny, const double lo[2], const double hi[2], Mesh &m);
void make_interval(int n, double x0, double x1, Mesh &m);
// (nx x ny x nz)-cell hexahedral lattice on the box [lo,hi]: vertex (ix,iy,iz) -> (iz*(ny+1)+iy)*(nx+1)+ix,
// cell (ix,iy,iz) -> (iz*ny+iy)*nx+ix with the tensor-product vertex order of basix (x fastest) and
// facets 0 z=lo, 1 y=lo, 2 x=lo, 3 x=hi, 4 y=hi, 5 z=hi [upstream, from memory]
void make_box_hex(int nx, int ny, int nz, const double lo[3], const double hi[3], Mesh &m);
// z-slab of the (nx x ny x gnz)-cell box: cell layers [k0, k0+nk); z coordinates from the GLOBAL formula; the
// z = lo / z = hi faces carry exterior facets only where the slab touches the true boundary
void make_box_hex_slab(int nx, int ny, int gnz, int k0, int nk, const double lo[3], const double hi[3], Mesh &m);
// unstructured mesh from caller arrays (what the reference reads from msh2xdmf output, utils_dolfinx.py:69-123):
// kind = MeshKind, cells with nvpc vertices in basix order; exterior facets = facets owned by exactly one cell,
// sorted by (cell, local facet); `lattice` is false (no geometric multigrid hierarchy)
void make_from_arrays(int kind, int gdim, int64_t nverts, const double *coords, int64_t ncells, const int32_t *cells, Mesh &m);
// periodic polar lattice on the annulus r0 <= r <= r1: node (ir, ith) -> ir*nth + ith
void make_annulus_tri(int nr, int nth, double r0, double r1, Mesh &m);

// A finite-element space on a mesh: dof = block*node + comp, local index a*block+comp.
struct Space {
    int element = EL_VERTEX;
    int block = 1;
    int64_t ndofs = 0;
    int ndpc = 0;  // dofs per cell
    void init(const Mesh &m, int element_, int block_);
    inline void cell_dofs(const Mesh &m, int64_t cell, int32_t *out) const {
        if (element == EL_DG0) {
            for (int c = 0; c < block; ++c) out[c] = (int32_t)(cell * block + c);
        } else if (element == EL_P2) {
            for (int a = 0; a < 3; ++a) {
                out[a] = m.cells[cell * 3 + a];
                out[3 + a] = (int32_t)(m.nverts + m.cell_edges[cell * 3 + a]);
            }
        } else if (element == EL_RMP) {
            for (int a = 0; a < 12; ++a) out[a] = cell_dof(m, cell, a);
        } else {
            const int32_t *v = &m.cells[cell * m.nvpc];
            for (int a = 0; a < m.nvpc; ++a)
                for (int c = 0; c < block; ++c) out[a * block + c] = v[a] * block + c;
        }
    }
    // dof carried by local index a of `cell`
    inline int32_t cell_dof(const Mesh &m, int64_t cell, int a) const {
        if (element == EL_DG0) return (int32_t)(cell * block + a);
        if (element == EL_P2) return a < 3 ? m.cells[cell * 3 + a] : (int32_t)(m.nverts + m.cell_edges[cell * 3 + a - 3]);
        if (element == EL_RMP) {
            if (a < 3) return m.cells[cell * 3 + a];
            if (a < 6) return (int32_t)(m.nverts + m.cell_edges[cell * 3 + a - 3]);
            return (int32_t)(m.nverts + m.nedges + 2 * (int64_t)m.cells[cell * 3 + (a - 6) / 2] + (a - 6) % 2);
        }
        return m.cells[cell * m.nvpc + a / block] * block + a % block;
    }
};

// One integral of a form: all cells, or a list of facets (each owned by a cell).
// Element tensors of block b live in scratch at  off + k*ne + e  (SoA planes).
struct IntegralBlock {
    int64_t ne = 0;
    const int32_t *ent_cell = nullptr;  // nullptr -> entity e is cell e
};

// Sorted segmented-reduction map for vectors: out[i] = sum scratch[src[ptr[i]..ptr[i+1])]
struct VecMap {
    int64_t n = 0, ncontrib = 0, scratch_len = 0;
    std::vector<int32_t> ptr, src;
};

// CSR pattern + matrix gather map + transpose bookkeeping.
struct Pattern {
    int64_t nrows = 0, ncols = 0, nnz = 0, ncontrib = 0, scratch_len = 0;
    std::vector<int32_t> rowptr, col;
    std::vector<int32_t> gptr, gsrc;  // gptr has nnz+1 entries
    bool square_symmetric = false;    // structurally symmetric: transpose shares rowptr/col
    std::vector<int32_t> t_rowptr, t_col;  // transposed CSR (empty when square_symmetric)
    std::vector<int32_t> t_perm;           // transposed entry tt takes vals[t_perm[tt]]
    // SpMV row blocks: block k covers rows [rb[k], rb[k+1]) with <= kSpmvCap nnz and <= kSpmvRows rows
    std::vector<int32_t> rb, t_rb;
    // 3x3 block view of a square pattern whose rows / columns come in complete node triples (vector states in 3-D):
    // block row I has blocks b_rowptr[I] .. b_rowptr[I+1], block k couples node I with node b_col[k]; its 9 values
    // (row-major) are the CSR values at b_perm[9k .. 9k+9).  b_rb: CTA work units (<= kBsrBlocks blocks each).
    std::vector<int32_t> b_rowptr, b_col, b_perm, b_rb;
};
constexpr int kBsrBlocks = 224;   // 3x3 blocks staged per CTA (224 * 72 B = 15.75 KB of values)
constexpr int kBsrRows = 64;      // block rows per CTA
// fills the BSR view of `p` (no-op unless the pattern is made of complete 3x3 blocks)
void build_bsr3(Pattern &p);

constexpr int kSpmvCap = 2048;   // products staged in shared memory per row block
constexpr int kSpmvRows = 512;   // rows per block (row pointers staged too)
void build_rowblocks(const std::vector<int32_t> &rowptr, int64_t nrows, std::vector<int32_t> &rb);

// Build the vector map of `space` over `blocks` (K = ndpc planes per block).
void build_vecmap(const Mesh &m, const Space &space, const std::vector<IntegralBlock> &blocks, VecMap &out);
// Build pattern + gather map for rows in `rs`, cols in `cs` (K = ndpc_r*ndpc_c planes per block).
void build_pattern(const Mesh &m, const Space &rs, const Space &cs, const std::vector<IntegralBlock> &blocks,
                   Pattern &out);

// m-point Gauss-Legendre on [0,1] (own Newton iteration; no tables).
void gauss_legendre_01(int m, std::vector<double> &x, std::vector<double> &w);

struct LayoutError {
    int code;
    std::string msg;
};

}  // namespace femo
