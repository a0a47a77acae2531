// Internal definitions shared by the translation units of libferrite_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ferrite_b200.h"

// ---- error handling ---------------------------------------------------------------------
void fb2_set_error(const char* fmt, ...);
int fb2_fail(int code, const char* fmt, ...);

#define FB2_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            return fb2_fail(e__ == cudaErrorMemoryAllocation ? FB2_ERR_OOM : FB2_ERR_CUDA,      \
                            "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,  \
                            __LINE__);                                                          \
        }                                                                                       \
    } while (0)

#define FB2_CHECK(cond, code, ...)                  \
    do {                                            \
        if (!(cond)) return fb2_fail(code, __VA_ARGS__); \
    } while (0)

#define FB2_NEED_DEVICE(ctx)                                                                      \
    do {                                                                                          \
        if ((ctx)->device < 0)                                                                    \
            return fb2_fail(FB2_ERR_CUDA, "%s needs a CUDA device but the context is host-only; " \
                            "libferrite_b200 has no CPU fallback", __func__);                     \
    } while (0)

#define FB2_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != FB2_OK) return rc__; \
    } while (0)

// ---- reference-shape / interpolation tables (tables.cpp) -----------------------------------
struct RefShapeInfo {
    int celltype;
    int rdim;
    int nvertices;
    int nedges;
    int nfaces;
    int edges[12][2];      // 0-based local vertex ids
    int faces[6][4];       // 0-based, -1 padded
    int face_nverts[6];
    int face_edges[6][4];  // edge numbers (0-based) around each face
};
const RefShapeInfo* fb2_refshape(int celltype);

struct LagrangeInfo {
    int celltype;
    int order;
    int nbase;
    int rdim;
    double refcoords[27][3];
    int nvertexdofs;  // per vertex (0 or 1)
    int nedgedofs;    // interior dofs per edge
    int nfacedofs;    // interior dofs per face
    int nvolumedofs;
    int edge_first;   // 0-based index of the first edge-interior dof
    int face_first;
    int vol_first;
};
// returns false if (celltype, order) is outside the supported menu
bool fb2_lagrange(int celltype, int order, LagrangeInfo* out);
// N[n], dN[n][rdim] at xi
void fb2_lagrange_eval(const LagrangeInfo& ip, const double* xi, double* N, double* dN);
// local dof lists (0-based scalar basis indices) of boundary entities: kind = FB2_BC_FACET/FACE/EDGE/VERTEX
std::vector<std::vector<int>> fb2_boundarydof_indices(const LagrangeInfo& ip, int kind);
// quadrature: default rule of the reference per shape; points are nq x rdim
bool fb2_quadrature(int celltype, int order, std::vector<double>* w, std::vector<double>* pts);

// ---- objects --------------------------------------------------------------------------------
uint64_t fb2_next_uid();   // process-wide, starts at 1, never repeats (api.cu)
struct fb2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 148;
    int64_t launches = 0;
    int* d_errflag = nullptr;   // [0] = code, [1] = cell id
    int* h_errflag = nullptr;   // pinned
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    // copy streams + events of the streamed host-buffer entry point (lazy)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_pool[64] = {};
    bool ev_ready = false;
};

struct fb2_grid {
    fb2_ctx* ctx = nullptr;
    int celltype = 0;
    int64_t ncells = 0, nnodes = 0;
    int nnpc = 0, sdim = 0;
    std::vector<int64_t> cells;   // nnpc x ncells, 1-based
    std::vector<double> xyz;      // sdim x nnodes
    std::map<std::string, std::vector<int64_t>> facetsets;  // flattened (cell, facet) pairs, 1-based, sorted
    bool generated = false;
    int64_t nel[3] = {1, 1, 1};
    double left[3] = {0, 0, 0}, right[3] = {0, 0, 0};
    // device
    int64_t ncells_pad = 0;       // multiple of 32
    int32_t* d_conn = nullptr;    // SoA [nnpc][ncells_pad], 0-based node ids
    double* d_xyz = nullptr;      // [nnodes][xstride]
    double* d_xyz_stage = nullptr;  // staging buffer of fb2_grid_upload_coordinates_async
    int xstride = 0;              // 4 for sdim 3, 2 for sdim 2, 1 for sdim 1
    // structured view of a hexahedral grid that is NOT stored in generate_grid order (the local grid of a block partition:
    // cells ordered [interface | interior | halo]): the cells form a box of sv_nel cells, box position -> cell id in
    // sv_cellmap (-1 = no such cell), and node (a, b, c) of the box has the id a + (sv_nel[0]+1) (b + (sv_nel[1]+1) c).
    bool structured = false;
    int64_t sv_nel[3] = {0, 0, 0};
    std::vector<int32_t> sv_cellmap;
    int32_t* d_sv_cellmap = nullptr;
};
int fb2_grid_upload(fb2_grid* g);
int fb2_host_threads();   // threads for the independent host loops of the set-up (host_grid.cpp)
// node coordinates of generate_grid / perturb, one definition for fb2_grid_generate and the rank-local set-up (host_grid.cpp)
void fb2_generated_corners(int dim, const double* lo, const double* hi, int* nc, double* refcoords, double* corner);
void fb2_generated_node(int dim, const int64_t* nn, int nc, const double* refcoords, const double* corner, const int64_t* idx, double* x);
double fb2_perturb_delta(int64_t id, int d, double amplitude, double h);
int fb2_grid_upload_xyz(fb2_grid* g);

struct fb2_dh {
    fb2_grid* grid = nullptr;
    std::vector<fb2_field> fields;
    std::vector<LagrangeInfo> ips;
    int64_t ndofs = 0;
    int ndpc = 0;
    std::vector<int32_t> cell_dofs;  // host, ndpc x ncells, 0-based
    int32_t* d_cell_dofs = nullptr;  // SoA [ndpc][ncells_pad], 0-based
    uint64_t generation = 0;         // bumped by renumber!: caches derived from cell_dofs compare it
    int field_offset(int f) const {
        int o = 0;
        for (int i = 0; i < f; ++i) o += ips[i].nbase * fields[i].vdim;
        return o;
    }
};

struct fb2_pattern {
    fb2_dh* dh = nullptr;
    int64_t n = 0, nnz = 0;
    int64_t* d_colptr = nullptr;   // [n+1], 0-based offsets
    int32_t* d_rowval = nullptr;   // [nnz], 0-based
    int64_t* d_diag = nullptr;     // [n] position of the diagonal entry, -1 if absent
    bool structurally_symmetric = false;
    int max_col_len = 0;
    // solver support (solver.cu, lazy): position of the transposed entry (j,i) for every stored (i,j); -1 if absent
    int64_t* d_tperm = nullptr;
    int tperm_state = 0;           // 0 not built, 1 built and complete (structurally symmetric), 2 built with gaps
    double* d_work = nullptr;      // CG work vectors
    size_t work_count = 0;
};

struct fb2_cv {
    fb2_ctx* ctx = nullptr;
    uint64_t uid = 0;             // never reused: identifies the tables resident in the device's __constant__ bank
    int celltype = 0;
    int rdim = 0, nq = 0, nb = 0, vdim = 1, ngeo = 0;
    int ip_order = 0, geo_order = 0, qr_order = 0;
    // host tables, q-major: N[q][i], dN[q][i][d], M[q][j], dM[q][j][d], w[q], pts[q][d]
    std::vector<double> N, dN, M, dM, w, pts;
    double* d_tables = nullptr;   // packed [w | N | dN | M | dM] on the device
    int tables_device = -1;       // device that holds d_tables (the first one that used this CellValues); -1 = none yet
    size_t tables_count = 0;
};

struct fb2_part;

// Tile schedule of the scalar thread-per-cell kernels (tiles.cu)
struct TileSchedule {
    int TC = 0;               // cells per tile
    int64_t ntiles = 0;
    int nslots = 0;           // NSYM + NB shared-memory slots per cell
    int max_ent = 0, max_src = 0;  // max (padded) entries / sources of a tile: shared-memory staging sizes
    int64_t nentries = 0;
    int32_t* d_conn = nullptr;        // [ntiles][nnpc][TC] node ids, tile-ordered SoA (padded with the last cell)
    int32_t* d_ncells = nullptr;      // [ntiles]
    int32_t* d_cell_ids = nullptr;    // [ntiles][TC] grid cell id (error reporting)
    int64_t* d_tile_base = nullptr;   // [ntiles] smallest nzval position touched by the tile
    int64_t* d_ent_ptr = nullptr;     // [ntiles+1]
    uint2* d_rec = nullptr;           // per entry {target | flags, first source | nsources << 16}, see tiles.cu
    int64_t* d_src_ptr = nullptr;     // [ntiles+1]
    uint16_t* d_src = nullptr;        // shared-memory slot index: slot * TC + cell_local
};
void fb2_tiles_free(TileSchedule* S);

struct fb2_assembler {
    TileSchedule* tiles = nullptr;
    bool tiles_failed = false;
    int32_t* d_wfirst = nullptr;   // warp list of k_cell_scalar (lazy): first cell of every warp, 4 warps = 4 grid rows
    uint8_t* d_wcount = nullptr;   //   number of cells of the warp (0 for padding warps)
    int64_t nwarps = 0;
    bool map_complete = false;     // every (cell, i, j) has a pattern entry: unchecked scatter allowed
    // selective zero fill of the marching-tile kernel (lazy): sorted list of the columns that receive reduce-adds, for the
    // chunk length it was built for
    int march_zlz = 0;
    int64_t march_nzcols = 0;
    int32_t* d_march_zcols = nullptr;
    int march_state = 0;           // marching-tile kernel: 0 not checked, 1 usable (continuous Q1 numbering on a generated
                                   // hexahedral grid: every grid node carries ONE dof), 2 not usable
    fb2_dh* dh = nullptr;
    fb2_pattern* pat = nullptr;
    fb2_cv* cv = nullptr;
    int n = 0;                     // dofs per cell covered by the element (= ndpc)
    uint16_t* d_map = nullptr;     // [n*n][ncells_pad]: offset of row dof_i inside column dof_j, e = j*n + i
    uint16_t* d_mapc = nullptr;    // cell-major copy for k_cell_blocks: [ncells][ceil8(n*n)] (lazy)
    int32_t* d_dofc = nullptr;     // cell-major dofs for the CTA kernels: [ncells][n] (built with d_mapc)
    int64_t* d_basec = nullptr;    //   and their column bases colptr[dof]
    double* d_cmat = nullptr;      // stiffness tensor of FB2_ELEM_ELASTICITY_GENERAL (81 doubles, lazy)
    double h_cmat[81] = {};        //   and its host copy (k_march_vec takes it as a kernel argument)
    uint8_t* d_mapb = nullptr;     // byte-packed copy for the marching-tile kernel: [ceil(n*n/16)][ncells_pad][16] (lazy)
    uint16_t* d_map8 = nullptr;    // packed copy for the thread-per-cell kernels: [ceil(n*n/8)][ncells_pad][8] (lazy)
    uint32_t* d_mapv = nullptr;    // lane-major byte map of k_march_vec: [ncells][5][32] words (lazy)
    // Marching kernels on a partition-local grid, exchange mode: the CTAs (tile x chunk) that hold interface cells are launched
    // first (march_part = 1), the others (2) run while the interface columns are exchanged; 0 = one launch over all CTAs.
    // The two CTA lists are cached for the chunk structure / tile size they were built for.
    int march_part = 0;
    int march_overwrite = 0;       // part 2: nzval was zero-filled in front of part 1 (plain stores for tile-interior columns allowed)
    int64_t march_iface = 0;       // cells [0, march_iface) of the local grid are the interface cells
    int32_t* d_cta_list[2] = {nullptr, nullptr};
    int32_t* d_box_zcols = nullptr;   // k_march_vec on a partition-local box: columns that need start_assemble's zero fill (see
    int64_t box_nzcols = -1;          // fb2_march_split_zero_fill), built with the CTA lists; -1 = none
    int64_t cta_count[2] = {0, 0};
    int64_t cta_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int marchv_state = 0;          // k_march_vec: 0 not checked, 1 usable (every grid node carries the three dofs of ONE vector
                                   // field, the same in all its cells), 2 not usable
    // colouring (lazy)
    int ncolors = 0;
    std::vector<int32_t> cell_color;
    std::vector<int64_t> color_ptr;      // [ncolors+1]
    int32_t* d_color_cells = nullptr;    // cells sorted by colour
    // cell subset (partitioned assembly): an index list, or the range [cell_first, cell_first + ncells_active) when
    // d_cells is null and ncells_active > 0; otherwise all cells
    int32_t* d_cells = nullptr;
    int64_t cell_first = 0;
    int64_t ncells_active = 0;
    // slab schedule of the streamed host-buffer entry point (lazy): cell range, nodes needed so far, columns complete
    std::vector<int64_t> slab_cell, slab_node, slab_col, slab_pos;
    // zero-fill schedule of the device path (lazy): cell ranges and the nzval prefix each range needs zeroed
    std::vector<int64_t> zf_cell, zf_pos;
    int zf_state = 0;   // 0 not built, 1 usable, 2 not worthwhile (numbering does not follow the cell order)
    // scratch for the host-buffer entry point
    double* d_nzval = nullptr;
    double* d_f = nullptr;
    double* d_u = nullptr;
};

struct DirichletBC {
    int field = 0, kind = 0;
    std::vector<int> comps;                // 1-based
    std::vector<int64_t> entities;         // flattened pairs or node ids (1-based)
    std::vector<double> points;            // sdim x npoints (evaluation order of update!)
    std::vector<int64_t> point_dofs;       // ncomp x npoints global dofs (0-based)
};

struct fb2_ch {
    fb2_dh* dh = nullptr;
    bool closed = false;
    std::vector<DirichletBC> bcs;
    std::vector<int64_t> insertion;          // prescribed dofs in insertion order (0-based)
    std::vector<int64_t> prescribed;         // sorted, 0-based
    std::vector<double> inhom;
    std::vector<int32_t> dofmap;             // dof -> index in prescribed (or -1), built at close
    int32_t* d_prescribed = nullptr;
    double* d_inhom = nullptr;
    uint8_t* d_isconstrained = nullptr;
    double* d_scratch = nullptr;             // reduction scratch
    bool inhom_dirty = true;
    // AffineConstraint(dof, [master => coeff, ...], b) (src/Dofs/ConstraintHandler.jl:114-131): u_dof = sum coeff u_master + b.
    // Host: per constrained dof (0-based) its masters / coefficients / b, in insertion order of the masters; after close!
    // aff_ptr / aff_dof / aff_coef are aligned with `prescribed` (dofcoefficients of the reference, empty = none).
    struct Affine { std::vector<int64_t> masters; std::vector<double> coefs; double b = 0.0; };
    std::map<int64_t, Affine> affine;
    std::vector<uint8_t> aff_is;             // per prescribed dof: constrained by an AffineConstraint (possibly without masters)
    std::vector<int32_t> aff_ptr, aff_dof;
    std::vector<double> aff_coef;
    bool has_affine = false;                 // some constrained dof has masters
    int32_t* d_aff_ptr = nullptr;            // [np + 1]
    int32_t* d_aff_dof = nullptr;
    double* d_aff_coef = nullptr;
    int32_t* d_aff_of = nullptr;             // [ndofs] index into prescribed of a dof constrained by an AffineConstraint, else -1
    int32_t* d_aff_list = nullptr;           // indices into prescribed of the dofs with masters
    int64_t n_aff = 0;
};

// element assembly (element_assembly.cu)
struct fb2_ea {
    fb2_dh* dh = nullptr;          // the caller's DofHandler (borrowed)
    fb2_cv* cv = nullptr;          // borrowed
    fb2_dh* bdh = nullptr;         // broken twin
    fb2_pattern* bpat = nullptr;   // block-diagonal pattern of the twin
    fb2_assembler* basm = nullptr;
    int n = 0;                     // dofs per cell
    double* d_ub = nullptr;        // state in the broken numbering (lazy)
    double* d_work = nullptr;      // CG work vectors of fb2_ea_cg (lazy)
    size_t work_count = 0;
    struct EaSplit* split = nullptr;   // fb2_apply_assemble: boundary-layer problem of the last ConstraintHandler (lazy)
};

// ---- device-side helpers implemented in .cu files ------------------------------------------
int fb2_pattern_build_device(fb2_pattern* p);
// same builder on an explicit table of (pseudo-)cells: d_cell_dofs is SoA [ndpc][ncells_pad]
int fb2_pattern_build_device_from(fb2_pattern* p, const int32_t* d_cell_dofs, int64_t ncells, int64_t ncells_pad, int ndpc);
int fb2_pattern_finalize(fb2_pattern* p);   // diag index, max col len
int fb2_map_build(fb2_assembler* a);
int fb2_map_build_packed(fb2_assembler* a);
int fb2_map_build_bytes(fb2_assembler* a);
int fb2_map_build_cellmajor(fb2_assembler* a);
int fb2_map_build_vec(fb2_assembler* a);
// zero fill of nzval in front of a split marching launch (fb2_assemble_distributed): column-wise where the kernel allows it
// (*done = true), else the caller runs a memset
int fb2_march_split_zero_fill(fb2_assembler* a, int element, int64_t ncells_own, double* nzval_dev, bool* done);
int fb2_tiles_build(fb2_assembler* a, int TC);
int fb2_warplist_build(fb2_assembler* a);
int fb2_launch_assemble(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* u_dev,
                        double* nzval_dev, double* f_dev, const fb2_asm_opts* opts);
int fb2_check_device_error(fb2_ctx* ctx);
// the marching-tile kernel (march_kernels.cuh) will take `element` on this assembler (structured hexahedral grid, continuous
// Q1 numbering, atomic scatter, default variant)
bool fb2_march_applicable(fb2_assembler* a, int element, const fb2_asm_opts* opts);
int fb2_coloring_build(fb2_assembler* a);
int fb2_ch_sync_device(fb2_ch* ch);      // upload the inhomogeneities if update! changed them
