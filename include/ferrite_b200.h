/*
 * libferrite_b200.so -- C ABI of the B200-native global FE assembly path.
 *
 * The reference (Ferrite.jl, pure Julia) has no FFI; its extension points for a
 * custom matrix/assembler back-end are the generic functions listed in
 * docs/src/devdocs/assembly.md:18-51.  Each entry point below names the
 * reference interface it replaces (path:line under the reference tree).  A Julia
 * shim binds these with `ccall`; the Python mirror in
 * ferrite.jl_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every function returns an int status (FB2_OK == 0); no C++ exception crosses
 *     the boundary; fb2_last_error() returns the message of the last failure on
 *     the calling thread.
 *   - all index arrays that cross the boundary are 1-based int64 exactly as the
 *     reference stores them (cell node ids, cell_dofs, colptr, rowval,
 *     prescribed_dofs); floating point is double.
 *   - host pointers are borrowed for the duration of the call only.
 *   - handles are opaque, created and destroyed by the library.  A context is bound
 *     to one CUDA device and one stream; calls on one context are not re-entrant
 *     (like the reference's per-task assemblers, src/assembler.jl:273-276).
 *   - there is no CPU fallback: every compute entry point fails with
 *     FB2_ERR_CUDA when no CUDA device is usable.
 */
#ifndef FERRITE_B200_H
#define FERRITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ----------------------------------------------------------------- */
enum {
    FB2_OK = 0,
    FB2_ERR_BAD_ARG = 1,
    FB2_ERR_CUDA = 2,
    FB2_ERR_OOM = 3,
    FB2_ERR_DETJ_NOT_POSITIVE = 4,     /* throw_detJ_not_pos, src/FEValues/common_values.jl:5 */
    FB2_ERR_MISSING_PATTERN_ENTRY = 5, /* _missing_sparsity_pattern_error, src/assembler.jl:459-467 */
    FB2_ERR_UNSUPPORTED = 6,
    FB2_ERR_NCCL = 7,
    FB2_ERR_INTERNAL = 8
};

/* ---- enums ------------------------------------------------------------------------ */
/* cell types (src/Grid/grid.jl:282-338): Line, Triangle, Quadrilateral, Tetrahedron, Hexahedron */
enum { FB2_LINE = 1, FB2_TRIANGLE = 2, FB2_QUADRILATERAL = 3, FB2_TETRAHEDRON = 4, FB2_HEXAHEDRON = 5 };

/* element-routine menu (the reference's element routine is user Julia code; these are the
 * tutorial kernels) */
enum {
    FB2_ELEM_HEAT = 1,       /* Ke = int k grad(Ni).grad(Nj), fe = int s Ni; heat_equation.jl:143-164 */
    FB2_ELEM_MASS = 2,       /* Ke = int rho Ni.Nj, fe = 0 */
    FB2_ELEM_ELASTICITY = 3, /* Ke = int eps_i:C:eps_j (isotropic C), fe = int Ni.b; threaded_assembly.jl:105-119 */
    FB2_ELEM_NEOHOOKE = 4,   /* tangent + residual of Psi = mu/2(Ic-3-2lnJ)+lam/2(J-1)^2; hyperelasticity.jl:162-176,241-276 */
    FB2_ELEM_ELASTICITY_GENERAL = 5 /* Ke = int grad(dN_i) : C : grad(N_j) for ANY stiffness tensor C with minor symmetries
                                       (orthotropic, fully anisotropic, ...), fe = int Ni.b: the routine as the reference
                                       writes it, linear_elasticity.jl:266-281, benchmark/helper.jl:249-262.  The isotropic
                                       FB2_ELEM_ELASTICITY (tensor-core SYRK) is the fast special case. */
};

/* scatter strategies (src/assembler.jl:174-231 `atomic` flag; threaded_assembly.jl:232-265 colouring) */
enum { FB2_SCATTER_ATOMIC = 0, FB2_SCATTER_COLORED = 1 };

/* boundary entity kinds of a Dirichlet set (FacetIndex/FaceIndex/EdgeIndex/VertexIndex, node set) */
enum { FB2_BC_FACET = 0, FB2_BC_FACE = 1, FB2_BC_EDGE = 2, FB2_BC_VERTEX = 3, FB2_BC_NODE = 4 };

/* ---- POD parameter structs -------------------------------------------------------- */
typedef struct { int order; int vdim; } fb2_field; /* Lagrange{refshape(cell), order}()^vdim, add!(dh, name, ip) src/Dofs/DofHandler.jl:420-439 */

typedef struct { double k; double source; } fb2_heat_params;
typedef struct { double rho; } fb2_mass_params;
typedef struct { double lambda; double mu; double b[3]; } fb2_elasticity_params; /* also for FB2_ELEM_NEOHOOKE */
/* C[((i*3 + j)*3 + k)*3 + l] = C_ijkl (SymmetricTensor{4,dim}, dim = 2 uses the indices < 2), body force b */
typedef struct { double C[81]; double b[3]; } fb2_elasticity_general_params;

typedef struct {
    int fillzero;     /* start_assemble(K, f; fillzero) src/assembler.jl:287-291 */
    int scatter_mode; /* FB2_SCATTER_* */
    int variant;      /* 0 = default kernel for the element.  >0 = measured alternatives, all parity-tested:
                         1 DFMA block kernel, 2 unrolled quadrature loop, 4 branch-free scatter, 5 tile kernel,
                         6 x+y face merge, 7 no sector pairing, 8 warp-specialised groups, 9 coordinates in shared
                         memory, 12 zero fill overlapped with the assembly; 30 / 32 the thread-per-cell / warp-per-cell kernel
                         where a marching-tile kernel (k_march_hex / k_march_vec) is the default, 31 table-driven integration
                         inside k_march_hex; 20 / 21 are measurement-only (wrong results): integration without scatter /
                         scatter without integration */
    int reserved;
} fb2_asm_opts;

typedef struct fb2_ctx fb2_ctx;
typedef struct fb2_grid fb2_grid;
typedef struct fb2_dh fb2_dh;
typedef struct fb2_pattern fb2_pattern;
typedef struct fb2_cv fb2_cv;
typedef struct fb2_assembler fb2_assembler;
typedef struct fb2_ch fb2_ch;

/* ---- context ---------------------------------------------------------------------- */
const char* fb2_version(void);
const char* fb2_last_error(void);
int fb2_ctx_create(int device, fb2_ctx** out);
int fb2_ctx_destroy(fb2_ctx* ctx);
int fb2_ctx_synchronize(fb2_ctx* ctx);
/* run all subsequent work of this context on an existing CUDA stream (cudaStream_t as void*) */
int fb2_ctx_set_stream(fb2_ctx* ctx, void* cuda_stream);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
int fb2_ctx_launch_count(fb2_ctx* ctx, int64_t* out);
/* device memory owned by the library on behalf of the caller (nzval, f, u ...) */
int fb2_device_alloc(fb2_ctx* ctx, size_t nbytes, void** dev_ptr);
int fb2_device_free(fb2_ctx* ctx, void* dev_ptr);
int fb2_memcpy_h2d(fb2_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);
int fb2_memcpy_d2h(fb2_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes);
/* dependent-chain-free FP64 FMA microbenchmark on the context's device: measured roofline denominator
 * for the FP64-bound kernels (TFLOP/s) */
int fb2_measure_fp64_peak(fb2_ctx* ctx, double* tflops);

/* ---- grid: Grid / generate_grid ---------------------------------------------------- */
/* Grid(cells, nodes), src/Grid/grid.jl:385-393.  cells: nnpc x ncells column-major (= Julia
 * Vector{Hexahedron}), 1-based node ids; xyz: sdim x nnodes (= Vector{Node{sdim,Float64}}). */
int fb2_grid_from_host(fb2_ctx* ctx, int celltype, int64_t ncells, int64_t nnodes, int sdim,
                       const int64_t* cells, const double* xyz, fb2_grid** out);
/* generate_grid(CellType, nel, left, right), src/Grid/grid_generators.jl:8-37 (Line), :78-112
 * (Quadrilateral), :383-417 (Triangle), :159-203 (Hexahedron), :474-537 (Tetrahedron); node
 * coordinates as _generate_nodes :550-578.  Also creates the named facet sets. */
int fb2_grid_generate(fb2_ctx* ctx, int celltype, const int64_t* nel, const double* left,
                      const double* right, fb2_grid** out);
/* deterministic interior-node perturbation x += amplitude*h*(hash(node)-1/2) of a generated
 * grid (cf. perturb_standard_grid!, test/test_utils.jl:282); synthetic benchmark input only */
int fb2_grid_perturb(fb2_grid* grid, double amplitude);
/* replace the node coordinates (sdim x nnodes, host) on the host copy and on the device */
int fb2_grid_set_coordinates(fb2_grid* grid, const double* xyz);
/* device-only, stream-ordered coordinate update from (ideally pinned) host memory: the per-step input of
 * the end-to-end path (moving meshes).  The host copy used by Dirichlet set-up is not touched. */
int fb2_grid_upload_coordinates_async(fb2_grid* grid, const double* xyz_host);
int fb2_grid_info(fb2_grid* grid, int* celltype, int64_t* ncells, int64_t* nnodes, int* nnpc, int* sdim);
int fb2_grid_export(fb2_grid* grid, int64_t* cells, double* xyz);
/* getfacetset(grid, name): pairs = 2 x n (cell, local facet), sorted; pass pairs = NULL to query n */
int fb2_grid_facetset(fb2_grid* grid, const char* name, int64_t* n, int64_t* pairs);
int fb2_grid_destroy(fb2_grid* grid);

/* ---- DofHandler: add! / close! / celldofs / ndofs ------------------------------------- */
/* DofHandler(grid); add!(dh, :f_k, Lagrange{..,order}()^vdim) for each field; close!(dh):
 * src/Dofs/DofHandler.jl:169,420-439,474-569,576-794.  The numbering is bit-identical to the reference. */
int fb2_dh_close(fb2_grid* grid, int nfields, const fb2_field* fields, fb2_dh** out);
/* arrays-in mode: adopt the reference's own dh.cell_dofs (ndofs_per_cell x ncells, 1-based),
 * src/Dofs/DofHandler.jl:126-131 */
int fb2_dh_from_host(fb2_grid* grid, int nfields, const fb2_field* fields, int64_t ndofs,
                     int ndofs_per_cell, const int64_t* cell_dofs, fb2_dh** out);
/* renumber!(dh, order): src/Dofs/DofRenumbering.jl:79-125,167-246.  order: 0 = the permutation perm_in (1-based, dof i
 * becomes perm_in[i]), FB2_ORDER_FIELDWISE / FB2_ORDER_COMPONENTWISE with optional 1-based target blocks (one per field /
 * per component; NULL = declaration order), FB2_ORDER_METIS = DofOrder.Ext{Metis}() (ext/FerriteMetis.jl:29-92: METIS_NodeND
 * on the dof coupling graph, no coupling matrix).  perm_out (nullable, ndofs entries) receives the permutation for
 * fb2_ch_renumber.  Renumber before allocate_matrix: patterns and assemblers of the old numbering are stale. */
enum { FB2_ORDER_PERMUTATION = 0, FB2_ORDER_FIELDWISE = 1, FB2_ORDER_COMPONENTWISE = 2, FB2_ORDER_METIS = 3 };
int fb2_dh_renumber(fb2_dh* dh, int order, const int64_t* target_blocks, int ntargets, const int64_t* perm_in, int64_t* perm_out);
int fb2_dh_info(fb2_dh* dh, int64_t* ndofs, int* ndofs_per_cell, int* nfields);
/* celldofs!(dofs, dh, i) for all cells: ndofs_per_cell x ncells, 1-based (src/Dofs/DofHandler.jl:248-253) */
int fb2_dh_export(fb2_dh* dh, int64_t* cell_dofs);
/* dof_range(dh, field): 1-based inclusive range of local dofs (src/Dofs/DofHandler.jl:1173-1187) */
int fb2_dh_dof_range(fb2_dh* dh, int field, int* first, int* last);
int fb2_dh_destroy(fb2_dh* dh);

/* ---- sparsity pattern: allocate_matrix ------------------------------------------------ */
/* allocate_matrix(dh) -> SparseMatrixCSC pattern, built on the device:
 * src/Dofs/sparsity_pattern.jl:628-645,370-398,1136-1249,951-991.  colptr/rowval are bit-identical. */
int fb2_pattern_create(fb2_dh* dh, fb2_pattern** out);
/* arrays-in mode: adopt K.colptr (n+1) / K.rowval (nnz), 1-based */
int fb2_pattern_from_host(fb2_dh* dh, const int64_t* colptr, const int64_t* rowval, fb2_pattern** out);
/* allocate_matrix(dh, ch): the condensed pattern for a ConstraintHandler with affine / periodic constraints, i.e. the entries of
 * allocate_matrix(dh) plus those `_condense!` writes into (src/Dofs/sparsity_pattern.jl:782-844); built on the device */
int fb2_pattern_create_condensed(fb2_dh* dh, fb2_ch* ch, fb2_pattern** out);
int fb2_pattern_info(fb2_pattern* p, int64_t* n, int64_t* nnz);
int fb2_pattern_export(fb2_pattern* p, int64_t* colptr, int64_t* rowval);
int fb2_pattern_destroy(fb2_pattern* p);

/* ---- CellValues ------------------------------------------------------------------------- */
/* CellValues(QuadratureRule{refshape}(qr_order), Lagrange{refshape,ip_order}()^vdim, Lagrange{refshape,geo_order}()):
 * src/FEValues/CellValues.jl:57-83; quadrature src/Quadrature/quadrature.jl:64-138 (default rule per shape) */
int fb2_cellvalues_create(fb2_ctx* ctx, int celltype, int qr_order, int ip_order, int vdim, int geo_order, fb2_cv** out);
/* arrays-in mode: the reference's own tables.  N: n x nq, dNdxi: rdim x n x nq, M: ngeo x nq,
 * dMdxi: rdim x ngeo x nq (column-major, i.e. cv.fun_values.Nxi/dNdxi FunctionValues.jl:46-49,
 * cv.geo_mapping.M/dMdxi GeometryMapping.jl:38-39), w: nq (cv.qr.weights) */
int fb2_cellvalues_from_tables(fb2_ctx* ctx, int celltype, int nq, int n, int vdim, int ngeo, const double* N,
                               const double* dNdxi, const double* M, const double* dMdxi, const double* w, fb2_cv** out);
int fb2_cellvalues_info(fb2_cv* cv, int* nq, int* nbase_scalar, int* vdim, int* ngeo, int* rdim);
int fb2_cellvalues_export(fb2_cv* cv, double* N, double* dNdxi, double* M, double* dMdxi, double* w, double* points);
int fb2_cellvalues_destroy(fb2_cv* cv);

/* ---- assembler: start_assemble / assemble! / finish_assemble ---------------------------- */
/* start_assemble(K, f) (src/assembler.jl:287-291) for the whole cell loop: builds the
 * cell-local -> nzval index map (replaces the per-cell sort + merge walk of _assemble_inner!,
 * src/assembler.jl:347-457) and, for FB2_SCATTER_COLORED, create_coloring (src/Grid/coloring.jl:308-317). */
int fb2_assembler_create(fb2_dh* dh, fb2_pattern* p, fb2_cv* cv, fb2_assembler** out);
/* the whole `for cell in CellIterator(dh) ... assemble!(assembler, celldofs(cell), Ke, fe)` loop
 * (docs/src/literate-tutorials/heat_equation.jl:181-204; src/iterators.jl:72-93; src/FEValues/CellValues.jl:122-140)
 * nzval_dev (nnz) and f_dev (ndofs, nullable) are device pointers; u_dev (ndofs, nullable) is the
 * current solution for nonlinear elements. */
int fb2_assemble(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* u_dev,
                 double* nzval_dev, double* f_dev, const fb2_asm_opts* opts);
/* same call with HOST buffers: uploads u (if any), assembles into library-owned device buffers and
 * downloads nzval/f into the caller's SparseMatrixCSC.nzval / f vectors */
int fb2_assemble_host(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* u_host,
                      double* nzval_host, double* f_host, const fb2_asm_opts* opts);
/* Same result, pipelined: cells are assembled in slabs; the node coordinates (xyz_host, sdim x nnodes like
 * Vector{Vec}, nullable = keep the coordinates on the device) of the next slab go up and the matrix columns
 * completed by the previous slabs come down while a slab is assembled.  Needs fillzero and the atomic scatter.
 * Host buffers should be pinned (cudaHostRegister / CUDA.pin) for the copies to overlap. */
int fb2_assemble_host_streamed(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* xyz_host,
                               const double* u_host, double* nzval_host, double* f_host, const fb2_asm_opts* opts);
/* name of the kernel family the last fb2_assemble* call on this thread's process launched for the cell loop ("k_march_hex",
 * "k_cell_scalar", "k_cell_syrk", "k_cell_blocks", "k_tile_scalar", ...): lets tests and benchmarks state WHICH hand-written
 * kernel produced a number (no reference counterpart) */
const char* fb2_last_kernel(void);
/* MultiFieldCellValues(qr, (u = ip_u, p = ip_p)) + the mixed u-p element of the incompressible-elasticity tutorial
 * (src/FEValues/CellValues.jl:229-298; docs/src/literate-tutorials/incompressible_elasticity.jl:266-311): cv_u / cv_p are the
 * CellValues of the two fields on ONE quadrature rule and geometric interpolation (checked); `a` = fb2_assembler_create(dh,
 * pattern, NULL) of the two-field DofHandler; field_u / field_p = field indices in add! order.  Integrates K_uu = int 2G
 * dev3d(eps_i):dev3d(eps_j), K_pu = -int psi_i div phi_j, K_pp = -int psi_i psi_j / K (inv_bulk = 1/K, 0 for the
 * incompressible limit) and scatters with assemble! semantics.  f_dev (nullable) is only zero-filled; the traction term
 * comes from fb2_assemble_facets on the displacement field. */
int fb2_assemble_mixed_up(fb2_assembler* a, fb2_cv* cv_u, fb2_cv* cv_p, int field_u, int field_p, double shear_G, double inv_bulk,
                          double* nzval_dev, double* f_dev, const fb2_asm_opts* opts);
/* create_coloring(grid): number of colours and, optionally, the colour of every cell (0-based colour ids) */
int fb2_assembler_coloring(fb2_assembler* a, int* ncolors, int32_t* cell_color);
/* scatter-only entry: assemble!(assembler, dofs, Ke, fe) for a batch of precomputed element matrices
 * (Ke: n x n x ncells column-major, fe: n x ncells, host) -- src/assembler.jl:322-331 */
int fb2_scatter_host(fb2_assembler* a, const double* Ke, const double* fe, double* nzval_dev, double* f_dev,
                     const fb2_asm_opts* opts);
/* the same from device-resident element matrices (layout of fb2_ea_assemble) */
int fb2_scatter_device(fb2_assembler* a, const double* Ke_dev, const double* fe_dev, double* nzval_dev, double* f_dev,
                       const fb2_asm_opts* opts);
int fb2_assembler_destroy(fb2_assembler* a);

/* ---- ConstraintHandler: Dirichlet / close! / update! / apply! ---------------------------- */
int fb2_ch_create(fb2_dh* dh, fb2_ch** out);
/* add!(ch, Dirichlet(field, set, f, components)): src/Dofs/ConstraintHandler.jl:965-1008,404-493.
 * entities: 2 x n (cell, local entity) for FB2_BC_FACET/FACE/EDGE/VERTEX, or n node ids for FB2_BC_NODE;
 * components: 1-based, sorted; ncomponents = 0 means all. Returns the index of the condition in *ibc. */
int fb2_ch_add_dirichlet(fb2_ch* ch, int field, int kind, int64_t n, const int64_t* entities, int ncomponents,
                         const int* components, int* ibc);
/* add!(ch, AffineConstraint(dof, [master => coeff, ...], b)): u_dof = sum coeff u_master + b (src/Dofs/ConstraintHandler.jl:114-131,
 * 383-401).  dof / masters 1-based; n = 0 prescribes the value b.  close! rejects nested constraints ("nested affine constraints
 * currently not supported", :338-361); restriction of this library: masters must be unconstrained dofs.  fb2_apply then runs
 * add_inhomogeneities!, `_condense!` (:782-867) and the row / column zeroing on the device; fb2_apply_vector sets
 * u_dof = sum coeff u_master + b (:686-700).  Use fb2_pattern_create_condensed for the matrix. */
int fb2_ch_add_affine(fb2_ch* ch, int64_t dof, int n, const int64_t* masters, const double* coefs, double b);
/* add!(ch, PeriodicDirichlet(field, collect_periodic_facets(grid, mirror_set, image_set), components)) for facet sets that are
 * translates of each other (src/Dofs/ConstraintHandler.jl:1032-1300): the dofs on the mirror facets are constrained to the dofs
 * at the matching positions of the image facets (u_mirror = u_image, :1046-1047), as affine constraints; pairs = 2 x n
 * (cell, local facet), 1-based; ncomponents = 0 means all */
int fb2_ch_add_periodic(fb2_ch* ch, int field, int64_t n_mirror, const int64_t* mirror_pairs, int64_t n_image,
                        const int64_t* image_pairs, int ncomponents, const int* components);
/* ch.dofcoefficients of the closed handler, aligned with ch.prescribed_dofs (src/Dofs/ConstraintHandler.jl:160-165): ptr
 * (nprescribed + 1 offsets, 0-based), masters (1-based), coefs; *ntotal receives the number of (master, coeff) pairs; any
 * output pointer may be NULL */
int fb2_ch_affine_export(fb2_ch* ch, int64_t* ntotal, int64_t* ptr, int64_t* masters, double* coefs);
/* close!(ch): src/Dofs/ConstraintHandler.jl:303-361 (sorts prescribed dofs, builds isconstrained) */
int fb2_ch_close(fb2_ch* ch);
/* arrays-in mode: adopt a closed reference ConstraintHandler (ch.prescribed_dofs sorted, ch.inhomogeneities),
 * src/Dofs/ConstraintHandler.jl:160-162 */
int fb2_ch_from_host(fb2_dh* dh, int64_t n, const int64_t* prescribed_dofs, const double* inhomogeneities, fb2_ch** out);
/* update!(ch, t) in arrays-in mode: the reference's ch.inhomogeneities after its own update! (src/Dofs/ConstraintHandler.jl:303-361) */
int fb2_ch_set_inhomogeneities(fb2_ch* ch, int64_t n, const double* inhomogeneities);
/* update!(ch, t) (src/Dofs/ConstraintHandler.jl:504-580) is split around the user's Julia function f(x,t):
 * fb2_ch_bc_points returns the dof locations x (sdim x npoints) of condition ibc in the reference's
 * evaluation order (BCValues, src/FEValues/FacetValues.jl:185-236); the caller evaluates f and hands the
 * values (ncomponents x npoints) to fb2_ch_bc_set_values. */
int fb2_ch_bc_points(fb2_ch* ch, int ibc, int64_t* npoints, double* x);
int fb2_ch_bc_set_values(fb2_ch* ch, int ibc, int64_t npoints, const double* values);
/* renumber!(dh, ch, perm): the ConstraintHandler half, src/Dofs/DofRenumbering.jl:92-125 */
int fb2_ch_renumber(fb2_ch* ch, const int64_t* perm);
int fb2_ch_info(fb2_ch* ch, int64_t* nprescribed);
int fb2_ch_export(fb2_ch* ch, int64_t* prescribed_dofs, double* inhomogeneities);
/* apply!(K, f, ch) / apply_zero!(K, f, ch): src/Dofs/ConstraintHandler.jl:710-740,755-768,931-958.
 * f_dev may be NULL (apply!(K, ch)). *meandiag (nullable) receives the mean |diagonal| used. */
int fb2_apply(fb2_ch* ch, fb2_pattern* p, double* nzval_dev, double* f_dev, int applyzero, double* meandiag);
/* apply!(u, ch) / apply_zero!(u, ch): src/Dofs/ConstraintHandler.jl:686-700 */
int fb2_apply_vector(fb2_ch* ch, double* u_dev, int applyzero);
/* get_rhs_data(ch, A) / apply_rhs!(data, f, ch, applyzero): src/Dofs/ConstraintHandler.jl:191-240.  The mean diagonal and the
 * prescribed columns of the matrix are captured BEFORE apply!; afterwards every new right-hand side gets the boundary
 * conditions of the current update! without touching K (time stepping with one factorisation). */
typedef struct fb2_rhsdata fb2_rhsdata;
int fb2_rhsdata_create(fb2_ch* ch, fb2_pattern* p, const double* nzval_dev, fb2_rhsdata** out);
int fb2_rhsdata_info(fb2_rhsdata* data, double* meandiag, int64_t* nprescribed, int64_t* nstored);
int fb2_apply_rhs(fb2_rhsdata* data, double* f_dev, fb2_ch* ch, int applyzero);
int fb2_rhsdata_destroy(fb2_rhsdata* data);
int fb2_ch_destroy(fb2_ch* ch);

/* reinit!(cv, cell) for a batch of cells outside the fused loop (post-processing): src/FEValues/CellValues.jl:122-140.
 * cells: n 1-based cell ids (NULL = cells 1..n).  Outputs on the device, the reference's per-cell arrays back to back:
 * dNdx[cell][q][i][d] (rdim x n x nq per cell, column-major = cv.fun_values.dNdx) and detJdV[cell][q]. */
int fb2_reinit_cells(fb2_cv* cv, fb2_grid* grid, const int64_t* cells, int64_t n, double* dNdx_dev, double* detJdV_dev);

/* spatial_coordinate(cv, q, x) for a batch of cells (src/FEValues/common_values.jl:363-372): x_dev is n x nq x sdim.
 * cells: 1-based ids, NULL = the first n cells. */
int fb2_spatial_coordinates(fb2_cv* cv, fb2_grid* grid, const int64_t* cells, int64_t n, double* x_dev);
/* function_value / function_gradient (src/FEValues/common_values.jl:177-227) of the dof vector u at every quadrature
 * point of every cell: values[cell][q][c] (vdim x nq per cell) and gradients[cell][q][c][d] (dim x vdim x nq per cell,
 * column-major); either output may be NULL. */
int fb2_function_values(fb2_cv* cv, fb2_dh* dh, const double* u_dev, double* values_dev, double* gradients_dev);

/* ---- FacetValues and the Neumann / traction facet loop ------------------------ */
typedef struct fb2_fv fb2_fv;
typedef struct fb2_fset fb2_fset;
/* FacetValues(FacetQuadratureRule{refshape}(qr_order), Lagrange{refshape,ip_order}()^vdim):
 * src/FEValues/FacetValues.jl:39-88; FacetQuadratureRule src/Quadrature/quadrature.jl:205-238 with
 * facet_to_element_transformation src/FEValues/facet_integrals.jl:102-217 */
int fb2_facetvalues_create(fb2_ctx* ctx, int celltype, int qr_order, int ip_order, int vdim, int geo_order, fb2_fv** out);
int fb2_facetvalues_info(fb2_fv* fv, int* nfacets, int* nq, int* nbase_scalar, int* vdim, int* rdim);
/* tables per local facet: w nq x nfacets, points rdim x nq x nfacets (cell reference coordinates), N n x nq x nfacets
 * (column-major); any pointer may be NULL */
int fb2_facetvalues_export(fb2_fv* fv, double* w, double* points, double* N);
int fb2_facetvalues_destroy(fb2_fv* fv);
/* a FacetIndex set, e.g. getfacetset(grid, "top") or a union: pairs = 2 x n (cell, local facet), 1-based */
int fb2_facetset_create(fb2_grid* grid, const int64_t* pairs, int64_t n, fb2_fset** out);
int fb2_facetset_destroy(fb2_fset* set);
enum { FB2_FACET_FLUX = 1,            /* scalar field: fe[i] += q N_i dGamma; params = {q} */
       FB2_FACET_TRACTION = 2,        /* fe[(i,c)] += t_c N_i dGamma; params = t[vdim] */
       FB2_FACET_NORMAL_TRACTION = 3  /* fe[(i,c)] += p n_c N_i dGamma, n = outward unit normal; params = {p}
                                         (hyperelasticity.jl:278-291 is p = -tn) */ };
/* for (cell, facet) in set: reinit!(fv, cell, facet) (src/FEValues/FacetValues.jl:128-154: J, weighted_normal
 * src/FEValues/facet_integrals.jl:122-239, detJ = |weighted normal| > 0, dGamma = detJ w); integrate fe;
 * assemble!(f, celldofs(cell), fe) (src/assembler.jl:338-345).  Adds onto f_dev (no zero fill). */
int fb2_assemble_facets(fb2_dh* dh, fb2_fv* fv, fb2_fset* set, int kind, const double* params, int nparams, double* f_dev);

/* ---- the step after the path: CSR view, SpMV, conjugate gradients on the device ---------------- */
/* y = K x (transpose = 0) or y = K^T x (transpose = 1) with K = (pattern, nzval_dev) in CSC.  Both are gathers with a
 * fixed summation order (K x goes through the transpose permutation of a structurally symmetric pattern; other
 * patterns fall back to a column scatter with FP64 atomics).  x_dev and y_dev must not alias. */
int fb2_spmv(fb2_pattern* p, const double* nzval_dev, const double* x_dev, double* y_dev, int transpose);
/* CSR values of a structurally symmetric pattern (every pattern of allocate_matrix is): with rowptr = colptr and
 * colval = rowval, nzval_csr[k] = K[row(k), colval[k]].  Counterpart of assembling into SparseMatrixCSR,
 * ext/FerriteSparseMatrixCSR.jl:9-95. */
int fb2_csr_values(fb2_pattern* p, const double* nzval_csc_dev, double* nzval_csr_dev);
/* x_dev holds the initial guess and receives the solution of K x = b.  Stops when ||r|| <= max(reltol ||r0||, abstol)
 * or after maxiter iterations (the stopping rule of IterativeSolvers.cg!, hyperelasticity.jl:418).  jacobi != 0:
 * diagonal preconditioner.  symmetric != 0: K is symmetric in value and K^T p is used for K p (cheapest).
 * iters / resnorm (nullable) receive the iteration count and the final residual norm. */
int fb2_cg(fb2_pattern* p, const double* nzval_dev, const double* b_dev, double* x_dev, double reltol, double abstol,
           int maxiter, int jacobi, int symmetric, int* iters, double* resnorm);

/* ---- element assembly, matrix-free operator, apply_local! / apply_assemble! -------------- */
/* All element matrices are kept instead of being summed into a CSC (docs/src/literate-howto/gpu_assembly.jl:265-304):
 * Kes is n x n x ncells (column-major per cell, Kes[c*n*n + j*n + i] = Ke_c[i, j]), fes is n x ncells; both device buffers
 * owned by the caller.  The same fused kernels as fb2_assemble compute them. */
typedef struct fb2_ea fb2_ea;
int fb2_ea_create(fb2_dh* dh, fb2_cv* cv, fb2_ea** out);
int fb2_ea_info(fb2_ea* ea, int64_t* ncells, int* ndofs_per_cell);
/* element routine of the kernel menu for every cell; u_dev (global numbering) for FB2_ELEM_NEOHOOKE; fes_dev nullable */
int fb2_ea_assemble(fb2_ea* ea, int element, const void* params, size_t params_bytes, const double* u_dev, double* Kes_dev,
                    double* fes_dev);
/* y = sum_e P_e' Ke P_e x, the operator of gpu_assembly.jl:287-304 (y is overwritten; equals K * x of the assembled matrix) */
int fb2_ea_mul(fb2_ea* ea, const double* Kes_dev, const double* x_dev, double* y_dev);
/* apply_local!(Ke, fe, celldofs(cell), ch; apply_zero) for every cell, Dirichlet constraints
 * (src/Dofs/ConstraintHandler.jl:1750-1822): fe -= v Ke[:, l], rows/columns of prescribed local dofs zeroed, the diagonal
 * set to meandiag(Ke) and fe[l] = v meandiag(Ke).  fes_dev nullable. */
int fb2_ea_apply_local(fb2_ea* ea, fb2_ch* ch, double* Kes_dev, double* fes_dev, int applyzero);
/* apply_assemble!(assembler, ch, celldofs(cell), Ke, fe; apply_zero) for every cell (src/assembler.jl:491-503):
 * fb2_ea_assemble + fb2_ea_apply_local + fb2_scatter_device into the caller's nzval / f */
int fb2_apply_assemble(fb2_assembler* a, fb2_ea* ea, fb2_ch* ch, int element, const void* params, size_t params_bytes,
                       const double* u_dev, double* nzval_dev, double* f_dev, int applyzero, const fb2_asm_opts* opts);
/* the global right-hand side of the stored element vectors: f = sum_e P_e' fe (f_dev: ndofs doubles, overwritten) */
int fb2_ea_rhs(fb2_ea* ea, const double* fes_dev, double* f_dev);
/* diagonal of the operator: diag[dof(c, i)] = sum over cells of Ke_c[i, i] (diag_dev: ndofs doubles, overwritten) */
int fb2_ea_diag(fb2_ea* ea, const double* Kes_dev, double* diag_dev);
/* fb2_cg on the matrix-free operator (no global matrix): the solver loop gpu_assembly.jl:287-304 is written for.  With
 * Kes / fes after fb2_ea_apply_local the Dirichlet conditions are part of the operator; b = sum_e P_e' fe. */
int fb2_ea_cg(fb2_ea* ea, const double* Kes_dev, const double* b_dev, double* x_dev, double reltol, double abstol, int maxiter,
              int jacobi, int* iters, double* resnorm);
int fb2_ea_destroy(fb2_ea* ea);

/* ---- partitioned multi-GPU assembly (new capability; the reference is single-process) ------ */
typedef struct fb2_part fb2_part;
enum { FB2_DIST_EXCHANGE = 0, /* assemble own cells, exchange interface columns over NCCL */
       FB2_DIST_HALO = 1,     /* assemble own + halo cells redundantly, no communication */
       FB2_DIST_OWN_ONLY = 2  /* assemble own cells only; the host drives pack / transport / unpack_add / mask */ };
/* Host logic only (works on a host-only context).  `dh` is the GLOBAL DofHandler in the reference's numbering.
 * Cells go to ranks as px*py*pz blocks for generate_grid input (dims, nullable = automatic) or as contiguous
 * ranges otherwise; a dof (= matrix column) is owned by the lowest rank among the cells touching it.  Rank
 * `rank` gets a local problem = own cells + halo cells (every cell touching an owned dof), local node / dof
 * numbering by ascending global id, and per-peer exchange lists.  Local cells are ordered [own cells touching
 * a dof shared with another rank | other own cells | halo cells], each group by ascending global id, so that
 * every group is a contiguous range and the interface exchange can overlap the interior cells. */
int fb2_partition_create(fb2_dh* dh, int nparts, int rank, const int* dims, fb2_part** out);
/* same plan from any partitioner's cell -> rank array (ncells entries, 0-based ranks), e.g. METIS_PartMeshDual as used by
 * ext/FerriteMetis.jl */
int fb2_partition_create_from_owners(fb2_dh* dh, int nparts, int rank, const int32_t* cell_owner, fb2_part** out);
/* cell -> rank from METIS_PartMeshDual (the CUDA toolkit's libmetis_static.a, linked statically): general grids */
int fb2_partition_create_metis(fb2_dh* dh, int nparts, int rank, fb2_part** out);
/* The same plan as fb2_partition_create for generate_grid(Hexahedron, nel, left, right) [+ fb2_grid_perturb(perturb)] with one
   Lagrange field of order 1 and `vdim` components, built WITHOUT the global grid and DofHandler: close!'s first-appearance
   numbering (src/Dofs/DofHandler.jl:576-738), ownership and node coordinates have closed forms on that grid, so every rank
   derives its part directly (set-up time independent of the number of ranks).  `host_ctx`: a host-only context that owns the
   metadata-only global problem the plan refers to.  dims: px, py, pz or NULL. */
int fb2_partition_create_generated(fb2_ctx* host_ctx, const int64_t* nel, const double* left, const double* right, double perturb,
                                   int vdim, int nparts, int rank, const int* dims, fb2_part** out);
int fb2_partition_info(fb2_part* part, int64_t* ncells_local, int64_t* ncells_own, int64_t* nnodes_local,
                       int64_t* ndofs_local, int64_t* ndofs_owned);
/* global ids (1-based) of the local cells / nodes / dofs, own-cell flags, owner rank of each local dof */
int fb2_partition_export(fb2_part* part, int64_t* cells_global, uint8_t* cell_is_own, int64_t* l2g_node,
                         int64_t* l2g_dof, int32_t* dof_owner);
/* the local sub-grid / DofHandler of this rank on context `ctx` (a device, or host-only for tests) */
int fb2_partition_local_grid(fb2_part* part, fb2_ctx* ctx, fb2_grid** out);
int fb2_partition_local_dh(fb2_part* part, fb2_grid* local_grid, fb2_dh** out);
/* per-peer exchange plan: number of nz values (and f values) sent to / received from `peer` */
int fb2_partition_peer_counts(fb2_part* part, int peer, int64_t* nz_send, int64_t* f_send, int64_t* nz_recv, int64_t* f_recv);
/* the lists themselves as local 0-based dof ids (row, col), sorted by global (col, row); for tests and for hosts
 * that drive the transport themselves */
int fb2_partition_peer_lists(fb2_part* part, int peer, int32_t* send_rows, int32_t* send_cols, int32_t* recv_rows,
                             int32_t* recv_cols, int32_t* send_f, int32_t* recv_f);
/* device: resolve the lists to positions in the local nzval of assembler `a` (local pattern), upload the own-cell
 * subset and the owned-column mask */
int fb2_partition_bind(fb2_part* part, fb2_assembler* a);
/* pack the partial sums this rank holds for columns owned by `peer` (nz values, then f values) into send_dev
 * (NULL = the plan's internal buffer) */
int fb2_partition_pack(fb2_part* part, int peer, const double* nzval_dev, const double* f_dev, double* send_dev);
/* add the partial sums received from `peer` (recv_dev, NULL = internal buffer) into the owned columns */
int fb2_partition_unpack_add(fb2_part* part, int peer, const double* recv_dev, double* nzval_dev, double* f_dev);
/* zero every column of nzval / entry of f this rank does not own (the owned part is then final) */
int fb2_partition_mask_unowned(fb2_part* part, double* nzval_dev, double* f_dev);
int fb2_partition_destroy(fb2_part* part);

/* NCCL transport inside the library (libnccl.so.2 is resolved with dlopen at run time).  The 128-byte unique id is
 * created on one rank and distributed by the caller (e.g. a torch.distributed / MPI broadcast). */
int fb2_comm_unique_id(void* id128 /* 128 bytes out */);
int fb2_comm_init_rank(fb2_ctx* ctx, const void* id128, int nranks, int rank);
int fb2_comm_destroy(fb2_ctx* ctx);
/* pack for all peers, one grouped ncclSend/ncclRecv, unpack-add; stream-ordered on the context's stream */
int fb2_partition_exchange(fb2_part* part, double* nzval_dev, double* f_dev);
/* assemble! for a partition: FB2_DIST_EXCHANGE = own cells + fb2_partition_exchange, FB2_DIST_HALO = own + halo
 * cells; then mask what is not owned.  nzval_dev / f_dev are the LOCAL matrix / vector of this rank. */
int fb2_assemble_distributed(fb2_assembler* a, fb2_part* part, int mode, int element, const void* params, size_t params_bytes,
                             const double* u_dev, double* nzval_dev, double* f_dev, const fb2_asm_opts* opts);

#ifdef __cplusplus
}
#endif
#endif /* FERRITE_B200_H */
