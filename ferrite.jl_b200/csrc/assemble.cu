// Host side of the assembly path: table upload, kernel selection, launches, colouring, scatter-only.
//
// fb2_assemble = start_assemble (zero fill, src/assembler.jl:287-291 + src/arrayutils.jl:118-152) followed by
// one fused kernel per launch (or one per colour) that covers the reference's whole cell loop; see
// assemble_kernels.cuh for the kernels and DESIGN.md for the per-kernel byte / flop accounting.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cub/cub.cuh>

#include "assemble_kernels.cuh"
#include "march_kernels.cuh"
#include "march_vec_kernels.cuh"

const char* g_fb2_last_kernel = "";
extern "C" const char* fb2_last_kernel(void) { return g_fb2_last_kernel; }

namespace {

// the zero fill of nzval that start_assemble owes, for every kernel that does not fuse it
int pay_zero_fill(fb2_assembler* a, AsmArgs& A) {
    if (A.zero_pending) {
        FB2_CUDA(cudaMemsetAsync(A.nzval, 0, (size_t)a->pat->nnz * sizeof(double), a->dh->grid->ctx->stream));
        A.zero_pending = 0;
    }
    return FB2_OK;
}


// c_tab is one module-global array per DEVICE, whatever the number of contexts on it: the owner is tracked per device
// (by the CellValues' never-reused uid), not per context.
uint64_t g_const_tables_uid[64] = {};

int upload_tables(fb2_assembler* a, AsmArgs* A) {
    fb2_cv* cv = a->cv;
    fb2_ctx* ctx = a->dh->grid->ctx;   // everything of an assembly runs on the grid's context / stream
    const int nq = cv->nq, nb = cv->nb, ng = cv->ngeo, rd = cv->rdim;
    A->nq = nq;
    A->o_w = 0;
    A->o_N = A->o_w + nq;
    A->o_dN = A->o_N + nq * nb;
    A->o_M = A->o_dN + nq * nb * rd;
    A->o_dM = A->o_M + nq * ng;
    const int total = A->o_dM + nq * ng * rd;
    // tables larger than the constant bank (e.g. Q2 hexahedra with a 4x4x4 rule) are served from the global copy alone:
    // only the thread-per-cell kernels read c_tab, and those shapes are handled by the CTA kernels anyway
    A->const_ok = total <= FB2_TAB_MAX;
    uint64_t& owner = g_const_tables_uid[ctx->device & 63];
    if (owner != cv->uid || cv->d_tables == nullptr) {
        std::vector<double> h(total);
        memcpy(h.data() + A->o_w, cv->w.data(), sizeof(double) * nq);
        memcpy(h.data() + A->o_N, cv->N.data(), sizeof(double) * nq * nb);
        memcpy(h.data() + A->o_dN, cv->dN.data(), sizeof(double) * nq * nb * rd);
        memcpy(h.data() + A->o_M, cv->M.data(), sizeof(double) * nq * ng);
        memcpy(h.data() + A->o_dM, cv->dM.data(), sizeof(double) * nq * ng * rd);
        // synchronous w.r.t. the host buffer; ordered on the stream w.r.t. earlier kernels
        if (A->const_ok) {
            FB2_CUDA(cudaMemcpyToSymbolAsync(c_tab, h.data(), sizeof(double) * total, 0, cudaMemcpyHostToDevice, ctx->stream));
            FB2_CUDA(cudaStreamSynchronize(ctx->stream));
            owner = cv->uid;
        }
        if (cv->d_tables == nullptr) {   // global copy: the CTA kernels index the tables per lane, which a constant bank serialises
            FB2_CUDA(cudaMalloc(&cv->d_tables, sizeof(double) * total));
            FB2_CUDA(cudaMemcpy(cv->d_tables, h.data(), sizeof(double) * total, cudaMemcpyHostToDevice));
            cv->tables_device = ctx->device;
            cv->tables_count = total;
        }
    }
    A->tab = cv->d_tables;
    return FB2_OK;
}

template <int DIM, int NGEO, int NB, int NQ, int ELEM>
int launch_scalar(fb2_ctx* ctx, const AsmArgs& A, bool atomic, int variant = 0, bool unchecked = false) {
    const int bs = 128;
    const unsigned grid = (unsigned)((A.ncount + bs - 1) / bs);
    // NQ >= 8: keep the quadrature loop rolled (code fits the instruction cache; 4 % faster on Q1 hex); variant 2
    // selects the fully unrolled body for comparison
    constexpr bool ROLL = NQ >= 8;
    if (atomic && variant == 2) k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 1, false><<<grid, bs, 0, ctx->stream>>>(A);
    // variant 4: branch-free scatter for complete maps; measured slower on Q1 hex (2.98 vs 2.86 ms: the kernel is bound
    // by L2 atomic throughput, not by instruction issue), so the checked scatter stays the default
    else if (atomic && unchecked && variant == 4) k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 1, ROLL, false><<<grid, bs, 0, ctx->stream>>>(A);
    // variant 8: warp-specialised compute / scatter groups (Q1 quadrilaterals and hexahedra)
    else if (atomic && variant == 8 && A.wfirst == nullptr && ((DIM == 3 && NB == 8 && NGEO == 8) || (DIM == 2 && NB == 4 && NGEO == 4))) {
        if constexpr ((DIM == 3 && NB == 8 && NGEO == 8) || (DIM == 2 && NB == 4 && NGEO == 4)) {
            auto k = unchecked ? k_cell_scalar_ws<DIM, NGEO, NB, NQ, ELEM, ROLL, false> : k_cell_scalar_ws<DIM, NGEO, NB, NQ, ELEM, ROLL, true>;
            const size_t smem = WsSmem<NB>::total;
            FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t nbatch = (A.ncount + 127) / 128;
            const unsigned g2 = (unsigned)std::min<int64_t>(nbatch, 2 * (int64_t)ctx->sm_count);
            k<<<g2, 256, smem, ctx->stream>>>(A, nbatch);
        }
    }
    // variant 9: node coordinates in shared memory instead of 48 registers, 3 CTAs per SM (168 registers, 120 B of
    // spills).  C2: 2.34 ms vs 2.30 ms -- 50 % more resident warps buy nothing, so the kernel is not bound by per-warp
    // latency but by the FP64 pipe and the L2 atomic path taking turns (4 CTAs at 128 registers: 4.7 ms, spills).
    else if (atomic && variant == 9) k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 3, ROLL, true, true, true><<<grid, bs, 0, ctx->stream>>>(A);
    // variants 20 / 21 (measurement only, results are wrong): integration without scatter / scatter without integration
    else if (atomic && (variant == 20 || variant == 21) && DIM == 3 && NB == 8 && NGEO == 8 && ELEM == FB2_ELEM_HEAT) {
        if constexpr (DIM == 3 && NB == 8 && NGEO == 8 && ELEM == FB2_ELEM_HEAT) {
            if (variant == 20) k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 1, ROLL, true, true, false, 1><<<grid, bs, 0, ctx->stream>>>(A);
            else k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 1, ROLL, true, true, false, 2><<<grid, bs, 0, ctx->stream>>>(A);
        }
    }
    // variant 7: without the lane-parity sector pairing of the scatter (A/B measurement)
    else if (atomic && variant == 7) k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 1, ROLL, true, false><<<grid, bs, 0, ctx->stream>>>(A);
    else if (atomic) k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, true, 1, ROLL><<<grid, bs, 0, ctx->stream>>>(A);
    else k_cell_scalar<DIM, NGEO, NB, NQ, ELEM, false><<<grid, bs, 0, ctx->stream>>>(A);
    ctx->launches++;
    g_fb2_last_kernel = "k_cell_scalar";
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

template <int DIM, int NGEO, int NBS, int VDIM, int ELEM>
int launch_blocks(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic) {
    constexpr int TB = TileOf<NBS, VDIM>::TB;
    constexpr int NT = (NBS + TB - 1) / TB;
    FB2_TRY(fb2_map_build_cellmajor(a));
    A.mapc = a->d_mapc;
    A.dofc = a->d_dofc;
    A.basec = a->d_basec;
    // cells per CTA: phase-B items (row node x column tile) are looped over 256 threads; take as many cells as fit in
    // ~100 KB of shared memory (two CTAs per SM overlap one CTA's geometry phase / barriers with the other's phase B),
    // preferring a count whose item total fills the last loop iteration
    const int per_cell = NBS * NT;
    int cells = 1, best_eff = 0;
    for (int c = 1; c <= 32; ++c) {
        if (fb2_blocks_smem<DIM, NBS, VDIM, ELEM>(A.nq, c).total > 100 * 1024) break;
        const int items = c * per_cell, slots = (items + 255) / 256 * 256;
        const int eff = items * 1000 / slots;
        if (eff >= best_eff || items <= 256) { best_eff = std::max(eff, best_eff); cells = c; }
        if (items >= 1024) break;
    }
    const BlockSmem L = fb2_blocks_smem<DIM, NBS, VDIM, ELEM>(A.nq, cells);
    FB2_CHECK(L.total <= 227 * 1024, FB2_ERR_UNSUPPORTED, "element needs %zu bytes of shared memory per cell", L.total);
    const int bs = std::min(ELEM == FB2_ELEM_NEOHOOKE ? 192 : 256, (per_cell * cells + 31) / 32 * 32);
    const unsigned grid = (unsigned)((A.ncount + cells - 1) / cells);
    if (atomic) {
        auto k = k_cell_blocks<DIM, NGEO, NBS, VDIM, ELEM, true>;
        FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        k<<<grid, bs, L.total, ctx->stream>>>(A, cells);
    } else {
        auto k = k_cell_blocks<DIM, NGEO, NBS, VDIM, ELEM, false>;
        FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        k<<<grid, bs, L.total, ctx->stream>>>(A, cells);
    }
    g_fb2_last_kernel = "k_cell_blocks";
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

// isotropic elasticity on the FP64 tensor cores (k_cell_syrk)
template <int DIM, int NGEO, int NBS>
int launch_syrk(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic) {
    using S = SyrkOf<NBS, DIM>;
    FB2_TRY(fb2_map_build_cellmajor(a));
    A.mapc = a->d_mapc;
    A.dofc = a->d_dofc;
    A.basec = a->d_basec;
    const SyrkSmem L = fb2_syrk_smem<NBS, DIM>(A.nq);
    const size_t smem = L.cell * S::CELLS;
    FB2_CHECK(smem <= 227 * 1024, FB2_ERR_UNSUPPORTED, "element needs %zu bytes of shared memory", smem);
    const unsigned grid = (unsigned)((A.ncount + S::CELLS - 1) / S::CELLS);
    if (atomic) {
        auto k = k_cell_syrk<DIM, NGEO, NBS, true>;
        FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, S::NTHR, smem, ctx->stream>>>(A);
    } else {
        auto k = k_cell_syrk<DIM, NGEO, NBS, false>;
        FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, S::NTHR, smem, ctx->stream>>>(A);
    }
    g_fb2_last_kernel = "k_cell_syrk";
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

int dispatch_syrk(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic, int celltype, int nbs, int vdim, bool* handled) {
    *handled = true;
#define CASE(CT, DIM, NGEO, NBS) \
    if (celltype == CT && nbs == NBS && vdim == DIM) return launch_syrk<DIM, NGEO, NBS>(a, ctx, A, atomic);
    CASE(FB2_TRIANGLE, 2, 3, 3)
    CASE(FB2_TRIANGLE, 2, 3, 6)
    CASE(FB2_QUADRILATERAL, 2, 4, 4)
    CASE(FB2_QUADRILATERAL, 2, 4, 9)
    CASE(FB2_TETRAHEDRON, 3, 4, 4)
    CASE(FB2_TETRAHEDRON, 3, 4, 10)
    CASE(FB2_HEXAHEDRON, 3, 8, 8)
    CASE(FB2_HEXAHEDRON, 3, 8, 27)
#undef CASE
    *handled = false;
    return FB2_OK;
}

// shape dispatch helpers -------------------------------------------------------------------------
template <int ELEM, int VDIM_IS_DIM>
int dispatch_blocks(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic, int celltype, int nbs, int vdim) {
#define CASE(CT, DIM, NGEO, NBS)                                                                  \
    if (celltype == CT && nbs == NBS) {                                                           \
        if (VDIM_IS_DIM) { if (vdim == DIM) return launch_blocks<DIM, NGEO, NBS, DIM, ELEM>(a, ctx, A, atomic); } \
        else { if (vdim == 1) return launch_blocks<DIM, NGEO, NBS, 1, ELEM>(a, ctx, A, atomic); }    \
    }
    CASE(FB2_TRIANGLE, 2, 3, 3)
    CASE(FB2_TRIANGLE, 2, 3, 6)
    CASE(FB2_QUADRILATERAL, 2, 4, 4)
    CASE(FB2_QUADRILATERAL, 2, 4, 9)
    CASE(FB2_TETRAHEDRON, 3, 4, 4)
    CASE(FB2_TETRAHEDRON, 3, 4, 10)
    CASE(FB2_HEXAHEDRON, 3, 8, 8)
    CASE(FB2_HEXAHEDRON, 3, 8, 27)
#undef CASE
    return fb2_fail(FB2_ERR_UNSUPPORTED, "no kernel for cell type %d with %d scalar basis functions and vdim %d", celltype, nbs, vdim);
}

int dispatch_neohooke(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic, int celltype, int nbs, int vdim) {
    if (vdim == 3) {
        if (celltype == FB2_TETRAHEDRON && nbs == 4) return launch_blocks<3, 4, 4, 3, FB2_ELEM_NEOHOOKE>(a, ctx, A, atomic);
        if (celltype == FB2_TETRAHEDRON && nbs == 10) return launch_blocks<3, 4, 10, 3, FB2_ELEM_NEOHOOKE>(a, ctx, A, atomic);
        if (celltype == FB2_HEXAHEDRON && nbs == 8) return launch_blocks<3, 8, 8, 3, FB2_ELEM_NEOHOOKE>(a, ctx, A, atomic);
        if (celltype == FB2_HEXAHEDRON && nbs == 27) return launch_blocks<3, 8, 27, 3, FB2_ELEM_NEOHOOKE>(a, ctx, A, atomic);
    }
    return fb2_fail(FB2_ERR_UNSUPPORTED, "Neo-Hooke needs a 3-D cell with a 3-component field");
}

// fb2_hex8_heat hard-codes the tables of CellValues(QuadratureRule{RefHexahedron}(2), Lagrange{RefHexahedron,1}()): accept a
// CellValues (library-made or arrays-in) only if its tables agree with them to rounding.
bool cv_is_q1hex_gauss2(const fb2_cv* cv) {
    if (cv->celltype != FB2_HEXAHEDRON || cv->nq != 8 || cv->nb != 8 || cv->ngeo != 8 || cv->rdim != 3) return false;
    const double tol = 4e-15;
    for (int q = 1; q < 8; ++q)
        if (cv->w[q] != cv->w[0]) return false;     // the analytic element takes one weight
    for (int q = 0; q < 8; ++q) {
        const int qb[3] = {q & 1, (q >> 1) & 1, (q >> 2) & 1};
        for (int i = 0; i < 8; ++i) {
            const int sb[3] = {fb2_hx(i), fb2_hy(i), fb2_hz(i)};
            double n[3], d[3];
            for (int a = 0; a < 3; ++a) { n[a] = fb2_q1n(sb[a], qb[a]); d[a] = sb[a] ? 0.5 : -0.5; }
            const double N = n[0] * n[1] * n[2];
            const double dN[3] = {d[0] * n[1] * n[2], n[0] * d[1] * n[2], n[0] * n[1] * d[2]};
            if (std::fabs(cv->N[q * 8 + i] - N) > tol || std::fabs(cv->M[q * 8 + i] - N) > tol) return false;
            for (int a = 0; a < 3; ++a)
                if (std::fabs(cv->dN[(q * 8 + i) * 3 + a] - dN[a]) > tol || std::fabs(cv->dM[(q * 8 + i) * 3 + a] - dN[a]) > tol) return false;
        }
    }
    return true;
}

// The marching-tile kernel keeps one matrix-column copy per tile NODE, so it needs a numbering in which every grid node
// carries exactly one dof of the (scalar, first-order) field -- true for close!(dh) and any renumber!, false for the
// broken twin of the element-assembly path (element_assembly.cu), whose cells own private dofs.  Checked once per assembler.
bool march_usable(fb2_assembler* a) {
    if (a->march_state == 0) {
        const fb2_dh* dh = a->dh;
        const fb2_grid* g = dh->grid;
        bool ok = dh->ndpc == 8 && g->nnpc == 8 && (int64_t)dh->cell_dofs.size() == g->ncells * 8;
        if (ok) {
            std::vector<int32_t> node_dof((size_t)g->nnodes, -1);
            for (int64_t c = 0; c < g->ncells && ok; ++c)
                for (int i = 0; i < 8; ++i) {
                    int32_t& nd = node_dof[(size_t)g->cells[(size_t)c * 8 + i] - 1];
                    const int32_t d = dh->cell_dofs[(size_t)c * 8 + i];
                    if (nd < 0) nd = d;
                    else if (nd != d) { ok = false; break; }
                }
        }
        a->march_state = ok ? 1 : 2;
    }
    return a->march_state == 1;
}

// Columns the marching-tile kernel adds to with reduce-adds (see k_march_mark), as a sorted device list; cached per chunk length.
int march_zero_list(fb2_assembler* a, int lz, int tx = 8, int ty = 4, int vdim = 1, int nfull = 1 << 30, int lt = 1) {
    const int zkey = lz + 1024 * lt + (nfull < 1024 ? nfull << 20 : 0);
    if (a->d_march_zcols && a->march_zlz == zkey) return FB2_OK;
    const fb2_dh* dh = a->dh;
    const fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int nx = (int)g->nel[0], ny = (int)g->nel[1], nz = (int)g->nel[2];
    const int64_t nn = (int64_t)(nx + 1) * (ny + 1) * (nz + 1), nd = dh->ndofs;
    cudaFree(a->d_march_zcols);
    a->d_march_zcols = nullptr;
    uint8_t* d_flag = nullptr;
    int32_t* d_out = nullptr;
    int64_t* d_num = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int64_t num = 0;
    cudaError_t e = cudaMalloc(&d_flag, (size_t)nd);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)nd * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_num, sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, (size_t)nd, st);
    if (e == cudaSuccess) {
        k_march_mark<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(dh->d_cell_dofs, g->ncells_pad, nx, ny, nz, lz, std::min(nfull, nz / std::max(lz, 1) + 1), lt, tx, ty, vdim, d_flag);
        cub::CountingInputIterator<int32_t> ids(0);
        e = cub::DeviceSelect::Flagged(nullptr, tmp_bytes, ids, d_flag, d_out, d_num, (int)nd, st);
        if (e == cudaSuccess) e = cudaMalloc(&d_tmp, tmp_bytes);
        if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, ids, d_flag, d_out, d_num, (int)nd, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&num, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e == cudaSuccess) e = cudaMalloc(&a->d_march_zcols, std::max<size_t>((size_t)num, 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(a->d_march_zcols, d_out, (size_t)num * sizeof(int32_t), cudaMemcpyDeviceToDevice);
    cudaFree(d_flag); cudaFree(d_out); cudaFree(d_num); cudaFree(d_tmp);
    if (e != cudaSuccess) return fb2_fail(FB2_ERR_CUDA, "march_zero_list: %s", cudaGetErrorString(e));
    ctx->launches += 2;
    a->march_zlz = zkey;
    a->march_nzcols = num;
    return FB2_OK;
}

// k_march_vec keeps three matrix-column copies per tile NODE: it needs a numbering in which every grid node carries the three
// dofs of one first-order vector field, the same in every cell around it (true for close!(dh) and any renumber!, false for
// the broken twin of the element-assembly path).  Checked once per assembler.
bool marchv_usable(fb2_assembler* a) {
    if (a->marchv_state == 0) {
        const fb2_dh* dh = a->dh;
        const fb2_grid* g = dh->grid;
        bool ok = dh->ndpc == 24 && g->nnpc == 8 && dh->fields.size() == 1 && dh->fields[0].vdim == 3 &&
                  (int64_t)dh->cell_dofs.size() == g->ncells * 24;
        if (ok) {
            std::vector<int32_t> node_dof((size_t)g->nnodes * 3, -1);
            for (int64_t c = 0; c < g->ncells && ok; ++c)
                for (int i = 0; i < 24; ++i) {
                    int32_t& nd = node_dof[((size_t)g->cells[(size_t)c * 8 + i / 3] - 1) * 3 + i % 3];
                    const int32_t d = dh->cell_dofs[(size_t)c * 24 + i];
                    if (nd < 0) nd = d;
                    else if (nd != d) { ok = false; break; }
                }
        }
        a->marchv_state = ok ? 1 : 2;
    }
    return a->marchv_state == 1;
}

// CTA lists of a split marching launch on a partition-local box (see fb2_assembler::march_part): list 0 = the CTAs (tile x
// chunk) that hold an interface cell, list 1 = the other CTAs with own cells.  Built on the host from the cell map.
// Chunks of k_march_hex: about a dozen CTAs per resident slot keep the tail of the launch short; chunks of >= 4 layers keep the
// share of first / last planes (reduce-adds instead of stores) small (FB2_MARCH_LZ overrides).  CTAs of equal length run in
// synchronised waves, and the last, partly filled wave costs a whole chunk time (4 % of the launch on a 201 x 200 x 200 box:
// 10.56 waves of 20 layers): the last ~15 % of the layers are cut into chunks of a third of the length (first nfull chunks:
// lz layers, then lt).  FB2_MARCH_LT (tests) forces short chunks behind the first half of the full ones.
void marchh_plan(const fb2_ctx* ctx, int tiles_x, int tiles_y, int nzl, int* lz, int* nfull, int* lt, int* nchunks) {
    const int64_t tiles = (int64_t)tiles_x * tiles_y, resident = (int64_t)ctx->sm_count * 8;
    const int64_t want = std::max<int64_t>(1, 12 * resident / tiles);
    *lz = (int)std::max<int64_t>(4, (nzl + want - 1) / want);
    if (const char* e = getenv("FB2_MARCH_LZ")) *lz = std::max(1, atoi(e));
    *nchunks = (nzl + *lz - 1) / *lz;
    *nfull = *nchunks;
    *lt = *lz;
    if (!getenv("FB2_MARCH_LZ") && *nchunks >= 4 && *lz >= 12) {
        *nfull = std::max(1, (int)(0.85 * nzl / *lz));
        *lt = std::max(4, *lz / 3);
        *nchunks = *nfull + (nzl - *nfull * *lz + *lt - 1) / *lt;
    }
    if (const char* e = getenv("FB2_MARCH_LT")) {
        *nfull = std::max(1, ((nzl + *lz - 1) / *lz) / 2);
        *lt = std::min(*lz, std::max(1, atoi(e)));
        *nchunks = *nfull + (std::max(0, nzl - *nfull * *lz) + *lt - 1) / *lt;
    }
}

// chunk length of k_march_vec: about two dozen CTAs per resident slot keep the tail of the launch short; chunks of >= 8
// layers keep the share of first / last planes (reduce-adds instead of stores) small.  FB2_MARCH_LZ overrides (tuning).
void marchv_plan(const fb2_ctx* ctx, int tiles_x, int tiles_y, int nzl, int* lz, int* nchunks) {
    const int64_t tiles = (int64_t)tiles_x * tiles_y, resident = (int64_t)ctx->sm_count * 2;
    const int64_t want = std::max<int64_t>(1, (24 * resident + tiles - 1) / tiles);
    *lz = (int)std::max<int64_t>(8, (nzl + want - 1) / want);
    if (const char* e = getenv("FB2_MARCH_LZ")) *lz = std::max(1, atoi(e));
    *nchunks = (nzl + *lz - 1) / *lz;
}

int march_cta_lists(fb2_assembler* a, const MarchArgs& M, int tile_x, int tile_y, int nchunks, int64_t ncells_own, bool zero_list = false) {
    const fb2_grid* g = a->dh->grid;
    const int64_t key[8] = {M.lz, M.nfull, M.lt, tile_x, tile_y, a->march_iface, ncells_own, nchunks + (zero_list ? ((int64_t)1 << 32) : 0)};
    if (a->d_cta_list[0] && memcmp(key, a->cta_key, sizeof(key)) == 0) return FB2_OK;
    FB2_CHECK(g->structured && !g->sv_cellmap.empty(), FB2_ERR_INTERNAL, "march_cta_lists: the grid has no cell map");
    const int64_t nx = g->sv_nel[0], ny = g->sv_nel[1], nz = g->sv_nel[2];
    const int64_t tiles = (int64_t)M.tiles_x * M.tiles_y;
    std::vector<uint8_t> flag((size_t)(tiles * nchunks), 0);
    for (int64_t z = 0; z < nz; ++z) {
        const int64_t zr = z - M.z0;
        const int64_t ch = M.nfull > 0 && zr >= (int64_t)M.nfull * M.lz ? M.nfull + (zr - (int64_t)M.nfull * M.lz) / M.lt : zr / M.lz;
        if (zr < 0 || ch >= nchunks) continue;
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t x = 0; x < nx; ++x) {
                const int32_t c = g->sv_cellmap[(size_t)(x + nx * (y + ny * z))];
                if (c < 0 || c >= ncells_own) continue;
                uint8_t& fl = flag[(size_t)(ch * tiles + (y / tile_y) * M.tiles_x + x / tile_x)];
                fl |= c < a->march_iface ? 3 : 1;
            }
    }
    std::vector<int32_t> list[2];
    for (size_t i = 0; i < flag.size(); ++i) {
        if (flag[i] & 2) list[0].push_back((int32_t)i);
        else if (flag[i] & 1) list[1].push_back((int32_t)i);
    }
    for (int k = 0; k < 2; ++k) {
        cudaFree(a->d_cta_list[k]);
        a->d_cta_list[k] = nullptr;
        FB2_CUDA(cudaMalloc(&a->d_cta_list[k], std::max<size_t>(list[k].size(), 1) * sizeof(int32_t)));
        FB2_CUDA(cudaMemcpy(a->d_cta_list[k], list[k].data(), list[k].size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        a->cta_count[k] = (int64_t)list[k].size();
    }
    // Columns that need start_assemble's zero fill when tile-interior columns are written with plain stores (three dofs per
    // node, k_march_vec): every node that is NOT strictly inside a CTA (tile x chunk) with own cells -- nodes on tile faces,
    // on the planes between chunks, and the nodes of regions without own cells (nobody writes their columns).
    cudaFree(a->d_box_zcols);
    a->d_box_zcols = nullptr;
    a->box_nzcols = -1;
    if (zero_list && (a->dh->ndpc == 24 || a->dh->ndpc == 8)) {
        const int vd = a->dh->ndpc / 8;
        const fb2_dh* dh = a->dh;
        std::vector<int32_t> zc;
        for (int64_t z = 0; z <= nz; ++z) {
            const int64_t zr = z - M.z0;
            // strictly between the first and the last node plane of a chunk
            const int64_t zt = zr - (int64_t)M.nfull * M.lz;
            const bool zin = zr > 0 && zr < M.z1 - M.z0 && ((M.nfull <= 0 || zt <= 0) ? zr % M.lz != 0 : zt % M.lt != 0);
            for (int64_t y = 0; y <= ny; ++y)
                for (int64_t x = 0; x <= nx; ++x) {
                    // the node's dofs from any local cell around it; written by a plain store iff the node is strictly inside
                    // a tile and a chunk and an own cell touches it (the kernel then holds a copy of its columns)
                    int32_t cany = -1;
                    int lnany = 0;
                    bool own_touch = false;
                    for (int dz = 0; dz < 2; ++dz)
                        for (int dy = 0; dy < 2; ++dy)
                            for (int dx = 0; dx < 2; ++dx) {
                                const int64_t cx = x - dx, cy = y - dy, cz = z - dz;
                                if (cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz) continue;
                                const int32_t c = g->sv_cellmap[(size_t)(cx + nx * (cy + ny * cz))];
                                if (c < 0) continue;
                                if (cany < 0) { cany = c; lnany = fb2_hexnode(dx, dy, dz); }
                                if (c < ncells_own) own_touch = true;
                            }
                    if (cany < 0) continue;
                    const bool stored = zin && own_touch && x % tile_x != 0 && y % tile_y != 0;
                    if (stored) continue;
                    for (int k = 0; k < vd; ++k) zc.push_back(dh->cell_dofs[(size_t)cany * 8 * vd + lnany * vd + k]);
                }
        }
        std::sort(zc.begin(), zc.end());
        FB2_CUDA(cudaMalloc(&a->d_box_zcols, std::max<size_t>(zc.size(), 1) * sizeof(int32_t)));
        FB2_CUDA(cudaMemcpy(a->d_box_zcols, zc.data(), zc.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        a->box_nzcols = (int64_t)zc.size();
    }
    memcpy(a->cta_key, key, sizeof(key));
    return FB2_OK;
}

// conditions of k_march_vec that do not depend on the launch: element tables, structured grid, column length, numbering
bool marchv_static_ok(fb2_assembler* a) {
    const fb2_grid* g = a->dh->grid;
    const fb2_cv* cv = a->cv;
    if (!cv || cv->vdim != 3 || !cv_is_q1hex_gauss2(cv)) return false;
    const bool gen = g->generated && g->celltype == FB2_HEXAHEDRON;
    const bool boxed = !gen && g->structured && g->d_sv_cellmap != nullptr;
    if (!(gen || boxed)) return false;
    const int64_t* snel = gen ? g->nel : g->sv_nel;
    if (snel[0] >= (1 << 28) || snel[1] >= (1 << 28) || a->pat->max_col_len >= 255) return false;
    if (fb2_mvec_smem(fb2_mvec_cap(std::max(a->pat->max_col_len, 1))) > 113 * 1024) return false;
    return marchv_usable(a);
}

// Isotropic elasticity on trilinear hexahedra of a structured grid: the marching-tile kernel k_march_vec.  *handled = false
// when it does not apply (the caller falls back to k_cell_syrk).
int try_march_vec(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, int accumulate, bool* handled, bool general = false) {
    *handled = false;
    fb2_grid* g = a->dh->grid;
    const fb2_cv* cv = a->cv;
    if (!marchv_static_ok(a)) return FB2_OK;
    const bool gen = g->generated && g->celltype == FB2_HEXAHEDRON;
    if (A.cells != nullptr || A.ncount <= 0) return FB2_OK;
    const int64_t* snel = gen ? g->nel : g->sv_nel;
    const int64_t lay = snel[0] * snel[1];
    if (gen && !(A.cell_first % lay == 0 && A.ncount % lay == 0)) return FB2_OK;
    if ((reinterpret_cast<uintptr_t>(A.nzval) & 15) != 0) return FB2_OK;   // the bulk (TMA) flush needs 16-byte aligned matrix storage
    MarchArgs M;
    memset(&M, 0, sizeof(M));
    M.nx = (int)snel[0];
    M.ny = (int)snel[1];
    M.z0 = gen ? (int)(A.cell_first / lay) : 0;
    M.z1 = gen ? M.z0 + (int)(A.ncount / lay) : (int)snel[2];
    M.cellmap = gen ? nullptr : g->d_sv_cellmap;
    M.cell_lo = gen ? 0 : A.cell_first;
    M.cell_hi = gen ? g->ncells : A.cell_first + A.ncount;
    M.tiles_x = (M.nx + 3) / 4;
    M.tiles_y = (M.ny + 3) / 4;
    M.cap = fb2_mvec_cap(std::max(a->pat->max_col_len, 1));
    M.overwrite = accumulate ? 0 : 1;
    const size_t smem = fb2_mvec_smem(M.cap);
    if (smem > 113 * 1024 || M.cap >= 65536) return FB2_OK;
    FB2_TRY(fb2_map_build_vec(a));
    M.mapv = a->d_mapv;
    const int64_t tiles = (int64_t)M.tiles_x * M.tiles_y;
    int nchunks = 1;
    marchv_plan(ctx, M.tiles_x, M.tiles_y, M.z1 - M.z0, &M.lz, &nchunks);
    A.p[5] = 0.125 * cv->w[0];   // the common quadrature weight / 8
    if (const char* e = getenv("FB2_MVEC_DBG")) M.dbg = atoi(e);
    // FB2_MVEC_FLUSH=thread: flush with coalesced per-thread stores / REDs instead of the TMA engine (A/B measurement)
    const char* ef = getenv("FB2_MVEC_FLUSH");
    const bool tmaf = !(ef && strcmp(ef, "thread") == 0);
    auto k = general ? (a->map_complete ? k_march_vec<false, true, true> : k_march_vec<true, true, true>)
             : tmaf  ? (a->map_complete ? k_march_vec<false, true, false> : k_march_vec<true, true, false>)
                     : (a->map_complete ? k_march_vec<false, false, false> : k_march_vec<true, false, false>);
    MarchCmat CM;
    memcpy(CM.c, a->h_cmat, sizeof(CM.c));
    FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // start_assemble's zero fill of nzval, if still owed: only the columns that receive reduce-adds need it (nodes on tile
    // faces and on the planes where two chunks meet, 47 % of nzval at 128^3); with three columns = 1.9 KB per node the
    // column-wise fill beats the memset over everything here (FB2_MARCH_ZSEL=0 keeps the memset)
    if (A.zero_pending) {
        const char* ez = getenv("FB2_MARCH_ZSEL");
        if (!(ez && atoi(ez) == 0) && gen && M.z0 == 0 && M.z1 == g->nel[2] && a->dh->ndofs < (int64_t)1 << 31) {
            FB2_TRY(march_zero_list(a, M.lz, 4, 4, 3));
            if (a->march_nzcols > 0)
                k_zero_columns<<<(unsigned)((a->march_nzcols + 255) / 256), 256, 0, ctx->stream>>>(a->d_march_zcols, a->march_nzcols,
                                                                                              a->pat->d_colptr, A.nzval, a->pat->nnz);
            ctx->launches++;
            A.zero_pending = 0;
        } else {
            FB2_TRY(pay_zero_fill(a, A));
        }
    }
    int64_t nctas = tiles * nchunks;
    if (a->march_part != 0 && !gen) {   // split launch of the partitioned exchange path
        M.nfull = 0;
        M.lt = M.lz;
        FB2_TRY(march_cta_lists(a, M, 4, 4, nchunks, A.cell_first + A.ncount, true));
        M.ctalist = a->d_cta_list[a->march_part - 1];
        nctas = a->cta_count[a->march_part - 1];
    }
    if (nctas > 0) k<<<(unsigned)nctas, 128, smem, ctx->stream>>>(A, M, CM);
    g_fb2_last_kernel = "k_march_vec";
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    *handled = true;
    return FB2_OK;
}

// tile kernel (tiles.cu): whole grid, atomic mode only; falls back to the per-cell kernel when no schedule exists
template <int DIM, int NGEO, int NB, int NQ, int ELEM>
int launch_tiles_or_cells(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic, int variant, int accumulate) {
    constexpr int NSYM = NB * (NB + 1) / 2;
    constexpr int TC = (NSYM + NB) <= 44 ? 128 : 64;   // one cell per thread; two CTAs per SM overlap their phases
    // The tile kernel is opt-in (variant 5): on B200 it is slower than the per-cell kernel for Q1 hex (phase 2 costs as
    // many instructions as it saves in L2 traffic); kept because it removes 80 % of the RED traffic.
    if (atomic && variant == 5 && A.cells == nullptr && A.ncount == a->dh->grid->ncells && a->dh->grid->ncells >= 8 * TC) {
        FB2_TRY(fb2_tiles_build(a, TC));
        if (a->tiles) {
            const TileSchedule* S = a->tiles;
            TileArgs T{S->d_conn, S->d_ncells, S->d_cell_ids, S->d_tile_base, S->d_ent_ptr, S->d_rec, S->d_src_ptr, S->d_src,
                       S->max_ent, S->max_src, accumulate};
            const size_t smem = (size_t)(NSYM + NB) * TC * sizeof(double) + (size_t)S->max_ent * sizeof(uint2) +
                                (size_t)S->max_src * sizeof(uint16_t);
            if (smem <= 227 * 1024) {
                auto k = k_tile_scalar<DIM, NGEO, NB, NQ, ELEM, TC, 2>;
                FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k<<<(unsigned)S->ntiles, TC, smem, ctx->stream>>>(A, T);
                g_fb2_last_kernel = "k_tile_scalar";
                ctx->launches++;
                FB2_CUDA(cudaGetLastError());
                return FB2_OK;
            }
        }
    }
    FB2_TRY(fb2_map_build_packed(a));
    A.map8 = reinterpret_cast<const uint4*>(a->d_map8);
    // Default for Q1 hexahedra of generate_grid (atomic mode, whole layers): the marching-tile kernel (march_kernels.cuh).
    // variant 30 forces the thread-per-cell kernel for comparison.
    if constexpr (DIM == 3 && NGEO == 8 && NB == 8 && NQ == 8) {
        fb2_grid* g = a->dh->grid;
        // structured view: generate_grid order (cells x fastest; any whole-layer range), or a box with a cell map (the local
        // grid of a block partition; any range of cell ids)
        const bool gen = g->generated && g->celltype == FB2_HEXAHEDRON;
        const bool boxed = !gen && g->structured && g->d_sv_cellmap != nullptr;
        const int64_t* snel = gen ? g->nel : g->sv_nel;
        const int64_t lay = (gen || boxed) ? snel[0] * snel[1] : 1;
        const bool range_ok = gen ? (A.cell_first % lay == 0 && A.ncount % lay == 0) : boxed;
        if (atomic && (variant == 0 || variant == 31) && (gen || boxed) && A.cells == nullptr && A.ncount > 0 && range_ok &&
            snel[0] < (1 << 28) && snel[1] < (1 << 28) && march_usable(a)) {
            MarchArgs M;
            memset(&M, 0, sizeof(M));
            M.nx = (int)snel[0];
            M.ny = (int)snel[1];
            M.z0 = gen ? (int)(A.cell_first / lay) : 0;
            M.z1 = gen ? M.z0 + (int)(A.ncount / lay) : (int)snel[2];
            M.cellmap = gen ? nullptr : g->d_sv_cellmap;
            M.cell_lo = gen ? 0 : A.cell_first;
            M.cell_hi = gen ? g->ncells : A.cell_first + A.ncount;
            M.tiles_x = (M.nx + 7) / 8;
            M.tiles_y = (M.ny + 3) / 4;
            M.cap = fb2_march_cap(std::max(a->pat->max_col_len, 1));
            M.overwrite = accumulate ? 0 : 1;
            const size_t smem = fb2_march_smem(M.cap);
            // the bulk (TMA) flush needs 16-byte aligned matrix storage and columns short enough for the byte-packed map
            if (smem <= 100 * 1024 && M.cap < 65536 && a->pat->max_col_len < 255 && (reinterpret_cast<uintptr_t>(A.nzval) & 15) == 0) {
                FB2_TRY(fb2_map_build_bytes(a));
                M.mapb = a->d_mapb;
                const int64_t tiles = (int64_t)M.tiles_x * M.tiles_y;
                int nchunks = 1;
                marchh_plan(ctx, M.tiles_x, M.tiles_y, M.z1 - M.z0, &M.lz, &M.nfull, &M.lt, &nchunks);
                // variant 31: table-driven integration inside the marching kernel (A/B against the analytic element)
                const bool analytic = ELEM == FB2_ELEM_HEAT && variant != 31 && cv_is_q1hex_gauss2(a->cv);
                if (analytic) A.p[2] = 0.125 * a->cv->w[0];   // the common quadrature weight / 8 (see fb2_hex8_heat)
                auto k = a->map_complete ? (analytic ? k_march_hex<ELEM, false, true> : k_march_hex<ELEM, false, false>)
                                         : (analytic ? k_march_hex<ELEM, true, true> : k_march_hex<ELEM, true, false>);
                FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                // eight single-warp CTAs per SM need 8 x 28 KB: ask for the full shared-memory carveout
                FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                // start_assemble's zero fill of nzval, if still owed.  Only the columns that receive reduce-adds need it
                // (38 % of nzval on C2: nodes on tile faces and on the planes between chunks).  Zeroed column by column this
                // took as long as the memset over everything (0.28 vs 0.30 ms); as runs of adjacent columns rounded to whole
                // 32-byte sectors (k_zero_columns) it takes 0.22 ms: C2 step 1.93 -> 1.85 ms.  FB2_MARCH_ZSEL=0 keeps the memset.
                if (A.zero_pending) {
                    const char* ez = getenv("FB2_MARCH_ZSEL");
                    if (!(ez && atoi(ez) == 0) && gen && M.z0 == 0 && M.z1 == g->nel[2] && a->dh->ndofs < (int64_t)1 << 31) {
                        FB2_TRY(march_zero_list(a, M.lz, 8, 4, 1, M.nfull, M.lt));
                        if (a->march_nzcols > 0)
                            k_zero_columns<<<(unsigned)((a->march_nzcols + 255) / 256), 256, 0, ctx->stream>>>(a->d_march_zcols, a->march_nzcols,
                                                                                                          a->pat->d_colptr, A.nzval, a->pat->nnz);
                        ctx->launches++;
                    } else {
                        FB2_CUDA(cudaMemsetAsync(A.nzval, 0, (size_t)a->pat->nnz * sizeof(double), ctx->stream));
                    }
                    A.zero_pending = 0;
                }
                int64_t nctas = tiles * nchunks;
                if (a->march_part != 0 && !gen) {   // split launch of the partitioned exchange path
                    FB2_TRY(march_cta_lists(a, M, 8, 4, nchunks, A.cell_first + A.ncount, true));
                    M.ctalist = a->d_cta_list[a->march_part - 1];
                    nctas = a->cta_count[a->march_part - 1];
                }
                if (nctas > 0) k<<<(unsigned)nctas, 32, smem, ctx->stream>>>(A, M);
                g_fb2_last_kernel = "k_march_hex";
                ctx->launches++;
                FB2_CUDA(cudaGetLastError());
                return FB2_OK;
            }
        }
    }
    FB2_TRY(pay_zero_fill(a, A));   // the marching-tile kernel was not applicable
    // variant 6: launch over the warp list, which adds the y-merge through shared memory.  Measured on C2 it removes a
    // further ~20 % of the REDs but costs two CTA barriers and 12 % padding lanes (200 = 6*32 + 8): 2.78 ms vs 2.41 ms
    // with the x-merge alone, so it is not the default.
    if (atomic && A.cells == nullptr && A.ncount == a->dh->grid->ncells && variant == 6) {
        FB2_TRY(fb2_warplist_build(a));
        A.wfirst = a->d_wfirst;
        A.wcount = a->d_wcount;
        A.ncount = a->nwarps * 32;
    }
    return launch_scalar<DIM, NGEO, NB, NQ, ELEM>(ctx, A, atomic, variant == 30 ? 0 : variant, a->map_complete);
}

template <int ELEM>
bool try_scalar(fb2_assembler* a, fb2_ctx* ctx, AsmArgs& A, bool atomic, int variant, int accumulate, int celltype, int nb,
                int nq, int* rc) {
    if (!A.const_ok) return false;   // the thread-per-cell kernels need the tables in the constant bank
#define CASE(CT, DIM, NGEO, NB, NQ)                                              \
    if (celltype == CT && nb == NB && nq == NQ) {                                \
        *rc = launch_tiles_or_cells<DIM, NGEO, NB, NQ, ELEM>(a, ctx, A, atomic, variant, accumulate); \
        return true;                                                             \
    }
    CASE(FB2_QUADRILATERAL, 2, 4, 4, 4)
    CASE(FB2_HEXAHEDRON, 3, 8, 8, 8)
    CASE(FB2_TRIANGLE, 2, 3, 3, 1)
    CASE(FB2_TRIANGLE, 2, 3, 3, 3)
    CASE(FB2_TRIANGLE, 2, 3, 6, 3)
    CASE(FB2_TETRAHEDRON, 3, 4, 4, 1)
    CASE(FB2_TETRAHEDRON, 3, 4, 4, 4)
    CASE(FB2_TETRAHEDRON, 3, 4, 10, 4)
#undef CASE
    return false;
}

// assemble! of stored element matrices.  Ke is cell-major (Ke[c][e], e = j*n + i) while the offset map and the dofs are
// entry-major SoA ([e][cell]); a CTA therefore takes 32 cells x EC entries, reads the Ke tile with consecutive lanes along e,
// turns it in shared memory and scatters with consecutive lanes along the cells, so that both sides are coalesced.
constexpr int SB_EC = 128;   // entries per tile
__global__ void __launch_bounds__(256) k_scatter_batch(const double* __restrict__ Ke, const double* __restrict__ fe,
                                                       const int32_t* __restrict__ cell_dofs, const int64_t* __restrict__ colptr,
                                                       const uint16_t* __restrict__ map, int64_t ncells, int64_t ncells_pad, int n,
                                                       double* __restrict__ nzval, double* __restrict__ f, int* errflag) {
    __shared__ double s_v[32][SB_EC + 1];
    const int per = n * n;
    const int64_t c0 = (int64_t)blockIdx.x * 32;
    const int ncl = (int)min((int64_t)32, ncells - c0);
    const int e0 = blockIdx.y * SB_EC;
    const int ne = min(SB_EC, per - e0);
    for (int idx = threadIdx.x; idx < ncl * ne; idx += 256) {
        const int cl = idx / ne, ee = idx - cl * ne;
        s_v[cl][ee] = __ldcs(Ke + (size_t)(c0 + cl) * per + e0 + ee);
    }
    __syncthreads();
    const int cl = threadIdx.x & 31;
    if (cl < ncl) {
        const int64_t c = c0 + cl;
        for (int ee = threadIdx.x >> 5; ee < ne; ee += 8) {
            const double v = s_v[cl][ee];
            if (v != 0.0) {
                const int e = e0 + ee;
                const unsigned off = map[(size_t)e * ncells_pad + c];
                if (off == 0xFFFFu) fb2_flag_error(errflag, FB2_ERR_MISSING_PATTERN_ENTRY, c);
                else atomicAdd(nzval + colptr[cell_dofs[(size_t)(e / n) * ncells_pad + c]] + off, v);
            }
        }
        if (f != nullptr && fe != nullptr && blockIdx.y == 0)
            for (int i = threadIdx.x >> 5; i < n; i += 8) atomicAdd(f + cell_dofs[(size_t)i * ncells_pad + c], fe[(size_t)c * n + i]);
    }
}

}  // namespace

// Warp list of k_cell_scalar: every warp covers <= 32 consecutive cells; the 4 warps of a CTA cover the same x-range of
// 4 consecutive grid rows when the grid comes from generate_grid (cells are stored row by row, src/Grid/
// grid_generators.jl:96-98,170-178), so that the kernel can merge shared faces in x (shuffles) and y (shared memory).
// Other grids get consecutive chunks (the merges then fire wherever neighbouring lanes happen to share a face).
int fb2_warplist_build(fb2_assembler* a) {
    if (a->d_wfirst) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    std::vector<int32_t> first;
    std::vector<uint8_t> count;
    const bool rows = g->generated && (g->celltype == FB2_HEXAHEDRON || g->celltype == FB2_QUADRILATERAL);
    if (rows) {
        const int64_t nx = g->nel[0], ny = g->nel[1], nz = g->celltype == FB2_HEXAHEDRON ? g->nel[2] : 1;
        for (int64_t k = 0; k < nz; ++k)
            for (int64_t j0 = 0; j0 < ny; j0 += 4)
                for (int64_t i0 = 0; i0 < nx; i0 += 32)
                    for (int r = 0; r < 4; ++r) {
                        const int64_t j = j0 + r;
                        if (j < ny) {
                            first.push_back((int32_t)(i0 + nx * (j + ny * k)));
                            count.push_back((uint8_t)std::min<int64_t>(32, nx - i0));
                        } else {
                            first.push_back(0);
                            count.push_back(0);
                        }
                    }
    } else {
        for (int64_t c = 0; c < g->ncells; c += 32) {
            first.push_back((int32_t)c);
            count.push_back((uint8_t)std::min<int64_t>(32, g->ncells - c));
        }
        while (first.size() % 4) { first.push_back(0); count.push_back(0); }
    }
    a->nwarps = (int64_t)first.size();
    FB2_CUDA(cudaSetDevice(g->ctx->device));
    FB2_CUDA(cudaMalloc(&a->d_wfirst, first.size() * sizeof(int32_t)));
    FB2_CUDA(cudaMalloc(&a->d_wcount, count.size()));
    FB2_CUDA(cudaMemcpy(a->d_wfirst, first.data(), first.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    FB2_CUDA(cudaMemcpy(a->d_wcount, count.data(), count.size(), cudaMemcpyHostToDevice));
    return FB2_OK;
}

int fb2_march_split_zero_fill(fb2_assembler* a, int element, int64_t ncells_own, double* nzval_dev, bool* done) {
    *done = false;
    const char* ez = getenv("FB2_MARCH_ZSEL");
    if (ez && atoi(ez) == 0) return FB2_OK;
    const bool vec = element == FB2_ELEM_ELASTICITY || element == FB2_ELEM_ELASTICITY_GENERAL;
    const bool scalar = element == FB2_ELEM_HEAT || element == FB2_ELEM_MASS;
    if (vec ? !marchv_static_ok(a) : !(scalar && fb2_march_applicable(a, element, nullptr))) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    if (g->generated || !g->structured || a->dh->ndofs >= (int64_t)1 << 31 || (reinterpret_cast<uintptr_t>(nzval_dev) & 15) != 0) return FB2_OK;
    // the same tiles and chunks as the launch will use for the own cells of the box
    MarchArgs M;
    memset(&M, 0, sizeof(M));
    M.nx = (int)g->sv_nel[0];
    M.ny = (int)g->sv_nel[1];
    M.z0 = 0;
    M.z1 = (int)g->sv_nel[2];
    const int tx = vec ? 4 : 8, ty = 4;
    M.tiles_x = (M.nx + tx - 1) / tx;
    M.tiles_y = (M.ny + ty - 1) / ty;
    int nchunks = 1;
    if (vec) { marchv_plan(ctx, M.tiles_x, M.tiles_y, M.z1 - M.z0, &M.lz, &nchunks); M.lt = M.lz; }
    else marchh_plan(ctx, M.tiles_x, M.tiles_y, M.z1 - M.z0, &M.lz, &M.nfull, &M.lt, &nchunks);
    FB2_TRY(march_cta_lists(a, M, tx, ty, nchunks, ncells_own, true));
    if (a->box_nzcols < 0) return FB2_OK;
    if (a->box_nzcols > 0)
        k_zero_columns<<<(unsigned)((a->box_nzcols + 255) / 256), 256, 0, ctx->stream>>>(a->d_box_zcols, a->box_nzcols, a->pat->d_colptr, nzval_dev,
                                                                                     a->pat->nnz);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    *done = true;
    return FB2_OK;
}

int fb2_map_build_vec(fb2_assembler* a) {
    if (a->d_mapv) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CHECK(a->n == 24 && a->pat->max_col_len < 255, FB2_ERR_UNSUPPORTED, "lane-major byte map: 24 dofs per cell and columns shorter than 255 entries");
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int64_t total = g->ncells * 160;
    FB2_CUDA(cudaMalloc(&a->d_mapv, std::max<int64_t>(total, 1) * sizeof(uint32_t)));
    k_pack_map_vec<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(a->d_map, g->ncells, g->ncells_pad, a->d_mapv);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    return FB2_OK;
}

bool fb2_march_applicable(fb2_assembler* a, int element, const fb2_asm_opts* opts) {
    const fb2_cv* cv = a->cv;
    const fb2_grid* g = a->dh->grid;
    if (cv && (element == FB2_ELEM_ELASTICITY || element == FB2_ELEM_ELASTICITY_GENERAL))
        return !(opts && (opts->scatter_mode != FB2_SCATTER_ATOMIC || opts->variant != 0)) && marchv_static_ok(a);
    if (!cv || !(element == FB2_ELEM_HEAT || element == FB2_ELEM_MASS)) return false;
    if (opts && (opts->scatter_mode != FB2_SCATTER_ATOMIC || !(opts->variant == 0 || opts->variant == 31))) return false;
    if (cv->celltype != FB2_HEXAHEDRON || cv->nb != 8 || cv->nq != 8 || cv->vdim != 1 || cv->ngeo != 8) return false;
    const bool gen = g->generated && g->celltype == FB2_HEXAHEDRON, boxed = !gen && g->structured && g->d_sv_cellmap != nullptr;
    if (!(gen || boxed) || a->pat->max_col_len >= 255) return false;
    return march_usable(a);
}

int fb2_check_device_error(fb2_ctx* ctx) {
    FB2_CUDA(cudaMemcpyAsync(ctx->h_errflag, ctx->d_errflag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    const int code = ctx->h_errflag[0], cell = ctx->h_errflag[1];
    if (code == 0) return FB2_OK;
    FB2_CUDA(cudaMemsetAsync(ctx->d_errflag, 0, 2 * sizeof(int), ctx->stream));
    if (code == FB2_ERR_DETJ_NOT_POSITIVE)
        return fb2_fail(code, "det(J) is not positive in cell %d (1-based); check the node ordering of the cell", cell + 1);
    if (code == FB2_ERR_MISSING_PATTERN_ENTRY)
        return fb2_fail(code, "a non-zero element matrix entry of cell %d (1-based) is missing in the sparsity pattern", cell + 1);
    return fb2_fail(FB2_ERR_INTERNAL, "device error flag %d in cell %d", code, cell + 1);
}

int fb2_coloring_build(fb2_assembler* a) {
    // Greedy colouring: no two cells of one colour share a node (src/Grid/coloring.jl:108-160).  The colours
    // need not equal the reference's, only be valid; cells keep ascending order inside each colour.
    if (a->ncolors > 0) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    const int64_t nc = g->ncells;
    const int nnpc = g->nnpc;
    std::vector<uint64_t> mask_lo((size_t)g->nnodes, 0), mask_hi((size_t)g->nnodes, 0);
    a->cell_color.assign((size_t)nc, 0);
    int ncol = 0;
    for (int64_t c = 0; c < nc; ++c) {
        const int64_t* cell = &g->cells[(size_t)c * nnpc];
        uint64_t lo = 0, hi = 0;
        for (int k = 0; k < nnpc; ++k) { lo |= mask_lo[cell[k] - 1]; hi |= mask_hi[cell[k] - 1]; }
        int col = -1;
        if (~lo) col = __builtin_ctzll(~lo);
        else if (~hi) col = 64 + __builtin_ctzll(~hi);
        FB2_CHECK(col >= 0, FB2_ERR_UNSUPPORTED, "colouring needs more than 128 colours");
        a->cell_color[c] = col;
        ncol = std::max(ncol, col + 1);
        for (int k = 0; k < nnpc; ++k) {
            if (col < 64) mask_lo[cell[k] - 1] |= 1ull << col;
            else mask_hi[cell[k] - 1] |= 1ull << (col - 64);
        }
    }
    a->ncolors = ncol;
    a->color_ptr.assign(ncol + 1, 0);
    for (int64_t c = 0; c < nc; ++c) a->color_ptr[a->cell_color[c] + 1]++;
    for (int k = 0; k < ncol; ++k) a->color_ptr[k + 1] += a->color_ptr[k];
    std::vector<int32_t> order((size_t)nc);
    std::vector<int64_t> cur(a->color_ptr.begin(), a->color_ptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c) order[cur[a->cell_color[c]]++] = (int32_t)c;
    FB2_CUDA(cudaSetDevice(g->ctx->device));
    FB2_CUDA(cudaMalloc(&a->d_color_cells, (size_t)nc * sizeof(int32_t)));
    FB2_CUDA(cudaMemcpy(a->d_color_cells, order.data(), (size_t)nc * sizeof(int32_t), cudaMemcpyHostToDevice));
    return FB2_OK;
}

static int launch_one(fb2_assembler* a, AsmArgs& A, int element, bool atomic, int variant, int accumulate) {
    fb2_cv* cv = a->cv;
    fb2_ctx* ctx = a->dh->grid->ctx;
    const int ct = cv->celltype, nbs = cv->nb, vdim = cv->vdim;
    int rc = FB2_OK;
    // only the marching-tile kernel (Q1 hexahedra, heat / mass, default variant) takes the fill over
    const bool may_fuse = atomic && ct == FB2_HEXAHEDRON && nbs == 8 && cv->nq == 8 &&
                          (((element == FB2_ELEM_HEAT || element == FB2_ELEM_MASS) && (variant == 0 || variant == 31) && vdim == 1) ||
                           ((element == FB2_ELEM_ELASTICITY || element == FB2_ELEM_ELASTICITY_GENERAL) && variant == 0 && vdim == 3));
    if (!may_fuse) FB2_TRY(pay_zero_fill(a, A));
    switch (element) {
        case FB2_ELEM_HEAT:
            FB2_CHECK(vdim == 1, FB2_ERR_BAD_ARG, "the heat element needs a scalar field");
            if (variant != 1 && try_scalar<FB2_ELEM_HEAT>(a, ctx, A, atomic, variant, accumulate, ct, nbs, cv->nq, &rc)) return rc;
            return dispatch_blocks<FB2_ELEM_HEAT, 0>(a, ctx, A, atomic, ct, nbs, vdim);
        case FB2_ELEM_MASS:
            FB2_CHECK(vdim == 1, FB2_ERR_UNSUPPORTED, "the mass element is implemented for scalar fields");
            if (variant != 1 && try_scalar<FB2_ELEM_MASS>(a, ctx, A, atomic, variant, accumulate, ct, nbs, cv->nq, &rc)) return rc;
            return dispatch_blocks<FB2_ELEM_MASS, 0>(a, ctx, A, atomic, ct, nbs, vdim);
        case FB2_ELEM_ELASTICITY:
            if (variant == 0 && atomic) {   // structured trilinear hexahedra: the marching-tile kernel (variant 32 = without it)
                bool handled = false;
                rc = try_march_vec(a, ctx, A, accumulate, &handled);
                if (rc != FB2_OK || handled) return rc;
            }
            FB2_TRY(pay_zero_fill(a, A));   // the marching-tile kernel was not applicable
            if (variant != 1) {   // variant 1: the DFMA block kernel
                bool handled = false;
                rc = dispatch_syrk(a, ctx, A, atomic, ct, nbs, vdim, &handled);
                if (handled) return rc;
            }
            return dispatch_blocks<FB2_ELEM_ELASTICITY, 1>(a, ctx, A, atomic, ct, nbs, vdim);
        case FB2_ELEM_ELASTICITY_GENERAL:
            FB2_CHECK(vdim == cv->rdim, FB2_ERR_BAD_ARG, "the elasticity element needs a vector field with as many components as space dimensions");
            if (variant == 0 && atomic) {   // structured trilinear hexahedra: the marching-tile kernel with C as an argument
                bool handled = false;
                rc = try_march_vec(a, ctx, A, accumulate, &handled, true);
                if (rc != FB2_OK || handled) return rc;
            }
            FB2_TRY(pay_zero_fill(a, A));
            return dispatch_blocks<FB2_ELEM_ELASTICITY_GENERAL, 1>(a, ctx, A, atomic, ct, nbs, vdim);
        case FB2_ELEM_NEOHOOKE:
            FB2_CHECK(A.u != nullptr, FB2_ERR_BAD_ARG, "the Neo-Hooke element needs the current solution u");
            return dispatch_neohooke(a, ctx, A, atomic, ct, nbs, vdim);
    }
    return fb2_fail(FB2_ERR_BAD_ARG, "unknown element id %d", element);
}

int fb2_launch_assemble(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* u_dev,
                        double* nzval_dev, double* f_dev, const fb2_asm_opts* opts) {
    fb2_dh* dh = a->dh;
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    if (opts) o = *opts;
    AsmArgs A;
    memset(&A, 0, sizeof(A));
    FB2_TRY(upload_tables(a, &A));
    A.conn = g->d_conn;
    A.xyz = g->d_xyz;
    A.cell_dofs = dh->d_cell_dofs;
    A.colptr = a->pat->d_colptr;
    A.map = a->d_map;
    A.map8 = reinterpret_cast<const uint4*>(a->d_map8);
    A.ncells_pad = g->ncells_pad;
    A.nzval = nzval_dev;
    A.f = f_dev;
    A.u = u_dev;
    A.errflag = ctx->d_errflag;
    switch (element) {
        case FB2_ELEM_HEAT: {
            fb2_heat_params p = {1.0, 1.0};
            if (params) { FB2_CHECK(params_bytes == sizeof(p), FB2_ERR_BAD_ARG, "heat params: expected %zu bytes", sizeof(p)); memcpy(&p, params, sizeof(p)); }
            A.p[0] = p.k; A.p[1] = p.source;
            break;
        }
        case FB2_ELEM_MASS: {
            fb2_mass_params p = {1.0};
            if (params) { FB2_CHECK(params_bytes == sizeof(p), FB2_ERR_BAD_ARG, "mass params: expected %zu bytes", sizeof(p)); memcpy(&p, params, sizeof(p)); }
            A.p[0] = p.rho; A.p[1] = 0.0;
            break;
        }
        case FB2_ELEM_ELASTICITY:
        case FB2_ELEM_NEOHOOKE: {
            fb2_elasticity_params p;
            FB2_CHECK(params && params_bytes == sizeof(p), FB2_ERR_BAD_ARG, "elasticity params: expected %zu bytes", sizeof(p));
            memcpy(&p, params, sizeof(p));
            A.p[0] = p.lambda; A.p[1] = p.mu; A.p[2] = p.b[0]; A.p[3] = p.b[1]; A.p[4] = p.b[2];
            break;
        }
        case FB2_ELEM_ELASTICITY_GENERAL: {
            fb2_elasticity_general_params p;
            FB2_CHECK(params && params_bytes == sizeof(p), FB2_ERR_BAD_ARG, "general elasticity params: expected %zu bytes", sizeof(p));
            memcpy(&p, params, sizeof(p));
            // minor symmetries C_ijkl = C_jikl = C_ijlk (SymmetricTensor{4}); the kernel relies on them
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    for (int k = 0; k < 3; ++k)
                        for (int l = 0; l < 3; ++l) {
                            const double c = p.C[((i * 3 + j) * 3 + k) * 3 + l];
                            const double tol = 1e-12 * (std::fabs(c) + 1e-300);
                            FB2_CHECK(std::fabs(c - p.C[((j * 3 + i) * 3 + k) * 3 + l]) <= tol && std::fabs(c - p.C[((i * 3 + j) * 3 + l) * 3 + k]) <= tol,
                                      FB2_ERR_BAD_ARG, "general elasticity: C lacks the minor symmetries at (%d,%d,%d,%d)", i + 1, j + 1, k + 1, l + 1);
                        }
            if (!a->d_cmat) FB2_CUDA(cudaMalloc(&a->d_cmat, 81 * sizeof(double)));
            memcpy(a->h_cmat, p.C, sizeof(a->h_cmat));
            FB2_CUDA(cudaMemcpyAsync(a->d_cmat, p.C, 81 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            FB2_CUDA(cudaStreamSynchronize(ctx->stream));   // p lives on this stack frame
            A.cmat = a->d_cmat;
            A.p[2] = p.b[0]; A.p[3] = p.b[1]; A.p[4] = p.b[2];
            break;
        }
        default: return fb2_fail(FB2_ERR_BAD_ARG, "unknown element id %d", element);
    }
    // Zero fill overlapped with the assembly: the cells are launched in ZS ranges, and the part of nzval a range writes
    // is zeroed on a second stream while the previous range is assembled (the fill is DRAM-bound, the kernel is not).
    // Needs a numbering that follows the cell order (the reference's does).  Opt-in (variant 12): on C2 it is SLOWER than
    // the plain memset in front of one launch (2.61 vs 2.54 ms per step) -- the fill's DRAM writes slow the concurrent
    // kernel down by more than the 0.25 ms they hide, and four launches have four tails.
    constexpr int ZS = 4;
    const bool whole_grid = a->d_cells == nullptr && a->ncells_active == 0;
    if (o.fillzero && o.scatter_mode == FB2_SCATTER_ATOMIC && whole_grid && g->ncells >= 1024 && o.variant == 12 && a->zf_state != 2) {
        if (a->zf_state == 0) {
            const int64_t nc = g->ncells;
            const int ndpc = dh->ndpc;
            std::vector<int64_t> cell(ZS + 1), col(ZS + 1, 0), pos(ZS + 1, 0);
            for (int k = 0; k <= ZS; ++k) cell[k] = std::min<int64_t>(nc, ((nc * k / ZS) + 127) / 128 * 128);
            cell[ZS] = nc;
            int64_t mx = 0;
            for (int k = 0; k < ZS; ++k) {
                for (int64_t c = cell[k]; c < cell[k + 1]; ++c)
                    for (int i = 0; i < ndpc; ++i) mx = std::max<int64_t>(mx, dh->cell_dofs[(size_t)c * ndpc + i] + 1);
                col[k + 1] = mx;   // columns [0, mx) are touched by ranges <= k
            }
            col[ZS] = dh->ndofs;
            for (int k = 0; k <= ZS; ++k)
                FB2_CUDA(cudaMemcpyAsync(&pos[k], a->pat->d_colptr + col[k], sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            FB2_CUDA(cudaStreamSynchronize(ctx->stream));
            a->zf_cell = cell;
            a->zf_pos = pos;
            // worthwhile only if the first range does not already need most of the matrix
            a->zf_state = pos[1] <= a->pat->nnz / 2 ? 1 : 2;
        }
        if (a->zf_state == 1) {
            if (!ctx->h2d_stream) FB2_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
            if (!ctx->ev_ready) {
                for (cudaEvent_t& e : ctx->ev_pool) FB2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->ev_ready = true;
            }
            cudaStream_t aux = ctx->h2d_stream;
            cudaEvent_t* ev = ctx->ev_pool + 32;   // [0] start, [1..ZS] fills, [ZS+1] aux idle
            FB2_CUDA(cudaEventRecord(ev[0], ctx->stream));
            FB2_CUDA(cudaStreamWaitEvent(aux, ev[0], 0));
            if (f_dev) FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)dh->ndofs * sizeof(double), aux));
            for (int k = 0; k < ZS; ++k) {
                const int64_t p0 = a->zf_pos[k], p1 = a->zf_pos[k + 1];
                if (p1 > p0) FB2_CUDA(cudaMemsetAsync(nzval_dev + p0, 0, (size_t)(p1 - p0) * sizeof(double), aux));
                FB2_CUDA(cudaEventRecord(ev[1 + k], aux));
            }
            A.cells = nullptr;
            int rc = FB2_OK;
            for (int k = 0; k < ZS && rc == FB2_OK; ++k) {
                FB2_CUDA(cudaStreamWaitEvent(ctx->stream, ev[1 + k], 0));
                A.cell_first = a->zf_cell[k];
                A.ncount = a->zf_cell[k + 1] - a->zf_cell[k];
                if (A.ncount > 0) rc = launch_one(a, A, element, true, o.variant, 1);
            }
            return rc;
        }
    }
    if (o.fillzero) {
        // the fill of nzval is owed to the kernel launch below (launch_one pays it with a memset unless the kernel fuses it)
        const bool subset_now = a->d_cells != nullptr || a->ncells_active > 0;
        if (o.scatter_mode == FB2_SCATTER_COLORED || subset_now) FB2_CUDA(cudaMemsetAsync(nzval_dev, 0, (size_t)a->pat->nnz * sizeof(double), ctx->stream));
        else A.zero_pending = 1;
        if (f_dev) FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)dh->ndofs * sizeof(double), ctx->stream));
    }
    if (o.scatter_mode == FB2_SCATTER_COLORED) {
        FB2_CHECK(a->d_cells == nullptr && a->ncells_active == 0, FB2_ERR_UNSUPPORTED, "coloured scatter on a partitioned assembler is not supported");
        FB2_TRY(fb2_coloring_build(a));
        for (int c = 0; c < a->ncolors; ++c) {
            A.cells = a->d_color_cells + a->color_ptr[c];
            A.ncount = a->color_ptr[c + 1] - a->color_ptr[c];
            if (A.ncount == 0) continue;
            FB2_TRY(launch_one(a, A, element, false, o.variant, !o.fillzero));
        }
        return FB2_OK;
    }
    A.cells = a->d_cells;
    const bool subset = a->d_cells != nullptr || a->ncells_active > 0;
    A.cell_first = subset && !a->d_cells ? a->cell_first : 0;
    A.ncount = subset ? a->ncells_active : g->ncells;
    if (A.ncount == 0) return pay_zero_fill(a, A);
    // split marching launch (fb2_assemble_distributed): the caller paid the zero fill in front of both parts
    const int accumulate = a->march_part != 0 ? (a->march_overwrite ? 0 : 1) : !o.fillzero;
    return launch_one(a, A, element, true, o.variant, accumulate);
}

// assemble!(a, celldofs(cell), Ke, fe) for every cell from device-resident element matrices (the layout of fb2_ea_assemble:
// Ke[c*n*n + j*n + i], fe[c*n + i]); src/assembler.jl:322-331,347-457 (zero skip, missing-entry error)
extern "C" int fb2_scatter_device(fb2_assembler* a, const double* Ke_dev, const double* fe_dev, double* nzval_dev, double* f_dev,
                                  const fb2_asm_opts* opts) {
    FB2_CHECK(a && Ke_dev && nzval_dev, FB2_ERR_BAD_ARG, "fb2_scatter_device: null argument");
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    if (opts) o = *opts;
    const int n = a->n;
    if (o.fillzero) {
        FB2_CUDA(cudaMemsetAsync(nzval_dev, 0, (size_t)a->pat->nnz * sizeof(double), ctx->stream));
        if (f_dev) FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)a->dh->ndofs * sizeof(double), ctx->stream));
    }
    const int64_t total = (int64_t)n * n * g->ncells;
    if (total == 0) return FB2_OK;
    const dim3 grid((unsigned)((g->ncells + 31) / 32), (unsigned)((n * n + SB_EC - 1) / SB_EC));
    k_scatter_batch<<<grid, 256, 0, ctx->stream>>>(Ke_dev, f_dev ? fe_dev : nullptr, a->dh->d_cell_dofs,
                                                                           a->pat->d_colptr, a->d_map, g->ncells, g->ncells_pad, n,
                                                                           nzval_dev, f_dev, ctx->d_errflag);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return fb2_check_device_error(ctx);
}

extern "C" int fb2_scatter_host(fb2_assembler* a, const double* Ke, const double* fe, double* nzval_dev, double* f_dev,
                                const fb2_asm_opts* opts) {
    FB2_CHECK(a && Ke && nzval_dev, FB2_ERR_BAD_ARG, "fb2_scatter_host: null argument");
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int n = a->n;
    const size_t nk = (size_t)n * n * g->ncells, nf = (size_t)n * g->ncells;
    double *d_Ke = nullptr, *d_fe = nullptr;
    FB2_CUDA(cudaMalloc(&d_Ke, std::max<size_t>(nk, 1) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_Ke, Ke, nk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && fe && f_dev) {
        e = cudaMalloc(&d_fe, std::max<size_t>(nf, 1) * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_fe, fe, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    }
    int rc = FB2_OK;
    if (e != cudaSuccess) rc = fb2_fail(FB2_ERR_CUDA, "fb2_scatter_host: %s", cudaGetErrorString(e));
    else rc = fb2_scatter_device(a, d_Ke, d_fe, nzval_dev, f_dev, opts);
    cudaFree(d_Ke);
    cudaFree(d_fe);
    return rc;
}
