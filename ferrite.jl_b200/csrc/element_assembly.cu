// Element assembly: every element matrix is kept (Kes[:, :, cell], fes[:, cell]) instead of being summed into a CSC.
//
// Reference: the "element assembly" strategy of docs/src/literate-howto/gpu_assembly.jl:265-304 (operator y = sum_e P_e' Ke P_e x
// without a global matrix), and the element-local constraint path apply_local! / apply_assemble!
// (src/Dofs/ConstraintHandler.jl:1750-1822, src/assembler.jl:491-503), which needs Ke and fe of a cell before they are scattered.
//
// Design: no second family of element kernels.  The element matrices of a DofHandler ARE the matrix of its "broken" twin: give
// cell c the private dofs c*n .. c*n+n-1 and the pattern becomes block diagonal with dense n x n blocks, so the CSC value of
// (row c*n+i, column c*n+j) sits at c*n*n + j*n + i -- column-major Ke of cell c -- and f[c*n+i] is fe[i].  The fused assembly
// kernels (assemble_kernels.cuh) therefore write Kes/fes directly when they are launched on the broken DofHandler; their face
// merges never fire (no two cells share a dof) and the scatter map degenerates to the identity.
// New kernels here: k_ea_gather (state vector into the broken numbering), k_ea_mul (the matrix-free operator),
// k_ea_apply_local (apply_local! on every cell that touches a prescribed dof).
#include "common.h"

#include <algorithm>

// fb2_apply_assemble touches element matrices only where apply_local! changes them: the cells that own a prescribed dof.
// They form a sub-grid (own connectivity, the parent's coordinates) with its own element assembly; all other cells go
// through the fused kernel of the caller's assembler, restricted to the interior list.
struct EaSplit {
    const fb2_ch* ch = nullptr;    // cache key: handle, number of prescribed dofs, checksum of the dof list
    size_t np = 0;
    uint64_t sig = 0;
    uint64_t dh_generation = 0;    // numbering the split was built for (renumber! invalidates it)
    int64_t ninterior = 0, nboundary = 0;
    int32_t* d_interior = nullptr; // cells without a prescribed dof, ascending
    fb2_grid* sgrid = nullptr;     // boundary cells; d_xyz aliases the parent's coordinates (follows coordinate updates)
    fb2_dh* sdh = nullptr;         // their cell dofs in the global numbering
    fb2_ea* sea = nullptr;         // element matrices of the boundary cells
    fb2_assembler* sasm = nullptr; // scatter of the boundary cells into the caller's pattern
    const fb2_pattern* pat = nullptr;
    double* d_Kes = nullptr;
    double* d_fes = nullptr;
};

namespace {

inline unsigned nblocks(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ub[c*n + i] = u[dof(c, i)]
__global__ void k_ea_gather(const double* __restrict__ u, const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad,
                            int n, double* __restrict__ ub) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * n) return;
    const int64_t c = t % ncells;
    const int i = (int)(t / ncells);
    ub[c * n + i] = u[cell_dofs[(size_t)i * ncells_pad + c]];
}

// y[dof(c, i)] += sum_j Ke_c[i, j] x[dof(c, j)]: thread per (cell, row), rows fastest, so that a warp reads whole columns of
// consecutive cells (Ke_c[:, j] is contiguous) and every Ke entry is read exactly once: HBM-bound, 8 n^2 bytes per cell.
__global__ void __launch_bounds__(256) k_ea_mul(const double* __restrict__ Kes, const double* __restrict__ x,
                                                const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int n,
                                                double* __restrict__ y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * n) return;
    const int64_t c = t / n;
    const int i = (int)(t - c * n);
    const double* __restrict__ K = Kes + (size_t)c * n * n + i;
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    for (; j + 1 < n; j += 2) {
        const double x0 = __ldg(x + cell_dofs[(size_t)j * ncells_pad + c]);
        const double x1 = __ldg(x + cell_dofs[(size_t)(j + 1) * ncells_pad + c]);
        s0 = fma(__ldcs(K + (size_t)j * n), x0, s0);
        s1 = fma(__ldcs(K + (size_t)(j + 1) * n), x1, s1);
    }
    if (j < n) s0 = fma(__ldcs(K + (size_t)j * n), __ldg(x + cell_dofs[(size_t)j * ncells_pad + c]), s0);
    atomicAdd(y + cell_dofs[(size_t)i * ncells_pad + c], s0 + s1);
}

// diag[dof(c, i)] += Ke_c[i, i]: the diagonal of the operator (Jacobi preconditioner of the matrix-free CG)
__global__ void k_ea_diag(const double* __restrict__ Kes, const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int n,
                          double* __restrict__ diag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * n) return;
    const int64_t c = t / n;
    const int i = (int)(t - c * n);
    atomicAdd(diag + cell_dofs[(size_t)i * ncells_pad + c], Kes[(size_t)c * n * n + (size_t)i * n + i]);
}

// f[dof(c, i)] += fe_c[i]
__global__ void k_ea_rhs(const double* __restrict__ fes, const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int n,
                         double* __restrict__ f) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * n) return;
    const int64_t c = t / n;
    const int i = (int)(t - c * n);
    const double v = fes[t];
    if (v != 0.0) atomicAdd(f + cell_dofs[(size_t)i * ncells_pad + c], v);
}

__device__ __forceinline__ int64_t ea_find(const int32_t* __restrict__ prescribed, int64_t np, int32_t d) {
    int64_t lo = 0, hi = np;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (prescribed[mid] < d) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// _apply_local! without affine constraints (src/Dofs/ConstraintHandler.jl:1762-1822), thread per cell; cells that touch no
// prescribed dof leave after n byte loads, so the cost is that of the boundary layer.
//   1. fe -= v * Ke[:, l] for every prescribed local dof l with inhomogeneity v (skipped for apply_zero or v == 0)
//   2. m = meandiag(Ke) of the unmodified matrix
//   4. column l and row l zeroed, Ke[l, l] = m, fe[l] = v * m (0 for apply_zero)
__global__ void k_ea_apply_local(double* __restrict__ Kes, double* __restrict__ fes, const int32_t* __restrict__ cell_dofs,
                                 int64_t ncells, int64_t ncells_pad, int n, const uint8_t* __restrict__ isconstrained,
                                 const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np, int applyzero) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    bool any = false;
    for (int l = 0; l < n; ++l) any |= isconstrained[cell_dofs[(size_t)l * ncells_pad + c]] != 0;
    if (!any) return;
    double* K = Kes + (size_t)c * n * n;
    double* fe = fes ? fes + (size_t)c * n : nullptr;
    if (!applyzero && fe) {
        for (int l = 0; l < n; ++l) {
            const int32_t d = cell_dofs[(size_t)l * ncells_pad + c];
            if (!isconstrained[d]) continue;
            const double v = inhom[ea_find(prescribed, np, d)];
            if (v != 0.0)
                for (int j = 0; j < n; ++j) fe[j] -= v * K[(size_t)l * n + j];
        }
    }
    double m = 0.0;
    for (int i = 0; i < n; ++i) m += fabs(K[(size_t)i * n + i]);
    m /= n;
    for (int l = 0; l < n; ++l) {
        const int32_t d = cell_dofs[(size_t)l * ncells_pad + c];
        if (!isconstrained[d]) continue;
        for (int r = 0; r < n; ++r) K[(size_t)l * n + r] = 0.0;
        for (int j = 0; j < n; ++j) K[(size_t)j * n + l] = 0.0;
        K[(size_t)l * n + l] = m;
        if (fe) fe[l] = applyzero ? 0.0 : inhom[ea_find(prescribed, np, d)] * m;
    }
}

}  // namespace

static void split_free(EaSplit* sp) {
    if (!sp) return;
    if (sp->sasm) fb2_assembler_destroy(sp->sasm);
    if (sp->sea) fb2_ea_destroy(sp->sea);
    if (sp->sdh) fb2_dh_destroy(sp->sdh);
    if (sp->sgrid) {
        sp->sgrid->d_xyz = nullptr;   // borrowed from the parent grid
        fb2_grid_destroy(sp->sgrid);
    }
    cudaFree(sp->d_interior);
    cudaFree(sp->d_Kes);
    cudaFree(sp->d_fes);
    delete sp;
}

static uint64_t ch_signature(const fb2_ch* ch) {
    uint64_t h = 1469598103934665603ull;
    for (int64_t d : ch->prescribed) h = (h ^ (uint64_t)d) * 1099511628211ull;
    return h;
}

static int split_build(fb2_ea* ea, fb2_assembler* a, fb2_ch* ch) {
    EaSplit* sp = ea->split;
    const uint64_t sig = ch_signature(ch);
    if (sp && sp->ch == ch && sp->np == ch->prescribed.size() && sp->sig == sig && sp->pat == a->pat && sp->dh_generation == a->dh->generation) return FB2_OK;
    split_free(sp);
    ea->split = nullptr;
    sp = new EaSplit();
    sp->ch = ch;
    sp->np = ch->prescribed.size();
    sp->sig = sig;
    sp->dh_generation = a->dh->generation;
    sp->pat = a->pat;
    fb2_dh* dh = ea->dh;
    fb2_grid* g = dh->grid;
    const int n = ea->n, nnpc = g->nnpc;
    std::vector<uint8_t> isc((size_t)dh->ndofs, 0);
    for (int64_t d : ch->prescribed) isc[(size_t)d] = 1;
    std::vector<int32_t> interior, boundary;
    for (int64_t c = 0; c < g->ncells; ++c) {
        bool any = false;
        for (int i = 0; i < n; ++i) any |= isc[(size_t)dh->cell_dofs[(size_t)c * n + i]] != 0;
        (any ? boundary : interior).push_back((int32_t)c);
    }
    sp->ninterior = (int64_t)interior.size();
    sp->nboundary = (int64_t)boundary.size();
    int rc = FB2_OK;
    cudaError_t e = cudaSuccess;
    if (sp->ninterior) {
        e = cudaMalloc(&sp->d_interior, interior.size() * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpy(sp->d_interior, interior.data(), interior.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && sp->nboundary) {
        const int64_t nb = sp->nboundary;
        fb2_grid* s = new fb2_grid();
        sp->sgrid = s;
        s->ctx = g->ctx;
        s->celltype = g->celltype;
        s->ncells = nb;
        s->nnodes = g->nnodes;
        s->nnpc = nnpc;
        s->sdim = g->sdim;
        s->xstride = g->xstride;
        s->ncells_pad = (nb + 31) / 32 * 32;
        s->cells.resize((size_t)nb * nnpc);
        std::vector<int32_t> conn((size_t)nnpc * s->ncells_pad);
        std::vector<int64_t> cd((size_t)nb * n);
        for (int64_t k = 0; k < s->ncells_pad; ++k) {
            const int64_t c = boundary[(size_t)std::min<int64_t>(k, nb - 1)];   // padding replicates the last cell
            for (int j = 0; j < nnpc; ++j) {
                const int64_t node = g->cells[(size_t)c * nnpc + j];
                conn[(size_t)j * s->ncells_pad + k] = (int32_t)(node - 1);
                if (k < nb) s->cells[(size_t)k * nnpc + j] = node;
            }
            if (k < nb)
                for (int i = 0; i < n; ++i) cd[(size_t)k * n + i] = (int64_t)dh->cell_dofs[(size_t)c * n + i] + 1;
        }
        e = cudaMalloc(&s->d_conn, conn.size() * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpy(s->d_conn, conn.data(), conn.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
        s->d_xyz = g->d_xyz;
        if (e == cudaSuccess) {
            rc = fb2_dh_from_host(s, (int)dh->fields.size(), dh->fields.data(), dh->ndofs, n, cd.data(), &sp->sdh);
            if (rc == FB2_OK) rc = fb2_ea_create(sp->sdh, ea->cv, &sp->sea);
            if (rc == FB2_OK) rc = fb2_assembler_create(sp->sdh, a->pat, nullptr, &sp->sasm);
            if (rc == FB2_OK) e = cudaMalloc(&sp->d_Kes, (size_t)nb * n * n * sizeof(double));
            if (rc == FB2_OK && e == cudaSuccess) e = cudaMalloc(&sp->d_fes, (size_t)nb * n * sizeof(double));
        }
    }
    if (rc == FB2_OK && e != cudaSuccess)
        rc = fb2_fail(e == cudaErrorMemoryAllocation ? FB2_ERR_OOM : FB2_ERR_CUDA, "fb2_apply_assemble: %s", cudaGetErrorString(e));
    if (rc != FB2_OK) { split_free(sp); return rc; }
    ea->split = sp;
    return FB2_OK;
}

extern "C" int fb2_ea_destroy(fb2_ea* ea) {
    if (!ea) return FB2_OK;
    if (ea->basm) fb2_assembler_destroy(ea->basm);
    if (ea->bpat) fb2_pattern_destroy(ea->bpat);
    if (ea->bdh) fb2_dh_destroy(ea->bdh);
    cudaFree(ea->d_ub);
    cudaFree(ea->d_work);
    split_free(ea->split);
    delete ea;
    return FB2_OK;
}

// the broken twin is built by the first fb2_ea_assemble: fb2_apply_assemble only needs the twin of its boundary layer
static int ea_ensure_twin(fb2_ea* ea) {
    if (ea->basm) return FB2_OK;
    fb2_grid* g = ea->dh->grid;
    const int n = ea->n;
    const int64_t nc = g->ncells;
    std::vector<int64_t> cd((size_t)nc * n);
    for (size_t i = 0; i < cd.size(); ++i) cd[i] = (int64_t)i + 1;
    if (!ea->bdh) FB2_TRY(fb2_dh_from_host(g, (int)ea->dh->fields.size(), ea->dh->fields.data(), nc * n, n, cd.data(), &ea->bdh));
    if (!ea->bpat) FB2_TRY(fb2_pattern_create(ea->bdh, &ea->bpat));
    FB2_CHECK(ea->bpat->nnz == nc * n * n, FB2_ERR_INTERNAL, "element assembly: block-diagonal pattern has %lld entries, expected %lld",
              (long long)ea->bpat->nnz, (long long)(nc * n * n));
    return fb2_assembler_create(ea->bdh, ea->bpat, ea->cv, &ea->basm);
}

extern "C" int fb2_ea_create(fb2_dh* dh, fb2_cv* cv, fb2_ea** out) {
    FB2_CHECK(dh && cv && out, FB2_ERR_BAD_ARG, "fb2_ea_create: null argument");
    fb2_grid* g = dh->grid;
    FB2_NEED_DEVICE(g->ctx);
    const int n = dh->ndpc;
    FB2_CHECK(g->ncells > 0, FB2_ERR_BAD_ARG, "fb2_ea_create: the grid has no cells");
    FB2_CHECK(g->ncells * n < (int64_t)2147483647, FB2_ERR_UNSUPPORTED, "fb2_ea_create: ncells * ndofs_per_cell must stay below 2^31-1");
    FB2_CHECK(cv->celltype == g->celltype, FB2_ERR_BAD_ARG, "fb2_ea_create: CellValues are for another cell type");
    FB2_CHECK(cv->nb * cv->vdim == n, FB2_ERR_UNSUPPORTED, "fb2_ea_create: the element must cover all %d dofs of a cell (CellValues has %d)", n,
              cv->nb * cv->vdim);
    fb2_ea* ea = new fb2_ea();
    ea->dh = dh;
    ea->cv = cv;
    ea->n = n;
    *out = ea;
    return FB2_OK;
}

extern "C" int fb2_ea_info(fb2_ea* ea, int64_t* ncells, int* n) {
    FB2_CHECK(ea, FB2_ERR_BAD_ARG, "fb2_ea_info: null handle");
    if (ncells) *ncells = ea->dh->grid->ncells;
    if (n) *n = ea->n;
    return FB2_OK;
}

extern "C" int fb2_ea_assemble(fb2_ea* ea, int element, const void* params, size_t params_bytes, const double* u_dev, double* Kes_dev,
                               double* fes_dev) {
    FB2_CHECK(ea && Kes_dev, FB2_ERR_BAD_ARG, "fb2_ea_assemble: null argument");
    FB2_TRY(ea_ensure_twin(ea));
    fb2_grid* g = ea->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const double* ub = nullptr;
    if (u_dev) {
        const int64_t tot = g->ncells * ea->n;
        if (!ea->d_ub) FB2_CUDA(cudaMalloc(&ea->d_ub, (size_t)tot * sizeof(double)));
        k_ea_gather<<<nblocks(tot, 256), 256, 0, ctx->stream>>>(u_dev, ea->dh->d_cell_dofs, g->ncells, g->ncells_pad, ea->n, ea->d_ub);
        ctx->launches++;
        FB2_CUDA(cudaGetLastError());
        ub = ea->d_ub;
    }
    const fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    return fb2_launch_assemble(ea->basm, element, params, params_bytes, ub, Kes_dev, fes_dev, &o);
}

extern "C" int fb2_ea_mul(fb2_ea* ea, const double* Kes_dev, const double* x_dev, double* y_dev) {
    FB2_CHECK(ea && Kes_dev && x_dev && y_dev, FB2_ERR_BAD_ARG, "fb2_ea_mul: null argument");
    FB2_CHECK(x_dev != y_dev, FB2_ERR_BAD_ARG, "fb2_ea_mul: x and y must not alias");
    fb2_grid* g = ea->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaMemsetAsync(y_dev, 0, (size_t)ea->dh->ndofs * sizeof(double), ctx->stream));
    const int64_t tot = g->ncells * ea->n;
    k_ea_mul<<<nblocks(tot, 256), 256, 0, ctx->stream>>>(Kes_dev, x_dev, ea->dh->d_cell_dofs, g->ncells, g->ncells_pad, ea->n, y_dev);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_ea_diag(fb2_ea* ea, const double* Kes_dev, double* diag_dev) {
    FB2_CHECK(ea && Kes_dev && diag_dev, FB2_ERR_BAD_ARG, "fb2_ea_diag: null argument");
    fb2_grid* g = ea->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaMemsetAsync(diag_dev, 0, (size_t)ea->dh->ndofs * sizeof(double), ctx->stream));
    const int64_t tot = g->ncells * ea->n;
    k_ea_diag<<<nblocks(tot, 256), 256, 0, ctx->stream>>>(Kes_dev, ea->dh->d_cell_dofs, g->ncells, g->ncells_pad, ea->n, diag_dev);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_ea_rhs(fb2_ea* ea, const double* fes_dev, double* f_dev) {
    FB2_CHECK(ea && fes_dev && f_dev, FB2_ERR_BAD_ARG, "fb2_ea_rhs: null argument");
    fb2_grid* g = ea->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)ea->dh->ndofs * sizeof(double), ctx->stream));
    const int64_t tot = g->ncells * ea->n;
    k_ea_rhs<<<nblocks(tot, 256), 256, 0, ctx->stream>>>(fes_dev, ea->dh->d_cell_dofs, g->ncells, g->ncells_pad, ea->n, f_dev);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_ea_apply_local(fb2_ea* ea, fb2_ch* ch, double* Kes_dev, double* fes_dev, int applyzero) {
    FB2_CHECK(ea && ch && Kes_dev, FB2_ERR_BAD_ARG, "fb2_ea_apply_local: null argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_ea_apply_local: the ConstraintHandler is not closed");
    FB2_CHECK(!ch->has_affine, FB2_ERR_UNSUPPORTED, "fb2_ea_apply_local: affine constraints need the global apply! (`_condense_local!` is not implemented)");
    FB2_CHECK(ch->dh == ea->dh || ch->dh->ndofs == ea->dh->ndofs, FB2_ERR_BAD_ARG,
              "fb2_ea_apply_local: the ConstraintHandler belongs to another DofHandler");
    fb2_grid* g = ea->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(fb2_ch_sync_device(ch));
    const int64_t np = (int64_t)ch->prescribed.size();
    if (np == 0) return FB2_OK;
    k_ea_apply_local<<<nblocks(g->ncells, 128), 128, 0, ctx->stream>>>(Kes_dev, fes_dev, ea->dh->d_cell_dofs, g->ncells, g->ncells_pad,
                                                                       ea->n, ch->d_isconstrained, ch->d_prescribed, ch->d_inhom, np,
                                                                       applyzero);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

// apply_assemble!(assembler, ch, celldofs(cell), Ke, fe; apply_zero) for every cell (src/assembler.jl:491-503).  apply_local!
// changes Ke / fe only in cells that own a prescribed dof, so the interior cells run through the fused kernel of `a`
// (restricted to the interior list) and only the boundary layer takes the three-step route element matrices ->
// apply_local! -> assemble!.  `a` is the assembler of the caller's matrix (fb2_assembler_create(dh, pattern, cv)).
extern "C" int fb2_apply_assemble(fb2_assembler* a, fb2_ea* ea, fb2_ch* ch, int element, const void* params, size_t params_bytes,
                                  const double* u_dev, double* nzval_dev, double* f_dev, int applyzero, const fb2_asm_opts* opts) {
    FB2_CHECK(a && ea && ch && nzval_dev && f_dev, FB2_ERR_BAD_ARG, "fb2_apply_assemble: null argument");
    FB2_CHECK(a->dh == ea->dh, FB2_ERR_BAD_ARG, "fb2_apply_assemble: assembler and element assembly belong to different DofHandlers");
    FB2_CHECK(a->cv, FB2_ERR_BAD_ARG, "fb2_apply_assemble: this assembler was created without CellValues (scatter-only)");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_apply_assemble: the ConstraintHandler is not closed");
    FB2_CHECK(ch->dh->ndofs == ea->dh->ndofs, FB2_ERR_BAD_ARG, "fb2_apply_assemble: the ConstraintHandler belongs to another DofHandler");
    FB2_CHECK(a->d_cells == nullptr && a->ncells_active == 0, FB2_ERR_UNSUPPORTED, "fb2_apply_assemble on a partitioned assembler is not supported");
    fb2_grid* g = ea->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    if (opts) o = *opts;
    FB2_TRY(split_build(ea, a, ch));
    EaSplit* sp = ea->split;
    if (o.fillzero) {
        FB2_CUDA(cudaMemsetAsync(nzval_dev, 0, (size_t)a->pat->nnz * sizeof(double), ctx->stream));
        FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)a->dh->ndofs * sizeof(double), ctx->stream));
    }
    const fb2_asm_opts acc = {0, FB2_SCATTER_ATOMIC, o.variant, 0};
    if (sp->ninterior) {
        a->d_cells = sp->d_interior;   // borrowed for this launch
        a->ncells_active = sp->ninterior;
        const int rc = fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, &acc);
        a->d_cells = nullptr;
        a->ncells_active = 0;
        FB2_TRY(rc);
    }
    if (sp->nboundary) {
        FB2_TRY(fb2_ea_assemble(sp->sea, element, params, params_bytes, u_dev, sp->d_Kes, sp->d_fes));
        FB2_TRY(fb2_ea_apply_local(sp->sea, ch, sp->d_Kes, sp->d_fes, applyzero));
        FB2_TRY(fb2_scatter_device(sp->sasm, sp->d_Kes, sp->d_fes, nzval_dev, f_dev, &acc));
    }
    return fb2_check_device_error(ctx);
}
