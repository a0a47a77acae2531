// Context, CellValues and assembler entry points of the C ABI (see include/ferrite_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "common.h"

static thread_local char g_err[1024] = "";

void fb2_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fb2_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

uint64_t fb2_next_uid() {
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}

extern "C" const char* fb2_version(void) { return "ferrite_b200 0.1.0 (sm_100a)"; }
extern "C" const char* fb2_last_error(void) { return g_err; }

// ---- context -------------------------------------------------------------------------------------
extern "C" int fb2_ctx_create(int device, fb2_ctx** out) {
    FB2_CHECK(out, FB2_ERR_BAD_ARG, "fb2_ctx_create: null output pointer");
    if (device == -1) {  // host-only context: grid / dof / constraint set-up logic without a device
        fb2_ctx* ctx = new fb2_ctx();
        ctx->device = -1;
        ctx->own_stream = false;
        *out = ctx;
        return FB2_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fb2_fail(FB2_ERR_CUDA, "no usable CUDA device (%s); libferrite_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    FB2_CHECK(device >= 0 && device < ndev, FB2_ERR_BAD_ARG, "fb2_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    FB2_CUDA(cudaSetDevice(device));
    fb2_ctx* ctx = new fb2_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    FB2_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    FB2_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    FB2_CUDA(cudaMalloc(&ctx->d_errflag, 2 * sizeof(int)));
    FB2_CUDA(cudaMemset(ctx->d_errflag, 0, 2 * sizeof(int)));
    FB2_CUDA(cudaMallocHost(&ctx->h_errflag, 2 * sizeof(int)));
    *out = ctx;
    return FB2_OK;
}

extern "C" int fb2_comm_destroy(fb2_ctx* ctx);

extern "C" int fb2_ctx_destroy(fb2_ctx* ctx) {
    if (!ctx) return FB2_OK;
    if (ctx->device < 0) { delete ctx; return FB2_OK; }
    cudaSetDevice(ctx->device);
    fb2_comm_destroy(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->ev_ready) for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    cudaFree(ctx->d_errflag);
    cudaFreeHost(ctx->h_errflag);
    delete ctx;
    return FB2_OK;
}

extern "C" int fb2_ctx_synchronize(fb2_ctx* ctx) {
    FB2_CHECK(ctx, FB2_ERR_BAD_ARG, "fb2_ctx_synchronize: null context");
    if (ctx->device < 0) return FB2_OK;
    FB2_CUDA(cudaSetDevice(ctx->device));
    return fb2_check_device_error(ctx);  // synchronises the stream and surfaces kernel-side errors
}

extern "C" int fb2_ctx_set_stream(fb2_ctx* ctx, void* cuda_stream) {
    FB2_CHECK(ctx, FB2_ERR_BAD_ARG, "fb2_ctx_set_stream: null context");
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return FB2_OK;
}

extern "C" int fb2_ctx_launch_count(fb2_ctx* ctx, int64_t* out) {
    FB2_CHECK(ctx && out, FB2_ERR_BAD_ARG, "fb2_ctx_launch_count: null argument");
    *out = ctx->launches;
    return FB2_OK;
}

extern "C" int fb2_device_alloc(fb2_ctx* ctx, size_t nbytes, void** dev_ptr) {
    FB2_CHECK(ctx && dev_ptr, FB2_ERR_BAD_ARG, "fb2_device_alloc: null argument");
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaMalloc(dev_ptr, nbytes ? nbytes : 8));
    return FB2_OK;
}

extern "C" int fb2_device_free(fb2_ctx* ctx, void* dev_ptr) {
    FB2_CHECK(ctx, FB2_ERR_BAD_ARG, "fb2_device_free: null context");
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    FB2_CUDA(cudaFree(dev_ptr));
    return FB2_OK;
}

extern "C" int fb2_memcpy_h2d(fb2_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes) {
    FB2_CHECK(ctx && dst_dev && src_host, FB2_ERR_BAD_ARG, "fb2_memcpy_h2d: null argument");
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    return FB2_OK;
}

extern "C" int fb2_memcpy_d2h(fb2_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes) {
    FB2_CHECK(ctx && dst_host && src_dev, FB2_ERR_BAD_ARG, "fb2_memcpy_d2h: null argument");
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return fb2_check_device_error(ctx);
}

// ---- CellValues -------------------------------------------------------------------------------------
static int celltype_rdim(int ct) {
    const RefShapeInfo* rs = fb2_refshape(ct);
    return rs ? rs->rdim : 0;
}

extern "C" int fb2_cellvalues_create(fb2_ctx* ctx, int celltype, int qr_order, int ip_order, int vdim, int geo_order, fb2_cv** out) {
    FB2_CHECK(ctx && out, FB2_ERR_BAD_ARG, "fb2_cellvalues_create: null argument");
    LagrangeInfo ip, geo;
    FB2_CHECK(fb2_lagrange(celltype, ip_order, &ip), FB2_ERR_UNSUPPORTED, "CellValues: Lagrange order %d on cell type %d not supported", ip_order, celltype);
    FB2_CHECK(fb2_lagrange(celltype, geo_order, &geo), FB2_ERR_UNSUPPORTED, "CellValues: geometric order %d not supported", geo_order);
    FB2_CHECK(vdim >= 1 && vdim <= 3, FB2_ERR_BAD_ARG, "CellValues: vdim must be 1..3");
    fb2_cv* cv = new fb2_cv();
    cv->uid = fb2_next_uid();
    cv->ctx = ctx;
    cv->celltype = celltype;
    cv->rdim = ip.rdim;
    cv->nb = ip.nbase;
    cv->vdim = vdim;
    cv->ngeo = geo.nbase;
    cv->ip_order = ip_order;
    cv->geo_order = geo_order;
    cv->qr_order = qr_order;
    if (!fb2_quadrature(celltype, qr_order, &cv->w, &cv->pts)) {
        delete cv;
        return fb2_fail(FB2_ERR_UNSUPPORTED, "QuadratureRule of order %d on cell type %d not supported", qr_order, celltype);
    }
    cv->nq = (int)cv->w.size();
    const int rd = cv->rdim;
    cv->N.resize((size_t)cv->nq * cv->nb);
    cv->dN.resize((size_t)cv->nq * cv->nb * rd);
    cv->M.resize((size_t)cv->nq * cv->ngeo);
    cv->dM.resize((size_t)cv->nq * cv->ngeo * rd);
    for (int q = 0; q < cv->nq; ++q) {
        fb2_lagrange_eval(ip, &cv->pts[(size_t)q * rd], &cv->N[(size_t)q * cv->nb], &cv->dN[(size_t)q * cv->nb * rd]);
        fb2_lagrange_eval(geo, &cv->pts[(size_t)q * rd], &cv->M[(size_t)q * cv->ngeo], &cv->dM[(size_t)q * cv->ngeo * rd]);
    }
    *out = cv;
    return FB2_OK;
}

extern "C" int fb2_cellvalues_from_tables(fb2_ctx* ctx, int celltype, int nq, int n, int vdim, int ngeo, const double* N,
                                          const double* dNdxi, const double* M, const double* dMdxi, const double* w, fb2_cv** out) {
    FB2_CHECK(ctx && N && dNdxi && M && dMdxi && w && out, FB2_ERR_BAD_ARG, "fb2_cellvalues_from_tables: null argument");
    const int rd = celltype_rdim(celltype);
    FB2_CHECK(rd > 0, FB2_ERR_BAD_ARG, "fb2_cellvalues_from_tables: unknown cell type %d", celltype);
    FB2_CHECK(nq >= 1 && n >= 1 && ngeo >= 1 && vdim >= 1 && vdim <= 3, FB2_ERR_BAD_ARG, "fb2_cellvalues_from_tables: bad sizes");
    fb2_cv* cv = new fb2_cv();
    cv->uid = fb2_next_uid();
    cv->ctx = ctx;
    cv->celltype = celltype;
    cv->rdim = rd;
    cv->nq = nq;
    cv->nb = n;
    cv->vdim = vdim;
    cv->ngeo = ngeo;
    cv->w.assign(w, w + nq);
    cv->pts.assign((size_t)nq * rd, 0.0);
    cv->N.resize((size_t)nq * n);
    cv->dN.resize((size_t)nq * n * rd);
    cv->M.resize((size_t)nq * ngeo);
    cv->dM.resize((size_t)nq * ngeo * rd);
    // inputs are column-major Julia arrays: N[i, q], dNdxi[d, i, q] -> exactly the q-major layout used here
    memcpy(cv->N.data(), N, cv->N.size() * sizeof(double));
    memcpy(cv->dN.data(), dNdxi, cv->dN.size() * sizeof(double));
    memcpy(cv->M.data(), M, cv->M.size() * sizeof(double));
    memcpy(cv->dM.data(), dMdxi, cv->dM.size() * sizeof(double));
    *out = cv;
    return FB2_OK;
}

extern "C" int fb2_cellvalues_info(fb2_cv* cv, int* nq, int* nbase_scalar, int* vdim, int* ngeo, int* rdim) {
    FB2_CHECK(cv, FB2_ERR_BAD_ARG, "fb2_cellvalues_info: null handle");
    if (nq) *nq = cv->nq;
    if (nbase_scalar) *nbase_scalar = cv->nb;
    if (vdim) *vdim = cv->vdim;
    if (ngeo) *ngeo = cv->ngeo;
    if (rdim) *rdim = cv->rdim;
    return FB2_OK;
}

extern "C" int fb2_cellvalues_export(fb2_cv* cv, double* N, double* dNdxi, double* M, double* dMdxi, double* w, double* points) {
    FB2_CHECK(cv, FB2_ERR_BAD_ARG, "fb2_cellvalues_export: null handle");
    if (N) memcpy(N, cv->N.data(), cv->N.size() * sizeof(double));
    if (dNdxi) memcpy(dNdxi, cv->dN.data(), cv->dN.size() * sizeof(double));
    if (M) memcpy(M, cv->M.data(), cv->M.size() * sizeof(double));
    if (dMdxi) memcpy(dMdxi, cv->dM.data(), cv->dM.size() * sizeof(double));
    if (w) memcpy(w, cv->w.data(), cv->w.size() * sizeof(double));
    if (points) memcpy(points, cv->pts.data(), cv->pts.size() * sizeof(double));
    return FB2_OK;
}

extern "C" int fb2_cellvalues_destroy(fb2_cv* cv) {
    if (!cv) return FB2_OK;
    if (cv->d_tables) cudaFree(cv->d_tables);
    delete cv;
    return FB2_OK;
}

// ---- assembler -----------------------------------------------------------------------------------------
extern "C" int fb2_assembler_create(fb2_dh* dh, fb2_pattern* p, fb2_cv* cv, fb2_assembler** out) {
    FB2_CHECK(dh && p && out, FB2_ERR_BAD_ARG, "fb2_assembler_create: null argument");
    FB2_CHECK(p->dh == dh || p->n == dh->ndofs, FB2_ERR_BAD_ARG, "fb2_assembler_create: pattern does not belong to this DofHandler");
    FB2_NEED_DEVICE(dh->grid->ctx);
    if (cv) {
        FB2_CHECK(cv->celltype == dh->grid->celltype, FB2_ERR_BAD_ARG, "fb2_assembler_create: CellValues are for another cell type");
        FB2_CHECK(cv->nb * cv->vdim == dh->ndpc, FB2_ERR_UNSUPPORTED,
                  "fb2_assembler_create: the element must cover all %d dofs of a cell (CellValues has %d)", dh->ndpc, cv->nb * cv->vdim);
        FB2_CHECK(cv->ngeo == dh->grid->nnpc, FB2_ERR_BAD_ARG, "fb2_assembler_create: geometric interpolation has %d nodes, cells have %d", cv->ngeo, dh->grid->nnpc);
        FB2_CHECK(cv->rdim == dh->grid->sdim, FB2_ERR_UNSUPPORTED, "embedded elements (rdim %d in sdim %d) are not supported", cv->rdim, dh->grid->sdim);
        // kernels, zero fills and table uploads all run on the GRID's context (stream), whatever context the CellValues were
        // created on; their device copy belongs to the first device that uses them
        FB2_CHECK(cv->tables_device < 0 || cv->tables_device == dh->grid->ctx->device, FB2_ERR_BAD_ARG,
                  "fb2_assembler_create: these CellValues are already in use on device %d, the grid lives on device %d (create one CellValues per device)",
                  cv->tables_device, dh->grid->ctx->device);
    }
    fb2_assembler* a = new fb2_assembler();
    a->dh = dh;
    a->pat = p;
    a->cv = cv;
    a->n = dh->ndpc;
    int rc = fb2_map_build(a);
    if (rc != FB2_OK) { fb2_assembler_destroy(a); return rc; }
    *out = a;
    return FB2_OK;
}

extern "C" int fb2_assemble(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* u_dev,
                            double* nzval_dev, double* f_dev, const fb2_asm_opts* opts) {
    FB2_CHECK(a && nzval_dev, FB2_ERR_BAD_ARG, "fb2_assemble: null argument");
    FB2_CHECK(a->cv, FB2_ERR_BAD_ARG, "fb2_assemble: this assembler was created without CellValues (scatter-only)");
    return fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, opts);
}

extern "C" int fb2_assemble_host(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* u_host,
                                 double* nzval_host, double* f_host, const fb2_asm_opts* opts) {
    FB2_CHECK(a && nzval_host, FB2_ERR_BAD_ARG, "fb2_assemble_host: null argument");
    FB2_CHECK(a->cv, FB2_ERR_BAD_ARG, "fb2_assemble_host: this assembler was created without CellValues (scatter-only)");
    fb2_ctx* ctx = a->dh->grid->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const size_t nnz = (size_t)a->pat->nnz, n = (size_t)a->dh->ndofs;
    if (!a->d_nzval) FB2_CUDA(cudaMalloc(&a->d_nzval, nnz * sizeof(double)));
    if (!a->d_f) FB2_CUDA(cudaMalloc(&a->d_f, n * sizeof(double)));
    if (u_host) {
        if (!a->d_u) FB2_CUDA(cudaMalloc(&a->d_u, n * sizeof(double)));
        FB2_CUDA(cudaMemcpyAsync(a->d_u, u_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    if (opts) o = *opts;
    if (!o.fillzero) {  // accumulate onto the caller's current values
        FB2_CUDA(cudaMemcpyAsync(a->d_nzval, nzval_host, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (f_host) FB2_CUDA(cudaMemcpyAsync(a->d_f, f_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    FB2_TRY(fb2_launch_assemble(a, element, params, params_bytes, u_host ? a->d_u : nullptr, a->d_nzval, f_host ? a->d_f : nullptr, &o));
    FB2_CUDA(cudaMemcpyAsync(nzval_host, a->d_nzval, nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (f_host) FB2_CUDA(cudaMemcpyAsync(f_host, a->d_f, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return fb2_check_device_error(ctx);
}

// Streamed variant of fb2_assemble_host: the cells are cut into slabs (contiguous ranges in cell order); the
// coordinates a slab needs are uploaded while the previous slab is assembled, and the matrix columns no later slab
// touches are downloaded while the following slabs are assembled.  With the reference's cell / dof numbering (both
// follow the grid sweep) every slab completes a contiguous block of columns; for an arbitrary numbering the schedule
// degenerates to upload-all / assemble / download-all, which is still correct.
__global__ void k_repack_xyz_range(const double* __restrict__ src, int64_t node0, int64_t node1, int sdim, int xstride, double* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + node0 * xstride;
    if (t >= node1 * xstride) return;
    const int64_t node = t / xstride;
    const int d = (int)(t - node * xstride);
    dst[t] = d < sdim ? src[node * sdim + d] : 0.0;
}

static int build_slab_schedule(fb2_assembler* a, int nslabs) {
    if (!a->slab_cell.empty()) return FB2_OK;
    fb2_dh* dh = a->dh;
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    const int64_t nc = g->ncells;
    const int nnpc = g->nnpc, ndpc = dh->ndpc;
    std::vector<int64_t> cell(nslabs + 1), node(nslabs + 1, 0), col(nslabs + 1, 0);
    for (int k = 0; k <= nslabs; ++k) cell[k] = std::min<int64_t>(nc, ((nc * k / nslabs) + 127) / 128 * 128);
    cell[nslabs] = nc;
    // nodes needed by slabs <= k: 1 + the largest node id touched so far
    int64_t mx = 0;
    for (int k = 0; k < nslabs; ++k) {
        for (int64_t c = cell[k]; c < cell[k + 1]; ++c)
            for (int j = 0; j < nnpc; ++j) mx = std::max<int64_t>(mx, g->cells[(size_t)c * nnpc + j]);   // 1-based = count
        node[k + 1] = mx;
    }
    node[nslabs] = g->nnodes;
    // columns complete after slab k: every column below the smallest dof touched by a later slab
    int64_t mn = dh->ndofs;
    col[nslabs] = dh->ndofs;
    for (int k = nslabs - 1; k >= 1; --k) {
        for (int64_t c = cell[k]; c < cell[k + 1]; ++c)
            for (int i = 0; i < ndpc; ++i) mn = std::min<int64_t>(mn, dh->cell_dofs[(size_t)c * ndpc + i]);
        col[k] = mn;
    }
    col[0] = 0;
    std::vector<int64_t> pos(nslabs + 1);
    for (int k = 0; k <= nslabs; ++k)
        FB2_CUDA(cudaMemcpyAsync(&pos[k], a->pat->d_colptr + col[k], sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    a->slab_cell = cell; a->slab_node = node; a->slab_col = col; a->slab_pos = pos;
    return FB2_OK;
}

extern "C" int fb2_assemble_host_streamed(fb2_assembler* a, int element, const void* params, size_t params_bytes, const double* xyz_host,
                                          const double* u_host, double* nzval_host, double* f_host, const fb2_asm_opts* opts) {
    FB2_CHECK(a && nzval_host, FB2_ERR_BAD_ARG, "fb2_assemble_host_streamed: null argument");
    FB2_CHECK(a->cv, FB2_ERR_BAD_ARG, "fb2_assemble_host_streamed: this assembler was created without CellValues (scatter-only)");
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    if (opts) o = *opts;
    FB2_CHECK(o.fillzero && o.scatter_mode == FB2_SCATTER_ATOMIC && a->d_cells == nullptr && a->ncells_active == 0, FB2_ERR_UNSUPPORTED,
              "fb2_assemble_host_streamed: needs fillzero, the atomic scatter and an unpartitioned assembler");
    FB2_CUDA(cudaSetDevice(ctx->device));
    int NS = 8;
    if (const char* e = getenv("FB2_HOST_SLABS")) NS = std::max(1, std::min(20, atoi(e)));   // tuning override
    if ((int)a->slab_cell.size() != NS + 1) a->slab_cell.clear();
    const size_t nnz = (size_t)a->pat->nnz, n = (size_t)a->dh->ndofs;
    if (!a->d_nzval) FB2_CUDA(cudaMalloc(&a->d_nzval, nnz * sizeof(double)));
    if (!a->d_f) FB2_CUDA(cudaMalloc(&a->d_f, n * sizeof(double)));
    if (!ctx->h2d_stream) FB2_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) FB2_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    if (!ctx->ev_ready) {
        for (cudaEvent_t& e : ctx->ev_pool) FB2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_ready = true;
    }
    FB2_TRY(build_slab_schedule(a, NS));
    cudaStream_t sm = ctx->stream, sh = ctx->h2d_stream, sd = ctx->d2h_stream;
    cudaEvent_t* ev = ctx->ev_pool;   // [0] start, [1..NS] uploads, [NS+1..2NS] kernels, [2NS+1] downloads done
    FB2_CUDA(cudaEventRecord(ev[0], sm));
    FB2_CUDA(cudaStreamWaitEvent(sh, ev[0], 0));
    FB2_CUDA(cudaStreamWaitEvent(sd, ev[0], 0));
    if (u_host) {
        if (!a->d_u) FB2_CUDA(cudaMalloc(&a->d_u, n * sizeof(double)));
        FB2_CUDA(cudaMemcpyAsync(a->d_u, u_host, n * sizeof(double), cudaMemcpyHostToDevice, sm));
    }
    FB2_CUDA(cudaMemsetAsync(a->d_nzval, 0, nnz * sizeof(double), sm));
    if (f_host) FB2_CUDA(cudaMemsetAsync(a->d_f, 0, n * sizeof(double), sm));
    if (xyz_host) {
        const size_t tot = (size_t)g->nnodes * g->sdim;
        if (!g->d_xyz_stage) FB2_CUDA(cudaMalloc(&g->d_xyz_stage, tot * sizeof(double)));
        for (int k = 0; k < NS; ++k) {
            const int64_t n0 = a->slab_node[k], n1 = a->slab_node[k + 1];
            if (n1 > n0) {
                FB2_CUDA(cudaMemcpyAsync(g->d_xyz_stage + n0 * g->sdim, xyz_host + n0 * g->sdim, (size_t)(n1 - n0) * g->sdim * sizeof(double),
                                         cudaMemcpyHostToDevice, sh));
                const int64_t cnt = (n1 - n0) * g->xstride;
                k_repack_xyz_range<<<(unsigned)((cnt + 255) / 256), 256, 0, sh>>>(g->d_xyz_stage, n0, n1, g->sdim, g->xstride, g->d_xyz);
                ctx->launches++;
            }
            FB2_CUDA(cudaEventRecord(ev[1 + k], sh));
        }
    }
    o.fillzero = 0;
    int rc = FB2_OK;
    for (int k = 0; k < NS && rc == FB2_OK; ++k) {
        if (xyz_host) FB2_CUDA(cudaStreamWaitEvent(sm, ev[1 + k], 0));
        a->cell_first = a->slab_cell[k];
        a->ncells_active = a->slab_cell[k + 1] - a->slab_cell[k];
        if (a->ncells_active > 0)
            rc = fb2_launch_assemble(a, element, params, params_bytes, u_host ? a->d_u : nullptr, a->d_nzval, f_host ? a->d_f : nullptr, &o);
        if (rc != FB2_OK) break;
        FB2_CUDA(cudaEventRecord(ev[1 + NS + k], sm));
        FB2_CUDA(cudaStreamWaitEvent(sd, ev[1 + NS + k], 0));
        const int64_t p0 = a->slab_pos[k], p1 = a->slab_pos[k + 1], c0 = a->slab_col[k], c1 = a->slab_col[k + 1];
        if (p1 > p0) FB2_CUDA(cudaMemcpyAsync(nzval_host + p0, a->d_nzval + p0, (size_t)(p1 - p0) * sizeof(double), cudaMemcpyDeviceToHost, sd));
        if (f_host && c1 > c0) FB2_CUDA(cudaMemcpyAsync(f_host + c0, a->d_f + c0, (size_t)(c1 - c0) * sizeof(double), cudaMemcpyDeviceToHost, sd));
    }
    a->cell_first = 0;
    a->ncells_active = 0;
    cudaEventRecord(ev[2 * NS + 1], sd);
    cudaStreamWaitEvent(sm, ev[2 * NS + 1], 0);
    cudaStreamWaitEvent(sm, ev[NS], 0);   // (uploads that no slab waited for)
    FB2_TRY(rc);
    return fb2_check_device_error(ctx);
}

extern "C" int fb2_assembler_coloring(fb2_assembler* a, int* ncolors, int32_t* cell_color) {
    FB2_CHECK(a, FB2_ERR_BAD_ARG, "fb2_assembler_coloring: null handle");
    FB2_TRY(fb2_coloring_build(a));
    if (ncolors) *ncolors = a->ncolors;
    if (cell_color) memcpy(cell_color, a->cell_color.data(), a->cell_color.size() * sizeof(int32_t));
    return FB2_OK;
}

extern "C" int fb2_assembler_destroy(fb2_assembler* a) {
    if (!a) return FB2_OK;
    cudaSetDevice(a->dh->grid->ctx->device);
    cudaFree(a->d_map);
    cudaFree(a->d_map8);
    cudaFree(a->d_mapb);
    cudaFree(a->d_mapv);
    cudaFree(a->d_cta_list[0]);
    cudaFree(a->d_cta_list[1]);
    cudaFree(a->d_box_zcols);
    cudaFree(a->d_cmat);
    cudaFree(a->d_march_zcols);
    cudaFree(a->d_mapc);
    cudaFree(a->d_dofc);
    cudaFree(a->d_basec);
    cudaFree(a->d_wfirst);
    cudaFree(a->d_wcount);
    fb2_tiles_free(a->tiles);
    cudaFree(a->d_color_cells);
    cudaFree(a->d_cells);
    cudaFree(a->d_nzval);
    cudaFree(a->d_f);
    cudaFree(a->d_u);
    delete a;
    return FB2_OK;
}
