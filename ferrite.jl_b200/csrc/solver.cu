// The step after the path: CSR view of the assembled matrix, SpMV and a conjugate-gradient solver that
// consume nzval where the assembly kernels left it (HBM), so a Newton / time step never moves K to the host.
//
// Reference call sites: `IterativeSolvers.cg!(ddu, K, g; maxiter = 1000)` docs/src/literate-tutorials/hyperelasticity.jl:418
// and `u = K \ f` heat_equation.jl:217; CSR storage ext/FerriteSparseMatrixCSR.jl:9-95.
//
// K is stored CSC.  y = K^T x is a gather over the stored columns (sub-warp per column, contiguous reads of nzval /
// rowval); y = K x uses the transpose permutation (CSR values of a structurally symmetric pattern are
// csc[perm[k]] with rowptr = colptr, colval = rowval), also a gather: no atomics, deterministic summation order.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.h"

namespace {

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// perm[k] = position of entry (col_of(k), rowval[k]) read as (row, col), i.e. of the transposed entry
__global__ void k_transpose_perm(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t n,
                                 int64_t* __restrict__ perm, int* __restrict__ gaps) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int j = (int)w;
    for (int64_t k = colptr[j] + lane; k < colptr[j + 1]; k += 32) {
        const int i = rowval[k];                 // stored entry (i, j); look for (j, i) in column i
        int64_t lo = colptr[i], hi = colptr[i + 1], pos = -1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const int r = rowval[mid];
            if (r == j) { pos = mid; break; }
            if (r < j) lo = mid + 1; else hi = mid;
        }
        perm[k] = pos;
        if (pos < 0) atomicAdd(gaps, 1);
    }
}

// TPC threads per column; USE_PERM: values through the transpose permutation (y = K x), else y = K^T x
template <int TPC, bool USE_PERM>
__global__ void k_spmv_gather(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                              const int64_t* __restrict__ perm, const double* __restrict__ nzval,
                              const double* __restrict__ x, double* __restrict__ y, int64_t n) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t col = gid / TPC;
    const int sub = (int)(gid % TPC);
    double s = 0.0;
    if (col < n) {
        const int64_t k1 = colptr[col + 1];
#pragma unroll 4
        for (int64_t k = colptr[col] + sub; k < k1; k += TPC) {
            const double v = USE_PERM ? nzval[perm[k]] : nzval[k];
            s = fma(v, x[rowval[k]], s);
        }
    }
#pragma unroll
    for (int o = TPC / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, TPC);
    if (col < n && sub == 0) y[col] = s;
}

__global__ void k_spmv_scatter(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                               const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y, int64_t n) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const double xj = x[w];
    for (int64_t k = colptr[w] + lane; k < colptr[w + 1]; k += 32) atomicAdd(y + rowval[k], nzval[k] * xj);
}

__global__ void k_csr_values(const int64_t* __restrict__ perm, const double* __restrict__ csc, double* __restrict__ csr, int64_t nnz) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) csr[k] = csc[perm[k]];
}

// ---- CG building blocks: fixed-shape two-stage reductions (deterministic), scalars stay on the device ----------------
constexpr int RB = 1024;   // reduction blocks

// partial[b] = sum over the block's grid-stride slice of a[i] * b[i]
__global__ void k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fma(a[i], b[i], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__device__ double block_sum_partials(const double* partial, int nb) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// scalars: [0] rz, [1] alpha, [2] beta, [3] rr, [4] pAp
__global__ void k_cg_alpha(const double* partial, int nb, double* sc) {
    const double pAp = block_sum_partials(partial, nb);
    if (threadIdx.x == 0) { sc[4] = pAp; sc[1] = sc[0] / pAp; }
}

// x += alpha p; r -= alpha Ap; z = Minv r; partials of r.z and r.r
__global__ void k_cg_update(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, const double* __restrict__ p,
                            const double* __restrict__ Ap, const double* __restrict__ dinv, int64_t n, const double* __restrict__ sc,
                            double* __restrict__ prz, double* __restrict__ prr) {
    __shared__ double sh1[256], sh2[256];
    const double alpha = sc[1];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, Ap[i], r[i]);
        r[i] = ri;
        const double zi = dinv ? dinv[i] * ri : ri;
        z[i] = zi;
        s1 = fma(ri, zi, s1);
        s2 = fma(ri, ri, s2);
    }
    sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sh1[threadIdx.x] += sh1[threadIdx.x + o]; sh2[threadIdx.x] += sh2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { prz[blockIdx.x] = sh1[0]; prr[blockIdx.x] = sh2[0]; }
}

__global__ void k_cg_beta(const double* prz, const double* prr, int nb, double* sc) {
    const double rz = block_sum_partials(prz, nb);
    const double rr = block_sum_partials(prr, nb);
    if (threadIdx.x == 0) { sc[2] = rz / sc[0]; sc[0] = rz; sc[3] = rr; }
}

__global__ void k_cg_p(double* __restrict__ p, const double* __restrict__ z, int64_t n, const double* __restrict__ sc) {
    const double beta = sc[2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}

// r = b - Ax (Ax given), z = Minv r, p = z; partials r.z, r.r
__global__ void k_cg_init(const double* __restrict__ b, const double* __restrict__ Ax, double* __restrict__ r, double* __restrict__ z,
                          double* __restrict__ p, const double* __restrict__ dinv, int64_t n, double* __restrict__ prz, double* __restrict__ prr) {
    __shared__ double sh1[256], sh2[256];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = b[i] - Ax[i];
        r[i] = ri;
        const double zi = dinv ? dinv[i] * ri : ri;
        z[i] = zi; p[i] = zi;
        s1 = fma(ri, zi, s1);
        s2 = fma(ri, ri, s2);
    }
    sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sh1[threadIdx.x] += sh1[threadIdx.x + o]; sh2[threadIdx.x] += sh2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { prz[blockIdx.x] = sh1[0]; prr[blockIdx.x] = sh2[0]; }
}

__global__ void k_cg_init_final(const double* prz, const double* prr, int nb, double* sc) {
    const double rz = block_sum_partials(prz, nb);
    const double rr = block_sum_partials(prr, nb);
    if (threadIdx.x == 0) { sc[0] = rz; sc[3] = rr; sc[1] = 0.0; sc[2] = 0.0; }
}

__global__ void k_diag_inverse(const int64_t* __restrict__ diag, const double* __restrict__ nzval, int64_t n, double* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = diag[i] >= 0 ? nzval[diag[i]] : 0.0;
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
}

int ensure_tperm(fb2_pattern* p) {
    if (p->tperm_state != 0) return FB2_OK;
    fb2_ctx* ctx = p->dh->grid->ctx;
    FB2_CUDA(cudaMalloc(&p->d_tperm, std::max<int64_t>(p->nnz, 1) * sizeof(int64_t)));
    int* d_gaps = nullptr;
    FB2_CUDA(cudaMalloc(&d_gaps, sizeof(int)));
    FB2_CUDA(cudaMemsetAsync(d_gaps, 0, sizeof(int), ctx->stream));
    if (p->n > 0) k_transpose_perm<<<nblk(p->n * 32, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, p->n, p->d_tperm, d_gaps);
    ctx->launches++;
    int gaps = 0;
    cudaError_t e = cudaMemcpyAsync(&gaps, d_gaps, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_gaps);
    FB2_CUDA(e);
    p->tperm_state = gaps == 0 ? 1 : 2;
    return FB2_OK;
}

template <bool USE_PERM>
int launch_gather(fb2_pattern* p, const double* nzval, const double* x, double* y) {
    fb2_ctx* ctx = p->dh->grid->ctx;
    const int64_t n = p->n;
    if (n == 0) return FB2_OK;
    double avg = (double)p->nnz / (double)n;
    // threads per column ~ column length / 12: few threads with four unrolled, independent (rowval -> x) load chains each
    // measured best on B200 (Q1 hex, 27 entries per column: 2 threads 5.46 TB/s, 4: 5.32, 8: 3.6-4.1, 16: 2.7, 32: 1.7)
    if (const char* e = getenv("FB2_SPMV_TPC")) avg = atoi(e) == 2 ? 1 : atoi(e) == 4 ? 40 : atoi(e) == 8 ? 80 : atoi(e) == 16 ? 200 : 1000;   // tuning override
    if (avg <= 36) k_spmv_gather<2, USE_PERM><<<nblk(n * 2, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, p->d_tperm, nzval, x, y, n);
    else if (avg <= 72) k_spmv_gather<4, USE_PERM><<<nblk(n * 4, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, p->d_tperm, nzval, x, y, n);
    else if (avg <= 144) k_spmv_gather<8, USE_PERM><<<nblk(n * 8, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, p->d_tperm, nzval, x, y, n);
    else if (avg <= 288) k_spmv_gather<16, USE_PERM><<<nblk(n * 16, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, p->d_tperm, nzval, x, y, n);
    else k_spmv_gather<32, USE_PERM><<<nblk(n * 32, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, p->d_tperm, nzval, x, y, n);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

}  // namespace

extern "C" int fb2_spmv(fb2_pattern* p, const double* nzval_dev, const double* x_dev, double* y_dev, int transpose) {
    FB2_CHECK(p && nzval_dev && x_dev && y_dev, FB2_ERR_BAD_ARG, "fb2_spmv: null argument");
    fb2_ctx* ctx = p->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(x_dev != y_dev, FB2_ERR_BAD_ARG, "fb2_spmv: x and y must not alias");
    FB2_CUDA(cudaSetDevice(ctx->device));
    if (transpose) return launch_gather<false>(p, nzval_dev, x_dev, y_dev);
    FB2_TRY(ensure_tperm(p));
    if (p->tperm_state == 1) return launch_gather<true>(p, nzval_dev, x_dev, y_dev);
    // not structurally symmetric: column scatter with FP64 atomics
    FB2_CUDA(cudaMemsetAsync(y_dev, 0, (size_t)p->n * sizeof(double), ctx->stream));
    if (p->n > 0) k_spmv_scatter<<<nblk(p->n * 32, 256), 256, 0, ctx->stream>>>(p->d_colptr, p->d_rowval, nzval_dev, x_dev, y_dev, p->n);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_csr_values(fb2_pattern* p, const double* nzval_csc_dev, double* nzval_csr_dev) {
    FB2_CHECK(p && nzval_csc_dev && nzval_csr_dev, FB2_ERR_BAD_ARG, "fb2_csr_values: null argument");
    fb2_ctx* ctx = p->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(nzval_csc_dev != nzval_csr_dev, FB2_ERR_BAD_ARG, "fb2_csr_values: in-place conversion is not supported");
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(ensure_tperm(p));
    FB2_CHECK(p->tperm_state == 1, FB2_ERR_UNSUPPORTED,
              "fb2_csr_values: the pattern is not structurally symmetric; rowptr = colptr, colval = rowval does not describe its CSR form");
    if (p->nnz > 0) k_csr_values<<<nblk(p->nnz, 256), 256, 0, ctx->stream>>>(p->d_tperm, nzval_csc_dev, nzval_csr_dev, p->nnz);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

namespace {

__global__ void k_reciprocal(double* __restrict__ d, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = d[i] != 0.0 ? 1.0 / d[i] : 1.0;
}

// Preconditioned CG on work vectors [r | z | p | Ap | dinv | partials]; A_times(v, out) applies the operator, fill_dinv (jacobi)
// writes the inverse diagonal.  Stopping rule of IterativeSolvers.cg!: ||r|| <= max(reltol ||r0||, abstol).
template <class Op, class Dinv>
int cg_run(fb2_ctx* ctx, int64_t n, double** work, size_t* work_count, Op A_times, int jacobi, Dinv fill_dinv, const double* b_dev,
           double* x_dev, double reltol, double abstol, int maxiter, int* iters, double* resnorm) {
    cudaStream_t st = ctx->stream;
    const size_t need = (size_t)5 * n + 3 * RB + 8;
    if (*work_count < need) {
        cudaFree(*work);
        *work = nullptr;
        *work_count = 0;
        FB2_CUDA(cudaMalloc(work, need * sizeof(double)));
        *work_count = need;
    }
    double *r = *work, *z = r + n, *pp = z + n, *Ap = pp + n, *dinv = Ap + n, *part1 = dinv + n, *part2 = part1 + RB, *part3 = part2 + RB,
           *sc = part3 + RB;
    const int nb = (int)std::min<int64_t>(RB, std::max<int64_t>(1, (n + 255) / 256));
    if (jacobi) FB2_TRY(fill_dinv(dinv));
    const double* dptr = jacobi ? dinv : nullptr;
    FB2_TRY(A_times(x_dev, Ap));
    k_cg_init<<<nb, 256, 0, st>>>(b_dev, Ap, r, z, pp, dptr, n, part1, part2);
    k_cg_init_final<<<1, 256, 0, st>>>(part1, part2, nb, sc);
    ctx->launches += 2;
    double h[5];
    FB2_CUDA(cudaMemcpyAsync(h, sc, 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
    FB2_CUDA(cudaStreamSynchronize(st));
    const double r0 = sqrt(h[3]);
    const double target = std::max(reltol * r0, abstol);
    double rn = r0;
    int it = 0;
    while (it < maxiter && rn > target) {
        FB2_TRY(A_times(pp, Ap));
        k_dot_partial<<<nb, 256, 0, st>>>(pp, Ap, n, part3);
        k_cg_alpha<<<1, 256, 0, st>>>(part3, nb, sc);
        k_cg_update<<<nb, 256, 0, st>>>(x_dev, r, z, pp, Ap, dptr, n, sc, part1, part2);
        k_cg_beta<<<1, 256, 0, st>>>(part1, part2, nb, sc);
        k_cg_p<<<nb, 256, 0, st>>>(pp, z, n, sc);
        ctx->launches += 5;
        FB2_CUDA(cudaMemcpyAsync(h, sc, 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
        FB2_CUDA(cudaStreamSynchronize(st));
        rn = sqrt(h[3]);
        ++it;
        if (!(h[4] > 0.0) && rn > target) {
            if (iters) *iters = it;
            if (resnorm) *resnorm = rn;
            return fb2_fail(FB2_ERR_BAD_ARG, "cg: p'Ap = %g is not positive after %d iterations (matrix not SPD?)", h[4], it);
        }
    }
    if (iters) *iters = it;
    if (resnorm) *resnorm = rn;
    return fb2_check_device_error(ctx);
}

}  // namespace

extern "C" int fb2_cg(fb2_pattern* p, const double* nzval_dev, const double* b_dev, double* x_dev, double reltol, double abstol,
                      int maxiter, int jacobi, int symmetric, int* iters, double* resnorm) {
    FB2_CHECK(p && nzval_dev && b_dev && x_dev, FB2_ERR_BAD_ARG, "fb2_cg: null argument");
    fb2_ctx* ctx = p->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(maxiter >= 0 && reltol >= 0 && abstol >= 0, FB2_ERR_BAD_ARG, "fb2_cg: bad tolerance / iteration limit");
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = p->n;
    if (!symmetric) FB2_TRY(ensure_tperm(p));
    auto A_times = [&](const double* v, double* out) -> int {
        // K symmetric in value: K v = K^T v, the plain gather; otherwise through the transpose permutation
        if (symmetric) return launch_gather<false>(p, nzval_dev, v, out);
        return fb2_spmv(p, nzval_dev, v, out, 0);
    };
    auto fill_dinv = [&](double* dinv) -> int {
        k_diag_inverse<<<nblk(n, 256), 256, 0, ctx->stream>>>(p->d_diag, nzval_dev, n, dinv);
        ctx->launches++;
        return FB2_OK;
    };
    return cg_run(ctx, n, &p->d_work, &p->work_count, A_times, jacobi, fill_dinv, b_dev, x_dev, reltol, abstol, maxiter, iters, resnorm);
}

// CG on the matrix-free operator of the element assembly: A p = sum_e P' Ke P p (fb2_ea_mul); no global matrix exists.
// With Kes / fes after fb2_ea_apply_local the Dirichlet conditions are part of the operator (b = sum_e P' fe).
extern "C" int fb2_ea_cg(fb2_ea* ea, const double* Kes_dev, const double* b_dev, double* x_dev, double reltol, double abstol,
                         int maxiter, int jacobi, int* iters, double* resnorm) {
    FB2_CHECK(ea && Kes_dev && b_dev && x_dev, FB2_ERR_BAD_ARG, "fb2_ea_cg: null argument");
    fb2_ctx* ctx = ea->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(maxiter >= 0 && reltol >= 0 && abstol >= 0, FB2_ERR_BAD_ARG, "fb2_ea_cg: bad tolerance / iteration limit");
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = ea->dh->ndofs;
    auto A_times = [&](const double* v, double* out) -> int { return fb2_ea_mul(ea, Kes_dev, v, out); };
    auto fill_dinv = [&](double* dinv) -> int {
        FB2_TRY(fb2_ea_diag(ea, Kes_dev, dinv));
        k_reciprocal<<<nblk(n, 256), 256, 0, ctx->stream>>>(dinv, n);
        ctx->launches++;
        return FB2_OK;
    };
    return cg_run(ctx, n, &ea->d_work, &ea->work_count, A_times, jacobi, fill_dinv, b_dev, x_dev, reltol, abstol, maxiter, iters, resnorm);
}
