// Measurement and transfer utilities: FP64 FMA peak microbenchmark (roofline denominator for the FP64-bound
// kernels; BASELINE.md section 3 asks for a measured value) and the asynchronous coordinate upload used by the
// end-to-end path (host xyz -> device staging buffer -> padded per-node records).
#include "common.h"

namespace {

__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double seed) {
    // 8 independent FMA chains per thread: enough ILP to saturate the FP64 pipe with 8 warps / SMSP resident
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains alive
}

__global__ void k_repack_xyz(const double* __restrict__ src, int64_t nnodes, int sdim, int xstride, double* __restrict__ dst) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnodes * xstride) return;
    int64_t node = t / xstride;
    int d = (int)(t - node * xstride);
    dst[t] = d < sdim ? src[node * sdim + d] : 0.0;
}

}  // namespace

extern "C" int fb2_measure_fp64_peak(fb2_ctx* ctx, double* tflops) {
    FB2_CHECK(ctx && tflops, FB2_ERR_BAD_ARG, "fb2_measure_fp64_peak: null argument");
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    double* d_out = nullptr;
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
    FB2_CUDA(cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    FB2_CUDA(cudaEventCreate(&e0));
    FB2_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        FB2_CUDA(cudaEventRecord(e0, ctx->stream));
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 1.0 + rep);
        ctx->launches++;
        FB2_CUDA(cudaEventRecord(e1, ctx->stream));
        FB2_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        FB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return FB2_OK;
}

// Device-only coordinate update for the end-to-end path: xyz_host (sdim x nnodes, ideally pinned) is copied
// asynchronously and repacked on the device.  The host copy kept for Dirichlet set-up is NOT touched; use
// fb2_grid_set_coordinates for a persistent change.
extern "C" int fb2_grid_upload_coordinates_async(fb2_grid* g, const double* xyz_host) {
    FB2_CHECK(g && xyz_host, FB2_ERR_BAD_ARG, "fb2_grid_upload_coordinates_async: null argument");
    FB2_NEED_DEVICE(g->ctx);
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)g->nnodes * g->sdim;
    if (!g->d_xyz_stage) FB2_CUDA(cudaMalloc(&g->d_xyz_stage, n * sizeof(double)));
    FB2_CUDA(cudaMemcpyAsync(g->d_xyz_stage, xyz_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t total = g->nnodes * g->xstride;
    k_repack_xyz<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(g->d_xyz_stage, g->nnodes, g->sdim, g->xstride, g->d_xyz);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}
