// Microbenchmark (measurement only): throughput of FP64 RED / STG / TMA bulk reduce-add on B200 as a function of how
// the 32 lanes of a warp spread over sectors.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_micro red_micro.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int STRIDE, bool RED>
__global__ void k_spread(double* a, int64_t n, double v) {
    // warp w handles block [w * 32 * STRIDE, (w+1) * 32 * STRIDE): lane l touches l * STRIDE + k, k = 0..STRIDE-1
    const int64_t nw = n / (32 * STRIDE);
    const int lane = threadIdx.x & 31;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nw; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        double* p = a + w * 32 * STRIDE + (int64_t)lane * STRIDE;
#pragma unroll
        for (int k = 0; k < STRIDE; ++k) {
            if (RED) atomicAdd(p + k, v); else p[k] = v;
        }
    }
}

// one elected thread per CTA issues bulk reduce-adds of BYTES from shared memory
template <int BYTES>
__global__ void k_bulk(double* a, int64_t n, double v) {
    extern __shared__ __align__(128) double s[];
    for (int i = threadIdx.x; i < BYTES / 8; i += blockDim.x) s[i] = v;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int64_t nchunk = n / (BYTES / 8);
    if (threadIdx.x == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(s);
        for (int64_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
            double* dst = a + c * (BYTES / 8);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

template <int STRIDE, bool RED>
void run_spread(double* a, int64_t n, const char* tag, int grid) {
    int64_t nn = n / (32 * STRIDE) * (32 * STRIDE);
    float ms = timeit([&] { k_spread<STRIDE, RED><<<grid, 256>>>(a, nn, 1.0); }, 3);
    printf("%-10s %s stride %2d: %8.3f ms  %7.2f G elem/s  %7.1f GB/s\n", tag, RED ? "RED" : "STG", STRIDE, ms, nn / ms * 1e-6, nn * 8.0 / ms * 1e-6);
}
template <int BYTES>
void run_bulk(double* a, int64_t n, const char* tag, int grid) {
    int64_t nn = n / (BYTES / 8) * (BYTES / 8);
    float ms = timeit([&] { k_bulk<BYTES><<<grid, 128, BYTES>>>(a, nn, 1.0); }, 3);
    printf("%-10s BULK %5d B: %8.3f ms  %7.2f G elem/s  %7.1f GB/s  %6.1f M ops/s\n", tag, BYTES, ms, nn / ms * 1e-6, nn * 8.0 / ms * 1e-6, nn / (BYTES / 8) / ms * 1e-3);
}

int main() {
    const int64_t nbig = 217081801, nsmall = 4 << 20;   // 1.74 GB (C2 nzval) and 32 MB (L2 resident)
    double* a;
    CK(cudaMalloc(&a, nbig * 8));
    CK(cudaMemset(a, 0, nbig * 8));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t n = pass ? nsmall : nbig;
        const char* tag = pass ? "L2(32MB)" : "DRAM(1.7G)";
        const int grid = sms * 8;
        run_spread<1, true>(a, n, tag, grid);
        run_spread<2, true>(a, n, tag, grid);
        run_spread<4, true>(a, n, tag, grid);
        run_spread<8, true>(a, n, tag, grid);
        run_spread<27, true>(a, n, tag, grid);
        run_spread<1, false>(a, n, tag, grid);
        run_spread<4, false>(a, n, tag, grid);
        run_spread<27, false>(a, n, tag, grid);
        run_bulk<208>(a, n, tag, sms * 4);
        run_bulk<432>(a, n, tag, sms * 4);
        run_bulk<1728>(a, n, tag, sms * 4);
        run_bulk<8192>(a, n, tag, sms * 4);
        run_bulk<208>(a, n, tag, sms * 16);
        run_bulk<1728>(a, n, tag, sms * 16);
        float ms = timeit([&] { cudaMemsetAsync(a, 0, n * 8); }, 3);
        printf("%-10s memset: %8.3f ms %7.1f GB/s\n", tag, ms, n * 8.0 / ms * 1e-6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
