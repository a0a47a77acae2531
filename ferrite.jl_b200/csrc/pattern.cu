// Sparsity pattern construction on the device + cell-local -> nzval offset map.
//
// Replaces allocate_matrix(dh) of the reference (src/Dofs/sparsity_pattern.jl:628-645): the union
// over cells of dofs x dofs plus the diagonal (:370-398, :1136-1249) emitted as a CSC matrix with
// ascending rows per column (:951-991).  The reference builds row -> cells (:1071-1105) and walks
// candidates with a marker array on the CPU; here one warp owns one column: it gathers the dofs of
// all cells touching the column's dof into shared memory, bitonic-sorts them and emits the unique
// values (count pass -> exclusive scan -> fill pass).  colptr/rowval are bit-identical to the
// reference because the result is a pure set function of cell_dofs.
//
// The map built by fb2_map_build replaces the per-cell sort + merge walk of _assemble_inner!
// (src/assembler.jl:347-457): for every cell and local (i, j) it stores the offset of row dof_i inside
// column dof_j as a uint16, so the scatter is nzval[colptr[dof_j] + off] += Ke[i, j].
#include <algorithm>
#include <climits>
#include <cub/cub.cuh>

#include "common.h"

namespace {

__global__ void k_count_incidence(const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int ndpc,
                                  int* __restrict__ cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = ncells * ndpc;
    if (t >= total) return;
    int i = (int)(t / ncells);
    int64_t c = t % ncells;
    atomicAdd(&cnt[cell_dofs[(size_t)i * ncells_pad + c]], 1);
}

__global__ void k_fill_incidence(const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int ndpc,
                                 const int64_t* __restrict__ ptr, int* __restrict__ cursor, int32_t* __restrict__ d2c) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = ncells * ndpc;
    if (t >= total) return;
    int i = (int)(t / ncells);
    int64_t c = t % ncells;
    int dof = cell_dofs[(size_t)i * ncells_pad + c];
    int pos = atomicAdd(&cursor[dof], 1);
    d2c[ptr[dof] + pos] = (int32_t)c;
}

// One warp per column.  FILL == false: write the number of unique rows to colcount[j].
// FILL == true: write the sorted unique rows to rowval[colptr[j]...] and the diagonal position.
template <bool FILL>
__global__ void k_pattern_columns(const int32_t* __restrict__ cell_dofs, int64_t ncells_pad, int ndpc, int64_t ndofs,
                                  const int64_t* __restrict__ d2c_ptr, const int32_t* __restrict__ d2c, int cap,
                                  int64_t* __restrict__ colcount, const int64_t* __restrict__ colptr,
                                  int32_t* __restrict__ rowval, int64_t* __restrict__ diag) {
    extern __shared__ int32_t smem[];
    const int warps = blockDim.x >> 5;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* buf = smem + (size_t)w * cap;
    for (int64_t j = (int64_t)blockIdx.x * warps + w; j < ndofs; j += (int64_t)gridDim.x * warps) {
        const int64_t b = d2c_ptr[j], e = d2c_ptr[j + 1];
        const int ncand = (int)(e - b) * ndpc + 1;
        int P = 32;
        while (P < ncand) P <<= 1;
        for (int t = lane; t < P; t += 32) {
            int32_t v = 0x7fffffff;
            if (t < ncand - 1) {
                int32_t cell = d2c[b + t / ndpc];
                v = cell_dofs[(size_t)(t % ndpc) * ncells_pad + cell];
            } else if (t == ncand - 1) {
                v = (int32_t)j;  // the diagonal is always stored
            }
            buf[t] = v;
        }
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1)
            for (int s = k >> 1; s > 0; s >>= 1) {
                for (int t = lane; t < (P >> 1); t += 32) {
                    int i = ((t & ~(s - 1)) << 1) | (t & (s - 1));  // index with bit s cleared
                    int p = i | s;
                    bool up = (i & k) == 0;
                    int32_t a = buf[i], c = buf[p];
                    if ((a > c) == up) { buf[i] = c; buf[p] = a; }
                }
                __syncwarp();
            }
        if (!FILL) {
            int cnt = 0;
            for (int t = lane; t < ncand; t += 32) cnt += (t == 0 || buf[t] != buf[t - 1]) ? 1 : 0;
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) colcount[j] = cnt;
        } else {
            int64_t base = colptr[j];
            for (int t0 = 0; t0 < ncand; t0 += 32) {
                int t = t0 + lane;
                bool flag = t < ncand && (t == 0 || buf[t] != buf[t - 1]);
                unsigned m = __ballot_sync(0xffffffffu, flag);
                if (flag) {
                    int64_t pos = base + __popc(m & ((1u << lane) - 1u));
                    rowval[pos] = buf[t];
                    if (buf[t] == (int32_t)j) diag[j] = pos;
                }
                base += __popc(m);
            }
        }
        __syncwarp();
    }
}

__global__ void k_collen_max(const int64_t* __restrict__ colptr, int64_t n, int* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int v = 0;
    if (j < n) v = (int)(colptr[j + 1] - colptr[j]);
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}

__global__ void k_find_diag(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t n,
                            int64_t* __restrict__ diag) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int64_t lo = colptr[j], hi = colptr[j + 1];
    int64_t pos = -1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t r = rowval[mid];
        if (r == (int32_t)j) { pos = mid; break; }
        if (r < (int32_t)j) lo = mid + 1; else hi = mid;
    }
    diag[j] = pos;
}

__global__ void k_build_map(const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int n,
                            const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                            uint16_t* __restrict__ map, int* __restrict__ nmissing) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)n * n * ncells_pad;
    if (t >= total) return;
    int64_t c = t % ncells_pad;
    int e = (int)(t / ncells_pad);
    if (c >= ncells) { map[t] = 0xFFFF; return; }
    int i = e % n, j = e / n;
    int32_t di = cell_dofs[(size_t)i * ncells_pad + c], dj = cell_dofs[(size_t)j * ncells_pad + c];
    int64_t b = colptr[dj], lo = b, hi = colptr[dj + 1];
    uint16_t off = 0xFFFF;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t r = rowval[mid];
        if (r == di) { off = (uint16_t)(mid - b); break; }
        if (r < di) lo = mid + 1; else hi = mid;
    }
    map[t] = off;
    if (off == 0xFFFF) atomicAdd(nmissing, 1);   // (cell, i, j) has no entry in the pattern
}

// Packed variant for the thread-per-cell kernels: 8 offsets per 16-byte chunk, chunk k of cell c at
// map8[k * ncells_pad + c] (one coalesced LDG.128 per chunk and warp).
__global__ void k_pack_map(const uint16_t* __restrict__ map, int64_t ncells_pad, int nn, int nchunks, uint16_t* __restrict__ map8) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)nchunks * 8 * ncells_pad;
    if (t >= total) return;
    int s = (int)(t & 7);
    int64_t r = t >> 3;
    int64_t c = r % ncells_pad;
    int k = (int)(r / ncells_pad);
    int e = k * 8 + s;
    map8[t] = e < nn ? map[(size_t)e * ncells_pad + c] : (uint16_t)0xFFFF;
}

// Byte-packed variant for the marching-tile kernel (columns of at most 254 entries): 16 offsets per 16-byte chunk, chunk k
// of cell c at mapb[(k * ncells_pad + c) * 16 ..], 0xFF = no such pattern entry.
__global__ void k_pack_map_bytes(const uint16_t* __restrict__ map, int64_t ncells_pad, int nn, int nchunks, uint8_t* __restrict__ mapb) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)nchunks * 16 * ncells_pad;
    if (t >= total) return;
    int s = (int)(t & 15);
    int64_t r = t >> 4;
    int64_t c = r % ncells_pad;
    int k = (int)(r / ncells_pad);
    int e = k * 16 + s;
    const unsigned v = e < nn ? map[(size_t)e * ncells_pad + c] : 0xFFFFu;
    mapb[t] = v >= 0xFFu ? (uint8_t)0xFF : (uint8_t)v;
}

// Cell-major variant for k_cell_blocks: the n*n offsets of one cell are contiguous (padded to a multiple of 8
// entries), so a CTA stages a cell's block with 16-byte cp.async copies.
__global__ void k_cellmajor_map(const uint16_t* __restrict__ map, int64_t ncells, int64_t ncells_pad, int nn, int stride,
                                uint16_t* __restrict__ mapc) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * stride) return;
    int64_t c = t / stride;
    int e = (int)(t - c * stride);
    mapc[t] = e < nn ? map[(size_t)e * ncells_pad + c] : (uint16_t)0xFFFF;
}

// cell-major dofs for the CTA kernels: a warp reads the n dofs of its cell as one contiguous run instead of n SoA rows
// and their column bases colptr[dof], so that the kernels do not chain a gather behind the dof load
__global__ void k_cellmajor_dofs(const int32_t* __restrict__ cell_dofs, int64_t ncells, int64_t ncells_pad, int n,
                                 const int64_t* __restrict__ colptr, int32_t* __restrict__ dofc, int64_t* __restrict__ basec) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * n) return;
    int64_t c = t / n;
    int i = (int)(t - c * n);
    const int d = cell_dofs[(size_t)i * ncells_pad + c];
    dofc[t] = d;
    basec[t] = colptr[d];
}

__global__ void k_to_onebased64(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int64_t)in[t] + 1;
}

__global__ void k_from_onebased64(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int32_t)(in[t] - 1);
}

__global__ void k_add_scalar64(int64_t* __restrict__ a, int64_t n, int64_t v) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] += v;
}

inline unsigned nblocks(int64_t total, int bs) { return (unsigned)((total + bs - 1) / bs); }

}  // namespace

int fb2_pattern_finalize(fb2_pattern* p) {
    fb2_ctx* ctx = p->dh->grid->ctx;
    cudaStream_t st = ctx->stream;
    int* d_max = nullptr;
    FB2_CUDA(cudaMalloc(&d_max, sizeof(int)));
    FB2_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    k_collen_max<<<nblocks(p->n, 256), 256, 0, st>>>(p->d_colptr, p->n, d_max);
    ctx->launches++;
    FB2_CUDA(cudaMemcpyAsync(&p->max_col_len, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    FB2_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_max);
    FB2_CHECK(p->max_col_len < 65535, FB2_ERR_UNSUPPORTED, "a matrix column has %d stored rows; the uint16 offset map needs < 65535", p->max_col_len);
    return FB2_OK;
}

int fb2_pattern_build_device(fb2_pattern* p) {
    return fb2_pattern_build_device_from(p, p->dh->d_cell_dofs, p->dh->grid->ncells, p->dh->grid->ncells_pad, p->dh->ndpc);
}

int fb2_pattern_build_device_from(fb2_pattern* p, const int32_t* d_cell_dofs, int64_t ncells, int64_t ncells_pad, int ndpc) {
    fb2_dh* dh = p->dh;
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = dh->ndofs;
    const int64_t ninc = ncells * ndpc;
    FB2_CHECK(ninc < (int64_t)INT_MAX, FB2_ERR_UNSUPPORTED, "pattern build: %lld cell-dof incidences exceed the 32-bit counters of the builder", (long long)ninc);
    p->n = n;

    int *d_cnt = nullptr, *d_cursor = nullptr;
    int64_t *d_ptr = nullptr, *d_colcount = nullptr;
    int32_t* d_d2c = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = FB2_OK;
    auto cleanup = [&]() {
        cudaFree(d_cnt); cudaFree(d_cursor); cudaFree(d_ptr); cudaFree(d_colcount); cudaFree(d_d2c); cudaFree(d_tmp);
    };
#define P_CUDA(call)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cleanup();                                                                                 \
            return fb2_fail(e__ == cudaErrorMemoryAllocation ? FB2_ERR_OOM : FB2_ERR_CUDA, "%s failed: %s", #call, \
                            cudaGetErrorString(e__));                                                  \
        }                                                                                              \
    } while (0)

    P_CUDA(cudaMalloc(&d_cnt, (n + 1) * sizeof(int)));
    P_CUDA(cudaMalloc(&d_cursor, (n + 1) * sizeof(int)));
    P_CUDA(cudaMalloc(&d_ptr, (n + 1) * sizeof(int64_t)));
    P_CUDA(cudaMalloc(&d_colcount, (n + 1) * sizeof(int64_t)));
    P_CUDA(cudaMalloc(&d_d2c, ninc * sizeof(int32_t)));
    P_CUDA(cudaMemsetAsync(d_cnt, 0, (n + 1) * sizeof(int), st));
    P_CUDA(cudaMemsetAsync(d_cursor, 0, (n + 1) * sizeof(int), st));
    k_count_incidence<<<nblocks(ninc, 256), 256, 0, st>>>(d_cell_dofs, ncells, ncells_pad, ndpc, d_cnt);
    ctx->launches++;
    // dof -> cells offsets (exclusive scan over n+1 entries; the last count is 0)
    P_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_ptr, n + 1, st));
    {
        size_t b2 = 0;
        P_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, d_colcount, d_colcount, n + 1, st));
        size_t b3 = 0;
        P_CUDA(cub::DeviceReduce::Max(nullptr, b3, d_cnt, d_cursor, n, st));
        tmp_bytes = std::max(tmp_bytes, std::max(b2, b3));
    }
    P_CUDA(cudaMalloc(&d_tmp, tmp_bytes));
    P_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_ptr, n + 1, st));
    ctx->launches++;
    // max cells per dof -> shared-memory capacity per warp
    int maxdeg = 0;
    {
        int* d_maxdeg = nullptr;
        P_CUDA(cudaMalloc(&d_maxdeg, sizeof(int)));
        cudaError_t e = cub::DeviceReduce::Max(d_tmp, tmp_bytes, d_cnt, d_maxdeg, n, st);
        ctx->launches++;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&maxdeg, d_maxdeg, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_maxdeg);
        P_CUDA(e);
    }
    k_fill_incidence<<<nblocks(ninc, 256), 256, 0, st>>>(d_cell_dofs, ncells, ncells_pad, ndpc, d_ptr, d_cursor, d_d2c);
    ctx->launches++;

    int cap = 32;
    while (cap < maxdeg * ndpc + 1) cap <<= 1;
    if (cap > 16384) {
        cleanup();
        return fb2_fail(FB2_ERR_UNSUPPORTED, "pattern build: a dof couples with up to %d candidates (> 16384)", maxdeg * ndpc + 1);
    }
    int warps = 8;
    while (warps > 1 && (size_t)warps * cap * sizeof(int32_t) > 96 * 1024) warps >>= 1;
    size_t smem = (size_t)warps * cap * sizeof(int32_t);
    P_CUDA(cudaFuncSetAttribute(k_pattern_columns<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P_CUDA(cudaFuncSetAttribute(k_pattern_columns<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)std::min<int64_t>((n + warps - 1) / warps, (int64_t)ctx->sm_count * 32);
    P_CUDA(cudaMemsetAsync(d_colcount, 0, (n + 1) * sizeof(int64_t), st));
    k_pattern_columns<false><<<grid, warps * 32, smem, st>>>(d_cell_dofs, ncells_pad, ndpc, n, d_ptr, d_d2c, cap,
                                                             d_colcount, nullptr, nullptr, nullptr);
    ctx->launches++;
    P_CUDA(cudaMalloc(&p->d_colptr, (n + 1) * sizeof(int64_t)));
    P_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_colcount, p->d_colptr, n + 1, st));
    ctx->launches++;
    int64_t nnz = 0;
    P_CUDA(cudaMemcpyAsync(&nnz, p->d_colptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaStreamSynchronize(st));
    p->nnz = nnz;
    P_CUDA(cudaMalloc(&p->d_rowval, nnz * sizeof(int32_t)));
    P_CUDA(cudaMalloc(&p->d_diag, n * sizeof(int64_t)));
    k_pattern_columns<true><<<grid, warps * 32, smem, st>>>(d_cell_dofs, ncells_pad, ndpc, n, d_ptr, d_d2c, cap,
                                                            nullptr, p->d_colptr, p->d_rowval, p->d_diag);
    ctx->launches++;
    P_CUDA(cudaGetLastError());
    P_CUDA(cudaStreamSynchronize(st));
#undef P_CUDA
    cleanup();
    p->structurally_symmetric = true;
    rc = fb2_pattern_finalize(p);
    return rc;
}

extern "C" int fb2_pattern_create(fb2_dh* dh, fb2_pattern** out) {
    FB2_CHECK(dh && out, FB2_ERR_BAD_ARG, "fb2_pattern_create: null argument");
    FB2_NEED_DEVICE(dh->grid->ctx);
    fb2_pattern* p = new fb2_pattern();
    p->dh = dh;
    int rc = fb2_pattern_build_device(p);
    if (rc != FB2_OK) { fb2_pattern_destroy(p); return rc; }
    *out = p;
    return FB2_OK;
}

// allocate_matrix(dh, ch): the pattern of allocate_matrix(dh) plus the entries `_condense!` writes (src/Dofs/sparsity_pattern.jl:
// 782-844 `_add_constraint_entries!`).  Those are, per cell, ext x ext with ext = (cell dofs that are not affinely constrained)
// + (masters of the cell's affinely constrained dofs) -- tests/test_oracle_goldens.py checks this identity against the
// reference's rule -- so the device builder runs unchanged on a table of pseudo-cells: the cells themselves plus one padded
// `ext` list for every cell that holds an affinely constrained dof (duplicates inside a list are harmless for a set union).
extern "C" int fb2_pattern_create_condensed(fb2_dh* dh, fb2_ch* ch, fb2_pattern** out) {
    FB2_CHECK(dh && ch && out, FB2_ERR_BAD_ARG, "fb2_pattern_create_condensed: null argument");
    FB2_CHECK(ch->closed && ch->dh == dh, FB2_ERR_BAD_ARG, "fb2_pattern_create_condensed: the ConstraintHandler must be closed and belong to this DofHandler");
    if (!ch->has_affine) return fb2_pattern_create(dh, out);
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int ndpc = dh->ndpc;
    const int64_t nc = g->ncells;
    std::vector<std::vector<int32_t>> ext;
    size_t W = (size_t)ndpc;
    for (int64_t c = 0; c < nc; ++c) {
        const int32_t* cd = &dh->cell_dofs[(size_t)c * ndpc];
        bool any = false;
        for (int i = 0; i < ndpc && !any; ++i) {
            const int32_t ip = ch->dofmap[cd[i]];
            any = ip >= 0 && ch->aff_ptr[ip + 1] > ch->aff_ptr[ip];
        }
        if (!any) continue;
        std::vector<int32_t> e;
        for (int i = 0; i < ndpc; ++i) {
            const int32_t ip = ch->dofmap[cd[i]];
            if (ip >= 0 && ch->aff_is[ip]) {      // an AffineConstraint (even one without masters) is replaced by its masters
                for (int t = ch->aff_ptr[ip]; t < ch->aff_ptr[ip + 1]; ++t) e.push_back(ch->aff_dof[t]);
            } else {
                e.push_back(cd[i]);
            }
        }
        std::sort(e.begin(), e.end());
        e.erase(std::unique(e.begin(), e.end()), e.end());
        if (e.empty()) continue;
        W = std::max(W, e.size());
        ext.push_back(std::move(e));
    }
    const int64_t ntot = nc + (int64_t)ext.size(), npad = (ntot + 31) / 32 * 32;
    std::vector<int32_t> tab((size_t)W * npad, 0);
    for (int64_t c = 0; c < nc; ++c) {
        const int32_t* cd = &dh->cell_dofs[(size_t)c * ndpc];
        for (size_t i = 0; i < W; ++i) tab[i * npad + c] = cd[i < (size_t)ndpc ? i : 0];
    }
    for (size_t k = 0; k < ext.size(); ++k)
        for (size_t i = 0; i < W; ++i) tab[i * npad + nc + k] = ext[k][i < ext[k].size() ? i : 0];
    int32_t* d_tab = nullptr;
    FB2_CUDA(cudaMalloc(&d_tab, tab.size() * sizeof(int32_t)));
    cudaError_t e = cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d_tab); return fb2_fail(FB2_ERR_CUDA, "fb2_pattern_create_condensed: %s", cudaGetErrorString(e)); }
    fb2_pattern* p = new fb2_pattern();
    p->dh = dh;
    int rc = fb2_pattern_build_device_from(p, d_tab, ntot, npad, (int)W);
    cudaFree(d_tab);
    if (rc != FB2_OK) { fb2_pattern_destroy(p); return rc; }
    *out = p;
    return FB2_OK;
}

extern "C" int fb2_pattern_from_host(fb2_dh* dh, const int64_t* colptr, const int64_t* rowval, fb2_pattern** out) {
    FB2_CHECK(dh && colptr && rowval && out, FB2_ERR_BAD_ARG, "fb2_pattern_from_host: null argument");
    fb2_ctx* ctx = dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = dh->ndofs;
    FB2_CHECK(colptr[0] == 1, FB2_ERR_BAD_ARG, "fb2_pattern_from_host: colptr must be 1-based");
    for (int64_t j = 0; j < n; ++j)
        FB2_CHECK(colptr[j + 1] >= colptr[j], FB2_ERR_BAD_ARG, "fb2_pattern_from_host: colptr not monotone at column %lld", (long long)j + 1);
    const int64_t nnz = colptr[n] - 1;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k) {
            FB2_CHECK(rowval[k] >= 1 && rowval[k] <= n, FB2_ERR_BAD_ARG, "fb2_pattern_from_host: row index out of range");
            FB2_CHECK(k == colptr[j] - 1 || rowval[k] > rowval[k - 1], FB2_ERR_BAD_ARG,
                      "fb2_pattern_from_host: rows of column %lld are not strictly ascending", (long long)j + 1);
        }
    fb2_pattern* p = new fb2_pattern();
    p->dh = dh;
    p->n = n;
    p->nnz = nnz;
    int64_t* d_tmp = nullptr;
    cudaError_t e = cudaMalloc(&p->d_colptr, (n + 1) * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_rowval, std::max<int64_t>(nnz, 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_diag, n * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_tmp, std::max<int64_t>(nnz, 1) * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_colptr, colptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tmp, rowval, nnz * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        cudaFree(d_tmp);
        fb2_pattern_destroy(p);
        return fb2_fail(e == cudaErrorMemoryAllocation ? FB2_ERR_OOM : FB2_ERR_CUDA, "fb2_pattern_from_host: %s", cudaGetErrorString(e));
    }
    k_add_scalar64<<<nblocks(n + 1, 256), 256, 0, st>>>(p->d_colptr, n + 1, -1);
    if (nnz > 0) k_from_onebased64<<<nblocks(nnz, 256), 256, 0, st>>>(d_tmp, nnz, p->d_rowval);
    k_find_diag<<<nblocks(n, 256), 256, 0, st>>>(p->d_colptr, p->d_rowval, n, p->d_diag);
    ctx->launches += 3;
    e = cudaStreamSynchronize(st);
    cudaFree(d_tmp);
    if (e != cudaSuccess) { fb2_pattern_destroy(p); return fb2_fail(FB2_ERR_CUDA, "fb2_pattern_from_host: %s", cudaGetErrorString(e)); }
    p->structurally_symmetric = false;
    int rc = fb2_pattern_finalize(p);
    if (rc != FB2_OK) { fb2_pattern_destroy(p); return rc; }
    *out = p;
    return FB2_OK;
}

extern "C" int fb2_pattern_info(fb2_pattern* p, int64_t* n, int64_t* nnz) {
    FB2_CHECK(p, FB2_ERR_BAD_ARG, "fb2_pattern_info: null handle");
    if (n) *n = p->n;
    if (nnz) *nnz = p->nnz;
    return FB2_OK;
}

extern "C" int fb2_pattern_export(fb2_pattern* p, int64_t* colptr, int64_t* rowval) {
    FB2_CHECK(p, FB2_ERR_BAD_ARG, "fb2_pattern_export: null handle");
    fb2_ctx* ctx = p->dh->grid->ctx;
    cudaStream_t st = ctx->stream;
    FB2_CUDA(cudaSetDevice(ctx->device));
    if (colptr) {
        FB2_CUDA(cudaMemcpyAsync(colptr, p->d_colptr, (p->n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        FB2_CUDA(cudaStreamSynchronize(st));
        for (int64_t j = 0; j <= p->n; ++j) colptr[j] += 1;
    }
    if (rowval && p->nnz > 0) {
        // widen on the device in chunks so that the staging buffer stays small
        const int64_t chunk = 1 << 26;
        int64_t* d_tmp = nullptr;
        FB2_CUDA(cudaMalloc(&d_tmp, std::min(chunk, p->nnz) * sizeof(int64_t)));
        for (int64_t s = 0; s < p->nnz; s += chunk) {
            int64_t m = std::min(chunk, p->nnz - s);
            k_to_onebased64<<<nblocks(m, 256), 256, 0, st>>>(p->d_rowval + s, m, d_tmp);
            ctx->launches++;
            cudaError_t e = cudaMemcpyAsync(rowval + s, d_tmp, m * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(d_tmp); return fb2_fail(FB2_ERR_CUDA, "fb2_pattern_export: %s", cudaGetErrorString(e)); }
        }
        cudaFree(d_tmp);
    }
    return FB2_OK;
}

extern "C" int fb2_pattern_destroy(fb2_pattern* p) {
    if (!p) return FB2_OK;
    cudaSetDevice(p->dh->grid->ctx->device);
    cudaFree(p->d_colptr);
    cudaFree(p->d_rowval);
    cudaFree(p->d_diag);
    cudaFree(p->d_tperm);
    cudaFree(p->d_work);
    delete p;
    return FB2_OK;
}

int fb2_map_build_packed(fb2_assembler* a) {
    if (a->d_map8) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int nn = a->n * a->n, nchunks = (nn + 7) / 8;
    const int64_t total = (int64_t)nchunks * 8 * g->ncells_pad;
    FB2_CUDA(cudaMalloc(&a->d_map8, total * sizeof(uint16_t)));
    k_pack_map<<<nblocks(total, 256), 256, 0, ctx->stream>>>(a->d_map, g->ncells_pad, nn, nchunks, a->d_map8);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    return FB2_OK;
}

int fb2_map_build_bytes(fb2_assembler* a) {
    if (a->d_mapb) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CHECK(a->pat->max_col_len < 255, FB2_ERR_UNSUPPORTED, "byte-packed offset map needs columns shorter than 255 entries");
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int nn = a->n * a->n, nchunks = (nn + 15) / 16;
    const int64_t total = (int64_t)nchunks * 16 * g->ncells_pad;
    FB2_CUDA(cudaMalloc(&a->d_mapb, total));
    k_pack_map_bytes<<<nblocks(total, 256), 256, 0, ctx->stream>>>(a->d_map, g->ncells_pad, nn, nchunks, a->d_mapb);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    return FB2_OK;
}

int fb2_map_build_cellmajor(fb2_assembler* a) {
    if (a->d_mapc) return FB2_OK;
    fb2_grid* g = a->dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int nn = a->n * a->n, stride = (nn + 7) / 8 * 8;
    const int64_t total = g->ncells * stride;
    FB2_CUDA(cudaMalloc(&a->d_mapc, total * sizeof(uint16_t)));
    k_cellmajor_map<<<nblocks(total, 256), 256, 0, ctx->stream>>>(a->d_map, g->ncells, g->ncells_pad, nn, stride, a->d_mapc);
    FB2_CUDA(cudaMalloc(&a->d_dofc, (size_t)g->ncells * a->n * sizeof(int32_t)));
    FB2_CUDA(cudaMalloc(&a->d_basec, (size_t)g->ncells * a->n * sizeof(int64_t)));
    k_cellmajor_dofs<<<nblocks(g->ncells * a->n, 256), 256, 0, ctx->stream>>>(a->dh->d_cell_dofs, g->ncells, g->ncells_pad, a->n,
                                                                               a->pat->d_colptr, a->d_dofc, a->d_basec);
    ctx->launches += 2;
    FB2_CUDA(cudaGetLastError());
    FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    return FB2_OK;
}

int fb2_map_build(fb2_assembler* a) {
    fb2_dh* dh = a->dh;
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int n = a->n;
    const int64_t total = (int64_t)n * n * g->ncells_pad;
    FB2_CUDA(cudaMalloc(&a->d_map, total * sizeof(uint16_t)));
    int* d_missing = nullptr;
    FB2_CUDA(cudaMalloc(&d_missing, sizeof(int)));
    FB2_CUDA(cudaMemsetAsync(d_missing, 0, sizeof(int), ctx->stream));
    k_build_map<<<nblocks(total, 256), 256, 0, ctx->stream>>>(dh->d_cell_dofs, g->ncells, g->ncells_pad, n, a->pat->d_colptr,
                                                              a->pat->d_rowval, a->d_map, d_missing);
    ctx->launches++;
    int missing = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_missing);
    FB2_CHECK(e == cudaSuccess, FB2_ERR_CUDA, "fb2_map_build: %s", cudaGetErrorString(e));
    // patterns built by fb2_pattern_create contain every (i, j) of every cell: the kernels then scatter without
    // the per-entry zero / missing checks of assemble! (src/assembler.jl:376-457)
    a->map_complete = missing == 0;
    return FB2_OK;
}
