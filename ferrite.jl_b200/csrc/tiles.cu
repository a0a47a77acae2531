// Tile schedules for the scalar thread-per-cell kernels (k_tile_scalar).
//
// Why: with one RED per local (i, j) the fused kernel is bound by L2 atomic sector throughput (2.4 scattered
// RED sectors per stored entry, profiles/r01_prof_c2_r1c.txt), not by FP64 or HBM.  A tile = TC spatially compact
// cells (Morton order of the cell centroids, any mesh).  The kernel keeps the tile's element matrices in shared
// memory, and a precomputed schedule tells it, for every distinct nzval / f entry the tile touches, which
// shared-memory slots to sum (fixed order => bitwise reproducible inside a tile).  Every entry is then written ONCE
// per tile, entries of a column are contiguous in CSC so the writes coalesce, and columns whose dof lies strictly
// inside the tile ("complete": all cells containing the dof are in the tile) are written with plain stores.
// The schedule replaces the reference's per-cell sort + merge walk (src/assembler.jl:347-457) at tile granularity.
#include <algorithm>
#include <cstring>
#include <parallel/algorithm>

#include "common.h"

namespace {

inline uint64_t spread3(uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

struct TileHost {
    std::vector<int32_t> col_dof;      // bit 31 = complete
    std::vector<uint32_t> ent_col;     // tile-local column of each entry
    std::vector<int32_t> ent_row;      // global row dof, or -1 for the f entry of the column
    std::vector<uint16_t> ent_srcend;  // first source of the entry
    std::vector<uint16_t> ent_nsrc;    // number of sources
    std::vector<uint16_t> src;
};

// Resolve every entry (column dof, row dof) to its position in nzval and pack the kernel's records; CTA per tile.
// Record = uint2 { x = target | flags, y = first source | (number of sources << 16) } with
//   nzval entry: x = (position - tile_base[t])          (30 bits)   | bit 31 if the column is complete in the tile
//   f entry:     x = dof                                (30 bits)   | bit 30 | bit 31 if complete
//   padding:     x = 0xFFFFFFFF
// status[0] counts entries missing in the pattern, status[1] offsets that do not fit 30 bits.
__global__ void k_resolve_entries(const int64_t* __restrict__ ent_ptr, const int64_t* __restrict__ col_ptr,
                                  const int32_t* __restrict__ col_dof, const uint32_t* __restrict__ ent_col,
                                  const int32_t* __restrict__ ent_row, const uint16_t* __restrict__ ent_srcbeg,
                                  const uint16_t* __restrict__ ent_nsrc,
                                  const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                                  int64_t* __restrict__ pos_tmp, int64_t* __restrict__ tile_base, uint2* __restrict__ rec,
                                  int* __restrict__ status) {
    const int64_t t = blockIdx.x;
    const int64_t e0 = ent_ptr[t], e1 = ent_ptr[t + 1], c0 = col_ptr[t];
    __shared__ unsigned long long s_min;
    if (threadIdx.x == 0) s_min = ~0ull;
    __syncthreads();
    unsigned long long mymin = ~0ull;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const uint32_t cl = ent_col[e];
        const int32_t row = ent_row[e];
        int64_t p = -1;
        if (cl != 0xFFFFu && row >= 0) {
            const int32_t j = col_dof[c0 + cl] & 0x7fffffff;
            int64_t lo = colptr[j], hi = colptr[j + 1];
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                int32_t r = rowval[mid];
                if (r == row) { p = mid; break; }
                if (r < row) lo = mid + 1; else hi = mid;
            }
            if (p < 0) atomicAdd(&status[0], 1);
            else mymin = min(mymin, (unsigned long long)p);
        }
        pos_tmp[e] = p;
    }
    atomicMin(&s_min, mymin);
    __syncthreads();
    const int64_t base = s_min == ~0ull ? 0 : (int64_t)s_min;
    if (threadIdx.x == 0) tile_base[t] = base;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const uint32_t cl = ent_col[e];
        uint2 r;
        r.x = 0xFFFFFFFFu;
        r.y = (unsigned)ent_srcbeg[e] | ((unsigned)ent_nsrc[e] << 16);
        if (cl != 0xFFFFu) {
            const int32_t cd = col_dof[c0 + cl];
            const unsigned complete = cd < 0 ? 0x80000000u : 0u;
            if (ent_row[e] < 0) {
                r.x = (unsigned)(cd & 0x3fffffff) | 0x40000000u | complete;
                if ((cd & 0x7fffffff) >= 0x40000000) atomicAdd(&status[1], 1);
            } else {
                const int64_t rel = pos_tmp[e] - base;
                if (rel < 0 || rel >= 0x40000000ll) atomicAdd(&status[1], 1);
                r.x = (unsigned)(rel & 0x3fffffff) | complete;
            }
        }
        rec[e] = r;
    }
}

template <typename T>
int upload(T** d, const std::vector<T>& h) {
    FB2_CUDA(cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty()) FB2_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FB2_OK;
}

}  // namespace

void fb2_tiles_free(TileSchedule* S) {
    if (!S) return;
    cudaFree(S->d_conn); cudaFree(S->d_ncells); cudaFree(S->d_cell_ids); cudaFree(S->d_tile_base);
    cudaFree(S->d_ent_ptr); cudaFree(S->d_rec); cudaFree(S->d_src_ptr); cudaFree(S->d_src);
    delete S;
}

// Build the schedule for all cells of the assembler's grid.  Returns FB2_OK and leaves a->tiles == nullptr when the
// schedule cannot be used (pattern lacks entries, too many columns per tile, ...): callers fall back to k_cell_scalar.
int fb2_tiles_build(fb2_assembler* a, int TC) {
    if (a->tiles || a->tiles_failed) return FB2_OK;
    fb2_dh* dh = a->dh;
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    const int64_t ncells = g->ncells;
    const int nnpc = g->nnpc, nb = dh->ndpc, sdim = g->sdim;
    const int nsym = nb * (nb + 1) / 2;
    if ((int64_t)(nsym + nb) * TC > 65535) { a->tiles_failed = true; return FB2_OK; }
    FB2_CUDA(cudaSetDevice(ctx->device));

    // ---- Morton order of the cell centroids ---------------------------------------------------------------------------
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t n = 0; n < g->nnodes; ++n)
        for (int d = 0; d < sdim; ++d) {
            lo[d] = std::min(lo[d], g->xyz[(size_t)n * sdim + d]);
            hi[d] = std::max(hi[d], g->xyz[(size_t)n * sdim + d]);
        }
    std::vector<std::pair<uint64_t, int32_t>> order((size_t)ncells);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncells; ++c) {
        uint64_t code = 0;
        for (int d = 0; d < sdim; ++d) {
            double s = 0;
            for (int k = 0; k < nnpc; ++k) s += g->xyz[(size_t)(g->cells[(size_t)c * nnpc + k] - 1) * sdim + d];
            s /= nnpc;
            double u = hi[d] > lo[d] ? (s - lo[d]) / (hi[d] - lo[d]) : 0.0;
            uint64_t q = (uint64_t)std::min(2097151.0, std::max(0.0, u * 2097152.0));
            code |= spread3(q) << d;
        }
        order[c] = {code, (int32_t)c};
    }
    __gnu_parallel::sort(order.begin(), order.end());
    const int64_t ntiles = (ncells + TC - 1) / TC;

    // ---- global degree of every dof ---------------------------------------------------------------------------------------
    std::vector<int32_t> deg((size_t)dh->ndofs, 0);
    for (size_t i = 0; i < dh->cell_dofs.size(); ++i) deg[dh->cell_dofs[i]]++;

    // ---- per-tile schedules ---------------------------------------------------------------------------------------------------
    std::vector<TileHost> tiles((size_t)ntiles);
    std::vector<int32_t> h_conn((size_t)ntiles * nnpc * TC), h_ncells((size_t)ntiles), h_cell_ids((size_t)ntiles * TC);
    int max_cols = 0, max_ent = 0, max_src = 0;
    bool overflow = false;
#pragma omp parallel
    {
        std::vector<int32_t> cols;
        std::vector<std::pair<uint64_t, uint16_t>> pairs;
        std::vector<int32_t> incount;
#pragma omp for schedule(dynamic, 16) reduction(max : max_cols, max_ent, max_src)
        for (int64_t t = 0; t < ntiles; ++t) {
            const int64_t c0 = t * TC, nc = std::min<int64_t>(TC, ncells - c0);
            h_ncells[t] = (int32_t)nc;
            for (int cl = 0; cl < TC; ++cl) {
                const int32_t cell = order[c0 + std::min<int64_t>(cl, nc - 1)].second;
                h_cell_ids[(size_t)t * TC + cl] = cell;
                for (int k = 0; k < nnpc; ++k)
                    h_conn[((size_t)t * nnpc + k) * TC + cl] = (int32_t)(g->cells[(size_t)cell * nnpc + k] - 1);
            }
            cols.clear();
            for (int cl = 0; cl < nc; ++cl) {
                const int32_t* cd = &dh->cell_dofs[(size_t)order[c0 + cl].second * nb];
                cols.insert(cols.end(), cd, cd + nb);
            }
            std::sort(cols.begin(), cols.end());
            incount.assign(cols.size(), 0);
            // unique + in-tile multiplicity
            size_t nu = 0;
            for (size_t i = 0; i < cols.size(); ++i) {
                if (nu == 0 || cols[i] != cols[nu - 1]) { cols[nu] = cols[i]; incount[nu] = 1; ++nu; }
                else incount[nu - 1]++;
            }
            cols.resize(nu);
            if (nu >= 32768) { overflow = true; continue; }
            max_cols = std::max(max_cols, (int)nu);
            TileHost& T = tiles[t];
            T.col_dof.resize(nu);
            for (size_t i = 0; i < nu; ++i) T.col_dof[i] = cols[i] | (incount[i] == deg[cols[i]] ? (int32_t)0x80000000 : 0);
            pairs.clear();
            for (int cl = 0; cl < nc; ++cl) {
                const int32_t* cd = &dh->cell_dofs[(size_t)order[c0 + cl].second * nb];
                int lc[32];
                for (int i = 0; i < nb; ++i) lc[i] = (int)(std::lower_bound(cols.begin(), cols.end(), cd[i]) - cols.begin());
                for (int j = 0; j < nb; ++j) {
                    for (int i = 0; i < nb; ++i) {
                        const int a0 = std::min(i, j), b0 = std::max(i, j);
                        const uint16_t slot = (uint16_t)((b0 * (b0 + 1) / 2 + a0) * TC + cl);
                        pairs.push_back({((uint64_t)lc[j] << 32) | (uint32_t)cd[i], slot});
                    }
                    // the f entry of the column sorts after all its rows
                    pairs.push_back({((uint64_t)lc[j] << 32) | 0xFFFFFFFFull, (uint16_t)((nsym + j) * TC + cl)});
                }
            }
            std::sort(pairs.begin(), pairs.end());
            T.src.reserve(pairs.size());
            std::vector<uint32_t> e_col;
            std::vector<int32_t> e_row;
            std::vector<uint16_t> e_beg, e_n;
            for (size_t i = 0; i < pairs.size(); ++i) {
                if (i == 0 || pairs[i].first != pairs[i - 1].first) {
                    e_col.push_back((uint32_t)(pairs[i].first >> 32));
                    const uint32_t r = (uint32_t)(pairs[i].first & 0xFFFFFFFFull);
                    e_row.push_back(r == 0xFFFFFFFFu ? -1 : (int32_t)r);
                    e_beg.push_back((uint16_t)i);
                    e_n.push_back(0);
                }
                e_n.back()++;
                T.src.push_back(pairs[i].second);
            }
            // Order the entries by their number of sources (descending, stable): the lanes of a warp then run the same
            // gather length (no SIMT divergence in phase 2).  Neighbouring entries of one class still share sectors often.
            std::vector<int> perm(e_col.size());
            for (size_t i = 0; i < perm.size(); ++i) perm[i] = (int)i;
            std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return e_n[x] > e_n[y]; });
            for (int i : perm) {
                T.ent_col.push_back(e_col[i]);
                T.ent_row.push_back(e_row[i]);
                T.ent_srcend.push_back(e_beg[i]);   // first source
                T.ent_nsrc.push_back(e_n[i]);
            }
            // pad to multiples of 8 so that every per-tile segment starts 16-byte aligned (cp.async staging);
            // padding entries carry the skip marker (column 0xFFFF) and no sources
            while (T.ent_col.size() % 8) {
                T.ent_col.push_back(0xFFFFu);
                T.ent_row.push_back(-1);
                T.ent_srcend.push_back(0);
                T.ent_nsrc.push_back(0);
            }
            while (T.src.size() % 8) T.src.push_back(0);
            max_ent = std::max(max_ent, (int)T.ent_col.size());
            max_src = std::max(max_src, (int)T.src.size());
        }
    }
    if (overflow) { a->tiles_failed = true; return FB2_OK; }

    // ---- concatenate + upload ----------------------------------------------------------------------------------------------------
    std::vector<int64_t> col_ptr((size_t)ntiles + 1, 0), ent_ptr((size_t)ntiles + 1, 0), src_ptr((size_t)ntiles + 1, 0);
    for (int64_t t = 0; t < ntiles; ++t) {
        col_ptr[t + 1] = col_ptr[t] + (int64_t)tiles[t].col_dof.size();
        ent_ptr[t + 1] = ent_ptr[t] + (int64_t)tiles[t].ent_col.size();
        src_ptr[t + 1] = src_ptr[t] + (int64_t)tiles[t].src.size();
    }
    std::vector<int32_t> col_dof((size_t)col_ptr[ntiles]), ent_row((size_t)ent_ptr[ntiles]);
    std::vector<uint32_t> ent_col((size_t)ent_ptr[ntiles]);
    std::vector<uint16_t> ent_srcend((size_t)ent_ptr[ntiles]), ent_nsrc((size_t)ent_ptr[ntiles]), src((size_t)src_ptr[ntiles]);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < ntiles; ++t) {
        const TileHost& T = tiles[t];
        std::copy(T.col_dof.begin(), T.col_dof.end(), col_dof.begin() + col_ptr[t]);
        std::copy(T.ent_col.begin(), T.ent_col.end(), ent_col.begin() + ent_ptr[t]);
        std::copy(T.ent_row.begin(), T.ent_row.end(), ent_row.begin() + ent_ptr[t]);
        std::copy(T.ent_srcend.begin(), T.ent_srcend.end(), ent_srcend.begin() + ent_ptr[t]);
        std::copy(T.ent_nsrc.begin(), T.ent_nsrc.end(), ent_nsrc.begin() + ent_ptr[t]);
        std::copy(T.src.begin(), T.src.end(), src.begin() + src_ptr[t]);
    }
    std::vector<TileHost>().swap(tiles);

    TileSchedule* S = new TileSchedule();
    S->TC = TC;
    S->ntiles = ntiles;
    S->nslots = nsym + nb;
    S->max_ent = max_ent;
    S->max_src = max_src;
    S->nentries = ent_ptr[ntiles];
    // temporaries of the resolve pass
    int64_t *d_col_ptr = nullptr, *d_pos = nullptr;
    int32_t *d_col_dof = nullptr, *d_ent_row = nullptr;
    uint32_t* d_ent_col = nullptr;
    uint16_t *d_srcend = nullptr, *d_nsrc = nullptr;
    int* d_status = nullptr;
    int rc = FB2_OK;
    auto free_tmp = [&]() {
        cudaFree(d_col_ptr); cudaFree(d_pos); cudaFree(d_col_dof); cudaFree(d_ent_row); cudaFree(d_ent_col); cudaFree(d_srcend);
        cudaFree(d_nsrc); cudaFree(d_status);
        d_col_ptr = d_pos = nullptr; d_col_dof = d_ent_row = nullptr; d_ent_col = nullptr; d_srcend = nullptr; d_nsrc = nullptr; d_status = nullptr;
    };
    auto fail = [&](int code) {
        free_tmp();
        fb2_tiles_free(S);
        return code;
    };
    if ((rc = upload(&S->d_conn, h_conn)) || (rc = upload(&S->d_ncells, h_ncells)) || (rc = upload(&S->d_cell_ids, h_cell_ids)) ||
        (rc = upload(&S->d_ent_ptr, ent_ptr)) || (rc = upload(&S->d_src_ptr, src_ptr)) || (rc = upload(&S->d_src, src)) ||
        (rc = upload(&d_col_ptr, col_ptr)) || (rc = upload(&d_col_dof, col_dof)) || (rc = upload(&d_srcend, ent_srcend)) || (rc = upload(&d_nsrc, ent_nsrc)) ||
        (rc = upload(&d_ent_col, ent_col)) || (rc = upload(&d_ent_row, ent_row)))
        return fail(rc);
    const size_t ne = std::max<size_t>(ent_col.size(), 1);
    cudaError_t e = cudaMalloc(&S->d_rec, ne * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMalloc(&S->d_tile_base, (size_t)ntiles * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_pos, ne * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_status, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, 2 * sizeof(int), ctx->stream);
    if (e != cudaSuccess) return fail(fb2_fail(FB2_ERR_OOM, "tile schedule: %s", cudaGetErrorString(e)));
    k_resolve_entries<<<(unsigned)ntiles, 256, 0, ctx->stream>>>(S->d_ent_ptr, d_col_ptr, d_col_dof, d_ent_col, d_ent_row, d_srcend, d_nsrc,
                                                               a->pat->d_colptr, a->pat->d_rowval, d_pos, S->d_tile_base, S->d_rec,
                                                               d_status);
    ctx->launches++;
    int status[2] = {0, 0};
    e = cudaMemcpyAsync(status, d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free_tmp();
    if (e != cudaSuccess) return fail(fb2_fail(FB2_ERR_CUDA, "tile schedule: %s", cudaGetErrorString(e)));
    if (status[0] > 0 || status[1] > 0) {
        // the pattern lacks entries (the per-cell kernel implements the zero-skip semantics) or a tile spans more
        // than 2^30 nzval positions: keep the per-cell kernel
        fb2_tiles_free(S);
        a->tiles_failed = true;
        return FB2_OK;
    }
    a->tiles = S;
    return FB2_OK;
}
