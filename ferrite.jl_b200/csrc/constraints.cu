// ConstraintHandler with Dirichlet conditions: host-side set-up, device-side apply!.
//
// Set-up mirrors src/Dofs/ConstraintHandler.jl of the reference: add!(ch, dbc) :965-1008, entity -> local
// dofs _local_facet_dofs_for_bc :431-444, _add! for boundary entities :404-428 and node sets :446-493,
// add_prescribed_dof! (a later condition overrides an earlier one) :383-401, close! (sort) :303-361,
// update!/_update! :504-580 with dof locations from BCValues (src/FEValues/FacetValues.jl:185-236).
//
// apply!(K, f, ch) (:710-740) on the device, in the reference's order of effects:
//   m = meandiag(K) (:952-958)                         -> k_meandiag_partial/k_meandiag_final (fixed tree order,
//                                                         deterministic; uses the diagonal-position table)
//   f -= K[:, d] * v_d for prescribed d, v_d != 0 (:755-768), zero_out_columns! (:931-938)
//                                                      -> k_apply_columns (one warp per prescribed column)
//   zero_out_rows! (:940-950)                          -> k_apply_rows_sym (pattern structurally symmetric: the
//                                                         stored columns of row r are the stored rows of column r)
//                                                         or k_apply_rows_stream (any pattern: one pass over nnz)
//   K[d,d] = m, f[d] = v_d * m (:731-738)              -> k_apply_diag
// The reference's GPU extension does the same with 4 KernelAbstractions kernels and column scans
// (ext/FerriteKAExt/constraints.jl:40-148).
#include <algorithm>
#include <cstring>

#include "common.h"

namespace {

__global__ void k_meandiag_partial(const double* __restrict__ nzval, const int64_t* __restrict__ diag, int64_t n,
                                   double* __restrict__ partial) {
    __shared__ double sh[256];
    // each block sums a fixed contiguous slice in a fixed order -> bitwise reproducible
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t b = (int64_t)blockIdx.x * per, e = min(n, b + per);
    double s = 0.0;
    for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        int64_t p = diag[i];
        if (p >= 0) s += fabs(nzval[p]);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void k_meandiag_final(double* __restrict__ partial, int nblocks, int64_t n) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[nblocks] = sh[0] / (double)n;
}

// one warp per prescribed column: f[row] -= v * K[row, d], then zero the column
__global__ void k_apply_columns(const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np,
                                const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                                double* __restrict__ nzval, double* __restrict__ f, int applyzero) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= np) return;
    const int d = prescribed[w];
    const double v = applyzero ? 0.0 : inhom[w];
    const int64_t b = colptr[d], e = colptr[d + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
        if (f != nullptr && v != 0.0) atomicAdd(f + rowval[k], -v * nzval[k]);
        nzval[k] = 0.0;
    }
}

// structurally symmetric pattern: for prescribed row r, visit the rows c of column r and zero entry (r, c)
__global__ void k_apply_rows_sym(const int32_t* __restrict__ prescribed, int64_t np, const int64_t* __restrict__ colptr,
                                 const int32_t* __restrict__ rowval, double* __restrict__ nzval) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= np) return;
    const int r = prescribed[w];
    const int64_t b = colptr[r], e = colptr[r + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
        const int c = rowval[k];
        int64_t lo = colptr[c], hi = colptr[c + 1];
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            int rr = rowval[mid];
            if (rr == r) { nzval[mid] = 0.0; break; }
            if (rr < r) lo = mid + 1; else hi = mid;
        }
    }
}

__global__ void k_apply_rows_stream(const uint8_t* __restrict__ isconstrained, const int32_t* __restrict__ rowval, int64_t nnz,
                                    double* __restrict__ nzval) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        if (isconstrained[rowval[k]]) nzval[k] = 0.0;
}

__global__ void k_apply_diag(const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np,
                             const int64_t* __restrict__ diag, const double* __restrict__ mean, double* __restrict__ nzval,
                             double* __restrict__ f, int applyzero) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const int d = prescribed[i];
    const double m = *mean;
    const int64_t p = diag[d];
    if (p >= 0) nzval[p] = m;
    if (f != nullptr) f[d] = (applyzero ? 0.0 : inhom[i]) * m;
}

__global__ void k_apply_vector(const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np,
                               double* __restrict__ u, int applyzero) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) u[prescribed[i]] = applyzero ? 0.0 : inhom[i];
}

// RHSData (src/Dofs/ConstraintHandler.jl:191-208): lengths of the prescribed columns, then their rows / values
__global__ void k_rhs_lengths(const int32_t* __restrict__ prescribed, int64_t np, const int64_t* __restrict__ colptr,
                              int64_t* __restrict__ len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) len[i] = colptr[prescribed[i] + 1] - colptr[prescribed[i]];
}

__global__ void k_rhs_copy(const int32_t* __restrict__ prescribed, int64_t np, const int64_t* __restrict__ colptr,
                           const int32_t* __restrict__ rowval, const double* __restrict__ nzval, const int64_t* __restrict__ ptr,
                           int32_t* __restrict__ rows, double* __restrict__ vals) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= np) return;
    const int64_t b = colptr[prescribed[w]], e = colptr[prescribed[w] + 1], o = ptr[w];
    for (int64_t k = b + lane; k < e; k += 32) {
        rows[o + k - b] = rowval[k];
        vals[o + k - b] = nzval[k];
    }
}

// apply_rhs! first loop (:219-226): f[row] -= v * K[row, d] from the stored columns, warp per prescribed dof
__global__ void k_rhs_columns(const double* __restrict__ inhom, int64_t np, const int64_t* __restrict__ ptr,
                              const int32_t* __restrict__ rows, const double* __restrict__ vals, double* __restrict__ f) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= np) return;
    const double v = inhom[w];
    if (v == 0.0) return;
    for (int64_t k = ptr[w] + lane; k < ptr[w + 1]; k += 32) atomicAdd(f + rows[k], -v * vals[k]);
}

// second loop (:227-237): f[pdof] = b * m
__global__ void k_rhs_prescribed(const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np, double m,
                                 double* __restrict__ f, int applyzero) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) f[prescribed[i]] = (applyzero ? 0.0 : inhom[i]) * m;
}

// ---- affine constraints (src/Dofs/ConstraintHandler.jl:782-867 `_condense!`, :686-700 apply!(u, ch)) ----------------------------
// position of entry (row, col), -1 if the pattern has none
__device__ __forceinline__ int64_t fb2_find_entry(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int row, int col) {
    int64_t lo = colptr[col], hi = colptr[col + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int r = rowval[mid];
        if (r == row) return mid;
        if (r < row) lo = mid + 1; else hi = mid;
    }
    return -1;
}

__device__ __forceinline__ void fb2_condense_add(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, double* __restrict__ nzval,
                                                 int row, int col, double v, int* errflag) {
    const int64_t p = fb2_find_entry(colptr, rowval, row, col);
    if (p >= 0) atomicAdd(nzval + p, v);
    else if (atomicCAS(&errflag[0], 0, FB2_ERR_MISSING_PATTERN_ENTRY) == 0) errflag[1] = -1;   // use allocate_matrix(dh, ch)
}

// number of stored entries (r, c) without a stored transpose (c, r): adopted patterns are not known to be symmetric
__global__ void k_count_asymmetric(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t n, int* __restrict__ count) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    int bad = 0;
    for (int64_t k = colptr[w] + lane; k < colptr[w + 1]; k += 32)
        if (fb2_find_entry(colptr, rowval, (int)w, rowval[k]) < 0) ++bad;
    if (bad) atomicAdd(count, bad);
}

// f -= K * g without touching K (add_inhomogeneities!, :770-780): one warp per prescribed column with a non-zero value
__global__ void k_add_inhomogeneities(const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np,
                                      const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                                      const double* __restrict__ nzval, double* __restrict__ f) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= np) return;
    const double v = inhom[w];
    if (v == 0.0) return;
    const int d = prescribed[w];
    for (int64_t k = colptr[d] + lane; k < colptr[d + 1]; k += 32) atomicAdd(f + rowval[k], -v * nzval[k]);
}

// `_condense!` of K: one warp per affinely constrained dof p (with masters).  Sources are the entries of row p and of column
// p, targets the entries of the masters' rows / columns; nested constraints are rejected at close!, so no target is a
// source and the adds commute (REDs).  The pattern is structurally symmetric, so the rows of column p list the columns
// that hold an entry in row p.
__global__ void k_condense(const int32_t* __restrict__ list, int64_t na, const int32_t* __restrict__ prescribed,
                           const int32_t* __restrict__ aptr, const int32_t* __restrict__ adof, const double* __restrict__ acoef,
                           const int32_t* __restrict__ aff_of, const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                           double* __restrict__ nzval, int* errflag) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= na) return;
    const int ip = list[w], p = prescribed[ip];
    const int cb = aptr[ip], ce = aptr[ip + 1];
    for (int64_t k = colptr[p] + lane; k < colptr[p + 1]; k += 32) {
        const int q = rowval[k];
        // column p, row q: K[q, p]
        const double kv = nzval[k];
        if (kv != 0.0) {
            const int iq = aff_of[q];
            if (iq < 0) {
                for (int t = cb; t < ce; ++t) fb2_condense_add(colptr, rowval, nzval, q, adof[t], acoef[t] * kv, errflag);
            } else {
                for (int t1 = aptr[iq]; t1 < aptr[iq + 1]; ++t1)
                    for (int t2 = cb; t2 < ce; ++t2)
                        fb2_condense_add(colptr, rowval, nzval, adof[t1], adof[t2], acoef[t1] * acoef[t2] * kv, errflag);
            }
        }
        // row p, column q (unconstrained columns only: constrained ones are handled by their own warp above): K[p, q]
        if (aff_of[q] < 0) {
            const int64_t pos = fb2_find_entry(colptr, rowval, p, q);
            if (pos >= 0) {
                const double kr = nzval[pos];
                if (kr != 0.0)
                    for (int t = cb; t < ce; ++t) fb2_condense_add(colptr, rowval, nzval, adof[t], q, acoef[t] * kr, errflag);
            }
        }
    }
}

// `_condense!` of f: f[master] += coeff f[p], f[p] = 0
__global__ void k_condense_rhs(const int32_t* __restrict__ list, int64_t na, const int32_t* __restrict__ prescribed,
                               const int32_t* __restrict__ aptr, const int32_t* __restrict__ adof, const double* __restrict__ acoef,
                               double* __restrict__ f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const int ip = list[i], p = prescribed[ip];
    const double fp = f[p];
    for (int t = aptr[ip]; t < aptr[ip + 1]; ++t) atomicAdd(f + adof[t], acoef[t] * fp);
    f[p] = 0.0;
}

// zero the prescribed columns (after the condensation has read them)
__global__ void k_zero_columns_of(const int32_t* __restrict__ prescribed, int64_t np, const int64_t* __restrict__ colptr, double* __restrict__ nzval) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= np) return;
    const int d = prescribed[w];
    for (int64_t k = colptr[d] + lane; k < colptr[d + 1]; k += 32) nzval[k] = 0.0;
}

// apply!(u, ch) with affine constraints: u_p = b_p + sum coeff u_master (masters are free dofs, untouched by this kernel)
__global__ void k_apply_vector_affine(const int32_t* __restrict__ prescribed, const double* __restrict__ inhom, int64_t np,
                                      const int32_t* __restrict__ aptr, const int32_t* __restrict__ adof, const double* __restrict__ acoef,
                                      double* __restrict__ u, int applyzero) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    double v = applyzero ? 0.0 : inhom[i];
    for (int t = aptr[i]; t < aptr[i + 1]; ++t) v += acoef[t] * u[adof[t]];
    u[prescribed[i]] = v;
}

inline unsigned nblocks(int64_t total, int bs) { return (unsigned)((total + bs - 1) / bs); }

void add_prescribed(fb2_ch* ch, int64_t dof) {
    ch->affine.erase(dof);   // a later constraint on the same dof replaces the earlier one
    if (ch->dofmap[dof] < 0) {
        ch->dofmap[dof] = (int32_t)ch->insertion.size();
        ch->insertion.push_back(dof);
    }
}

int upload_inhom(fb2_ch* ch) {
    if (!ch->inhom_dirty) return FB2_OK;
    fb2_ctx* ctx = ch->dh->grid->ctx;
    if (!ch->prescribed.empty()) {
        FB2_CUDA(cudaMemcpyAsync(ch->d_inhom, ch->inhom.data(), ch->inhom.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        FB2_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ch->inhom_dirty = false;
    return FB2_OK;
}

int finish_close(fb2_ch* ch) {
    fb2_ctx* ctx = ch->dh->grid->ctx;
    ch->closed = true;
    ch->inhom_dirty = true;
    if (ctx->device < 0) return FB2_OK;  // host-only context
    FB2_CUDA(cudaSetDevice(ctx->device));
    const size_t np = ch->prescribed.size();
    const int64_t n = ch->dh->ndofs;
    std::vector<int32_t> p32(np);
    std::vector<uint8_t> isc((size_t)n, 0);
    for (size_t i = 0; i < np; ++i) { p32[i] = (int32_t)ch->prescribed[i]; isc[ch->prescribed[i]] = 1; }
    FB2_CUDA(cudaMalloc(&ch->d_prescribed, std::max<size_t>(np, 1) * sizeof(int32_t)));
    FB2_CUDA(cudaMalloc(&ch->d_inhom, std::max<size_t>(np, 1) * sizeof(double)));
    FB2_CUDA(cudaMalloc(&ch->d_isconstrained, (size_t)n));
    FB2_CUDA(cudaMalloc(&ch->d_scratch, 2048 * sizeof(double)));
    if (np) FB2_CUDA(cudaMemcpy(ch->d_prescribed, p32.data(), np * sizeof(int32_t), cudaMemcpyHostToDevice));
    FB2_CUDA(cudaMemcpy(ch->d_isconstrained, isc.data(), (size_t)n, cudaMemcpyHostToDevice));
    if (ch->has_affine) {
        std::vector<int32_t> aff_of((size_t)n, -1), list;
        for (size_t i = 0; i < np; ++i) {
            if (ch->aff_is[i]) aff_of[ch->prescribed[i]] = (int32_t)i;     // an AffineConstraint without masters adds nothing
            if (ch->aff_ptr[i + 1] > ch->aff_ptr[i]) list.push_back((int32_t)i);
        }
        ch->n_aff = (int64_t)list.size();
        FB2_CUDA(cudaMalloc(&ch->d_aff_ptr, ch->aff_ptr.size() * sizeof(int32_t)));
        FB2_CUDA(cudaMalloc(&ch->d_aff_dof, std::max<size_t>(ch->aff_dof.size(), 1) * sizeof(int32_t)));
        FB2_CUDA(cudaMalloc(&ch->d_aff_coef, std::max<size_t>(ch->aff_coef.size(), 1) * sizeof(double)));
        FB2_CUDA(cudaMalloc(&ch->d_aff_of, (size_t)n * sizeof(int32_t)));
        FB2_CUDA(cudaMalloc(&ch->d_aff_list, std::max<size_t>(list.size(), 1) * sizeof(int32_t)));
        FB2_CUDA(cudaMemcpy(ch->d_aff_ptr, ch->aff_ptr.data(), ch->aff_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        FB2_CUDA(cudaMemcpy(ch->d_aff_dof, ch->aff_dof.data(), ch->aff_dof.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        FB2_CUDA(cudaMemcpy(ch->d_aff_coef, ch->aff_coef.data(), ch->aff_coef.size() * sizeof(double), cudaMemcpyHostToDevice));
        FB2_CUDA(cudaMemcpy(ch->d_aff_of, aff_of.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
        FB2_CUDA(cudaMemcpy(ch->d_aff_list, list.data(), list.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    ch->closed = true;
    ch->inhom_dirty = true;
    return FB2_OK;
}

// dofcoefficients aligned with the sorted prescribed dofs; nested constraints are rejected like the reference's close!
// (src/Dofs/ConstraintHandler.jl:338-361).  Restriction: masters must be unconstrained dofs.
int build_affine_tables(fb2_ch* ch) {
    const size_t np = ch->prescribed.size();
    ch->aff_ptr.assign(np + 1, 0);
    ch->aff_is.assign(np, 0);
    ch->aff_dof.clear();
    ch->aff_coef.clear();
    ch->has_affine = false;
    for (size_t i = 0; i < np; ++i) {
        auto it = ch->affine.find(ch->prescribed[i]);
        if (it != ch->affine.end()) {
            ch->aff_is[i] = 1;
            for (size_t k = 0; k < it->second.masters.size(); ++k) {
                const int64_t m = it->second.masters[k];
                if (ch->dofmap[m] >= 0) {
                    auto im = ch->affine.find(m);
                    if (im != ch->affine.end() && !im->second.masters.empty())
                        return fb2_fail(FB2_ERR_BAD_ARG, "nested affine constraints currently not supported");
                    return fb2_fail(FB2_ERR_UNSUPPORTED, "affine constraint of dof %lld: master dof %lld is itself prescribed (Dirichlet masters are not supported)",
                                    (long long)ch->prescribed[i] + 1, (long long)m + 1);
                }
                ch->aff_dof.push_back((int32_t)m);
                ch->aff_coef.push_back(it->second.coefs[k]);
                ch->has_affine = true;
            }
        }
        ch->aff_ptr[i + 1] = (int32_t)ch->aff_dof.size();
    }
    return FB2_OK;
}

// update!(ch, t) leaves the inhomogeneity of an affine constraint at its b (:555-560)
void set_affine_inhomogeneities(fb2_ch* ch) {
    for (const auto& kv : ch->affine) {
        const int32_t i = ch->dofmap[kv.first];
        if (i >= 0) ch->inhom[i] = kv.second.b;
    }
}

}  // namespace

extern "C" int fb2_ch_create(fb2_dh* dh, fb2_ch** out) {
    FB2_CHECK(dh && out, FB2_ERR_BAD_ARG, "fb2_ch_create: null argument");
    fb2_ch* ch = new fb2_ch();
    ch->dh = dh;
    ch->dofmap.assign((size_t)dh->ndofs, -1);
    *out = ch;
    return FB2_OK;
}

// dof locations and dofs of a boundary set in the reference's order (BCValues); prescribe = add the dofs to the handler
static int collect_bc(fb2_ch* ch, int field, int kind, int64_t n, const int64_t* entities, int ncomponents, const int* components,
                      bool prescribe, DirichletBC& bc) {
    FB2_CHECK(ch && (entities || n == 0), FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: null argument");
    FB2_CHECK(!ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: the ConstraintHandler is already closed");
    fb2_dh* dh = ch->dh;
    fb2_grid* g = dh->grid;
    FB2_CHECK(field >= 0 && field < (int)dh->fields.size(), FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: bad field index %d", field);
    FB2_CHECK(kind >= FB2_BC_FACET && kind <= FB2_BC_NODE, FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: bad entity kind %d", kind);
    const LagrangeInfo& ip = dh->ips[field];
    const int ncomp = dh->fields[field].vdim;
    bc.field = field;
    bc.kind = kind;
    if (ncomponents == 0) for (int c = 1; c <= ncomp; ++c) bc.comps.push_back(c);
    else {
        FB2_CHECK(components, FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: components is null");
        for (int k = 0; k < ncomponents; ++k) {
            FB2_CHECK(components[k] >= 1 && components[k] <= ncomp, FB2_ERR_BAD_ARG, "components not within range of the field (%d dimension(s))", ncomp);
            FB2_CHECK(k == 0 || components[k] > components[k - 1], FB2_ERR_BAD_ARG, "components must be sorted");
            bc.comps.push_back(components[k]);
        }
    }
    const int nc = (int)bc.comps.size();
    const int offset = dh->field_offset(field);
    const int sdim = g->sdim, nnpc = g->nnpc, ndpc = dh->ndpc;
    if (kind == FB2_BC_NODE) {
        bc.entities.assign(entities, entities + n);
        // first cell (ascending) containing a node defines its dofs: _add! for node sets :446-470
        std::vector<int32_t> node_dofs((size_t)nc * g->nnodes, -1);
        const int npts = std::min(ip.nbase, nnpc);
        for (int64_t ci = 0; ci < g->ncells; ++ci)
            for (int idx = 0; idx < npts; ++idx) {
                int64_t node = g->cells[(size_t)ci * nnpc + idx] - 1;
                if (node_dofs[(size_t)node * nc] >= 0) continue;
                for (int i = 0; i < nc; ++i)
                    node_dofs[(size_t)node * nc + i] = dh->cell_dofs[(size_t)ci * ndpc + offset + idx * ncomp + bc.comps[i] - 1];
            }
        for (int64_t k = 0; k < n; ++k) {
            int64_t node = entities[k] - 1;
            FB2_CHECK(node >= 0 && node < g->nnodes, FB2_ERR_BAD_ARG, "node id %lld out of range", (long long)entities[k]);
            if (node_dofs[(size_t)node * nc] < 0) continue;
            for (int d = 0; d < sdim; ++d) bc.points.push_back(g->xyz[(size_t)node * sdim + d]);
            for (int i = 0; i < nc; ++i) {
                int64_t dof = node_dofs[(size_t)node * nc + i];
                bc.point_dofs.push_back(dof);
                if (prescribe) add_prescribed(ch, dof);
            }
        }
    } else {
        bc.entities.assign(entities, entities + 2 * n);
        std::vector<std::vector<int>> table = fb2_boundarydof_indices(ip, kind);
        FB2_CHECK(!table.empty(), FB2_ERR_BAD_ARG, "cell type has no boundary entities of kind %d", kind);
        LagrangeInfo geo;
        fb2_lagrange(g->celltype, 1, &geo);
        // geometric shape values at the dof locations of each entity (BCValues)
        std::vector<std::vector<double>> Mtab(table.size());
        for (size_t e = 0; e < table.size(); ++e)
            for (int fdof : table[e]) {
                double N[27], dN[81];
                fb2_lagrange_eval(geo, ip.refcoords[fdof], N, dN);
                for (int i = 0; i < geo.nbase; ++i) Mtab[e].push_back(N[i]);
            }
        // add the dofs in the reference's order: all entities first (add!), points in update! order
        for (int64_t k = 0; k < n; ++k) {
            int64_t cell = entities[2 * k] - 1, ent = entities[2 * k + 1] - 1;
            FB2_CHECK(cell >= 0 && cell < g->ncells && ent >= 0 && ent < (int64_t)table.size(), FB2_ERR_BAD_ARG,
                      "boundary entity (%lld, %lld) out of range", (long long)entities[2 * k], (long long)entities[2 * k + 1]);
            const int32_t* cd = &dh->cell_dofs[(size_t)cell * ndpc];
            const std::vector<int>& loc = table[ent];
            for (size_t t = 0; t < loc.size(); ++t) {
                double x[3] = {0, 0, 0};
                for (int i = 0; i < geo.nbase; ++i) {
                    int64_t node = g->cells[(size_t)cell * nnpc + i] - 1;
                    for (int d = 0; d < sdim; ++d) x[d] += Mtab[ent][t * geo.nbase + i] * g->xyz[(size_t)node * sdim + d];
                }
                for (int d = 0; d < sdim; ++d) bc.points.push_back(x[d]);
                for (int i = 0; i < nc; ++i) {
                    int64_t dof = cd[offset + loc[t] * ncomp + bc.comps[i] - 1];
                    bc.point_dofs.push_back(dof);
                    if (prescribe) add_prescribed(ch, dof);
                }
            }
        }
    }
    return FB2_OK;
}

extern "C" int fb2_ch_add_dirichlet(fb2_ch* ch, int field, int kind, int64_t n, const int64_t* entities, int ncomponents,
                                    const int* components, int* ibc) {
    DirichletBC bc;
    FB2_TRY(collect_bc(ch, field, kind, n, entities, ncomponents, components, true, bc));
    ch->bcs.push_back(std::move(bc));
    if (ibc) *ibc = (int)ch->bcs.size() - 1;
    return FB2_OK;
}

// add!(ch, PeriodicDirichlet(field, collect_periodic_facets(grid, mirror_set, image_set), components)) for facet sets that
// are translates of each other (src/Dofs/ConstraintHandler.jl:1032-1300): "degrees-of-freedom on the mirror facet are
// constrained to the corresponding degrees-of-freedom on the image facet" (:1046-1047) -- every dof on the MIRROR facets is
// tied to the dof at the same position (up to the translation between the two sets) on the IMAGE facets, u_mirror = u_image,
// as an affine constraint.  Chains over several periodic directions are unwound (a constrained dof never becomes a master),
// the first constraint of a dof wins; Dirichlet conditions added afterwards replace the periodic constraint of their dofs.
extern "C" int fb2_ch_add_periodic(fb2_ch* ch, int field, int64_t n_mirror, const int64_t* mirror_pairs, int64_t n_image,
                                   const int64_t* image_pairs, int ncomponents, const int* components) {
    // below: `bi` = the set whose dofs get constrained (the mirror facets), `bm` = the set that supplies the masters (the
    // image facets)
    DirichletBC bm, bi;
    FB2_TRY(collect_bc(ch, field, FB2_BC_FACET, n_image, image_pairs, ncomponents, components, false, bm));
    FB2_TRY(collect_bc(ch, field, FB2_BC_FACET, n_mirror, mirror_pairs, ncomponents, components, false, bi));
    const int sdim = ch->dh->grid->sdim;
    const size_t nc = bm.comps.size(), nm = bm.points.size() / sdim, ni = bi.points.size() / sdim;
    FB2_CHECK(nm > 0 && ni > 0, FB2_ERR_BAD_ARG, "fb2_ch_add_periodic: empty facet set");
    double cm[3] = {0, 0, 0}, ci[3] = {0, 0, 0}, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t k = 0; k < nm; ++k)
        for (int d = 0; d < sdim; ++d) { cm[d] += bm.points[k * sdim + d] / nm; lo[d] = std::min(lo[d], bm.points[k * sdim + d]); hi[d] = std::max(hi[d], bm.points[k * sdim + d]); }
    for (size_t k = 0; k < ni; ++k)
        for (int d = 0; d < sdim; ++d) ci[d] += bi.points[k * sdim + d] / ni;
    double diam = 0.0;
    for (int d = 0; d < sdim; ++d) diam = std::max(diam, hi[d] - lo[d]);
    const double tol = 1e-9 * std::max(diam, 1e-300);
    // mirror points sorted by their first coordinate; an image point (shifted back) is looked up within the tolerance
    std::vector<size_t> order(nm);
    for (size_t k = 0; k < nm; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bm.points[a * sdim] < bm.points[b * sdim]; });
    std::vector<double> key(nm);
    for (size_t k = 0; k < nm; ++k) key[k] = bm.points[order[k] * sdim];
    for (size_t k = 0; k < ni; ++k) {
        double x[3];
        for (int d = 0; d < sdim; ++d) x[d] = bi.points[k * sdim + d] - (ci[d] - cm[d]);
        size_t j = std::lower_bound(key.begin(), key.end(), x[0] - tol) - key.begin();
        int64_t hit = -1;
        for (; j < nm && key[j] <= x[0] + tol; ++j) {
            bool same = true;
            for (int d = 1; d < sdim; ++d) same = same && std::fabs(bm.points[order[j] * sdim + d] - x[d]) <= tol;
            if (same) { hit = (int64_t)order[j]; break; }
        }
        FB2_CHECK(hit >= 0, FB2_ERR_BAD_ARG, "fb2_ch_add_periodic: mirror point %zu has no counterpart on the image set (the sets must be translates)", k);
        for (size_t c = 0; c < nc; ++c) {
            const int64_t di = bi.point_dofs[k * nc + c];
            int64_t dm = bm.point_dofs[(size_t)hit * nc + c];
            // unwind chains: follow the mirror while it is itself the image of an earlier periodic constraint
            for (int guard = 0; guard < 8; ++guard) {
                auto it = ch->affine.find(dm);
                if (it == ch->affine.end() || it->second.masters.size() != 1 || it->second.coefs[0] != 1.0 || it->second.b != 0.0) break;
                dm = it->second.masters[0];
            }
            if (di == dm || ch->dofmap[di] >= 0) continue;      // same dof, or already constrained: the first constraint wins
            // earlier images that were tied to `di` now follow its mirror
            for (auto& kv : ch->affine)
                for (int64_t& m : kv.second.masters)
                    if (m == di) m = dm;
            fb2_ch::Affine a;
            a.masters.push_back(dm);
            a.coefs.push_back(1.0);
            add_prescribed(ch, di);
            ch->affine[di] = std::move(a);
        }
    }
    return FB2_OK;
}

extern "C" int fb2_ch_close(fb2_ch* ch) {
    FB2_CHECK(ch, FB2_ERR_BAD_ARG, "fb2_ch_close: null handle");
    FB2_CHECK(!ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_close: already closed");
    ch->prescribed = ch->insertion;
    std::sort(ch->prescribed.begin(), ch->prescribed.end());
    for (size_t i = 0; i < ch->prescribed.size(); ++i) ch->dofmap[ch->prescribed[i]] = (int32_t)i;
    ch->inhom.assign(ch->prescribed.size(), std::nan(""));
    FB2_TRY(build_affine_tables(ch));
    set_affine_inhomogeneities(ch);
    return finish_close(ch);
}

// add!(ch, AffineConstraint(dof, [master => coeff, ...], b)): src/Dofs/ConstraintHandler.jl:114-131, 383-401.  dof / masters
// are 1-based; n = 0 prescribes u_dof = b (equivalent to a Dirichlet value)
extern "C" int fb2_ch_add_affine(fb2_ch* ch, int64_t dof, int n, const int64_t* masters, const double* coefs, double b) {
    FB2_CHECK(ch && (n == 0 || (masters && coefs)), FB2_ERR_BAD_ARG, "fb2_ch_add_affine: null argument");
    FB2_CHECK(!ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_add_affine: the ConstraintHandler is already closed");
    const int64_t nd = ch->dh->ndofs;
    FB2_CHECK(dof >= 1 && dof <= nd, FB2_ERR_BAD_ARG, "fb2_ch_add_affine: dof %lld outside 1..%lld", (long long)dof, (long long)nd);
    fb2_ch::Affine a;
    a.b = b;
    for (int k = 0; k < n; ++k) {
        FB2_CHECK(masters[k] >= 1 && masters[k] <= nd && masters[k] != dof, FB2_ERR_BAD_ARG, "fb2_ch_add_affine: bad master dof %lld", (long long)masters[k]);
        a.masters.push_back(masters[k] - 1);
        a.coefs.push_back(coefs[k]);
    }
    add_prescribed(ch, dof - 1);
    ch->affine[dof - 1] = std::move(a);
    return FB2_OK;
}

// dofcoefficients of the closed handler, aligned with prescribed_dofs: ptr (nprescribed + 1, 0-based offsets), masters
// (1-based), coefficients; pass NULL pointers to query the total count
extern "C" int fb2_ch_affine_export(fb2_ch* ch, int64_t* ntotal, int64_t* ptr, int64_t* masters, double* coefs) {
    FB2_CHECK(ch && ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_affine_export: handle is null or not closed");
    if (ntotal) *ntotal = (int64_t)ch->aff_dof.size();
    if (ptr) for (size_t i = 0; i < ch->aff_ptr.size(); ++i) ptr[i] = ch->aff_ptr[i];
    if (masters) for (size_t i = 0; i < ch->aff_dof.size(); ++i) masters[i] = (int64_t)ch->aff_dof[i] + 1;
    if (coefs && !ch->aff_coef.empty()) memcpy(coefs, ch->aff_coef.data(), ch->aff_coef.size() * sizeof(double));
    return FB2_OK;
}

extern "C" int fb2_ch_from_host(fb2_dh* dh, int64_t n, const int64_t* prescribed_dofs, const double* inhomogeneities, fb2_ch** out) {
    FB2_CHECK(dh && out && (n == 0 || (prescribed_dofs && inhomogeneities)), FB2_ERR_BAD_ARG, "fb2_ch_from_host: null argument");
    fb2_ch* ch = new fb2_ch();
    ch->dh = dh;
    ch->dofmap.assign((size_t)dh->ndofs, -1);
    for (int64_t i = 0; i < n; ++i) {
        int64_t d = prescribed_dofs[i] - 1;
        if (d < 0 || d >= dh->ndofs || (i > 0 && prescribed_dofs[i] <= prescribed_dofs[i - 1])) {
            delete ch;
            return fb2_fail(FB2_ERR_BAD_ARG, "fb2_ch_from_host: prescribed_dofs must be sorted, unique and within 1..ndofs");
        }
        ch->prescribed.push_back(d);
        ch->dofmap[d] = (int32_t)i;
        ch->inhom.push_back(inhomogeneities[i]);
    }
    ch->insertion = ch->prescribed;
    ch->aff_ptr.assign((size_t)n + 1, 0);
    ch->aff_is.assign((size_t)n, 0);
    int rc = finish_close(ch);
    if (rc != FB2_OK) { fb2_ch_destroy(ch); return rc; }
    *out = ch;
    return FB2_OK;
}

// renumber!(dh, ch, perm), the ConstraintHandler half (src/Dofs/DofRenumbering.jl:92-125): prescribed dofs are mapped through
// perm (1-based, dof i -> perm[i]) and re-sorted together with their inhomogeneities
extern "C" int fb2_ch_renumber(fb2_ch* ch, const int64_t* perm) {
    FB2_CHECK(ch && perm, FB2_ERR_BAD_ARG, "fb2_ch_renumber: null argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_renumber: close the ConstraintHandler first");
    FB2_CHECK(ch->affine.empty(), FB2_ERR_UNSUPPORTED, "fb2_ch_renumber: renumber before adding affine constraints");
    const int64_t n = ch->dh->ndofs;
    const size_t np = ch->prescribed.size();
    std::vector<std::pair<int64_t, double>> pv(np);
    for (size_t i = 0; i < np; ++i) {
        const int64_t v = perm[ch->prescribed[i]] - 1;
        FB2_CHECK(v >= 0 && v < n, FB2_ERR_BAD_ARG, "fb2_ch_renumber: permutation entry out of range");
        pv[i] = {v, ch->inhom[i]};
    }
    std::sort(pv.begin(), pv.end());
    for (int64_t& d : ch->insertion) d = perm[d] - 1;
    for (DirichletBC& bc : ch->bcs)
        for (int64_t& d : bc.point_dofs) d = perm[d] - 1;
    ch->dofmap.assign((size_t)n, -1);
    for (size_t i = 0; i < np; ++i) {
        ch->prescribed[i] = pv[i].first;
        ch->inhom[i] = pv[i].second;
        ch->dofmap[pv[i].first] = (int32_t)i;
    }
    if (ch->d_prescribed) { cudaFree(ch->d_prescribed); ch->d_prescribed = nullptr; }
    if (ch->d_inhom) { cudaFree(ch->d_inhom); ch->d_inhom = nullptr; }
    if (ch->d_isconstrained) { cudaFree(ch->d_isconstrained); ch->d_isconstrained = nullptr; }
    if (ch->d_scratch) { cudaFree(ch->d_scratch); ch->d_scratch = nullptr; }
    return finish_close(ch);
}

// arrays-in mode of update!(ch, t): the reference's ch.inhomogeneities (order of ch.prescribed_dofs) after its own update!
extern "C" int fb2_ch_set_inhomogeneities(fb2_ch* ch, int64_t n, const double* inhomogeneities) {
    FB2_CHECK(ch && inhomogeneities, FB2_ERR_BAD_ARG, "fb2_ch_set_inhomogeneities: null argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_set_inhomogeneities: the ConstraintHandler is not closed");
    FB2_CHECK(n == (int64_t)ch->prescribed.size(), FB2_ERR_BAD_ARG, "fb2_ch_set_inhomogeneities: %lld values for %lld prescribed dofs",
              (long long)n, (long long)ch->prescribed.size());
    ch->inhom.assign(inhomogeneities, inhomogeneities + n);
    ch->inhom_dirty = true;
    return FB2_OK;
}

extern "C" int fb2_ch_bc_points(fb2_ch* ch, int ibc, int64_t* npoints, double* x) {
    FB2_CHECK(ch && npoints && ibc >= 0 && ibc < (int)ch->bcs.size(), FB2_ERR_BAD_ARG, "fb2_ch_bc_points: bad argument");
    const DirichletBC& bc = ch->bcs[ibc];
    *npoints = (int64_t)bc.points.size() / ch->dh->grid->sdim;
    if (x) memcpy(x, bc.points.data(), bc.points.size() * sizeof(double));
    return FB2_OK;
}

extern "C" int fb2_ch_bc_set_values(fb2_ch* ch, int ibc, int64_t npoints, const double* values) {
    FB2_CHECK(ch && values && ibc >= 0 && ibc < (int)ch->bcs.size(), FB2_ERR_BAD_ARG, "fb2_ch_bc_set_values: bad argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_bc_set_values: close the ConstraintHandler first");
    const DirichletBC& bc = ch->bcs[ibc];
    const size_t nc = bc.comps.size();
    FB2_CHECK((size_t)npoints * nc == bc.point_dofs.size(), FB2_ERR_BAD_ARG, "fb2_ch_bc_set_values: expected %zu points", bc.point_dofs.size() / nc);
    for (size_t k = 0; k < bc.point_dofs.size(); ++k)
        if (!ch->affine.count(bc.point_dofs[k])) ch->inhom[ch->dofmap[bc.point_dofs[k]]] = values[k];   // a later affine constraint won the dof
    ch->inhom_dirty = true;
    return FB2_OK;
}

// inhomogeneities of the current update! on the device (used by element_assembly.cu)
int fb2_ch_sync_device(fb2_ch* ch) { return upload_inhom(ch); }

extern "C" int fb2_ch_info(fb2_ch* ch, int64_t* nprescribed) {
    FB2_CHECK(ch, FB2_ERR_BAD_ARG, "fb2_ch_info: null handle");
    if (nprescribed) *nprescribed = (int64_t)(ch->closed ? ch->prescribed.size() : ch->insertion.size());
    return FB2_OK;
}

extern "C" int fb2_ch_export(fb2_ch* ch, int64_t* prescribed_dofs, double* inhomogeneities) {
    FB2_CHECK(ch && ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_export: handle is null or not closed");
    if (prescribed_dofs) for (size_t i = 0; i < ch->prescribed.size(); ++i) prescribed_dofs[i] = ch->prescribed[i] + 1;
    if (inhomogeneities) memcpy(inhomogeneities, ch->inhom.data(), ch->inhom.size() * sizeof(double));
    return FB2_OK;
}

extern "C" int fb2_apply(fb2_ch* ch, fb2_pattern* p, double* nzval_dev, double* f_dev, int applyzero, double* meandiag) {
    FB2_CHECK(ch && p && nzval_dev, FB2_ERR_BAD_ARG, "fb2_apply: null argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_apply: the ConstraintHandler is not closed");
    FB2_CHECK(p->n == ch->dh->ndofs, FB2_ERR_BAD_ARG, "fb2_apply: matrix size does not match the DofHandler");
    fb2_ctx* ctx = ch->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(upload_inhom(ch));
    const int64_t np = (int64_t)ch->prescribed.size();
    const int nb = (int)std::min<int64_t>(1024, (p->n + 255) / 256);
    k_meandiag_partial<<<nb, 256, 0, st>>>(nzval_dev, p->d_diag, p->n, ch->d_scratch);
    k_meandiag_final<<<1, 256, 0, st>>>(ch->d_scratch, nb, p->n);
    ctx->launches += 2;
    if (np > 0 && ch->has_affine) {
        // apply! with affine constraints (:710-740): f -= K g, then `_condense!` while the constrained rows / columns still hold
        // their values, then zero them
        if (!p->structurally_symmetric) {   // an adopted pattern (fb2_pattern_from_host): check once
            int* d_cnt = nullptr;
            int cnt = 0;
            FB2_CUDA(cudaMalloc(&d_cnt, sizeof(int)));
            cudaMemsetAsync(d_cnt, 0, sizeof(int), st);
            k_count_asymmetric<<<nblocks(p->n * 32, 256), 256, 0, st>>>(p->d_colptr, p->d_rowval, p->n, d_cnt);
            cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st);
            cudaError_t e = cudaStreamSynchronize(st);
            cudaFree(d_cnt);
            FB2_CHECK(e == cudaSuccess, FB2_ERR_CUDA, "fb2_apply: %s", cudaGetErrorString(e));
            ctx->launches++;
            if (cnt == 0) p->structurally_symmetric = true;
        }
        FB2_CHECK(p->structurally_symmetric, FB2_ERR_UNSUPPORTED, "fb2_apply: affine constraints need a structurally symmetric pattern (allocate_matrix(dh, ch))");
        if (f_dev && !applyzero) k_add_inhomogeneities<<<nblocks(np * 32, 256), 256, 0, st>>>(ch->d_prescribed, ch->d_inhom, np, p->d_colptr, p->d_rowval, nzval_dev, f_dev);
        k_condense<<<nblocks(ch->n_aff * 32, 256), 256, 0, st>>>(ch->d_aff_list, ch->n_aff, ch->d_prescribed, ch->d_aff_ptr, ch->d_aff_dof, ch->d_aff_coef,
                                                                ch->d_aff_of, p->d_colptr, p->d_rowval, nzval_dev, ctx->d_errflag);
        if (f_dev) k_condense_rhs<<<nblocks(ch->n_aff, 256), 256, 0, st>>>(ch->d_aff_list, ch->n_aff, ch->d_prescribed, ch->d_aff_ptr, ch->d_aff_dof, ch->d_aff_coef, f_dev);
        k_zero_columns_of<<<nblocks(np * 32, 256), 256, 0, st>>>(ch->d_prescribed, np, p->d_colptr, nzval_dev);
        k_apply_rows_sym<<<nblocks(np * 32, 256), 256, 0, st>>>(ch->d_prescribed, np, p->d_colptr, p->d_rowval, nzval_dev);
        k_apply_diag<<<nblocks(np, 256), 256, 0, st>>>(ch->d_prescribed, ch->d_inhom, np, p->d_diag, ch->d_scratch + nb, nzval_dev, f_dev, applyzero);
        ctx->launches += 6;
        FB2_CUDA(cudaGetLastError());
        FB2_TRY(fb2_check_device_error(ctx));
    } else if (np > 0) {
        k_apply_columns<<<nblocks(np * 32, 256), 256, 0, st>>>(ch->d_prescribed, ch->d_inhom, np, p->d_colptr, p->d_rowval,
                                                              nzval_dev, f_dev, applyzero);
        if (p->structurally_symmetric)
            k_apply_rows_sym<<<nblocks(np * 32, 256), 256, 0, st>>>(ch->d_prescribed, np, p->d_colptr, p->d_rowval, nzval_dev);
        else
            k_apply_rows_stream<<<(unsigned)std::min<int64_t>((p->nnz + 255) / 256, (int64_t)ctx->sm_count * 16), 256, 0, st>>>(
                ch->d_isconstrained, p->d_rowval, p->nnz, nzval_dev);
        k_apply_diag<<<nblocks(np, 256), 256, 0, st>>>(ch->d_prescribed, ch->d_inhom, np, p->d_diag, ch->d_scratch + nb, nzval_dev,
                                                      f_dev, applyzero);
        ctx->launches += 3;
    }
    FB2_CUDA(cudaGetLastError());
    if (meandiag) {
        FB2_CUDA(cudaMemcpyAsync(meandiag, ch->d_scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
        FB2_CUDA(cudaStreamSynchronize(st));
    }
    return FB2_OK;
}

struct fb2_rhsdata {
    fb2_ctx* ctx = nullptr;
    int64_t n = 0, np = 0;
    double m = 0.0;              // meandiag(A)
    int64_t* d_ptr = nullptr;    // [np + 1]
    int32_t* d_rows = nullptr;   // A[:, prescribed_dofs] as compact CSC
    double* d_vals = nullptr;
};

extern "C" int fb2_rhsdata_destroy(fb2_rhsdata* r) {
    if (!r) return FB2_OK;
    cudaFree(r->d_ptr);
    cudaFree(r->d_rows);
    cudaFree(r->d_vals);
    delete r;
    return FB2_OK;
}

// get_rhs_data(ch, A): mean |diagonal| and the prescribed columns of the matrix as it is now (before apply!)
extern "C" int fb2_rhsdata_create(fb2_ch* ch, fb2_pattern* p, const double* nzval_dev, fb2_rhsdata** out) {
    FB2_CHECK(ch && p && nzval_dev && out, FB2_ERR_BAD_ARG, "fb2_rhsdata_create: null argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_rhsdata_create: the ConstraintHandler is not closed");
    FB2_CHECK(p->n == ch->dh->ndofs, FB2_ERR_BAD_ARG, "fb2_rhsdata_create: matrix size does not match the DofHandler");
    fb2_ctx* ctx = ch->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    FB2_CUDA(cudaSetDevice(ctx->device));
    fb2_rhsdata* r = new fb2_rhsdata();
    r->ctx = ctx;
    r->n = p->n;
    r->np = (int64_t)ch->prescribed.size();
    const int nb = (int)std::min<int64_t>(1024, (p->n + 255) / 256);
    k_meandiag_partial<<<nb, 256, 0, st>>>(nzval_dev, p->d_diag, p->n, ch->d_scratch);
    k_meandiag_final<<<1, 256, 0, st>>>(ch->d_scratch, nb, p->n);
    ctx->launches += 2;
    cudaError_t e = cudaMemcpyAsync(&r->m, ch->d_scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, st);
    std::vector<int64_t> ptr((size_t)r->np + 1, 0);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_ptr, ptr.size() * sizeof(int64_t));
    if (e == cudaSuccess && r->np > 0) {
        k_rhs_lengths<<<nblocks(r->np, 256), 256, 0, st>>>(ch->d_prescribed, r->np, p->d_colptr, r->d_ptr);
        ctx->launches++;
        e = cudaMemcpyAsync(ptr.data() + 1, r->d_ptr, (size_t)r->np * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    for (int64_t i = 0; i < r->np; ++i) ptr[i + 1] += ptr[i];
    const int64_t tot = ptr[r->np];
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_ptr, ptr.data(), ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMalloc(&r->d_rows, std::max<int64_t>(tot, 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&r->d_vals, std::max<int64_t>(tot, 1) * sizeof(double));
    if (e == cudaSuccess && r->np > 0) {
        k_rhs_copy<<<nblocks(r->np * 32, 256), 256, 0, st>>>(ch->d_prescribed, r->np, p->d_colptr, p->d_rowval, nzval_dev, r->d_ptr,
                                                            r->d_rows, r->d_vals);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // ptr (host) is read by the copy above
    if (e != cudaSuccess) {
        fb2_rhsdata_destroy(r);
        return fb2_fail(e == cudaErrorMemoryAllocation ? FB2_ERR_OOM : FB2_ERR_CUDA, "fb2_rhsdata_create: %s", cudaGetErrorString(e));
    }
    *out = r;
    return FB2_OK;
}

extern "C" int fb2_rhsdata_info(fb2_rhsdata* r, double* meandiag, int64_t* nprescribed, int64_t* nstored) {
    FB2_CHECK(r, FB2_ERR_BAD_ARG, "fb2_rhsdata_info: null handle");
    if (meandiag) *meandiag = r->m;
    if (nprescribed) *nprescribed = r->np;
    if (nstored) {
        FB2_CUDA(cudaMemcpy(nstored, r->d_ptr + r->np, sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    return FB2_OK;
}

// apply_rhs!(data, f, ch, applyzero): the boundary conditions of the current update! on a new right-hand side, K untouched
extern "C" int fb2_apply_rhs(fb2_rhsdata* r, double* f_dev, fb2_ch* ch, int applyzero) {
    FB2_CHECK(r && f_dev && ch, FB2_ERR_BAD_ARG, "fb2_apply_rhs: null argument");
    FB2_CHECK(ch->closed && (int64_t)ch->prescribed.size() == r->np && ch->dh->ndofs == r->n, FB2_ERR_BAD_ARG,
              "fb2_apply_rhs: the ConstraintHandler does not match the RHSData");
    fb2_ctx* ctx = ch->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(upload_inhom(ch));
    if (r->np == 0) return FB2_OK;
    if (!applyzero) {
        k_rhs_columns<<<nblocks(r->np * 32, 256), 256, 0, ctx->stream>>>(ch->d_inhom, r->np, r->d_ptr, r->d_rows, r->d_vals, f_dev);
        ctx->launches++;
    }
    k_rhs_prescribed<<<nblocks(r->np, 256), 256, 0, ctx->stream>>>(ch->d_prescribed, ch->d_inhom, r->np, r->m, f_dev, applyzero);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_apply_vector(fb2_ch* ch, double* u_dev, int applyzero) {
    FB2_CHECK(ch && u_dev, FB2_ERR_BAD_ARG, "fb2_apply_vector: null argument");
    FB2_CHECK(ch->closed, FB2_ERR_BAD_ARG, "fb2_apply_vector: the ConstraintHandler is not closed");
    fb2_ctx* ctx = ch->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(upload_inhom(ch));
    const int64_t np = (int64_t)ch->prescribed.size();
    if (np > 0 && ch->has_affine) {
        k_apply_vector_affine<<<nblocks(np, 256), 256, 0, ctx->stream>>>(ch->d_prescribed, ch->d_inhom, np, ch->d_aff_ptr, ch->d_aff_dof, ch->d_aff_coef, u_dev, applyzero);
        ctx->launches++;
        FB2_CUDA(cudaGetLastError());
    } else if (np > 0) {
        k_apply_vector<<<nblocks(np, 256), 256, 0, ctx->stream>>>(ch->d_prescribed, ch->d_inhom, np, u_dev, applyzero);
        ctx->launches++;
        FB2_CUDA(cudaGetLastError());
    }
    return FB2_OK;
}

extern "C" int fb2_ch_destroy(fb2_ch* ch) {
    if (!ch) return FB2_OK;
    if (ch->dh->grid->ctx->device >= 0) {
        cudaSetDevice(ch->dh->grid->ctx->device);
        cudaFree(ch->d_prescribed);
        cudaFree(ch->d_inhom);
        cudaFree(ch->d_isconstrained);
        cudaFree(ch->d_scratch);
        cudaFree(ch->d_aff_ptr); cudaFree(ch->d_aff_dof); cudaFree(ch->d_aff_coef); cudaFree(ch->d_aff_of); cudaFree(ch->d_aff_list);
    }
    delete ch;
    return FB2_OK;
}
