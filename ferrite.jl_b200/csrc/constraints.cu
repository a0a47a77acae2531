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

inline unsigned nblocks(int64_t total, int bs) { return (unsigned)((total + bs - 1) / bs); }

void add_prescribed(fb2_ch* ch, int64_t dof) {
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
    ch->closed = true;
    ch->inhom_dirty = true;
    return FB2_OK;
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

extern "C" int fb2_ch_add_dirichlet(fb2_ch* ch, int field, int kind, int64_t n, const int64_t* entities, int ncomponents,
                                    const int* components, int* ibc) {
    FB2_CHECK(ch && (entities || n == 0), FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: null argument");
    FB2_CHECK(!ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: the ConstraintHandler is already closed");
    fb2_dh* dh = ch->dh;
    fb2_grid* g = dh->grid;
    FB2_CHECK(field >= 0 && field < (int)dh->fields.size(), FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: bad field index %d", field);
    FB2_CHECK(kind >= FB2_BC_FACET && kind <= FB2_BC_NODE, FB2_ERR_BAD_ARG, "fb2_ch_add_dirichlet: bad entity kind %d", kind);
    const LagrangeInfo& ip = dh->ips[field];
    const int ncomp = dh->fields[field].vdim;
    DirichletBC bc;
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
                add_prescribed(ch, dof);
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
                    add_prescribed(ch, dof);
                }
            }
        }
    }
    ch->bcs.push_back(std::move(bc));
    if (ibc) *ibc = (int)ch->bcs.size() - 1;
    return FB2_OK;
}

extern "C" int fb2_ch_close(fb2_ch* ch) {
    FB2_CHECK(ch, FB2_ERR_BAD_ARG, "fb2_ch_close: null handle");
    FB2_CHECK(!ch->closed, FB2_ERR_BAD_ARG, "fb2_ch_close: already closed");
    ch->prescribed = ch->insertion;
    std::sort(ch->prescribed.begin(), ch->prescribed.end());
    for (size_t i = 0; i < ch->prescribed.size(); ++i) ch->dofmap[ch->prescribed[i]] = (int32_t)i;
    ch->inhom.assign(ch->prescribed.size(), std::nan(""));
    return finish_close(ch);
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
    for (size_t k = 0; k < bc.point_dofs.size(); ++k) ch->inhom[ch->dofmap[bc.point_dofs[k]]] = values[k];
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
    if (np > 0) {
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
    if (np > 0) {
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
    }
    delete ch;
    return FB2_OK;
}
