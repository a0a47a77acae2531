// Partitioned multi-GPU assembly.  New capability: the reference is single-process; the
// parity oracle for this path is the reference's serial global K, f.
//
// One process per GPU.  Every rank knows the global grid + DofHandler (host side, reference numbering) and builds:
//   * its cells: structured blocks px*py*pz for generate_grid input, contiguous ranges otherwise;
//   * dof/column ownership: the lowest rank among the cells touching a dof owns it;
//   * a LOCAL problem = own cells + halo cells (every cell touching an owned dof), with local node / dof numbering
//     (ascending global ids), so the ordinary single-GPU pattern / map / kernels run unchanged on it and the local
//     pattern of an owned column holds all of the column's global rows;
//   * per peer, the interface exchange lists.  K is CSC, so "interface-row contributions" are exchanged by column:
//     rank r sends, for every column j owned by peer o, the partial sums K[i, j] it produced from its own cells
//     (i in the dofs of r's own cells containing j); both sides derive the identical (j, i)-sorted list from the
//     global numbering, so only values travel.
// Two strategies: FB2_DIST_EXCHANGE assembles own cells and exchanges interface columns (NCCL grouped send/recv);
// FB2_DIST_HALO assembles own + halo cells redundantly and needs no communication at all.  Afterwards every rank
// holds the final values of its owned columns / dofs; everything it does not own is zeroed.
#include <dlfcn.h>

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>

#include "common.h"

struct PeerPlan {
    std::vector<int32_t> send_rows, send_cols, recv_rows, recv_cols;  // local dof ids, sorted by global (col, row)
    std::vector<int32_t> send_f, recv_f;                              // local dof ids, sorted by global id
    int64_t* d_send_pos = nullptr;
    int64_t* d_recv_pos = nullptr;
    int32_t* d_send_f = nullptr;
    int32_t* d_recv_f = nullptr;
    double* d_sendbuf = nullptr;
    double* d_recvbuf = nullptr;
};

struct fb2_part {
    fb2_dh* gdh = nullptr;
    bool owns_global = false;   // gdh (+ its grid) is the metadata-only global problem of fb2_partition_create_generated
    int nparts = 1, rank = 0;
    int dims[3] = {1, 1, 1};
    std::vector<int64_t> cells_global;   // local cells (own + halo), ascending global id (0-based)
    std::vector<uint8_t> cell_is_own;
    std::vector<int64_t> l2g_node, l2g_dof;  // 0-based global ids, ascending
    std::vector<int32_t> dof_owner;          // per local dof
    std::vector<int64_t> lcells;             // nnpc x nlocal, 1-based local node ids
    std::vector<double> lxyz;                // sdim x nlocal_nodes
    std::vector<int64_t> lcell_dofs;         // ndpc x nlocal, 1-based local dofs
    std::vector<PeerPlan> peers;
    int64_t ndofs_owned = 0, ncells_own = 0;
    // device binding
    fb2_assembler* bound = nullptr;
    int bound_device = -1;
    int32_t* d_col_owned = nullptr;   // list of local columns owned by other ranks
    int64_t n_unowned = 0;
    // local cells [0, n_iface) are the own cells touching an exchanged dof, [n_iface, ncells_own) the other own cells;
    // the interface exchange runs on a second stream while the interior cells are assembled
    int64_t n_iface = 0;
    cudaStream_t xstream = nullptr;
    cudaEvent_t ev_iface = nullptr, ev_done = nullptr;
};

namespace {

void default_dims(int nparts, const int64_t* nel, int dim, int* dims) {
    dims[0] = dims[1] = dims[2] = 1;
    int n = nparts;
    for (int p = 2; n > 1;) {
        if (n % p) { ++p; continue; }
        // give the factor to the direction with the most cells per block
        int best = 0;
        double bv = -1;
        for (int d = 0; d < dim; ++d) {
            double v = (double)nel[d] / dims[d];
            if (v > bv) { bv = v; best = d; }
        }
        dims[best] *= p;
        n /= p;
    }
}

inline int block_of(int64_t i, int64_t n, int p) {
    // balanced split: block b covers [floor(b*n/p), floor((b+1)*n/p))
    int b = (int)((i * p) / n);
    while (b + 1 < p && ((int64_t)(b + 1) * n) / p <= i) ++b;
    while (b > 0 && ((int64_t)b * n) / p > i) --b;
    return b;
}

__global__ void k_lookup_pos(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int64_t n,
                             const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t* __restrict__ pos) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int r = rows[t], c = cols[t];
    int64_t lo = colptr[c], hi = colptr[c + 1], p = -1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int rr = rowval[mid];
        if (rr == r) { p = mid; break; }
        if (rr < r) lo = mid + 1; else hi = mid;
    }
    pos[t] = p;
}

__global__ void k_pack(const int64_t* __restrict__ pos, int64_t nnz, const int32_t* __restrict__ fd, int64_t nf,
                       const double* __restrict__ nzval, const double* __restrict__ f, double* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nnz) out[t] = nzval[pos[t]];
    else if (t < nnz + nf) out[t] = f ? f[fd[t - nnz]] : 0.0;
}

__global__ void k_unpack_add(const int64_t* __restrict__ pos, int64_t nnz, const int32_t* __restrict__ fd, int64_t nf,
                             const double* __restrict__ in, double* __restrict__ nzval, double* __restrict__ f) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nnz) nzval[pos[t]] += in[t];               // positions are unique within one peer list
    else if (t < nnz + nf && f) f[fd[t - nnz]] += in[t];
}

// All peers in one launch: segment p = [nz values | f values] of the exchange with one peer.  An owned entry can receive
// partial sums from several peers, so the fused unpack adds with REDs.
struct XSeg {
    const int64_t* pos;
    const int32_t* fd;
    double* buf;
    int64_t nnz, nf, start;   // start = first thread of the segment
};
constexpr int XSEG_MAX = 64;
struct XSegs { XSeg s[XSEG_MAX]; int n; int64_t total; };

__global__ void k_pack_all(const XSegs S, const double* __restrict__ nzval, const double* __restrict__ f) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.total) return;
    int k = 0;
    while (k + 1 < S.n && t >= S.s[k + 1].start) ++k;
    const XSeg& sg = S.s[k];
    const int64_t i = t - sg.start;
    sg.buf[i] = i < sg.nnz ? nzval[sg.pos[i]] : (f ? f[sg.fd[i - sg.nnz]] : 0.0);
}

__global__ void k_unpack_all(const XSegs S, double* __restrict__ nzval, double* __restrict__ f) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.total) return;
    int k = 0;
    while (k + 1 < S.n && t >= S.s[k + 1].start) ++k;
    const XSeg& sg = S.s[k];
    const int64_t i = t - sg.start;
    const double v = sg.buf[i];
    if (i < sg.nnz) atomicAdd(nzval + sg.pos[i], v);
    else if (f) atomicAdd(f + sg.fd[i - sg.nnz], v);
}

__global__ void k_mask_unowned(const int32_t* __restrict__ unowned, int64_t n, const int64_t* __restrict__ colptr,
                               double* __restrict__ nzval, double* __restrict__ f) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int c = unowned[w];
    for (int64_t k = colptr[c] + lane; k < colptr[c + 1]; k += 32) nzval[k] = 0.0;
    if (lane == 0 && f) f[c] = 0.0;
}

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

static int partition_create_impl(fb2_dh* gdh, int nparts, int rank, const int* dims_in, const int32_t* cell_owner, fb2_part** out);

extern "C" int fb2_partition_create(fb2_dh* gdh, int nparts, int rank, const int* dims_in, fb2_part** out) {
    return partition_create_impl(gdh, nparts, rank, dims_in, nullptr, out);
}

// any partitioner's result (e.g. METIS_PartMeshDual as in ext/FerriteMetis.jl): cell_owner[c] = rank of cell c (0-based)
extern "C" int fb2_partition_create_from_owners(fb2_dh* gdh, int nparts, int rank, const int32_t* cell_owner, fb2_part** out) {
    FB2_CHECK(cell_owner, FB2_ERR_BAD_ARG, "fb2_partition_create_from_owners: null owner array");
    return partition_create_impl(gdh, nparts, rank, nullptr, cell_owner, out);
}

// METIS from the CUDA toolkit (libmetis_static.a, the copy cuSOLVER ships; built with 64-bit idx_t, verified by a probe):
// dual-graph partition of the mesh, two cells adjacent when they share a facet's worth of nodes.  Deterministic (default
// seed), so every rank derives the same cell -> rank array on its own.
extern "C" int METIS_PartMeshDual(int64_t* ne, int64_t* nn, int64_t* eptr, int64_t* eind, int64_t* vwgt, int64_t* vsize, int64_t* ncommon,
                                  int64_t* nparts, void* tpwgts, int64_t* options, int64_t* objval, int64_t* epart, int64_t* npart);
extern "C" int METIS_SetDefaultOptions(int64_t* options);

extern "C" int fb2_partition_create_metis(fb2_dh* gdh, int nparts, int rank, fb2_part** out) {
    FB2_CHECK(gdh && out, FB2_ERR_BAD_ARG, "fb2_partition_create_metis: null argument");
    FB2_CHECK(nparts >= 1 && rank >= 0 && rank < nparts, FB2_ERR_BAD_ARG, "fb2_partition_create_metis: bad nparts/rank");
    fb2_grid* g = gdh->grid;
    const int64_t nc = g->ncells;
    std::vector<int32_t> owner((size_t)nc, 0);
    if (nparts > 1) {
        const RefShapeInfo* rs = fb2_refshape(g->celltype);
        const int nv = rs->nvertices;
        int64_t ne = nc, nn = g->nnodes, ncommon = rs->rdim == 1 ? 1 : (rs->rdim == 2 ? 2 : rs->face_nverts[0]);
        int64_t np = nparts, objval = 0;
        std::vector<int64_t> eptr((size_t)nc + 1), eind((size_t)nc * nv), epart((size_t)nc), npart((size_t)nn);
        for (int64_t c = 0; c < nc; ++c) {
            eptr[c] = c * nv;
            for (int k = 0; k < nv; ++k) eind[(size_t)c * nv + k] = g->cells[(size_t)c * g->nnpc + k] - 1;   // vertices only
        }
        eptr[nc] = nc * nv;
        int64_t options[40];
        METIS_SetDefaultOptions(options);
        const int rc = METIS_PartMeshDual(&ne, &nn, eptr.data(), eind.data(), nullptr, nullptr, &ncommon, &np, nullptr, options, &objval,
                                          epart.data(), npart.data());
        FB2_CHECK(rc == 1, FB2_ERR_INTERNAL, "METIS_PartMeshDual failed with status %d", rc);
        for (int64_t c = 0; c < nc; ++c) owner[c] = (int32_t)epart[c];
    }
    return partition_create_impl(gdh, nparts, rank, nullptr, owner.data(), out);
}

static int partition_create_impl(fb2_dh* gdh, int nparts, int rank, const int* dims_in, const int32_t* cell_owner, fb2_part** out) {
    FB2_CHECK(gdh && out, FB2_ERR_BAD_ARG, "fb2_partition_create: null argument");
    FB2_CHECK(nparts >= 1 && rank >= 0 && rank < nparts, FB2_ERR_BAD_ARG, "fb2_partition_create: bad nparts/rank");
    fb2_grid* g = gdh->grid;
    const int64_t ncells = g->ncells;
    const int nnpc = g->nnpc, ndpc = gdh->ndpc, sdim = g->sdim;
    fb2_part* P = new fb2_part();
    P->gdh = gdh;
    P->nparts = nparts;
    P->rank = rank;
    // ---- cell -> rank --------------------------------------------------------------------------------------
    std::vector<int32_t> owner((size_t)ncells);
    if (cell_owner) {
        P->dims[0] = nparts;
        for (int64_t c = 0; c < ncells; ++c) {
            if (cell_owner[c] < 0 || cell_owner[c] >= nparts) {
                delete P;
                return fb2_fail(FB2_ERR_BAD_ARG, "fb2_partition_create_from_owners: owner %d of cell %lld is outside 0..%d", cell_owner[c], (long long)c + 1, nparts - 1);
            }
            owner[c] = cell_owner[c];
        }
    } else if (g->generated && g->celltype != FB2_LINE) {
        const int dim = g->sdim;
        if (dims_in) { for (int d = 0; d < 3; ++d) P->dims[d] = d < dim ? dims_in[d] : 1; }
        else default_dims(nparts, g->nel, dim, P->dims);
        if ((int64_t)P->dims[0] * P->dims[1] * P->dims[2] != nparts) {
            delete P;
            return fb2_fail(FB2_ERR_BAD_ARG, "fb2_partition_create: block layout does not multiply to nparts");
        }
        const int64_t nx = g->nel[0], ny = g->nel[1], nz = dim > 2 ? g->nel[2] : 1;
        const int per = (g->celltype == FB2_TRIANGLE) ? 2 : (g->celltype == FB2_TETRAHEDRON ? 6 : 1);
        // block of every cube, swept in storage order (no division per cell)
        std::vector<int> bx((size_t)nx), by((size_t)ny), bz((size_t)nz);
        for (int64_t i = 0; i < nx; ++i) bx[i] = block_of(i, nx, P->dims[0]);
        for (int64_t j = 0; j < ny; ++j) by[j] = block_of(j, ny, P->dims[1]);
        for (int64_t k = 0; k < nz; ++k) bz[k] = dim > 2 ? block_of(k, nz, P->dims[2]) : 0;
        const int nth = fb2_host_threads();
#pragma omp parallel for collapse(2) schedule(static) num_threads(nth)
        for (int64_t k = 0; k < nz; ++k)
            for (int64_t j = 0; j < ny; ++j) {
                const int bjk = P->dims[0] * (by[j] + P->dims[1] * bz[k]);
                int32_t* o = &owner[(size_t)(nx * (j + ny * k)) * per];
                for (int64_t i = 0; i < nx; ++i)
                    for (int t = 0; t < per; ++t) o[i * per + t] = bx[i] + bjk;
            }
    } else {
        P->dims[0] = nparts;
        for (int64_t c = 0; c < ncells; ++c) owner[c] = block_of(c, ncells, nparts);
    }
    // ---- dof -> rank: lowest rank among the cells touching the dof ---------------------------------------------
    std::vector<int32_t> gdof_owner((size_t)gdh->ndofs, nparts), gdof_maxowner((size_t)gdh->ndofs, -1);
    for (int64_t c = 0; c < ncells; ++c) {
        const int32_t* cd = &gdh->cell_dofs[(size_t)c * ndpc];
        const int32_t o = owner[c];
        for (int i = 0; i < ndpc; ++i) {
            if (o < gdof_owner[cd[i]]) gdof_owner[cd[i]] = o;
            if (o > gdof_maxowner[cd[i]]) gdof_maxowner[cd[i]] = o;
        }
    }
    // ---- local cells: own + every cell touching an owned dof ------------------------------------------------------
    // Order: [own cells touching an exchanged dof | other own cells | halo cells], each by ascending global id.  A dof is
    // exchanged when cells of more than one rank touch it.  Contiguous ranges let the kernels run without an index list,
    // and the interface cells can be assembled first so that their exchange overlaps the interior.
    {
        std::vector<int64_t> iface, inner, halo;
        for (int64_t c = 0; c < ncells; ++c) {
            const int32_t* cd = &gdh->cell_dofs[(size_t)c * ndpc];
            if (owner[c] == rank) {
                bool x = false;
                for (int i = 0; i < ndpc && !x; ++i) x = gdof_owner[cd[i]] != gdof_maxowner[cd[i]];
                (x ? iface : inner).push_back(c);
            } else {
                bool touch = false;
                for (int i = 0; i < ndpc && !touch; ++i) touch = gdof_owner[cd[i]] == rank;
                if (touch) halo.push_back(c);
            }
        }
        P->n_iface = (int64_t)iface.size();
        P->ncells_own = (int64_t)(iface.size() + inner.size());
        P->cells_global = iface;
        P->cells_global.insert(P->cells_global.end(), inner.begin(), inner.end());
        P->cells_global.insert(P->cells_global.end(), halo.begin(), halo.end());
        P->cell_is_own.assign(P->cells_global.size(), 0);
        std::fill(P->cell_is_own.begin(), P->cell_is_own.begin() + P->ncells_own, 1);
    }
    const int64_t nl = (int64_t)P->cells_global.size();
    if (nl == 0) { delete P; return fb2_fail(FB2_ERR_BAD_ARG, "fb2_partition_create: rank %d owns no cells", rank); }
    // ---- local node / dof numbering (ascending global ids) ---------------------------------------------------------------------
    std::vector<int32_t> g2l_node((size_t)g->nnodes, -1), g2l_dof((size_t)gdh->ndofs, -1);
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        for (int k = 0; k < nnpc; ++k) g2l_node[g->cells[(size_t)c * nnpc + k] - 1] = 0;
        for (int i = 0; i < ndpc; ++i) g2l_dof[gdh->cell_dofs[(size_t)c * ndpc + i]] = 0;
    }
    for (int64_t n = 0; n < g->nnodes; ++n)
        if (g2l_node[n] == 0) { g2l_node[n] = (int32_t)P->l2g_node.size(); P->l2g_node.push_back(n); }
    for (int64_t d = 0; d < gdh->ndofs; ++d)
        if (g2l_dof[d] == 0) {
            g2l_dof[d] = (int32_t)P->l2g_dof.size();
            P->l2g_dof.push_back(d);
            P->dof_owner.push_back(gdof_owner[d]);
            P->ndofs_owned += gdof_owner[d] == rank ? 1 : 0;
        }
    P->lcells.resize((size_t)nl * nnpc);
    P->lcell_dofs.resize((size_t)nl * ndpc);
    const int nthl = fb2_host_threads();
#pragma omp parallel for schedule(static) num_threads(nthl)
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        for (int k = 0; k < nnpc; ++k) P->lcells[(size_t)l * nnpc + k] = g2l_node[g->cells[(size_t)c * nnpc + k] - 1] + 1;
        for (int i = 0; i < ndpc; ++i) P->lcell_dofs[(size_t)l * ndpc + i] = g2l_dof[gdh->cell_dofs[(size_t)c * ndpc + i]] + 1;
    }
    P->lxyz.resize(P->l2g_node.size() * sdim);
    const int64_t nln = (int64_t)P->l2g_node.size();
#pragma omp parallel for schedule(static) num_threads(nthl)
    for (int64_t n = 0; n < nln; ++n)
        for (int d = 0; d < sdim; ++d) P->lxyz[(size_t)n * sdim + d] = g->xyz[(size_t)P->l2g_node[n] * sdim + d];
    // ---- exchange lists -----------------------------------------------------------------------------------------------
    P->peers.resize(nparts);
    std::vector<std::vector<uint64_t>> send_k(nparts), recv_k(nparts), send_fk(nparts), recv_fk(nparts);
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        const int32_t* cd = &gdh->cell_dofs[(size_t)c * ndpc];
        if (P->cell_is_own[l]) {
            for (int j = 0; j < ndpc; ++j) {
                const int o = gdof_owner[cd[j]];
                if (o == rank) continue;
                send_fk[o].push_back((uint64_t)cd[j]);
                for (int i = 0; i < ndpc; ++i) send_k[o].push_back(((uint64_t)cd[j] << 32) | (uint32_t)cd[i]);
            }
        } else {
            const int s = owner[c];
            for (int j = 0; j < ndpc; ++j) {
                if (gdof_owner[cd[j]] != rank) continue;
                recv_fk[s].push_back((uint64_t)cd[j]);
                for (int i = 0; i < ndpc; ++i) recv_k[s].push_back(((uint64_t)cd[j] << 32) | (uint32_t)cd[i]);
            }
        }
    }
    auto uniq = [](std::vector<uint64_t>& v) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    for (int p = 0; p < nparts; ++p) {
        uniq(send_k[p]); uniq(recv_k[p]); uniq(send_fk[p]); uniq(recv_fk[p]);
        PeerPlan& pp = P->peers[p];
        for (uint64_t k : send_k[p]) { pp.send_cols.push_back(g2l_dof[k >> 32]); pp.send_rows.push_back(g2l_dof[k & 0xffffffffu]); }
        for (uint64_t k : recv_k[p]) { pp.recv_cols.push_back(g2l_dof[k >> 32]); pp.recv_rows.push_back(g2l_dof[k & 0xffffffffu]); }
        for (uint64_t k : send_fk[p]) pp.send_f.push_back(g2l_dof[k]);
        for (uint64_t k : recv_fk[p]) pp.recv_f.push_back(g2l_dof[k]);
        std::vector<uint64_t>().swap(send_k[p]);
        std::vector<uint64_t>().swap(recv_k[p]);
    }
    *out = P;
    return FB2_OK;
}

// ---- rank-local set-up of a block partition of generate_grid(Hexahedron, nel) -------------------------------------------------
// fb2_partition_create needs the global grid and DofHandler on every rank (N = 8, 400^3 cells: 14 s per rank, nearly all of
// it global sweeps).  For the case the benchmarks use -- first-order hexahedra of generate_grid, one Lagrange field of order
// 1, px x py x pz blocks -- everything the plan holds has a closed form, so a rank can build its part directly:
//   * close!(dh) numbers dofs by first appearance in the cell sweep (src/Dofs/DofHandler.jl:576-738): node (i, j, k) first
//     appears in cell (max(i-1,0), max(j-1,0), max(k-1,0)), a cell (a, b, c) introduces (1 + [a=0]) (1 + [b=0]) (1 + [c=0]) nodes
//     in local vertex order, hence  rank(node) = #nodes introduced by earlier cells + position among the new nodes of its cell
//     (fb2_gen_node_rank), dofs vdim * rank + component;
//   * the owner of a dof = lowest rank of a cell touching it = block of that first cell; cells touching an owned dof = the block
//     extended by one cell on its high sides;
//   * node coordinates: the same function as fb2_grid_generate + fb2_grid_perturb (global node id in the hash).
// The result is array-for-array what fb2_partition_create derives from the global problem (tests/test_partition_host.py);
// the global grid / DofHandler the plan refers to hold metadata only.
namespace {
// sort on `nth` threads: sorted chunks, then pairwise merges
template <class T, class Cmp>
void fb2_psort(std::vector<T>& v, int nth, Cmp cmp) {
    const size_t n = v.size();
    if (nth <= 1 || n < ((size_t)1 << 16)) { std::sort(v.begin(), v.end(), cmp); return; }
    std::vector<size_t> b((size_t)nth + 1);
    for (int k = 0; k <= nth; ++k) b[k] = n * (size_t)k / (size_t)nth;
#pragma omp parallel for schedule(static, 1) num_threads(nth)
    for (int k = 0; k < nth; ++k) std::sort(v.begin() + b[k], v.begin() + b[k + 1], cmp);
    for (int w = 1; w < nth; w *= 2) {
#pragma omp parallel for schedule(static, 1) num_threads(nth)
        for (int k = 0; k < nth; k += 2 * w)
            if (k + w < nth) std::inplace_merge(v.begin() + b[k], v.begin() + b[k + w], v.begin() + b[std::min(k + 2 * w, nth)], cmp);
    }
}
inline int64_t gen_F(int64_t m) { return m + (m > 0 ? 1 : 0); }   // nodes (per direction) introduced by cell columns < m
inline int64_t gen_f(int64_t m) { return m == 0 ? 2 : 1; }
int64_t fb2_gen_node_rank(int64_t i, int64_t j, int64_t k, int64_t nx, int64_t ny) {
    static const int LV[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    const int64_t a = std::max<int64_t>(i - 1, 0), b = std::max<int64_t>(j - 1, 0), c = std::max<int64_t>(k - 1, 0);
    int64_t r = gen_F(c) * (ny + 1) * (nx + 1) + gen_f(c) * gen_F(b) * (nx + 1) + gen_f(c) * gen_f(b) * gen_F(a);
    for (int v = 0; v < 8; ++v) {
        const int da = LV[v][0], db = LV[v][1], dc = LV[v][2];
        if (a + da == i && b + db == j && c + dc == k) break;
        if ((da == 1 || a == 0) && (db == 1 || b == 0) && (dc == 1 || c == 0)) ++r;
    }
    return r;
}
}  // namespace

extern "C" int fb2_partition_create_generated(fb2_ctx* hctx, const int64_t* nel, const double* left, const double* right, double perturb,
                                              int vdim, int nparts, int rank, const int* dims_in, fb2_part** out) {
    FB2_CHECK(hctx && nel && out, FB2_ERR_BAD_ARG, "fb2_partition_create_generated: null argument");
    FB2_CHECK(nparts >= 1 && rank >= 0 && rank < nparts && vdim >= 1 && vdim <= 3, FB2_ERR_BAD_ARG, "fb2_partition_create_generated: bad nparts / rank / vdim");
    const int64_t nx = nel[0], ny = nel[1], nz = nel[2];
    FB2_CHECK(nx >= 1 && ny >= 1 && nz >= 1, FB2_ERR_BAD_ARG, "fb2_partition_create_generated: nel < 1");
    const int64_t nnodes = (nx + 1) * (ny + 1) * (nz + 1);
    FB2_CHECK(nnodes * vdim < (int64_t)2147483647, FB2_ERR_UNSUPPORTED, "more than 2^31-1 global dofs");
    // ---- the global problem, metadata only ----------------------------------------------------------------------------------
    fb2_grid* g = new fb2_grid();
    g->ctx = hctx;
    g->celltype = FB2_HEXAHEDRON;
    g->sdim = 3;
    g->nnpc = 8;
    g->generated = true;
    g->ncells = nx * ny * nz;
    g->nnodes = nnodes;
    double lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
    for (int d = 0; d < 3; ++d) {
        if (left) lo[d] = left[d];
        if (right) hi[d] = right[d];
        g->nel[d] = nel[d]; g->left[d] = lo[d]; g->right[d] = hi[d];
    }
    fb2_dh* gdh = new fb2_dh();
    gdh->grid = g;
    fb2_field fld = {1, vdim};
    LagrangeInfo ip;
    fb2_lagrange(FB2_HEXAHEDRON, 1, &ip);
    gdh->fields.push_back(fld);
    gdh->ips.push_back(ip);
    gdh->ndpc = 8 * vdim;
    gdh->ndofs = nnodes * vdim;
    fb2_part* P = new fb2_part();
    P->gdh = gdh;
    P->owns_global = true;
    P->nparts = nparts;
    P->rank = rank;
    if (dims_in) { for (int d = 0; d < 3; ++d) P->dims[d] = dims_in[d]; }
    else default_dims(nparts, g->nel, 3, P->dims);
    if ((int64_t)P->dims[0] * P->dims[1] * P->dims[2] != nparts) {
        fb2_partition_destroy(P);
        return fb2_fail(FB2_ERR_BAD_ARG, "fb2_partition_create: block layout does not multiply to nparts");
    }
    const int px = P->dims[0], py = P->dims[1];
    std::vector<int> bx((size_t)nx), by((size_t)ny), bz((size_t)nz);
    for (int64_t i = 0; i < nx; ++i) bx[i] = block_of(i, nx, P->dims[0]);
    for (int64_t j = 0; j < ny; ++j) by[j] = block_of(j, ny, P->dims[1]);
    for (int64_t k = 0; k < nz; ++k) bz[k] = block_of(k, nz, P->dims[2]);
    auto cell_owner = [&](int64_t i, int64_t j, int64_t k) { return bx[i] + px * (by[j] + py * bz[k]); };
    // owner of a node's dofs = lowest / highest rank among the cells around it
    auto node_owner = [&](int64_t i, int64_t j, int64_t k) { return cell_owner(std::max<int64_t>(i - 1, 0), std::max<int64_t>(j - 1, 0), std::max<int64_t>(k - 1, 0)); };
    auto node_maxowner = [&](int64_t i, int64_t j, int64_t k) { return cell_owner(std::min(i, nx - 1), std::min(j, ny - 1), std::min(k, nz - 1)); };
    // my block and the box of local cells (block + one cell on the high sides)
    const int rb[3] = {rank % px, (rank / px) % py, rank / (px * py)};
    const int64_t n3[3] = {nx, ny, nz};
    int64_t blo[3], bhi[3], ehi[3];
    for (int d = 0; d < 3; ++d) {
        blo[d] = ((int64_t)rb[d] * n3[d]) / P->dims[d];
        bhi[d] = ((int64_t)(rb[d] + 1) * n3[d]) / P->dims[d];
        ehi[d] = std::min(bhi[d] + 1, n3[d]);
    }
    if (blo[0] >= bhi[0] || blo[1] >= bhi[1] || blo[2] >= bhi[2]) {
        fb2_partition_destroy(P);
        return fb2_fail(FB2_ERR_BAD_ARG, "fb2_partition_create: rank %d owns no cells", rank);
    }
    static const int HV[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    // ---- local cells: [own, touching an exchanged dof | other own | halo], each by ascending global id ------------------------
    {
        std::vector<int64_t> iface, inner, halo;
        for (int64_t k = blo[2]; k < ehi[2]; ++k)
            for (int64_t j = blo[1]; j < ehi[1]; ++j)
                for (int64_t i = blo[0]; i < ehi[0]; ++i) {
                    const int64_t c = i + nx * (j + ny * k);
                    if (i < bhi[0] && j < bhi[1] && k < bhi[2]) {
                        bool x = false;
                        for (int v = 0; v < 8 && !x; ++v)
                            x = node_owner(i + HV[v][0], j + HV[v][1], k + HV[v][2]) != node_maxowner(i + HV[v][0], j + HV[v][1], k + HV[v][2]);
                        (x ? iface : inner).push_back(c);
                    } else {
                        bool touch = false;
                        for (int v = 0; v < 8 && !touch; ++v) touch = node_owner(i + HV[v][0], j + HV[v][1], k + HV[v][2]) == rank;
                        if (touch) halo.push_back(c);
                    }
                }
        P->n_iface = (int64_t)iface.size();
        P->ncells_own = (int64_t)(iface.size() + inner.size());
        P->cells_global = iface;
        P->cells_global.insert(P->cells_global.end(), inner.begin(), inner.end());
        P->cells_global.insert(P->cells_global.end(), halo.begin(), halo.end());
        P->cell_is_own.assign(P->cells_global.size(), 0);
        std::fill(P->cell_is_own.begin(), P->cell_is_own.begin() + P->ncells_own, 1);
    }
    const int64_t nl = (int64_t)P->cells_global.size();
    // ---- local nodes: the nodes of the local cells, ascending global id; a node marked by the cells that use it ------------------
    const int64_t L1[3] = {ehi[0] - blo[0] + 1, ehi[1] - blo[1] + 1, ehi[2] - blo[2] + 1};   // nodes of the box per direction
    const int64_t nbox = L1[0] * L1[1] * L1[2];
    std::vector<int32_t> box2l((size_t)nbox, -1);
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        const int64_t i = c % nx - blo[0], j = (c / nx) % ny - blo[1], k = c / (nx * ny) - blo[2];
        for (int v = 0; v < 8; ++v) box2l[(size_t)((i + HV[v][0]) + L1[0] * ((j + HV[v][1]) + L1[1] * (k + HV[v][2])))] = 0;
    }
    struct NodeRec { int64_t gdof0; int32_t lnode; int32_t owner; };
    std::vector<NodeRec> recs;
    for (int64_t bk = 0; bk < L1[2]; ++bk)
        for (int64_t bj = 0; bj < L1[1]; ++bj)
            for (int64_t bi = 0; bi < L1[0]; ++bi) {
                int32_t& m = box2l[(size_t)(bi + L1[0] * (bj + L1[1] * bk))];
                if (m != 0) continue;
                const int64_t i = bi + blo[0], j = bj + blo[1], k = bk + blo[2];
                m = (int32_t)P->l2g_node.size();
                P->l2g_node.push_back(i + (nx + 1) * (j + (ny + 1) * k));
                recs.push_back(NodeRec{(int64_t)vdim * fb2_gen_node_rank(i, j, k, nx, ny), m, (int32_t)node_owner(i, j, k)});
            }
    const int64_t nln = (int64_t)P->l2g_node.size();
    // ---- local dofs: ascending global dof id -----------------------------------------------------------------------------------------
    const int nth = fb2_host_threads();
    fb2_psort(recs, nth, [](const NodeRec& a, const NodeRec& b) { return a.gdof0 < b.gdof0; });
    std::vector<int32_t> node_dof0((size_t)nln);   // first local dof (0-based) of a local node
    P->l2g_dof.reserve((size_t)nln * vdim);
    P->dof_owner.reserve((size_t)nln * vdim);
    for (int64_t r = 0; r < nln; ++r) {
        node_dof0[recs[r].lnode] = (int32_t)(r * vdim);
        for (int t = 0; t < vdim; ++t) {
            P->l2g_dof.push_back(recs[r].gdof0 + t);
            P->dof_owner.push_back(recs[r].owner);
            P->ndofs_owned += recs[r].owner == rank ? 1 : 0;
        }
    }
    // ---- local arrays ------------------------------------------------------------------------------------------------------------------
    const int ndpc = 8 * vdim;
    P->lcells.resize((size_t)nl * 8);
    P->lcell_dofs.resize((size_t)nl * ndpc);
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        const int64_t i = c % nx - blo[0], j = (c / nx) % ny - blo[1], k = c / (nx * ny) - blo[2];
        for (int v = 0; v < 8; ++v) {
            const int32_t ln = box2l[(size_t)((i + HV[v][0]) + L1[0] * ((j + HV[v][1]) + L1[1] * (k + HV[v][2])))];
            P->lcells[(size_t)l * 8 + v] = ln + 1;
            for (int t = 0; t < vdim; ++t) P->lcell_dofs[(size_t)l * ndpc + v * vdim + t] = node_dof0[ln] + t + 1;
        }
    }
    P->lxyz.resize((size_t)nln * 3);
    {
        int nc = 0;
        double refc[8 * 3], corner[8 * 3];
        fb2_generated_corners(3, lo, hi, &nc, refc, corner);
        const int64_t nn[3] = {nx + 1, ny + 1, nz + 1};
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int64_t n = 0; n < nln; ++n) {
            const int64_t id = P->l2g_node[n];
            const int64_t idx[3] = {id % nn[0], (id / nn[0]) % nn[1], id / (nn[0] * nn[1])};
            double x[3];
            fb2_generated_node(3, nn, nc, refc, corner, idx, x);
            const bool interior = idx[0] > 0 && idx[0] < nn[0] - 1 && idx[1] > 0 && idx[1] < nn[1] - 1 && idx[2] > 0 && idx[2] < nn[2] - 1;
            if (perturb != 0.0 && interior)
                for (int d = 0; d < 3; ++d) x[d] += fb2_perturb_delta(id, d, perturb, (hi[d] - lo[d]) / (double)n3[d]);
            for (int d = 0; d < 3; ++d) P->lxyz[(size_t)n * 3 + d] = x[d];
        }
    }
    // ---- exchange lists (as in partition_create_impl, with local look-ups) --------------------------------------------------------------------
    P->peers.resize(nparts);
    std::vector<std::vector<uint64_t>> send_k(nparts), recv_k(nparts), send_fk(nparts), recv_fk(nparts);
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        const int64_t ci = c % nx, cj = (c / nx) % ny, ck = c / (nx * ny);
        int64_t cd[24];
        int own[24];
        for (int v = 0; v < 8; ++v) {
            const int32_t ln = P->lcell_dofs[(size_t)l * ndpc + v * vdim] - 1;   // first local dof of the node
            for (int t = 0; t < vdim; ++t) {
                cd[v * vdim + t] = P->l2g_dof[ln + t];
                own[v * vdim + t] = P->dof_owner[ln + t];
            }
        }
        if (P->cell_is_own[l]) {
            for (int j = 0; j < ndpc; ++j) {
                const int o = own[j];
                if (o == rank) continue;
                send_fk[o].push_back((uint64_t)cd[j]);
                for (int i = 0; i < ndpc; ++i) send_k[o].push_back(((uint64_t)cd[j] << 32) | (uint32_t)cd[i]);
            }
        } else {
            const int s = cell_owner(ci, cj, ck);
            for (int j = 0; j < ndpc; ++j) {
                if (own[j] != rank) continue;
                recv_fk[s].push_back((uint64_t)cd[j]);
                for (int i = 0; i < ndpc; ++i) recv_k[s].push_back(((uint64_t)cd[j] << 32) | (uint32_t)cd[i]);
            }
        }
    }
    auto uniq = [&](std::vector<uint64_t>& v) {
        fb2_psort(v, nth, [](uint64_t a, uint64_t b) { return a < b; });
        v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    auto g2l = [&](uint64_t gd) { return (int32_t)(std::lower_bound(P->l2g_dof.begin(), P->l2g_dof.end(), (int64_t)gd) - P->l2g_dof.begin()); };
    for (int p = 0; p < nparts; ++p) {
        uniq(send_k[p]); uniq(recv_k[p]); uniq(send_fk[p]); uniq(recv_fk[p]);
        PeerPlan& pp = P->peers[p];
        const int64_t ns = (int64_t)send_k[p].size(), nr = (int64_t)recv_k[p].size();
        pp.send_cols.resize((size_t)ns); pp.send_rows.resize((size_t)ns);
        pp.recv_cols.resize((size_t)nr); pp.recv_rows.resize((size_t)nr);
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int64_t q = 0; q < ns; ++q) { pp.send_cols[q] = g2l(send_k[p][q] >> 32); pp.send_rows[q] = g2l(send_k[p][q] & 0xffffffffu); }
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int64_t q = 0; q < nr; ++q) { pp.recv_cols[q] = g2l(recv_k[p][q] >> 32); pp.recv_rows[q] = g2l(recv_k[p][q] & 0xffffffffu); }
        for (uint64_t k : send_fk[p]) pp.send_f.push_back(g2l(k));
        for (uint64_t k : recv_fk[p]) pp.recv_f.push_back(g2l(k));
        std::vector<uint64_t>().swap(send_k[p]);
        std::vector<uint64_t>().swap(recv_k[p]);
    }
    *out = P;
    return FB2_OK;
}

extern "C" int fb2_partition_info(fb2_part* P, int64_t* ncells_local, int64_t* ncells_own, int64_t* nnodes_local,
                                  int64_t* ndofs_local, int64_t* ndofs_owned) {
    FB2_CHECK(P, FB2_ERR_BAD_ARG, "fb2_partition_info: null handle");
    if (ncells_local) *ncells_local = (int64_t)P->cells_global.size();
    if (ncells_own) *ncells_own = P->ncells_own;
    if (nnodes_local) *nnodes_local = (int64_t)P->l2g_node.size();
    if (ndofs_local) *ndofs_local = (int64_t)P->l2g_dof.size();
    if (ndofs_owned) *ndofs_owned = P->ndofs_owned;
    return FB2_OK;
}

extern "C" int fb2_partition_export(fb2_part* P, int64_t* cells_global, uint8_t* cell_is_own, int64_t* l2g_node, int64_t* l2g_dof,
                                    int32_t* dof_owner) {
    FB2_CHECK(P, FB2_ERR_BAD_ARG, "fb2_partition_export: null handle");
    if (cells_global) for (size_t i = 0; i < P->cells_global.size(); ++i) cells_global[i] = P->cells_global[i] + 1;
    if (cell_is_own) memcpy(cell_is_own, P->cell_is_own.data(), P->cell_is_own.size());
    if (l2g_node) for (size_t i = 0; i < P->l2g_node.size(); ++i) l2g_node[i] = P->l2g_node[i] + 1;
    if (l2g_dof) for (size_t i = 0; i < P->l2g_dof.size(); ++i) l2g_dof[i] = P->l2g_dof[i] + 1;
    if (dof_owner) memcpy(dof_owner, P->dof_owner.data(), P->dof_owner.size() * sizeof(int32_t));
    return FB2_OK;
}

// Structured view of the local grid of a block partition of a generate_grid hexahedral mesh: own + halo cells fill a box,
// and the local node numbering (ascending global ids) is the box's own x-fastest numbering.  Both facts are VERIFIED here
// cell by cell; if either fails the grid simply has no structured view (the generic kernels are used).
static void attach_structured_view(const fb2_part* P, fb2_grid* lg) {
    const fb2_grid* g = P->gdh->grid;
    if (!g->generated || g->celltype != FB2_HEXAHEDRON) return;
    const int64_t nx = g->nel[0], ny = g->nel[1];
    const int64_t nl = (int64_t)P->cells_global.size();
    int64_t lo[3] = {INT64_MAX, INT64_MAX, INT64_MAX}, hi[3] = {-1, -1, -1};
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        const int64_t ijk[3] = {c % nx, (c / nx) % ny, c / (nx * ny)};
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], ijk[d]); hi[d] = std::max(hi[d], ijk[d]); }
    }
    const int64_t L[3] = {hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1};
    if (L[0] * L[1] * L[2] > (int64_t)1 << 30) return;
    if ((int64_t)P->l2g_node.size() != (L[0] + 1) * (L[1] + 1) * (L[2] + 1)) return;   // the node set is not the full box
    std::vector<int32_t> map((size_t)(L[0] * L[1] * L[2]), -1);
    static const int hx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, hy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, hz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    for (int64_t l = 0; l < nl; ++l) {
        const int64_t c = P->cells_global[l];
        const int64_t b[3] = {c % nx - lo[0], (c / nx) % ny - lo[1], c / (nx * ny) - lo[2]};
        map[(size_t)(b[0] + L[0] * (b[1] + L[1] * b[2]))] = (int32_t)l;
        for (int v = 0; v < 8; ++v) {
            const int64_t expect = (b[0] + hx[v]) + (L[0] + 1) * ((b[1] + hy[v]) + (L[1] + 1) * (b[2] + hz[v]));
            if (P->lcells[(size_t)l * 8 + v] - 1 != expect) return;
        }
    }
    if (lg->ctx->device >= 0) {
        if (cudaSetDevice(lg->ctx->device) != cudaSuccess) return;
        if (cudaMalloc(&lg->d_sv_cellmap, map.size() * sizeof(int32_t)) != cudaSuccess) { lg->d_sv_cellmap = nullptr; return; }
        if (cudaMemcpy(lg->d_sv_cellmap, map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(lg->d_sv_cellmap);
            lg->d_sv_cellmap = nullptr;
            return;
        }
    }
    for (int d = 0; d < 3; ++d) lg->sv_nel[d] = L[d];
    lg->sv_cellmap.swap(map);
    lg->structured = true;
}

extern "C" int fb2_partition_local_grid(fb2_part* P, fb2_ctx* ctx, fb2_grid** out) {
    FB2_CHECK(P && ctx && out, FB2_ERR_BAD_ARG, "fb2_partition_local_grid: null argument");
    fb2_grid* g = P->gdh->grid;
    FB2_TRY(fb2_grid_from_host(ctx, g->celltype, (int64_t)P->cells_global.size(), (int64_t)P->l2g_node.size(), g->sdim,
                               P->lcells.data(), P->lxyz.data(), out));
    attach_structured_view(P, *out);
    return FB2_OK;
}

extern "C" int fb2_partition_local_dh(fb2_part* P, fb2_grid* local_grid, fb2_dh** out) {
    FB2_CHECK(P && local_grid && out, FB2_ERR_BAD_ARG, "fb2_partition_local_dh: null argument");
    return fb2_dh_from_host(local_grid, (int)P->gdh->fields.size(), P->gdh->fields.data(), (int64_t)P->l2g_dof.size(), P->gdh->ndpc,
                            P->lcell_dofs.data(), out);
}

extern "C" int fb2_partition_peer_counts(fb2_part* P, int peer, int64_t* nz_send, int64_t* f_send, int64_t* nz_recv, int64_t* f_recv) {
    FB2_CHECK(P && peer >= 0 && peer < P->nparts, FB2_ERR_BAD_ARG, "fb2_partition_peer_counts: bad argument");
    const PeerPlan& pp = P->peers[peer];
    if (nz_send) *nz_send = (int64_t)pp.send_rows.size();
    if (f_send) *f_send = (int64_t)pp.send_f.size();
    if (nz_recv) *nz_recv = (int64_t)pp.recv_rows.size();
    if (f_recv) *f_recv = (int64_t)pp.recv_f.size();
    return FB2_OK;
}

extern "C" int fb2_partition_peer_lists(fb2_part* P, int peer, int32_t* send_rows, int32_t* send_cols, int32_t* recv_rows,
                                        int32_t* recv_cols, int32_t* send_f, int32_t* recv_f) {
    FB2_CHECK(P && peer >= 0 && peer < P->nparts, FB2_ERR_BAD_ARG, "fb2_partition_peer_lists: bad argument");
    const PeerPlan& pp = P->peers[peer];
    auto cp = [](int32_t* dst, const std::vector<int32_t>& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(int32_t)); };
    cp(send_rows, pp.send_rows); cp(send_cols, pp.send_cols); cp(recv_rows, pp.recv_rows); cp(recv_cols, pp.recv_cols);
    cp(send_f, pp.send_f); cp(recv_f, pp.recv_f);
    return FB2_OK;
}

static void free_binding(fb2_part* P) {
    for (PeerPlan& pp : P->peers) {
        cudaFree(pp.d_send_pos); cudaFree(pp.d_recv_pos); cudaFree(pp.d_send_f); cudaFree(pp.d_recv_f);
        cudaFree(pp.d_sendbuf); cudaFree(pp.d_recvbuf);
        pp.d_send_pos = pp.d_recv_pos = nullptr; pp.d_send_f = pp.d_recv_f = nullptr; pp.d_sendbuf = pp.d_recvbuf = nullptr;
    }
    cudaFree(P->d_col_owned);
    P->d_col_owned = nullptr;
    if (P->xstream) cudaStreamDestroy(P->xstream);
    if (P->ev_iface) cudaEventDestroy(P->ev_iface);
    if (P->ev_done) cudaEventDestroy(P->ev_done);
    P->xstream = nullptr; P->ev_iface = P->ev_done = nullptr;
    P->bound = nullptr;
}

// Bind the plan to a local assembler: resolve the exchange lists to positions in the local nzval, upload the
// own-cell subset and the owned-column mask.
extern "C" int fb2_partition_bind(fb2_part* P, fb2_assembler* a) {
    FB2_CHECK(P && a, FB2_ERR_BAD_ARG, "fb2_partition_bind: null argument");
    fb2_ctx* ctx = a->dh->grid->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(a->dh->ndofs == (int64_t)P->l2g_dof.size() && a->dh->grid->ncells == (int64_t)P->cells_global.size(), FB2_ERR_BAD_ARG,
              "fb2_partition_bind: the assembler does not belong to this partition's local problem");
    FB2_CUDA(cudaSetDevice(ctx->device));
    free_binding(P);
    cudaStream_t st = ctx->stream;
    for (int p = 0; p < P->nparts; ++p) {
        PeerPlan& pp = P->peers[p];
        for (int dir = 0; dir < 2; ++dir) {
            const std::vector<int32_t>& rows = dir ? pp.recv_rows : pp.send_rows;
            const std::vector<int32_t>& cols = dir ? pp.recv_cols : pp.send_cols;
            const std::vector<int32_t>& fl = dir ? pp.recv_f : pp.send_f;
            const size_t n = rows.size();
            int64_t** d_pos = dir ? &pp.d_recv_pos : &pp.d_send_pos;
            int32_t** d_f = dir ? &pp.d_recv_f : &pp.d_send_f;
            double** d_buf = dir ? &pp.d_recvbuf : &pp.d_sendbuf;
            if (n + fl.size() == 0) continue;
            FB2_CUDA(cudaMalloc(d_buf, (n + fl.size()) * sizeof(double)));
            if (!fl.empty()) {
                FB2_CUDA(cudaMalloc(d_f, fl.size() * sizeof(int32_t)));
                FB2_CUDA(cudaMemcpyAsync(*d_f, fl.data(), fl.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            }
            if (n == 0) continue;
            int32_t *d_r = nullptr, *d_c = nullptr;
            std::vector<int64_t> h(n);
            cudaError_t e = cudaMalloc(&d_r, n * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaMalloc(&d_c, n * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaMalloc(d_pos, n * sizeof(int64_t));
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_r, rows.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, cols.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                k_lookup_pos<<<nblk((int64_t)n, 256), 256, 0, st>>>(d_r, d_c, (int64_t)n, a->pat->d_colptr, a->pat->d_rowval, *d_pos);
                ctx->launches++;
                // every listed entry must exist in the local pattern
                e = cudaMemcpyAsync(h.data(), *d_pos, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFree(d_r);     // the temporaries go on every path
            cudaFree(d_c);
            FB2_CHECK(e == cudaSuccess, e == cudaErrorMemoryAllocation ? FB2_ERR_OOM : FB2_ERR_CUDA, "fb2_partition_bind: %s", cudaGetErrorString(e));
            for (size_t t = 0; t < n; ++t)
                FB2_CHECK(h[t] >= 0, FB2_ERR_INTERNAL, "fb2_partition_bind: exchange entry missing in the local pattern (peer %d)", p);
        }
    }
    int lo_prio = 0, hi_prio = 0;
    FB2_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    FB2_CUDA(cudaStreamCreateWithPriority(&P->xstream, cudaStreamNonBlocking, hi_prio));
    FB2_CUDA(cudaEventCreateWithFlags(&P->ev_iface, cudaEventDisableTiming));
    FB2_CUDA(cudaEventCreateWithFlags(&P->ev_done, cudaEventDisableTiming));
    // list of the local columns this rank does NOT own (only these are visited by the mask kernel)
    std::vector<int32_t> unowned;
    for (size_t d = 0; d < P->dof_owner.size(); ++d)
        if (P->dof_owner[d] != P->rank) unowned.push_back((int32_t)d);
    P->n_unowned = (int64_t)unowned.size();
    FB2_CUDA(cudaMalloc(&P->d_col_owned, std::max<size_t>(unowned.size(), 1) * sizeof(int32_t)));
    FB2_CUDA(cudaMemcpy(P->d_col_owned, unowned.data(), unowned.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    P->bound = a;
    P->bound_device = ctx->device;
    return FB2_OK;
}

extern "C" int fb2_partition_pack(fb2_part* P, int peer, const double* nzval_dev, const double* f_dev, double* send_dev) {
    FB2_CHECK(P && P->bound && peer >= 0 && peer < P->nparts && nzval_dev, FB2_ERR_BAD_ARG, "fb2_partition_pack: bad argument or plan not bound");
    fb2_ctx* ctx = P->bound->dh->grid->ctx;
    PeerPlan& pp = P->peers[peer];
    const int64_t nnz = (int64_t)pp.send_rows.size(), nf = (int64_t)pp.send_f.size();
    if (nnz + nf == 0) return FB2_OK;
    FB2_CUDA(cudaSetDevice(ctx->device));
    k_pack<<<nblk(nnz + nf, 256), 256, 0, ctx->stream>>>(pp.d_send_pos, nnz, pp.d_send_f, nf, nzval_dev, f_dev, send_dev ? send_dev : pp.d_sendbuf);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_partition_unpack_add(fb2_part* P, int peer, const double* recv_dev, double* nzval_dev, double* f_dev) {
    FB2_CHECK(P && P->bound && peer >= 0 && peer < P->nparts && nzval_dev, FB2_ERR_BAD_ARG, "fb2_partition_unpack_add: bad argument or plan not bound");
    fb2_ctx* ctx = P->bound->dh->grid->ctx;
    PeerPlan& pp = P->peers[peer];
    const int64_t nnz = (int64_t)pp.recv_rows.size(), nf = (int64_t)pp.recv_f.size();
    if (nnz + nf == 0) return FB2_OK;
    FB2_CUDA(cudaSetDevice(ctx->device));
    k_unpack_add<<<nblk(nnz + nf, 256), 256, 0, ctx->stream>>>(pp.d_recv_pos, nnz, pp.d_recv_f, nf, recv_dev ? recv_dev : pp.d_recvbuf, nzval_dev, f_dev);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_partition_mask_unowned(fb2_part* P, double* nzval_dev, double* f_dev) {
    FB2_CHECK(P && P->bound && nzval_dev, FB2_ERR_BAD_ARG, "fb2_partition_mask_unowned: bad argument or plan not bound");
    fb2_ctx* ctx = P->bound->dh->grid->ctx;
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = P->n_unowned;
    if (n == 0) return FB2_OK;
    k_mask_unowned<<<nblk(n * 32, 256), 256, 0, ctx->stream>>>(P->d_col_owned, n, P->bound->pat->d_colptr, nzval_dev, f_dev);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

extern "C" int fb2_partition_destroy(fb2_part* P) {
    if (!P) return FB2_OK;
    // the bound assembler may already be gone (host languages finalise in any order): use the saved device id
    if (P->bound) { cudaSetDevice(P->bound_device); free_binding(P); }
    if (P->owns_global && P->gdh) { delete P->gdh->grid; delete P->gdh; }
    delete P;
    return FB2_OK;
}

// ---- NCCL, resolved at run time ----------------------------------------------------------------------------------------
// libnccl.so.2 is dlopen'ed instead of linked: inside a PyTorch process the already loaded (bundled) NCCL is reused,
// a Julia / C host gets the system library; only the stable point-to-point subset of the API is used.
namespace {
struct NcclApi {
    struct Id128 { char b[128]; };  // ncclUniqueId, passed by value
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.handle) return FB2_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    FB2_CHECK(h, FB2_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.handle = h;
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); FB2_CHECK(g_nccl.field, FB2_ERR_NCCL, "libnccl: missing symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return FB2_OK;
}
#define FB2_NCCL(call)                                                                                      \
    do {                                                                                                    \
        int r__ = (call);                                                                                   \
        if (r__ != 0) return fb2_fail(FB2_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__));   \
    } while (0)
}  // namespace

extern "C" int fb2_comm_unique_id(void* id128) {
    FB2_CHECK(id128, FB2_ERR_BAD_ARG, "fb2_comm_unique_id: null argument");
    FB2_TRY(nccl_load());
    FB2_NCCL(g_nccl.GetUniqueId(id128));
    return FB2_OK;
}

extern "C" int fb2_comm_init_rank(fb2_ctx* ctx, const void* id128, int nranks, int rank) {
    FB2_CHECK(ctx && id128, FB2_ERR_BAD_ARG, "fb2_comm_init_rank: null argument");
    FB2_NEED_DEVICE(ctx);
    FB2_TRY(nccl_load());
    FB2_CUDA(cudaSetDevice(ctx->device));
    NcclApi::Id128 id;
    memcpy(id.b, id128, 128);
    FB2_NCCL(g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
    return FB2_OK;
}

extern "C" int fb2_comm_destroy(fb2_ctx* ctx) {
    if (ctx && ctx->nccl_comm && g_nccl.handle) {
        g_nccl.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return FB2_OK;
}

// pack everything, one grouped send/recv with every peer, unpack-add everything (all on the context's stream)
extern "C" int fb2_partition_exchange(fb2_part* P, double* nzval_dev, double* f_dev) {
    FB2_CHECK(P && P->bound && nzval_dev, FB2_ERR_BAD_ARG, "fb2_partition_exchange: bad argument or plan not bound");
    fb2_ctx* ctx = P->bound->dh->grid->ctx;
    FB2_CHECK(ctx->nccl_comm, FB2_ERR_NCCL, "fb2_partition_exchange: call fb2_comm_init_rank first");
    FB2_CHECK(ctx->nranks == P->nparts && ctx->rank == P->rank, FB2_ERR_BAD_ARG, "fb2_partition_exchange: communicator and partition disagree");
    // pack for every peer in one launch (<= XSEG_MAX peers; more fall back to one launch per peer)
    XSegs SS, RS;
    SS.n = RS.n = 0;
    SS.total = RS.total = 0;
    bool fused = true;
    for (int p = 0; p < P->nparts && fused; ++p) {
        PeerPlan& pp = P->peers[p];
        const int64_t ns = (int64_t)(pp.send_rows.size() + pp.send_f.size()), nr = (int64_t)(pp.recv_rows.size() + pp.recv_f.size());
        if (ns) {
            if (SS.n == XSEG_MAX) { fused = false; break; }
            SS.s[SS.n++] = XSeg{pp.d_send_pos, pp.d_send_f, pp.d_sendbuf, (int64_t)pp.send_rows.size(), (int64_t)pp.send_f.size(), SS.total};
            SS.total += ns;
        }
        if (nr) {
            if (RS.n == XSEG_MAX) { fused = false; break; }
            RS.s[RS.n++] = XSeg{pp.d_recv_pos, pp.d_recv_f, pp.d_recvbuf, (int64_t)pp.recv_rows.size(), (int64_t)pp.recv_f.size(), RS.total};
            RS.total += nr;
        }
    }
    if (fused) {
        if (SS.total) {
            k_pack_all<<<nblk(SS.total, 256), 256, 0, ctx->stream>>>(SS, nzval_dev, f_dev);
            ctx->launches++;
        }
    } else {
        for (int p = 0; p < P->nparts; ++p) FB2_TRY(fb2_partition_pack(P, p, nzval_dev, f_dev, nullptr));
    }
    FB2_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < P->nparts; ++p) {
        PeerPlan& pp = P->peers[p];
        const size_t ns = pp.send_rows.size() + pp.send_f.size(), nr = pp.recv_rows.size() + pp.recv_f.size();
        if (ns) FB2_NCCL(g_nccl.Send(pp.d_sendbuf, ns, /*ncclFloat64*/ 8, p, ctx->nccl_comm, ctx->stream));
        if (nr) FB2_NCCL(g_nccl.Recv(pp.d_recvbuf, nr, /*ncclFloat64*/ 8, p, ctx->nccl_comm, ctx->stream));
    }
    FB2_NCCL(g_nccl.GroupEnd());
    if (fused) {
        if (RS.total) {
            k_unpack_all<<<nblk(RS.total, 256), 256, 0, ctx->stream>>>(RS, nzval_dev, f_dev);
            ctx->launches++;
        }
        FB2_CUDA(cudaGetLastError());
    } else {
        for (int p = 0; p < P->nparts; ++p) FB2_TRY(fb2_partition_unpack_add(P, p, nullptr, nzval_dev, f_dev));
    }
    return FB2_OK;
}

extern "C" int fb2_assemble_distributed(fb2_assembler* a, fb2_part* P, int mode, int element, const void* params, size_t params_bytes,
                                        const double* u_dev, double* nzval_dev, double* f_dev, const fb2_asm_opts* opts) {
    FB2_CHECK(a && P && nzval_dev, FB2_ERR_BAD_ARG, "fb2_assemble_distributed: null argument");
    FB2_CHECK(P->bound == a, FB2_ERR_BAD_ARG, "fb2_assemble_distributed: bind the partition to this assembler first");
    FB2_CHECK(mode == FB2_DIST_EXCHANGE || mode == FB2_DIST_HALO || mode == FB2_DIST_OWN_ONLY, FB2_ERR_BAD_ARG,
              "fb2_assemble_distributed: unknown mode %d", mode);
    fb2_ctx* ctx = a->dh->grid->ctx;
    // The marching-tile kernel sweeps the whole box of local cells in one launch (interface and interior cells are not
    // separate ranges of its tiles), so with it the exchange follows the assembly on the same stream.
    const bool marching = fb2_march_applicable(a, element, opts);
    if (!marching && mode == FB2_DIST_EXCHANGE && P->nparts > 1 && P->n_iface > 0 && P->n_iface < P->ncells_own && ctx->nccl_comm) {
        // interface cells first; their exchange (pack, NCCL send/recv, unpack, mask) runs on the second stream while the
        // interior cells are assembled.  Interior cells touch no entry of an exchanged row or column, so the two never
        // write the same address.
        fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
        if (opts) o = *opts;
        FB2_CHECK(o.scatter_mode == FB2_SCATTER_ATOMIC, FB2_ERR_UNSUPPORTED, "coloured scatter on a partitioned assembler is not supported");
        a->cell_first = 0;
        a->ncells_active = P->n_iface;
        int rc = fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, &o);
        if (rc == FB2_OK) {
            cudaEventRecord(P->ev_iface, ctx->stream);
            cudaStreamWaitEvent(P->xstream, P->ev_iface, 0);
            a->cell_first = P->n_iface;
            a->ncells_active = P->ncells_own - P->n_iface;
            o.fillzero = 0;
            rc = fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, &o);
        }
        a->cell_first = 0;
        a->ncells_active = 0;
        FB2_TRY(rc);
        cudaStream_t main_stream = ctx->stream;
        ctx->stream = P->xstream;
        rc = fb2_partition_exchange(P, nzval_dev, f_dev);
        if (rc == FB2_OK) rc = fb2_partition_mask_unowned(P, nzval_dev, f_dev);
        ctx->stream = main_stream;
        FB2_TRY(rc);
        FB2_CUDA(cudaEventRecord(P->ev_done, P->xstream));
        FB2_CUDA(cudaStreamWaitEvent(ctx->stream, P->ev_done, 0));
        return FB2_OK;
    }
    // Marching kernels, exchange mode: the CTAs (tile x chunk) that hold interface cells run first; their columns are then
    // complete on this rank, and pack / NCCL / unpack-add / mask run on the second stream while the other CTAs assemble
    // (those touch no exchanged row or column).  FB2_MARCH_SPLIT=0 keeps the exchange behind ONE launch; FB2_MARCH_SPLIT=2
    // splits the launch in every mode (tests of the CTA lists without a communicator).
    const char* esplit = getenv("FB2_MARCH_SPLIT");
    const int split_env = esplit ? atoi(esplit) : 1;
    const bool can_split = marching && P->n_iface > 0 && P->n_iface < P->ncells_own && a->dh->grid->structured && !a->dh->grid->generated;
    if (can_split && ((mode == FB2_DIST_EXCHANGE && P->nparts > 1 && ctx->nccl_comm && split_env != 0) || (split_env == 2 && mode != FB2_DIST_HALO))) {
        fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
        if (opts) o = *opts;
        const bool exchange = mode == FB2_DIST_EXCHANGE && P->nparts > 1 && ctx->nccl_comm;
        // start_assemble's zero fill, then both parts may write tile-interior columns with plain stores
        a->march_iface = P->n_iface;
        if (o.fillzero) {
            bool done = false;   // column-wise where the kernel writes the other columns with plain stores
            FB2_TRY(fb2_march_split_zero_fill(a, element, P->ncells_own, nzval_dev, &done));
            if (!done) FB2_CUDA(cudaMemsetAsync(nzval_dev, 0, (size_t)a->pat->nnz * sizeof(double), ctx->stream));
            if (f_dev) FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)a->dh->ndofs * sizeof(double), ctx->stream));
        }
        a->march_overwrite = o.fillzero ? 1 : 0;
        o.fillzero = 0;
        a->cell_first = 0;
        a->ncells_active = P->ncells_own;
        a->march_iface = P->n_iface;
        // part 1 (few CTAs) and the exchange behind it run on the second, high-priority stream NEXT TO part 2: launched one
        // after the other on one stream, the small launch would leave most SMs idle for the length of a chunk
        cudaStream_t main_stream = ctx->stream;
        if (exchange) {
            cudaEventRecord(P->ev_iface, main_stream);
            cudaStreamWaitEvent(P->xstream, P->ev_iface, 0);
            ctx->stream = P->xstream;
        }
        a->march_part = 1;
        int rc = fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, &o);
        if (rc == FB2_OK && exchange) {
            rc = fb2_partition_exchange(P, nzval_dev, f_dev);
            if (rc == FB2_OK) rc = fb2_partition_mask_unowned(P, nzval_dev, f_dev);
            if (rc == FB2_OK) cudaEventRecord(P->ev_done, P->xstream);
        }
        ctx->stream = main_stream;
        if (rc == FB2_OK) {
            a->march_part = 2;
            rc = fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, &o);
        }
        a->march_part = 0;
        a->ncells_active = 0;
        FB2_TRY(rc);
        if (exchange) {
            FB2_CUDA(cudaStreamWaitEvent(ctx->stream, P->ev_done, 0));
            return FB2_OK;
        }
        if (mode == FB2_DIST_OWN_ONLY) return FB2_OK;
        return fb2_partition_mask_unowned(P, nzval_dev, f_dev);
    }
    if (mode != FB2_DIST_HALO) {   // own cells = local cells [0, ncells_own)
        a->cell_first = 0;
        a->ncells_active = P->ncells_own;
    }
    int rc = fb2_launch_assemble(a, element, params, params_bytes, u_dev, nzval_dev, f_dev, opts);
    a->ncells_active = 0;
    FB2_TRY(rc);
    if (mode == FB2_DIST_OWN_ONLY) return FB2_OK;  // the host drives pack / transport / unpack_add / mask itself
    if (mode == FB2_DIST_EXCHANGE && P->nparts > 1) FB2_TRY(fb2_partition_exchange(P, nzval_dev, f_dev));
    return fb2_partition_mask_unowned(P, nzval_dev, f_dev);
}
