// DofHandler: dof distribution (host, bit-identical to the reference) + SoA upload.
//
// Replaces close!(dh) of the reference for one SubDofHandler covering the grid:
// src/Dofs/DofHandler.jl:493-569 (__close!), :576-676 (_close_subdofhandler!), :685-713
// (_distribute_dofs_for_cell!), :715-738 (add_vertex_dofs), :746-759 (get_or_create_dofs!),
// :761-794 (face / edge / volume dofs), :854-857 (sortedge), :1043-1057 (sortface_fast: a face is
// keyed by its three smallest node ids).  One counter, cells ascending, fields in add! order, per field
// vertices -> edge interiors -> face interiors -> cell interior; vdim copies of a scalar dof are
// consecutive.  Orders 1-2 never permute entity dofs (src/interpolations.jl:577-579).
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "common.h"

struct Key3 {
    int32_t a, b, c;
    bool operator==(const Key3& o) const { return a == o.a && b == o.b && c == o.c; }
};
struct Key3Hash {
    size_t operator()(const Key3& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.a * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)k.b + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
        h = (h ^ (h >> 29)) + (uint64_t)(uint32_t)k.c * 0x94D049BB133111EBull;
        return (size_t)(h ^ (h >> 32));
    }
};

static int upload_cell_dofs(fb2_dh* dh) {
    fb2_grid* g = dh->grid;
    if (g->ctx->device < 0) return FB2_OK;  // host-only context
    FB2_CUDA(cudaSetDevice(g->ctx->device));
    std::vector<int32_t> soa((size_t)dh->ndpc * g->ncells_pad);
    for (int i = 0; i < dh->ndpc; ++i) {
        int32_t* dst = soa.data() + (size_t)i * g->ncells_pad;
        for (int64_t c = 0; c < g->ncells; ++c) dst[c] = dh->cell_dofs[(size_t)c * dh->ndpc + i];
        for (int64_t c = g->ncells; c < g->ncells_pad; ++c) dst[c] = dst[g->ncells - 1];
    }
    FB2_CUDA(cudaMalloc(&dh->d_cell_dofs, soa.size() * sizeof(int32_t)));
    FB2_CUDA(cudaMemcpy(dh->d_cell_dofs, soa.data(), soa.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return FB2_OK;
}

static int init_fields(fb2_dh* dh, fb2_grid* grid, int nfields, const fb2_field* fields) {
    FB2_CHECK(nfields >= 1 && nfields <= 8, FB2_ERR_BAD_ARG, "DofHandler: need 1..8 fields");
    dh->grid = grid;
    dh->ndpc = 0;
    for (int f = 0; f < nfields; ++f) {
        LagrangeInfo ip;
        FB2_CHECK(fb2_lagrange(grid->celltype, fields[f].order, &ip), FB2_ERR_UNSUPPORTED,
                  "DofHandler: Lagrange order %d not supported on cell type %d", fields[f].order, grid->celltype);
        FB2_CHECK(fields[f].vdim >= 1 && fields[f].vdim <= 3, FB2_ERR_BAD_ARG, "DofHandler: vdim must be 1..3");
        dh->fields.push_back(fields[f]);
        dh->ips.push_back(ip);
        dh->ndpc += ip.nbase * fields[f].vdim;
    }
    return FB2_OK;
}

extern "C" int fb2_dh_close(fb2_grid* grid, int nfields, const fb2_field* fields, fb2_dh** out) {
    FB2_CHECK(grid && fields && out, FB2_ERR_BAD_ARG, "fb2_dh_close: null argument");
    fb2_dh* dh = new fb2_dh();
    int rc = init_fields(dh, grid, nfields, fields);
    if (rc != FB2_OK) { delete dh; return rc; }
    const RefShapeInfo* rs = fb2_refshape(grid->celltype);
    const int64_t ncells = grid->ncells;
    const int nnpc = grid->nnpc;
    dh->cell_dofs.resize((size_t)ncells * dh->ndpc);

    std::vector<std::vector<int32_t>> vertexdict(nfields);
    std::vector<std::unordered_map<uint64_t, int32_t>> edgedict(nfields);
    std::vector<std::unordered_map<Key3, int32_t, Key3Hash>> facedict(nfields);
    for (int f = 0; f < nfields; ++f) {
        vertexdict[f].assign((size_t)grid->nnodes, -1);
        if (dh->ips[f].nedgedofs > 0) edgedict[f].reserve((size_t)ncells * 2);
        if (dh->ips[f].nfacedofs > 0) facedict[f].reserve((size_t)ncells * (rs->rdim == 3 ? 2 : 1));
    }
    int64_t nextdof = 0;  // 0-based
    for (int64_t ci = 0; ci < ncells; ++ci) {
        const int64_t* cell = &grid->cells[(size_t)ci * nnpc];
        int32_t* row = &dh->cell_dofs[(size_t)ci * dh->ndpc];
        int col = 0;
        for (int f = 0; f < nfields; ++f) {
            const LagrangeInfo& ip = dh->ips[f];
            const int nc = dh->fields[f].vdim;
            // vertices
            if (ip.nvertexdofs > 0)
                for (int vi = 0; vi < rs->nvertices; ++vi) {
                    int64_t v = cell[vi] - 1;
                    int32_t first = vertexdict[f][v];
                    if (first < 0) {
                        first = (int32_t)nextdof;
                        vertexdict[f][v] = first;
                        nextdof += (int64_t)ip.nvertexdofs * nc;
                    }
                    for (int t = 0; t < ip.nvertexdofs * nc; ++t) row[col++] = first + t;
                }
            // edge interiors
            if (ip.nedgedofs > 0)
                for (int ei = 0; ei < rs->nedges; ++ei) {
                    uint64_t a = (uint64_t)cell[rs->edges[ei][0]], b = (uint64_t)cell[rs->edges[ei][1]];
                    uint64_t key = a < b ? (a << 32) | b : (b << 32) | a;
                    auto ins = edgedict[f].emplace(key, (int32_t)nextdof);
                    if (ins.second) nextdof += (int64_t)ip.nedgedofs * nc;
                    int32_t first = ins.first->second;
                    for (int t = 0; t < ip.nedgedofs * nc; ++t) row[col++] = first + t;
                }
            // face interiors
            if (ip.nfacedofs > 0)
                for (int fi = 0; fi < rs->nfaces; ++fi) {
                    int32_t ids[4];
                    int nfv = rs->face_nverts[fi];
                    for (int k = 0; k < nfv; ++k) ids[k] = (int32_t)cell[rs->faces[fi][k]];
                    for (int x = 1; x < nfv; ++x)  // insertion sort of <= 4 ids
                        for (int y = x; y > 0 && ids[y] < ids[y - 1]; --y) std::swap(ids[y], ids[y - 1]);
                    Key3 key{ids[0], ids[1], ids[2]};
                    auto ins = facedict[f].emplace(key, (int32_t)nextdof);
                    if (ins.second) nextdof += (int64_t)ip.nfacedofs * nc;
                    int32_t first = ins.first->second;
                    for (int t = 0; t < ip.nfacedofs * nc; ++t) row[col++] = first + t;
                }
            // cell interior
            for (int t = 0; t < ip.nvolumedofs * nc; ++t) row[col++] = (int32_t)nextdof++;
            if (nextdof >= (int64_t)2147483647) { delete dh; return fb2_fail(FB2_ERR_UNSUPPORTED, "more than 2^31-1 dofs per device"); }
        }
        if (col != dh->ndpc) {
            const int expected = dh->ndpc;
            delete dh;
            return fb2_fail(FB2_ERR_INTERNAL, "dof distribution produced %d dofs per cell, expected %d", col, expected);
        }
    }
    dh->ndofs = nextdof;
    rc = upload_cell_dofs(dh);
    if (rc != FB2_OK) { delete dh; return rc; }
    *out = dh;
    return FB2_OK;
}

extern "C" int fb2_dh_from_host(fb2_grid* grid, int nfields, const fb2_field* fields, int64_t ndofs, int ndofs_per_cell,
                                const int64_t* cell_dofs, fb2_dh** out) {
    FB2_CHECK(grid && fields && cell_dofs && out, FB2_ERR_BAD_ARG, "fb2_dh_from_host: null argument");
    FB2_CHECK(ndofs > 0 && ndofs < (int64_t)2147483647, FB2_ERR_UNSUPPORTED, "ndofs must be in 1..2^31-2");
    fb2_dh* dh = new fb2_dh();
    int rc = init_fields(dh, grid, nfields, fields);
    if (rc != FB2_OK) { delete dh; return rc; }
    if (dh->ndpc != ndofs_per_cell) {
        const int have = dh->ndpc;
        delete dh;
        return fb2_fail(FB2_ERR_BAD_ARG, "fb2_dh_from_host: fields give %d dofs per cell, caller says %d", have, ndofs_per_cell);
    }
    dh->ndofs = ndofs;
    size_t tot = (size_t)grid->ncells * dh->ndpc;
    dh->cell_dofs.resize(tot);
    for (size_t i = 0; i < tot; ++i) {
        if (cell_dofs[i] < 1 || cell_dofs[i] > ndofs) {
            delete dh;
            return fb2_fail(FB2_ERR_BAD_ARG, "fb2_dh_from_host: dof %lld out of range 1..%lld", (long long)cell_dofs[i], (long long)ndofs);
        }
        dh->cell_dofs[i] = (int32_t)(cell_dofs[i] - 1);
    }
    rc = upload_cell_dofs(dh);
    if (rc != FB2_OK) { delete dh; return rc; }
    *out = dh;
    return FB2_OK;
}

// METIS as shipped with the CUDA toolkit (libmetis_static.a, 64-bit idx_t; see partition.cu)
extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options, int64_t* perm, int64_t* iperm);
extern "C" int METIS_SetDefaultOptions(int64_t* options);

// renumber!(dh, order): src/Dofs/DofRenumbering.jl:79-125 (apply), :167-246 (FieldWise / ComponentWise permutations).
// kind 0: perm_in (1-based, dof i -> perm_in[i]); 1: FieldWise, 2: ComponentWise with optional target blocks (1-based block
// per field / per component; nullptr = one block each in declaration order).  The permutation is returned in perm_out
// (1-based, nullable) so that a ConstraintHandler can follow (fb2_ch_renumber).  Patterns, assemblers and partitions built
// from the old numbering are stale afterwards, exactly like in the reference.
extern "C" int fb2_dh_renumber(fb2_dh* dh, int kind, const int64_t* target_blocks, int ntargets, const int64_t* perm_in, int64_t* perm_out) {
    FB2_CHECK(dh, FB2_ERR_BAD_ARG, "fb2_dh_renumber: null handle");
    const int64_t n = dh->ndofs;
    const int64_t nc = dh->grid->ncells;
    const int ndpc = dh->ndpc;
    std::vector<int64_t> perm((size_t)n);   // 0-based: old -> new
    if (kind == 0) {
        FB2_CHECK(perm_in, FB2_ERR_BAD_ARG, "fb2_dh_renumber: a permutation is required");
        std::vector<uint8_t> seen((size_t)n, 0);
        for (int64_t i = 0; i < n; ++i) {
            const int64_t v = perm_in[i] - 1;
            FB2_CHECK(v >= 0 && v < n && !seen[v], FB2_ERR_BAD_ARG, "fb2_dh_renumber: input vector is not a permutation of length ndofs(dh)");
            seen[v] = 1;
            perm[i] = v;
        }
    } else if (kind == 1 || kind == 2) {
        // component -> block
        std::vector<int> fdim, coff(1, 0);
        for (const fb2_field& f : dh->fields) { fdim.push_back(f.vdim); coff.push_back(coff.back() + f.vdim); }
        const int ncomp = coff.back(), nfields = (int)fdim.size();
        std::vector<int> block((size_t)ncomp);
        if (!target_blocks) {
            for (int f = 0; f < nfields; ++f)
                for (int c = 0; c < fdim[f]; ++c) block[coff[f] + c] = kind == 1 ? f : coff[f] + c;
        } else {
            FB2_CHECK(ntargets == (kind == 1 ? nfields : ncomp), FB2_ERR_BAD_ARG,
                      kind == 1 ? "fb2_dh_renumber: length of target block vector does not match number of fields in DofHandler"
                                : "fb2_dh_renumber: length of target block vector does not match number of components in DofHandler");
            for (int f = 0; f < nfields; ++f)
                for (int c = 0; c < fdim[f]; ++c) block[coff[f] + c] = (int)target_blocks[kind == 1 ? f : coff[f] + c] - 1;
        }
        int nblocks = 0;
        for (int b : block) { FB2_CHECK(b >= 0, FB2_ERR_BAD_ARG, "fb2_dh_renumber: target blocks are 1-based"); nblocks = std::max(nblocks, b + 1); }
        std::vector<uint8_t> used((size_t)nblocks, 0);
        for (int b : block) used[b] = 1;
        for (uint8_t u : used) FB2_CHECK(u, FB2_ERR_BAD_ARG, "fb2_dh_renumber: target blocks must be continuous and in the range 1:maxblock");
        // dof -> block (every dof belongs to exactly one field component)
        std::vector<int> dofblock((size_t)n, -1);
        for (int64_t c = 0; c < nc; ++c) {
            int off = 0;
            for (int f = 0; f < nfields; ++f) {
                const int nloc = dh->ips[f].nbase * fdim[f];
                for (int j = 0; j < nloc; ++j) dofblock[dh->cell_dofs[(size_t)c * ndpc + off + j]] = block[coff[f] + j % fdim[f]];
                off += nloc;
            }
        }
        // stable: blocks in order, ascending old dof number inside a block
        std::vector<int64_t> start((size_t)nblocks + 1, 0);
        for (int64_t d = 0; d < n; ++d) { FB2_CHECK(dofblock[d] >= 0, FB2_ERR_INTERNAL, "fb2_dh_renumber: dof %lld belongs to no cell", (long long)d + 1); start[dofblock[d] + 1]++; }
        for (int b = 0; b < nblocks; ++b) start[b + 1] += start[b];
        for (int64_t d = 0; d < n; ++d) perm[d] = start[dofblock[d]]++;
    } else if (kind == 3) {
        // DofOrder.Ext{Metis}() (ext/FerriteMetis.jl:29-92): fill-reducing nested dissection of the dof coupling graph (two
        // dofs are adjacent when a cell holds both, no diagonal); new number of dof i = iperm[i]
        std::vector<int64_t> cnt((size_t)n + 1, 0);
        for (size_t i = 0; i < dh->cell_dofs.size(); ++i) cnt[dh->cell_dofs[i] + 1]++;
        for (int64_t d = 0; d < n; ++d) cnt[d + 1] += cnt[d];
        std::vector<int64_t> inc((size_t)cnt[n]), fill(cnt.begin(), cnt.end() - 1);   // dof -> cells
        for (int64_t c = 0; c < nc; ++c)
            for (int i = 0; i < ndpc; ++i) inc[fill[dh->cell_dofs[(size_t)c * ndpc + i]]++] = c;
        std::vector<int64_t> xadj((size_t)n + 1, 0), adjncy, row;
        for (int64_t d = 0; d < n; ++d) {
            row.clear();
            for (int64_t k = cnt[d]; k < cnt[d + 1]; ++k)
                for (int i = 0; i < ndpc; ++i) {
                    const int64_t e = dh->cell_dofs[(size_t)inc[k] * ndpc + i];
                    if (e != d) row.push_back(e);
                }
            std::sort(row.begin(), row.end());
            row.erase(std::unique(row.begin(), row.end()), row.end());
            adjncy.insert(adjncy.end(), row.begin(), row.end());
            xadj[d + 1] = (int64_t)adjncy.size();
        }
        int64_t nv = n, options[40];
        METIS_SetDefaultOptions(options);
        std::vector<int64_t> mperm((size_t)n), miperm((size_t)n);
        const int mrc = METIS_NodeND(&nv, xadj.data(), adjncy.data(), nullptr, options, mperm.data(), miperm.data());
        FB2_CHECK(mrc == 1, FB2_ERR_INTERNAL, "METIS_NodeND failed with status %d", mrc);
        for (int64_t d = 0; d < n; ++d) perm[d] = miperm[d];
    } else {
        return fb2_fail(FB2_ERR_BAD_ARG, "fb2_dh_renumber: unknown order %d", kind);
    }
    for (size_t i = 0; i < dh->cell_dofs.size(); ++i) dh->cell_dofs[i] = (int32_t)perm[dh->cell_dofs[i]];
    if (dh->d_cell_dofs) { cudaFree(dh->d_cell_dofs); dh->d_cell_dofs = nullptr; }
    dh->generation++;
    FB2_TRY(upload_cell_dofs(dh));
    if (perm_out) for (int64_t i = 0; i < n; ++i) perm_out[i] = perm[i] + 1;
    return FB2_OK;
}

extern "C" int fb2_dh_info(fb2_dh* dh, int64_t* ndofs, int* ndofs_per_cell, int* nfields) {
    FB2_CHECK(dh, FB2_ERR_BAD_ARG, "fb2_dh_info: null handle");
    if (ndofs) *ndofs = dh->ndofs;
    if (ndofs_per_cell) *ndofs_per_cell = dh->ndpc;
    if (nfields) *nfields = (int)dh->fields.size();
    return FB2_OK;
}

extern "C" int fb2_dh_export(fb2_dh* dh, int64_t* cell_dofs) {
    FB2_CHECK(dh && cell_dofs, FB2_ERR_BAD_ARG, "fb2_dh_export: null argument");
    for (size_t i = 0; i < dh->cell_dofs.size(); ++i) cell_dofs[i] = (int64_t)dh->cell_dofs[i] + 1;
    return FB2_OK;
}

extern "C" int fb2_dh_dof_range(fb2_dh* dh, int field, int* first, int* last) {
    FB2_CHECK(dh && field >= 0 && field < (int)dh->fields.size(), FB2_ERR_BAD_ARG, "fb2_dh_dof_range: bad field index");
    int off = dh->field_offset(field);
    if (first) *first = off + 1;
    if (last) *last = off + dh->ips[field].nbase * dh->fields[field].vdim;
    return FB2_OK;
}

extern "C" int fb2_dh_destroy(fb2_dh* dh) {
    if (!dh) return FB2_OK;
    if (dh->grid->ctx->device >= 0) {
        cudaSetDevice(dh->grid->ctx->device);
        cudaFree(dh->d_cell_dofs);
    }
    delete dh;
    return FB2_OK;
}
