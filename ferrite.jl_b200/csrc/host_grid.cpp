// Grid construction on the host + SoA upload.
//
// Replaces generate_grid (src/Grid/grid_generators.jl:8-37 Line, :78-112 Quadrilateral, :383-417
// Triangle, :159-203 Hexahedron, :474-537 Tetrahedron; nodes by _generate_nodes :550-578) and the
// Grid container (src/Grid/grid.jl:385-393).  Device layout: connectivity as int32 SoA
// [nnpc][ncells_pad] (coalesced per-cell loads), coordinates as one padded record per node
// (4 doubles in 3-D = one 32-byte sector per gathered node).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <thread>

#include <sched.h>

#include "common.h"

static int nnpc_of(int celltype) {
    switch (celltype) {
        case FB2_LINE: return 2;
        case FB2_TRIANGLE: return 3;
        case FB2_QUADRILATERAL: return 4;
        case FB2_TETRAHEDRON: return 4;
        case FB2_HEXAHEDRON: return 8;
    }
    return 0;
}

int fb2_grid_upload_xyz(fb2_grid* g) {
    if (g->ctx->device < 0) return FB2_OK;  // host-only context
    const int xs = g->xstride, sd = g->sdim;
    std::vector<double> tmp((size_t)g->nnodes * xs, 0.0);
    for (int64_t i = 0; i < g->nnodes; ++i)
        for (int d = 0; d < sd; ++d) tmp[(size_t)i * xs + d] = g->xyz[(size_t)i * sd + d];
    FB2_CUDA(cudaMemcpyAsync(g->d_xyz, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, g->ctx->stream));
    FB2_CUDA(cudaStreamSynchronize(g->ctx->stream));
    return FB2_OK;
}

int fb2_grid_upload(fb2_grid* g) {
    g->ncells_pad = (g->ncells + 31) / 32 * 32;
    g->xstride = g->sdim == 3 ? 4 : g->sdim;
    if (g->ctx->device < 0) return FB2_OK;  // host-only context
    FB2_CUDA(cudaSetDevice(g->ctx->device));
    std::vector<int32_t> conn((size_t)g->nnpc * g->ncells_pad);
    for (int j = 0; j < g->nnpc; ++j) {
        int32_t* dst = conn.data() + (size_t)j * g->ncells_pad;
        for (int64_t c = 0; c < g->ncells; ++c) dst[c] = (int32_t)(g->cells[(size_t)c * g->nnpc + j] - 1);
        // padding cells replicate the last cell so that padded lanes stay in bounds
        for (int64_t c = g->ncells; c < g->ncells_pad; ++c) dst[c] = dst[g->ncells - 1];
    }
    FB2_CUDA(cudaMalloc(&g->d_conn, conn.size() * sizeof(int32_t)));
    FB2_CUDA(cudaMemcpy(g->d_conn, conn.data(), conn.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    FB2_CUDA(cudaMalloc(&g->d_xyz, (size_t)g->nnodes * g->xstride * sizeof(double)));
    return fb2_grid_upload_xyz(g);
}

extern "C" int fb2_grid_from_host(fb2_ctx* ctx, int celltype, int64_t ncells, int64_t nnodes, int sdim,
                                  const int64_t* cells, const double* xyz, fb2_grid** out) {
    FB2_CHECK(ctx && cells && xyz && out, FB2_ERR_BAD_ARG, "fb2_grid_from_host: null argument");
    int nnpc = nnpc_of(celltype);
    FB2_CHECK(nnpc > 0, FB2_ERR_BAD_ARG, "fb2_grid_from_host: unknown cell type %d", celltype);
    FB2_CHECK(ncells > 0 && nnodes > 0 && sdim >= 1 && sdim <= 3, FB2_ERR_BAD_ARG, "fb2_grid_from_host: bad sizes");
    FB2_CHECK(nnodes < (int64_t)2147483647, FB2_ERR_UNSUPPORTED, "more than 2^31-1 nodes per device");
    for (int64_t i = 0; i < ncells * nnpc; ++i)
        FB2_CHECK(cells[i] >= 1 && cells[i] <= nnodes, FB2_ERR_BAD_ARG, "cell node id %lld out of range", (long long)cells[i]);
    fb2_grid* g = new fb2_grid();
    g->ctx = ctx;
    g->celltype = celltype;
    g->ncells = ncells;
    g->nnodes = nnodes;
    g->nnpc = nnpc;
    g->sdim = sdim;
    g->cells.assign(cells, cells + ncells * nnpc);
    g->xyz.assign(xyz, xyz + nnodes * sdim);
    int rc = fb2_grid_upload(g);
    if (rc != FB2_OK) { delete g; return rc; }
    *out = g;
    return FB2_OK;
}

static void sort_pairs(std::vector<int64_t>& v) {
    size_t n = v.size() / 2;
    std::vector<std::pair<int64_t, int64_t>> p(n);
    for (size_t i = 0; i < n; ++i) p[i] = {v[2 * i], v[2 * i + 1]};
    std::sort(p.begin(), p.end());
    p.erase(std::unique(p.begin(), p.end()), p.end());
    v.resize(2 * p.size());
    for (size_t i = 0; i < p.size(); ++i) { v[2 * i] = p[i].first; v[2 * i + 1] = p[i].second; }
}

// Threads for the embarrassingly parallel host loops of the set-up (independent iterations: results do not depend on the
// count).  torchrun exports OMP_NUM_THREADS=1 for every rank, so the count is taken from the CPUs this process may run on,
// shared between the ranks of the node: FB2_HOST_THREADS overrides.
int fb2_host_threads() {
    if (const char* e = getenv("FB2_HOST_THREADS")) return std::max(1, atoi(e));
    int ncpu = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) ncpu = CPU_COUNT(&set);
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    return std::max(1, std::min(8, ncpu / ranks));
}

// Node (idx) of generate_grid: xi = 2 idx / (nn - 1) - 1, x = sum_c M_c(xi) corner_c with the multilinear shape functions of the
// hypercube (src/Grid/grid_generators.jl:550-578).  ONE definition, also called by the rank-local set-up of partition.cu, so
// that both produce the same bits.  refcoords: [nc][3], corner: [8][3].
void fb2_generated_node(int dim, const int64_t* nn, int nc, const double* refcoords, const double* corner, const int64_t* idx, double* x) {
    const double scale = dim == 1 ? 2.0 : (dim == 2 ? 4.0 : 8.0);
    double xi[3];
    for (int d = 0; d < dim; ++d) xi[d] = 2.0 * (double)idx[d] / (double)(nn[d] - 1) - 1.0;
    x[0] = x[1] = x[2] = 0.0;
    for (int c = 0; c < nc; ++c) {
        double M = 1.0;
        for (int d = 0; d < dim; ++d) M = M * (refcoords[c * 3 + d] > 0 ? (1 + xi[d]) : (1 - xi[d]));
        M = M / scale;
        for (int d = 0; d < dim; ++d) x[d] += M * corner[c * 3 + d];
    }
}

// corners of the box in the vertex order of the linear hypercube (_extrema_to_corners :565-578)
void fb2_generated_corners(int dim, const double* lo, const double* hi, int* nc, double* refcoords, double* corner) {
    LagrangeInfo cube;
    fb2_lagrange(dim == 1 ? FB2_LINE : (dim == 2 ? FB2_QUADRILATERAL : FB2_HEXAHEDRON), 1, &cube);
    *nc = cube.nbase;
    for (int i = 0; i < cube.nbase; ++i)
        for (int d = 0; d < 3; ++d) {
            refcoords[i * 3 + d] = cube.refcoords[i][d];
            corner[i * 3 + d] = 0.0;
            if (d < dim) {
                double dxi = cube.refcoords[i][d] - (-1.0);
                corner[i * 3 + d] = lo[d] + ((hi[d] - lo[d]) / 2.0) * dxi;
            }
        }
}

static inline double hash01(uint64_t id, uint64_t salt);
// displacement of fb2_grid_perturb for component d of global node id (0-based); interior nodes only (the caller checks)
double fb2_perturb_delta(int64_t id, int d, double amplitude, double h);

extern "C" int fb2_grid_generate(fb2_ctx* ctx, int celltype, const int64_t* nel, const double* left, const double* right,
                                 fb2_grid** out) {
    FB2_CHECK(ctx && nel && out, FB2_ERR_BAD_ARG, "fb2_grid_generate: null argument");
    const RefShapeInfo* rs = fb2_refshape(celltype);
    FB2_CHECK(rs, FB2_ERR_BAD_ARG, "fb2_grid_generate: unknown cell type %d", celltype);
    const int dim = rs->rdim;
    int64_t n[3] = {1, 1, 1};
    double lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
        n[d] = nel[d];
        FB2_CHECK(n[d] >= 1, FB2_ERR_BAD_ARG, "fb2_grid_generate: nel[%d] < 1", d);
        if (left) lo[d] = left[d];
        if (right) hi[d] = right[d];
    }
    const int64_t nn[3] = {n[0] + 1, dim > 1 ? n[1] + 1 : 1, dim > 2 ? n[2] + 1 : 1};
    const int64_t nnodes = nn[0] * nn[1] * nn[2];
    FB2_CHECK(nnodes < (int64_t)2147483647, FB2_ERR_UNSUPPORTED, "more than 2^31-1 nodes per device");
    fb2_grid* g = new fb2_grid();
    g->ctx = ctx;
    g->celltype = celltype;
    g->sdim = dim;
    g->nnodes = nnodes;
    g->nnpc = nnpc_of(celltype);
    g->generated = true;
    for (int d = 0; d < 3; ++d) { g->nel[d] = n[d]; g->left[d] = lo[d]; g->right[d] = hi[d]; }

    int nc = 0;
    double refc[8 * 3], corner[8 * 3];
    fb2_generated_corners(dim, lo, hi, &nc, refc, corner);
    // nodes: xi = 2(idx-1)/(nn-1) - 1, x = sum_i M_i(xi) corner_i, first index fastest (:550-559)
    g->xyz.resize((size_t)nnodes * dim);
    const int nthreads = fb2_host_threads();
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < nn[2]; ++k)
        for (int64_t j = 0; j < nn[1]; ++j)
            for (int64_t i = 0; i < nn[0]; ++i) {
                const int64_t id = i + nn[0] * (j + nn[1] * k);
                const int64_t idx[3] = {i, j, k};
                double x[3];
                fb2_generated_node(dim, nn, nc, refc, corner, idx, x);
                for (int d = 0; d < dim; ++d) g->xyz[(size_t)id * dim + d] = x[d];
            }
    auto node = [&](int64_t i, int64_t j, int64_t k) { return 1 + i + nn[0] * (j + nn[1] * k); };
    auto& fs = g->facetsets;
    auto push = [](std::vector<int64_t>& v, int64_t c, int64_t f) { v.push_back(c); v.push_back(f); };
    if (celltype == FB2_LINE) {
        g->ncells = n[0];
        g->cells.resize((size_t)g->ncells * 2);
        for (int64_t i = 0; i < n[0]; ++i) { g->cells[2 * i] = i + 1; g->cells[2 * i + 1] = i + 2; }
        push(fs["left"], 1, 1);
        push(fs["right"], n[0], 2);
    } else if (celltype == FB2_QUADRILATERAL) {
        g->ncells = n[0] * n[1];
        g->cells.resize((size_t)g->ncells * 4);
        for (int64_t j = 0; j < n[1]; ++j)
            for (int64_t i = 0; i < n[0]; ++i) {
                int64_t c = i + n[0] * j;
                int64_t* p = &g->cells[(size_t)c * 4];
                p[0] = node(i, j, 0); p[1] = node(i + 1, j, 0); p[2] = node(i + 1, j + 1, 0); p[3] = node(i, j + 1, 0);
                if (j == 0) push(fs["bottom"], c + 1, 1);
                if (i == n[0] - 1) push(fs["right"], c + 1, 2);
                if (j == n[1] - 1) push(fs["top"], c + 1, 3);
                if (i == 0) push(fs["left"], c + 1, 4);
            }
    } else if (celltype == FB2_TRIANGLE) {
        g->ncells = 2 * n[0] * n[1];
        g->cells.resize((size_t)g->ncells * 3);
        for (int64_t j = 0; j < n[1]; ++j)
            for (int64_t i = 0; i < n[0]; ++i) {
                int64_t c = 2 * (i + n[0] * j);
                int64_t* p = &g->cells[(size_t)c * 3];
                p[0] = node(i, j, 0); p[1] = node(i + 1, j, 0); p[2] = node(i, j + 1, 0);
                p[3] = node(i + 1, j, 0); p[4] = node(i + 1, j + 1, 0); p[5] = node(i, j + 1, 0);
                if (j == 0) push(fs["bottom"], c + 1, 1);
                if (i == n[0] - 1) push(fs["right"], c + 2, 1);
                if (j == n[1] - 1) push(fs["top"], c + 2, 2);
                if (i == 0) push(fs["left"], c + 1, 3);
            }
    } else {
        const bool tet = celltype == FB2_TETRAHEDRON;
        const int64_t ncube = n[0] * n[1] * n[2];
        g->ncells = tet ? 6 * ncube : ncube;
        g->cells.resize((size_t)g->ncells * g->nnpc);
        static const int split[6][4] = {{1, 2, 4, 8}, {1, 5, 2, 8}, {2, 3, 4, 8}, {2, 7, 3, 8}, {2, 5, 6, 8}, {2, 6, 7, 8}};
        bool hex_filled = false;
        if (!tet) {   // connectivity in parallel; the facet sets (boundary cells only) in the serial sweep below
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads)
            for (int64_t k = 0; k < n[2]; ++k)
                for (int64_t j = 0; j < n[1]; ++j)
                    for (int64_t i = 0; i < n[0]; ++i) {
                        int64_t* p = &g->cells[(size_t)(i + n[0] * (j + n[1] * k)) * 8];
                        p[0] = node(i, j, k); p[1] = node(i + 1, j, k); p[2] = node(i + 1, j + 1, k); p[3] = node(i, j + 1, k);
                        p[4] = node(i, j, k + 1); p[5] = node(i + 1, j, k + 1); p[6] = node(i + 1, j + 1, k + 1); p[7] = node(i, j + 1, k + 1);
                    }
            hex_filled = true;
        }
        for (int64_t k = 0; k < n[2]; ++k)
            for (int64_t j = 0; j < n[1]; ++j)
                for (int64_t i = 0; i < n[0]; ++i) {
                    // hexahedra: only boundary cells are left to do (their facet sets); jump over the interior of the row
                    if (hex_filled && i == 1 && n[0] > 2 && k > 0 && k < n[2] - 1 && j > 0 && j < n[1] - 1) i = n[0] - 1;
                    int64_t c = i + n[0] * (j + n[1] * k);
                    int64_t v[8] = {node(i, j, k),     node(i + 1, j, k),     node(i + 1, j + 1, k),     node(i, j + 1, k),
                                    node(i, j, k + 1), node(i + 1, j, k + 1), node(i + 1, j + 1, k + 1), node(i, j + 1, k + 1)};
                    if (!tet) {
                        if (!hex_filled) memcpy(&g->cells[(size_t)c * 8], v, sizeof(v));
                        if (!(k == 0 || j == 0 || i == 0 || i == n[0] - 1 || j == n[1] - 1 || k == n[2] - 1)) continue;
                        if (k == 0) push(fs["bottom"], c + 1, 1);
                        if (j == 0) push(fs["front"], c + 1, 2);
                        if (i == n[0] - 1) push(fs["right"], c + 1, 3);
                        if (j == n[1] - 1) push(fs["back"], c + 1, 4);
                        if (i == 0) push(fs["left"], c + 1, 5);
                        if (k == n[2] - 1) push(fs["top"], c + 1, 6);
                    } else {
                        for (int t = 0; t < 6; ++t)
                            for (int m = 0; m < 4; ++m) g->cells[((size_t)c * 6 + t) * 4 + m] = v[split[t][m] - 1];
                        int64_t b = 6 * c;  // tet id = 6*cube + t (1-based: b + t)
                        if (i == 0) { push(fs["left"], b + 1, 4); push(fs["left"], b + 2, 2); }
                        if (i == n[0] - 1) { push(fs["right"], b + 4, 1); push(fs["right"], b + 6, 1); }
                        if (j == 0) { push(fs["front"], b + 2, 1); push(fs["front"], b + 5, 1); }
                        if (j == n[1] - 1) { push(fs["back"], b + 3, 3); push(fs["back"], b + 4, 3); }
                        if (k == 0) { push(fs["bottom"], b + 1, 1); push(fs["bottom"], b + 3, 1); }
                        if (k == n[2] - 1) { push(fs["top"], b + 5, 3); push(fs["top"], b + 6, 3); }
                    }
                }
    }
    for (auto& kv : fs) sort_pairs(kv.second);
    int rc = fb2_grid_upload(g);
    if (rc != FB2_OK) { delete g; return rc; }
    *out = g;
    return FB2_OK;
}

static inline double hash01(uint64_t id, uint64_t salt) {
    uint64_t x = id * 4ull + salt;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

double fb2_perturb_delta(int64_t id, int d, double amplitude, double h) { return amplitude * h * (hash01((uint64_t)id + 1, (uint64_t)d) - 0.5); }

extern "C" int fb2_grid_perturb(fb2_grid* g, double amplitude) {
    FB2_CHECK(g, FB2_ERR_BAD_ARG, "fb2_grid_perturb: null grid");
    FB2_CHECK(g->generated, FB2_ERR_BAD_ARG, "fb2_grid_perturb: only for grids made by fb2_grid_generate");
    const int dim = g->sdim;
    int64_t nn[3] = {g->nel[0] + 1, dim > 1 ? g->nel[1] + 1 : 1, dim > 2 ? g->nel[2] + 1 : 1};
    const int nthreads = fb2_host_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t id = 0; id < g->nnodes; ++id) {
        int64_t r = id, idx[3];
        for (int d = 0; d < dim; ++d) { idx[d] = r % nn[d]; r /= nn[d]; }
        bool interior = true;
        for (int d = 0; d < dim; ++d) interior = interior && idx[d] > 0 && idx[d] < nn[d] - 1;
        if (!interior) continue;
        for (int d = 0; d < dim; ++d) {
            double h = (g->right[d] - g->left[d]) / (double)g->nel[d];
            g->xyz[(size_t)id * dim + d] += fb2_perturb_delta(id, d, amplitude, h);
        }
    }
    return fb2_grid_upload_xyz(g);
}

extern "C" int fb2_grid_set_coordinates(fb2_grid* g, const double* xyz) {
    FB2_CHECK(g && xyz, FB2_ERR_BAD_ARG, "fb2_grid_set_coordinates: null argument");
    g->xyz.assign(xyz, xyz + (size_t)g->nnodes * g->sdim);
    return fb2_grid_upload_xyz(g);
}

extern "C" int fb2_grid_info(fb2_grid* g, int* celltype, int64_t* ncells, int64_t* nnodes, int* nnpc, int* sdim) {
    FB2_CHECK(g, FB2_ERR_BAD_ARG, "fb2_grid_info: null grid");
    if (celltype) *celltype = g->celltype;
    if (ncells) *ncells = g->ncells;
    if (nnodes) *nnodes = g->nnodes;
    if (nnpc) *nnpc = g->nnpc;
    if (sdim) *sdim = g->sdim;
    return FB2_OK;
}

extern "C" int fb2_grid_export(fb2_grid* g, int64_t* cells, double* xyz) {
    FB2_CHECK(g, FB2_ERR_BAD_ARG, "fb2_grid_export: null grid");
    if (cells) memcpy(cells, g->cells.data(), g->cells.size() * sizeof(int64_t));
    if (xyz) memcpy(xyz, g->xyz.data(), g->xyz.size() * sizeof(double));
    return FB2_OK;
}

extern "C" int fb2_grid_facetset(fb2_grid* g, const char* name, int64_t* n, int64_t* pairs) {
    FB2_CHECK(g && name && n, FB2_ERR_BAD_ARG, "fb2_grid_facetset: null argument");
    auto it = g->facetsets.find(name);
    FB2_CHECK(it != g->facetsets.end(), FB2_ERR_BAD_ARG, "fb2_grid_facetset: no facet set named \"%s\"", name);
    *n = (int64_t)it->second.size() / 2;
    if (pairs) memcpy(pairs, it->second.data(), it->second.size() * sizeof(int64_t));
    return FB2_OK;
}

extern "C" int fb2_grid_destroy(fb2_grid* g) {
    if (!g) return FB2_OK;
    if (g->ctx->device >= 0) {
        cudaSetDevice(g->ctx->device);
        cudaFree(g->d_conn);
        cudaFree(g->d_xyz);
        cudaFree(g->d_xyz_stage);
        cudaFree(g->d_sv_cellmap);
    }
    delete g;
    return FB2_OK;
}
