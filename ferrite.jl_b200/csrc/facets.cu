// FacetValues and the Neumann / traction facet loop.
//
// Reference: FacetQuadratureRule src/Quadrature/quadrature.jl:205-238 (create_facet_quad_rule
// src/FEValues/facet_integrals.jl:39-47), facet_to_element_transformation / weighted_normal
// src/FEValues/facet_integrals.jl:102-239, reinit!(fv, cell, x, facet) src/FEValues/FacetValues.jl:128-154, and the
// traction loop of docs/src/literate-tutorials/hyperelasticity.jl:278-291 with assemble!(f, dofs, fe)
// src/assembler.jl:338-345.
#include <cmath>
#include <cstring>

#include "common.h"

struct fb2_fv {
    fb2_ctx* ctx = nullptr;
    int celltype = 0, rdim = 0, nfacets = 0, nq = 0, nb = 0, vdim = 1, ngeo = 0;
    // host tables, facet-major then q-major: w[f][q], N[f][q][i], dM[f][q][j][d], pts[f][q][d]
    std::vector<double> w, N, dM, pts;
    double* d_tab = nullptr;   // [w | N | dM] on the device
};

struct fb2_fset {
    fb2_grid* grid = nullptr;
    int64_t n = 0;
    std::vector<int64_t> pairs;     // (cell, facet) 1-based, as given
    int32_t* d_cell = nullptr;      // 0-based
    int8_t* d_facet = nullptr;      // 0-based
};

namespace {

int facet_celltype(int celltype) {
    switch (celltype) {
        case FB2_TRIANGLE: case FB2_QUADRILATERAL: return FB2_LINE;
        case FB2_TETRAHEDRON: return FB2_TRIANGLE;
        case FB2_HEXAHEDRON: return FB2_QUADRILATERAL;
    }
    return 0;
}

// facet: 0-based; p: point of the facet's reference shape; out: point of the cell's reference shape
void facet_to_element(int celltype, int facet, const double* p, double* out) {
    const double x = p[0], y = p[1];
    switch (celltype) {
        case FB2_QUADRILATERAL: {
            const double t[4][2] = {{x, -1.0}, {1.0, x}, {-x, 1.0}, {-1.0, -x}};
            out[0] = t[facet][0]; out[1] = t[facet][1];
            break;
        }
        case FB2_TRIANGLE: {
            const double s = (x + 1.0) / 2;
            const double t[3][2] = {{1.0 - s, s}, {0.0, 1.0 - s}, {s, 0.0}};
            out[0] = t[facet][0]; out[1] = t[facet][1];
            break;
        }
        case FB2_HEXAHEDRON: {
            const double t[6][3] = {{y, x, -1.0}, {x, -1.0, y}, {1.0, x, y}, {-x, 1.0, y}, {-1.0, y, x}, {x, y, 1.0}};
            for (int d = 0; d < 3; ++d) out[d] = t[facet][d];
            break;
        }
        case FB2_TETRAHEDRON: {
            const double z = 1.0 - x - y;
            const double t[4][3] = {{z, y, 0.0}, {y, 0.0, z}, {x, y, z}, {0.0, z, y}};
            for (int d = 0; d < 3; ++d) out[d] = t[facet][d];
            break;
        }
    }
}

struct FacetArgs {
    const int32_t* conn;
    const double* xyz;
    const int32_t* cell_dofs;
    int64_t ncells_pad;
    const int32_t* cell;
    const int8_t* facet;
    int64_t n;
    const double* tab;
    int o_w, o_N, o_dM;
    int celltype, nq, nb, vdim, ngeo, xstride;
    int kind;
    double p[3];
    double* f;
    int* errflag;
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// J stored [a][b] = d x_a / d xi_b; returns the weighted normal of local facet `facet` (0-based)
template <int DIM>
__device__ void weighted_normal(int celltype, int facet, const double (&J)[DIM][DIM], double (&wn)[DIM]) {
    if (DIM == 2) {
        if (celltype == FB2_QUADRILATERAL) {
            const double s = facet < 2 ? 1.0 : -1.0;
            const int c = facet & 1;   // facets 1,3 use column 1; facets 2,4 column 2
            wn[0] = s * J[1][c];
            wn[1] = -s * J[0][c];
        } else {
            if (facet == 0) { wn[0] = -(J[1][0] - J[1][1]); wn[1] = J[0][0] - J[0][1]; }
            else if (facet == 1) { wn[0] = -J[1][1]; wn[1] = J[0][1]; }
            else { wn[0] = J[1][0]; wn[1] = -J[0][0]; }
        }
    } else {
        double a[3], b[3];
        int ia, ib;
        if (celltype == FB2_HEXAHEDRON) {
            const int A[6] = {1, 0, 1, 2, 2, 0}, B[6] = {0, 2, 2, 0, 1, 1};
            ia = A[facet]; ib = B[facet];
            for (int d = 0; d < 3; ++d) { a[d] = J[d][ia]; b[d] = J[d][ib]; }
        } else if (facet == 2) {
            for (int d = 0; d < 3; ++d) { a[d] = J[d][0] - J[d][2]; b[d] = J[d][1] - J[d][2]; }
        } else {
            const int A[4] = {1, 0, 0, 2}, B[4] = {0, 2, 0, 1};
            ia = A[facet]; ib = B[facet];
            for (int d = 0; d < 3; ++d) { a[d] = J[d][ia]; b[d] = J[d][ib]; }
        }
        double c[3];
        cross3(a, b, c);
        for (int d = 0; d < 3; ++d) wn[d] = c[d];
    }
}

// thread per (cell, facet) pair: per facet quadrature point J, weighted normal, |wn| > 0, dGamma; fe goes straight
// into f with FP64 REDs (a facet set is a surface: O(ncells^(2/3)) pairs, not a hot loop)
template <int DIM>
__global__ void k_facets(const FacetArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n) return;
    const int64_t cell = A.cell[t];
    const int facet = A.facet[t];
    double x[8][DIM];
    for (int j = 0; j < A.ngeo; ++j) {
        const int node = A.conn[(size_t)j * A.ncells_pad + cell];
        for (int a = 0; a < DIM; ++a) x[j][a] = A.xyz[(size_t)node * A.xstride + a];
    }
    const double* tw = A.tab + A.o_w + (size_t)facet * A.nq;
    const double* tN = A.tab + A.o_N + (size_t)facet * A.nq * A.nb;
    const double* tdM = A.tab + A.o_dM + (size_t)facet * A.nq * A.ngeo * DIM;
    for (int q = 0; q < A.nq; ++q) {
        double J[DIM][DIM];
        for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b) J[a][b] = 0.0;
        for (int j = 0; j < A.ngeo; ++j)
            for (int a = 0; a < DIM; ++a)
                for (int b = 0; b < DIM; ++b) J[a][b] = fma(x[j][a], tdM[(q * A.ngeo + j) * DIM + b], J[a][b]);
        double wn[DIM];
        weighted_normal<DIM>(A.celltype, facet, J, wn);
        double det = 0.0;
        for (int a = 0; a < DIM; ++a) det += wn[a] * wn[a];
        det = sqrt(det);
        if (!(det > 0.0)) {
            if (atomicCAS(&A.errflag[0], 0, FB2_ERR_DETJ_NOT_POSITIVE) == 0) A.errflag[1] = (int)cell;
            return;
        }
        const double dG = det * tw[q];
        double tr[3] = {0.0, 0.0, 0.0};   // traction (or scalar flux in tr[0]) at this point
        if (A.kind == FB2_FACET_FLUX) tr[0] = A.p[0];
        else if (A.kind == FB2_FACET_TRACTION) { for (int c = 0; c < A.vdim; ++c) tr[c] = A.p[c]; }
        else { for (int c = 0; c < DIM; ++c) tr[c] = A.p[0] * (wn[c] / det); }
        for (int i = 0; i < A.nb; ++i) {
            const double Ni = tN[q * A.nb + i];
            if (Ni == 0.0) continue;
            for (int c = 0; c < A.vdim; ++c) {
                const int dof = A.cell_dofs[(size_t)(i * A.vdim + c) * A.ncells_pad + cell];
                atomicAdd(A.f + dof, Ni * tr[c] * dG);
            }
        }
    }
}

}  // namespace

extern "C" int fb2_facetvalues_create(fb2_ctx* ctx, int celltype, int qr_order, int ip_order, int vdim, int geo_order, fb2_fv** out) {
    FB2_CHECK(ctx && out, FB2_ERR_BAD_ARG, "fb2_facetvalues_create: null argument");
    LagrangeInfo ip, geo;
    FB2_CHECK(fb2_lagrange(celltype, ip_order, &ip), FB2_ERR_UNSUPPORTED, "FacetValues: Lagrange order %d on cell type %d not supported", ip_order, celltype);
    FB2_CHECK(fb2_lagrange(celltype, geo_order, &geo), FB2_ERR_UNSUPPORTED, "FacetValues: geometric order %d not supported", geo_order);
    FB2_CHECK(vdim >= 1 && vdim <= 3, FB2_ERR_BAD_ARG, "FacetValues: vdim must be 1..3");
    const int fct = facet_celltype(celltype);
    FB2_CHECK(fct != 0, FB2_ERR_UNSUPPORTED, "FacetValues: cell type %d has no facet rule here", celltype);
    const RefShapeInfo* rs = fb2_refshape(celltype);
    std::vector<double> w, p;
    FB2_CHECK(fb2_quadrature(fct, qr_order, &w, &p), FB2_ERR_UNSUPPORTED, "FacetQuadratureRule of order %d on cell type %d not supported", qr_order, celltype);
    fb2_fv* fv = new fb2_fv();
    fv->ctx = ctx;
    fv->celltype = celltype;
    fv->rdim = ip.rdim;
    fv->nfacets = rs->rdim == 2 ? rs->nedges : rs->nfaces;
    fv->nq = (int)w.size();
    fv->nb = ip.nbase;
    fv->vdim = vdim;
    fv->ngeo = geo.nbase;
    const int rd = ip.rdim, frd = rd - 1, nf = fv->nfacets, nq = fv->nq;
    fv->w.resize((size_t)nf * nq);
    fv->pts.resize((size_t)nf * nq * rd);
    fv->N.resize((size_t)nf * nq * fv->nb);
    fv->dM.resize((size_t)nf * nq * fv->ngeo * rd);
    std::vector<double> dN((size_t)fv->nb * rd), M((size_t)fv->ngeo);
    for (int f = 0; f < nf; ++f)
        for (int q = 0; q < nq; ++q) {
            // the triangle's edges are parametrised over (0,1): half the line weights (quadrature.jl:231-232)
            fv->w[(size_t)f * nq + q] = celltype == FB2_TRIANGLE ? w[q] / 2 : w[q];
            double fp[2] = {p[(size_t)q * frd], frd > 1 ? p[(size_t)q * frd + 1] : 0.0};
            double* xi = &fv->pts[((size_t)f * nq + q) * rd];
            facet_to_element(celltype, f, fp, xi);
            fb2_lagrange_eval(ip, xi, &fv->N[((size_t)f * nq + q) * fv->nb], dN.data());
            fb2_lagrange_eval(geo, xi, M.data(), &fv->dM[((size_t)f * nq + q) * fv->ngeo * rd]);
        }
    *out = fv;
    return FB2_OK;
}

extern "C" int fb2_facetvalues_info(fb2_fv* fv, int* nfacets, int* nq, int* nbase_scalar, int* vdim, int* rdim) {
    FB2_CHECK(fv, FB2_ERR_BAD_ARG, "fb2_facetvalues_info: null handle");
    if (nfacets) *nfacets = fv->nfacets;
    if (nq) *nq = fv->nq;
    if (nbase_scalar) *nbase_scalar = fv->nb;
    if (vdim) *vdim = fv->vdim;
    if (rdim) *rdim = fv->rdim;
    return FB2_OK;
}

extern "C" int fb2_facetvalues_export(fb2_fv* fv, double* w, double* points, double* N) {
    FB2_CHECK(fv, FB2_ERR_BAD_ARG, "fb2_facetvalues_export: null handle");
    if (w) memcpy(w, fv->w.data(), fv->w.size() * sizeof(double));
    if (points) memcpy(points, fv->pts.data(), fv->pts.size() * sizeof(double));
    if (N) memcpy(N, fv->N.data(), fv->N.size() * sizeof(double));
    return FB2_OK;
}

extern "C" int fb2_facetvalues_destroy(fb2_fv* fv) {
    if (!fv) return FB2_OK;
    if (fv->d_tab) cudaFree(fv->d_tab);
    delete fv;
    return FB2_OK;
}

extern "C" int fb2_facetset_create(fb2_grid* g, const int64_t* pairs, int64_t n, fb2_fset** out) {
    FB2_CHECK(g && out && (pairs || n == 0) && n >= 0, FB2_ERR_BAD_ARG, "fb2_facetset_create: bad argument");
    const RefShapeInfo* rs = fb2_refshape(g->celltype);
    const int nf = rs->rdim == 1 ? 2 : (rs->rdim == 2 ? rs->nedges : rs->nfaces);
    for (int64_t k = 0; k < n; ++k) {
        FB2_CHECK(pairs[2 * k] >= 1 && pairs[2 * k] <= g->ncells && pairs[2 * k + 1] >= 1 && pairs[2 * k + 1] <= nf, FB2_ERR_BAD_ARG,
                  "fb2_facetset_create: (cell %lld, facet %lld) out of range", (long long)pairs[2 * k], (long long)pairs[2 * k + 1]);
    }
    fb2_fset* s = new fb2_fset();
    s->grid = g;
    s->n = n;
    s->pairs.assign(pairs, pairs + 2 * n);
    *out = s;
    return FB2_OK;
}

extern "C" int fb2_facetset_destroy(fb2_fset* s) {
    if (!s) return FB2_OK;
    if (s->d_cell) cudaFree(s->d_cell);
    if (s->d_facet) cudaFree(s->d_facet);
    delete s;
    return FB2_OK;
}

extern "C" int fb2_assemble_facets(fb2_dh* dh, fb2_fv* fv, fb2_fset* set, int kind, const double* params, int nparams, double* f_dev) {
    FB2_CHECK(dh && fv && set && f_dev, FB2_ERR_BAD_ARG, "fb2_assemble_facets: null argument");
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(set->grid == g, FB2_ERR_BAD_ARG, "fb2_assemble_facets: the facet set belongs to another grid");
    FB2_CHECK(fv->celltype == g->celltype && fv->rdim == g->sdim, FB2_ERR_BAD_ARG, "fb2_assemble_facets: FacetValues do not match the grid");
    FB2_CHECK(fv->ngeo == g->nnpc && fv->ngeo <= 8, FB2_ERR_UNSUPPORTED, "fb2_assemble_facets: geometric interpolation must match the cell's nodes");
    // the FacetValues integrate the FIRST field of the DofHandler (local dofs 0 .. nb*vdim-1); further fields (e.g. the pressure
    // of a mixed u-p problem, incompressible_elasticity.jl:295-309) take no part in the surface term
    FB2_CHECK(!dh->fields.empty() && dh->ips[0].nbase * dh->fields[0].vdim == fv->nb * fv->vdim, FB2_ERR_BAD_ARG,
              "fb2_assemble_facets: FacetValues must cover the first field of the DofHandler");
    FB2_CHECK(kind == FB2_FACET_FLUX || kind == FB2_FACET_TRACTION || kind == FB2_FACET_NORMAL_TRACTION, FB2_ERR_BAD_ARG,
              "fb2_assemble_facets: unknown kind %d", kind);
    const int need = kind == FB2_FACET_TRACTION ? fv->vdim : 1;
    FB2_CHECK(params && nparams == need, FB2_ERR_BAD_ARG, "fb2_assemble_facets: kind %d takes %d parameter(s)", kind, need);
    FB2_CHECK(kind != FB2_FACET_FLUX || fv->vdim == 1, FB2_ERR_BAD_ARG, "fb2_assemble_facets: a flux needs a scalar field");
    FB2_CHECK(kind != FB2_FACET_NORMAL_TRACTION || fv->vdim == fv->rdim, FB2_ERR_BAD_ARG, "fb2_assemble_facets: a normal traction needs vdim == dim");
    if (set->n == 0) return FB2_OK;
    FB2_CUDA(cudaSetDevice(ctx->device));
    FacetArgs A;
    memset(&A, 0, sizeof(A));
    const int nf = fv->nfacets, nq = fv->nq, rd = fv->rdim;
    A.o_w = 0;
    A.o_N = nf * nq;
    A.o_dM = A.o_N + nf * nq * fv->nb;
    if (!fv->d_tab) {
        std::vector<double> h((size_t)A.o_dM + (size_t)nf * nq * fv->ngeo * rd);
        memcpy(h.data() + A.o_w, fv->w.data(), fv->w.size() * sizeof(double));
        memcpy(h.data() + A.o_N, fv->N.data(), fv->N.size() * sizeof(double));
        memcpy(h.data() + A.o_dM, fv->dM.data(), fv->dM.size() * sizeof(double));
        FB2_CUDA(cudaMalloc(&fv->d_tab, h.size() * sizeof(double)));
        FB2_CUDA(cudaMemcpy(fv->d_tab, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (!set->d_cell) {
        std::vector<int32_t> c((size_t)set->n);
        std::vector<int8_t> f((size_t)set->n);
        for (int64_t k = 0; k < set->n; ++k) { c[k] = (int32_t)(set->pairs[2 * k] - 1); f[k] = (int8_t)(set->pairs[2 * k + 1] - 1); }
        FB2_CUDA(cudaMalloc(&set->d_cell, c.size() * sizeof(int32_t)));
        FB2_CUDA(cudaMalloc(&set->d_facet, f.size() * sizeof(int8_t)));
        FB2_CUDA(cudaMemcpy(set->d_cell, c.data(), c.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        FB2_CUDA(cudaMemcpy(set->d_facet, f.data(), f.size() * sizeof(int8_t), cudaMemcpyHostToDevice));
    }
    A.conn = g->d_conn;
    A.xyz = g->d_xyz;
    A.cell_dofs = dh->d_cell_dofs;
    A.ncells_pad = g->ncells_pad;
    A.cell = set->d_cell;
    A.facet = set->d_facet;
    A.n = set->n;
    A.tab = fv->d_tab;
    A.celltype = g->celltype;
    A.nq = nq; A.nb = fv->nb; A.vdim = fv->vdim; A.ngeo = fv->ngeo; A.xstride = g->xstride;
    A.kind = kind;
    for (int k = 0; k < nparams && k < 3; ++k) A.p[k] = params[k];
    A.f = f_dev;
    A.errflag = ctx->d_errflag;
    const unsigned grid = (unsigned)((set->n + 127) / 128);
    if (rd == 2) k_facets<2><<<grid, 128, 0, ctx->stream>>>(A);
    else if (rd == 3) k_facets<3><<<grid, 128, 0, ctx->stream>>>(A);
    else return fb2_fail(FB2_ERR_UNSUPPORTED, "fb2_assemble_facets: 1-D cells are not supported");
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}

// ---- reinit!(cv, cell) as a stand-alone operation (post-processing, user code outside the fused loop) ---------------------
// src/FEValues/CellValues.jl:122-140: per quadrature point J = sum_j x_j (x) dM_j/dxi, detJ > 0, detJdV = detJ w,
// dNdx = dNdxi . inv(J).  Thread per (cell, quadrature point); outputs in the reference's array layout:
// dNdx[cell][q][i][d] (= cv.fun_values.dNdx[i, q] per cell), detJdV[cell][q].
namespace {
// J = sum_j x_j (x) dM_j/dxi at quadrature point q of `cell`; returns det(J) and the inverse
template <int DIM>
__device__ double cell_jacobian_inverse(const int32_t* __restrict__ conn, const double* __restrict__ xyz, int64_t ncells_pad, int xstride,
                                        int64_t cell, const double* __restrict__ tdM, int q, int ngeo, double (&Ji)[DIM][DIM]) {
    double J[DIM][DIM];
    for (int a = 0; a < DIM; ++a)
        for (int b = 0; b < DIM; ++b) J[a][b] = 0.0;
    for (int j = 0; j < ngeo; ++j) {
        const int node = conn[(size_t)j * ncells_pad + cell];
        for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b) J[a][b] = fma(xyz[(size_t)node * xstride + a], tdM[(q * ngeo + j) * DIM + b], J[a][b]);
    }
    double det;
    if (DIM == 1) { det = J[0][0]; Ji[0][0] = 1.0 / det; }
    else if (DIM == 2) {
        const int Y = 1 % DIM;
        det = J[0][0] * J[Y][Y] - J[0][Y] * J[Y][0];
        const double r = 1.0 / det;
        Ji[0][0] = J[Y][Y] * r; Ji[0][Y] = -J[0][Y] * r; Ji[Y][0] = -J[Y][0] * r; Ji[Y][Y] = J[0][0] * r;
    } else {
        const int X = 0, Y = 1 % DIM, Z = 2 % DIM;
        const double c00 = J[Y][Y] * J[Z][Z] - J[Y][Z] * J[Z][Y], c01 = J[Y][X] * J[Z][Z] - J[Y][Z] * J[Z][X], c02 = J[Y][X] * J[Z][Y] - J[Y][Y] * J[Z][X];
        det = J[X][X] * c00 - J[X][Y] * c01 + J[X][Z] * c02;
        const double r = 1.0 / det;
        Ji[X][X] = c00 * r;  Ji[X][Y] = -(J[X][Y] * J[Z][Z] - J[X][Z] * J[Z][Y]) * r;  Ji[X][Z] = (J[X][Y] * J[Y][Z] - J[X][Z] * J[Y][Y]) * r;
        Ji[Y][X] = -c01 * r; Ji[Y][Y] = (J[X][X] * J[Z][Z] - J[X][Z] * J[Z][X]) * r;   Ji[Y][Z] = -(J[X][X] * J[Y][Z] - J[X][Z] * J[Y][X]) * r;
        Ji[Z][X] = c02 * r;  Ji[Z][Y] = -(J[X][X] * J[Z][Y] - J[X][Y] * J[Z][X]) * r;  Ji[Z][Z] = (J[X][X] * J[Y][Y] - J[X][Y] * J[Y][X]) * r;
    }
    return det;
}

template <int DIM>
__global__ void k_reinit_cells(const int32_t* __restrict__ conn, const double* __restrict__ xyz, int64_t ncells_pad, int xstride,
                               const int64_t* __restrict__ cells, int64_t n, const double* __restrict__ tab, int o_w, int o_dN, int o_dM,
                               int nq, int nb, int ngeo, double* __restrict__ dNdx, double* __restrict__ detJdV, int* errflag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nq) return;
    const int64_t k = t / nq;
    const int q = (int)(t - k * nq);
    const int64_t cell = cells ? cells[k] - 1 : k;
    double Ji[DIM][DIM];
    const double det = cell_jacobian_inverse<DIM>(conn, xyz, ncells_pad, xstride, cell, tab + o_dM, q, ngeo, Ji);
    if (!(det > 0.0)) {
        if (atomicCAS(&errflag[0], 0, FB2_ERR_DETJ_NOT_POSITIVE) == 0) errflag[1] = (int)cell;
        return;
    }
    detJdV[t] = det * tab[o_w + q];
    double* out = dNdx + (size_t)t * nb * DIM;
    for (int i = 0; i < nb; ++i)
        for (int b = 0; b < DIM; ++b) {
            double s = 0.0;
            for (int a = 0; a < DIM; ++a) s = fma(tab[o_dN + (q * nb + i) * DIM + a], Ji[a][b], s);
            out[i * DIM + b] = s;
        }
}

// spatial_coordinate(cv, q, x) (src/FEValues/common_values.jl:363-372): x_q = sum_j M_j(xi_q) x_j
__global__ void k_spatial_coordinates(const int32_t* __restrict__ conn, const double* __restrict__ xyz, int64_t ncells_pad, int xstride,
                                      int sdim, const int64_t* __restrict__ cells, int64_t n, const double* __restrict__ tab, int o_M,
                                      int nq, int ngeo, double* __restrict__ x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nq) return;
    const int64_t k = t / nq;
    const int q = (int)(t - k * nq);
    const int64_t cell = cells ? cells[k] - 1 : k;
    double s[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < ngeo; ++j) {
        const double M = tab[o_M + q * ngeo + j];
        const double* xj = xyz + (size_t)conn[(size_t)j * ncells_pad + cell] * xstride;
        for (int d = 0; d < sdim; ++d) s[d] = fma(M, xj[d], s[d]);
    }
    for (int d = 0; d < sdim; ++d) x[(size_t)t * sdim + d] = s[d];
}

// function_value / function_gradient of a dof vector at every quadrature point of every cell
// (src/FEValues/common_values.jl:177-227): val[c] = sum_a N_a u_(a,c), grad[c][d] = sum_a u_(a,c) dN_a/dx_d
template <int DIM>
__global__ void k_function_values(const int32_t* __restrict__ conn, const double* __restrict__ xyz, const int32_t* __restrict__ cell_dofs,
                                  int64_t ncells_pad, int xstride, int64_t n, const double* __restrict__ tab, int o_N, int o_dN, int o_dM,
                                  int nq, int nb, int vdim, int ngeo, const double* __restrict__ u, double* __restrict__ vals,
                                  double* __restrict__ grads, int* errflag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nq) return;
    const int64_t cell = t / nq;
    const int q = (int)(t - cell * nq);
    double Ji[DIM][DIM];
    const double det = cell_jacobian_inverse<DIM>(conn, xyz, ncells_pad, xstride, cell, tab + o_dM, q, ngeo, Ji);
    if (!(det > 0.0)) {
        if (atomicCAS(&errflag[0], 0, FB2_ERR_DETJ_NOT_POSITIVE) == 0) errflag[1] = (int)cell;
        return;
    }
    double v[3] = {0.0, 0.0, 0.0}, gr[3][DIM];
    for (int c = 0; c < 3; ++c)
        for (int d = 0; d < DIM; ++d) gr[c][d] = 0.0;
    for (int a = 0; a < nb; ++a) {
        double g[DIM];
        for (int b = 0; b < DIM; ++b) {
            double s = 0.0;
            for (int k = 0; k < DIM; ++k) s = fma(tab[o_dN + (q * nb + a) * DIM + k], Ji[k][b], s);
            g[b] = s;
        }
        const double Na = tab[o_N + q * nb + a];
        for (int c = 0; c < vdim; ++c) {
            const double uc = u[cell_dofs[(size_t)(a * vdim + c) * ncells_pad + cell]];
            v[c] = fma(Na, uc, v[c]);
            for (int d = 0; d < DIM; ++d) gr[c][d] = fma(uc, g[d], gr[c][d]);
        }
    }
    for (int c = 0; c < vdim; ++c) {
        if (vals) vals[(size_t)t * vdim + c] = v[c];
        if (grads) for (int d = 0; d < DIM; ++d) grads[((size_t)t * vdim + c) * DIM + d] = gr[c][d];
    }
}
}  // namespace

extern "C" int fb2_reinit_cells(fb2_cv* cv, fb2_grid* g, const int64_t* cells, int64_t n, double* dNdx_dev, double* detJdV_dev) {
    FB2_CHECK(cv && g && n >= 0, FB2_ERR_BAD_ARG, "fb2_reinit_cells: bad argument");
    if (n == 0) return FB2_OK;                       // empty batch: nothing to write (the outputs may be null)
    FB2_CHECK(dNdx_dev && detJdV_dev, FB2_ERR_BAD_ARG, "fb2_reinit_cells: null output");
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(cv->celltype == g->celltype && cv->rdim == g->sdim && cv->ngeo == g->nnpc, FB2_ERR_BAD_ARG, "fb2_reinit_cells: CellValues do not match the grid");
    if (cells) for (int64_t k = 0; k < n; ++k) FB2_CHECK(cells[k] >= 1 && cells[k] <= g->ncells, FB2_ERR_BAD_ARG, "fb2_reinit_cells: cell %lld out of range", (long long)cells[k]);
    else FB2_CHECK(n <= g->ncells, FB2_ERR_BAD_ARG, "fb2_reinit_cells: more cells than the grid has");
    FB2_CUDA(cudaSetDevice(ctx->device));
    const int nq = cv->nq, nb = cv->nb, ng = cv->ngeo, rd = cv->rdim;
    const int o_w = 0, o_N = nq, o_dN = o_N + nq * nb, o_M = o_dN + nq * nb * rd, o_dM = o_M + nq * ng;
    if (!cv->d_tables) {
        std::vector<double> h((size_t)o_dM + (size_t)nq * ng * rd);
        memcpy(h.data() + o_w, cv->w.data(), sizeof(double) * nq);
        memcpy(h.data() + o_N, cv->N.data(), sizeof(double) * nq * nb);
        memcpy(h.data() + o_dN, cv->dN.data(), sizeof(double) * nq * nb * rd);
        memcpy(h.data() + o_M, cv->M.data(), sizeof(double) * nq * ng);
        memcpy(h.data() + o_dM, cv->dM.data(), sizeof(double) * nq * ng * rd);
        FB2_CUDA(cudaMalloc(&cv->d_tables, h.size() * sizeof(double)));
        FB2_CUDA(cudaMemcpy(cv->d_tables, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        cv->tables_device = ctx->device;
        cv->tables_count = h.size();
    }
    int64_t* d_cells = nullptr;
    if (cells) {
        FB2_CUDA(cudaMalloc(&d_cells, n * sizeof(int64_t)));
        FB2_CUDA(cudaMemcpyAsync(d_cells, cells, n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    const unsigned grid = (unsigned)((n * nq + 127) / 128);
    if (rd == 1) k_reinit_cells<1><<<grid, 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, g->ncells_pad, g->xstride, d_cells, n, cv->d_tables, o_w, o_dN, o_dM, nq, nb, ng, dNdx_dev, detJdV_dev, ctx->d_errflag);
    else if (rd == 2) k_reinit_cells<2><<<grid, 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, g->ncells_pad, g->xstride, d_cells, n, cv->d_tables, o_w, o_dN, o_dM, nq, nb, ng, dNdx_dev, detJdV_dev, ctx->d_errflag);
    else k_reinit_cells<3><<<grid, 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, g->ncells_pad, g->xstride, d_cells, n, cv->d_tables, o_w, o_dN, o_dM, nq, nb, ng, dNdx_dev, detJdV_dev, ctx->d_errflag);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (d_cells) { cudaStreamSynchronize(ctx->stream); cudaFree(d_cells); }
    FB2_CUDA(e);
    return FB2_OK;
}

static int ensure_cv_tables(fb2_cv* cv) {
    if (cv->d_tables) return FB2_OK;
    const int nq = cv->nq, nb = cv->nb, ng = cv->ngeo, rd = cv->rdim;
    const int o_N = nq, o_dN = o_N + nq * nb, o_M = o_dN + nq * nb * rd, o_dM = o_M + nq * ng;
    std::vector<double> h((size_t)o_dM + (size_t)nq * ng * rd);
    memcpy(h.data(), cv->w.data(), sizeof(double) * nq);
    memcpy(h.data() + o_N, cv->N.data(), sizeof(double) * nq * nb);
    memcpy(h.data() + o_dN, cv->dN.data(), sizeof(double) * nq * nb * rd);
    memcpy(h.data() + o_M, cv->M.data(), sizeof(double) * nq * ng);
    memcpy(h.data() + o_dM, cv->dM.data(), sizeof(double) * nq * ng * rd);
    FB2_CUDA(cudaMalloc(&cv->d_tables, h.size() * sizeof(double)));
    FB2_CUDA(cudaMemcpy(cv->d_tables, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    { int dev = -1; cudaGetDevice(&dev); cv->tables_device = dev; }
    cv->tables_count = h.size();
    return FB2_OK;
}

extern "C" int fb2_spatial_coordinates(fb2_cv* cv, fb2_grid* g, const int64_t* cells, int64_t n, double* x_dev) {
    FB2_CHECK(cv && g && n >= 0, FB2_ERR_BAD_ARG, "fb2_spatial_coordinates: bad argument");
    if (n == 0) return FB2_OK;
    FB2_CHECK(x_dev, FB2_ERR_BAD_ARG, "fb2_spatial_coordinates: null output");
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(cv->celltype == g->celltype && cv->ngeo == g->nnpc, FB2_ERR_BAD_ARG, "fb2_spatial_coordinates: CellValues do not match the grid");
    if (cells) for (int64_t k = 0; k < n; ++k) FB2_CHECK(cells[k] >= 1 && cells[k] <= g->ncells, FB2_ERR_BAD_ARG, "fb2_spatial_coordinates: cell %lld out of range", (long long)cells[k]);
    else FB2_CHECK(n <= g->ncells, FB2_ERR_BAD_ARG, "fb2_spatial_coordinates: more cells than the grid has");
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(ensure_cv_tables(cv));
    const int nq = cv->nq, nb = cv->nb, ng = cv->ngeo, rd = cv->rdim;
    const int o_M = nq + nq * nb + nq * nb * rd;
    int64_t* d_cells = nullptr;
    if (cells) {
        FB2_CUDA(cudaMalloc(&d_cells, n * sizeof(int64_t)));
        FB2_CUDA(cudaMemcpyAsync(d_cells, cells, n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    k_spatial_coordinates<<<(unsigned)((n * nq + 127) / 128), 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, g->ncells_pad, g->xstride, g->sdim,
                                                                                     d_cells, n, cv->d_tables, o_M, nq, ng, x_dev);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (d_cells) { cudaStreamSynchronize(ctx->stream); cudaFree(d_cells); }
    FB2_CUDA(e);
    return FB2_OK;
}

extern "C" int fb2_function_values(fb2_cv* cv, fb2_dh* dh, const double* u_dev, double* values_dev, double* gradients_dev) {
    FB2_CHECK(cv && dh && u_dev && (values_dev || gradients_dev), FB2_ERR_BAD_ARG, "fb2_function_values: bad argument");
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    FB2_CHECK(cv->celltype == g->celltype && cv->rdim == g->sdim && cv->ngeo == g->nnpc, FB2_ERR_BAD_ARG, "fb2_function_values: CellValues do not match the grid");
    FB2_CHECK(dh->fields.size() == 1 && dh->ndpc == cv->nb * cv->vdim, FB2_ERR_BAD_ARG, "fb2_function_values: CellValues must cover the (single) field");
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(ensure_cv_tables(cv));
    const int nq = cv->nq, nb = cv->nb, ng = cv->ngeo, rd = cv->rdim;
    const int o_N = nq, o_dN = o_N + nq * nb, o_M = o_dN + nq * nb * rd, o_dM = o_M + nq * ng;
    const int64_t n = g->ncells;
    const unsigned grid = (unsigned)((n * nq + 127) / 128);
    if (rd == 1) k_function_values<1><<<grid, 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, dh->d_cell_dofs, g->ncells_pad, g->xstride, n, cv->d_tables, o_N, o_dN, o_dM, nq, nb, cv->vdim, ng, u_dev, values_dev, gradients_dev, ctx->d_errflag);
    else if (rd == 2) k_function_values<2><<<grid, 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, dh->d_cell_dofs, g->ncells_pad, g->xstride, n, cv->d_tables, o_N, o_dN, o_dM, nq, nb, cv->vdim, ng, u_dev, values_dev, gradients_dev, ctx->d_errflag);
    else k_function_values<3><<<grid, 128, 0, ctx->stream>>>(g->d_conn, g->d_xyz, dh->d_cell_dofs, g->ncells_pad, g->xstride, n, cv->d_tables, o_N, o_dN, o_dM, nq, nb, cv->vdim, ng, u_dev, values_dev, gradients_dev, ctx->d_errflag);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return FB2_OK;
}
