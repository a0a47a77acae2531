// Two-field (mixed u-p) assembly: MultiFieldCellValues + the element routine of the incompressible-elasticity tutorial.
//
// Reference: MultiFieldCellValues(qr, (u = ip_u, p = ip_p)) (src/FEValues/CellValues.jl:229-298) evaluates several
// interpolations on ONE quadrature rule and geometric mapping per reinit!; the element routine assemble_up!
// (docs/src/literate-tutorials/incompressible_elasticity.jl:266-311) integrates
//     K_uu[I,J] = int 2G dev3d(sym grad phi_I) : dev3d(sym grad phi_J),   K_pu[i,J] = -int psi_i div phi_J,
//     K_pp[i,j] = -int psi_i psi_j / K                                     (lower triangle, then symmetrised)
// with the local dof layout of a two-field DofHandler (all u dofs, then all p dofs: dof_range, src/Dofs/DofHandler.jl:1173-1187).
// dev3d: in 2-D (plane strain) the deviator is taken of the strain embedded in 3-D, i.e. the trace term keeps its factor 1/3.
//
// Here: one thread per (cell, local row).  The thread redoes the geometry of its cell per quadrature point (J from the
// geometric table shared by both fields, det > 0, inverse), forms the physical gradients of the displacement basis on the
// fly and accumulates its row of Ke in registers / local memory, then scatters it through the assembler's offset map
// (zero skip, missing-entry error as src/assembler.jl:347-457).  Mixed elements have 15 .. 34 dofs per cell and the path is
// not one of the benchmark configurations, so this kernel favours simplicity; the fused single-field kernels are elsewhere.
#include <cstring>

#include "common.h"

namespace {

struct MixedArgs {
    const int32_t* conn;
    const double* xyz;
    const int32_t* cell_dofs;
    const int64_t* colptr;
    const uint16_t* map;       // [n*n][ncells_pad], e = j * n + i
    int64_t ncells, ncells_pad;
    const double* tab_u;       // [w | N | dN | M | dM] of the displacement CellValues (geometry + weights are taken from it)
    const double* tab_p;       // same layout for the pressure CellValues
    int nq, nbu, nbp, ngeo, xstride, n, ou, op;   // ou / op: local offsets of the u / p dofs inside a cell
    int o_dNu, o_dMu, o_Np;    // table offsets
    double G, invK;
    double* nzval;
    int* errflag;
};

template <int DIM>
__global__ void k_cell_mixed_up(const MixedArgs A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.ncells * A.n) return;
    const int64_t cell = t / A.n;
    const int row = (int)(t - cell * A.n);
    constexpr int NMAX = 40, NGMAX = 8;
    double Ke[NMAX];
    for (int j = 0; j < A.n; ++j) Ke[j] = 0.0;
    double x[NGMAX][DIM];
    for (int j = 0; j < A.ngeo; ++j) {
        const int node = A.conn[(size_t)j * A.ncells_pad + cell];
        for (int a = 0; a < DIM; ++a) x[j][a] = A.xyz[(size_t)node * A.xstride + a];
    }
    const bool urow = row >= A.ou && row < A.ou + A.nbu * DIM;
    const int ra = urow ? (row - A.ou) / DIM : row - A.op, rc = urow ? (row - A.ou) % DIM : 0;
    const double* tw = A.tab_u;
    for (int q = 0; q < A.nq; ++q) {
        double J[DIM][DIM], Ji[DIM][DIM];
        for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b) J[a][b] = 0.0;
        for (int j = 0; j < A.ngeo; ++j)
            for (int a = 0; a < DIM; ++a)
                for (int b = 0; b < DIM; ++b) J[a][b] += x[j][a] * A.tab_u[A.o_dMu + (q * A.ngeo + j) * DIM + b];
        double det;
        if (DIM == 2) {
            det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            const double r = 1.0 / det;
            Ji[0][0] = J[1][1] * r; Ji[0][1] = -J[0][1] * r; Ji[1][0] = -J[1][0] * r; Ji[1][1] = J[0][0] * r;
        } else {
            const double c00 = J[1][1] * J[2 % DIM][2 % DIM] - J[1][2 % DIM] * J[2 % DIM][1], c01 = J[1][0] * J[2 % DIM][2 % DIM] - J[1][2 % DIM] * J[2 % DIM][0],
                         c02 = J[1][0] * J[2 % DIM][1] - J[1][1] * J[2 % DIM][0];
            det = J[0][0] * c00 - J[0][1] * c01 + J[0][2 % DIM] * c02;
            const double r = 1.0 / det;
            Ji[0][0] = c00 * r;
            Ji[0][1] = -(J[0][1] * J[2 % DIM][2 % DIM] - J[0][2 % DIM] * J[2 % DIM][1]) * r;
            Ji[0][2 % DIM] = (J[0][1] * J[1][2 % DIM] - J[0][2 % DIM] * J[1][1]) * r;
            Ji[1][0] = -c01 * r;
            Ji[1][1] = (J[0][0] * J[2 % DIM][2 % DIM] - J[0][2 % DIM] * J[2 % DIM][0]) * r;
            Ji[1][2 % DIM] = -(J[0][0] * J[1][2 % DIM] - J[0][2 % DIM] * J[1][0]) * r;
            Ji[2 % DIM][0] = c02 * r;
            Ji[2 % DIM][1] = -(J[0][0] * J[2 % DIM][1] - J[0][1] * J[2 % DIM][0]) * r;
            Ji[2 % DIM][2 % DIM] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
        }
        if (!(det > 0.0)) {
            if (atomicCAS(&A.errflag[0], 0, FB2_ERR_DETJ_NOT_POSITIVE) == 0) A.errflag[1] = (int)cell;
            return;
        }
        const double dO = det * tw[q];
        // physical gradient of displacement basis function b: g_b = dN_b/dxi . inv(J)
        auto grad = [&](int b, double (&g)[DIM]) {
            for (int d = 0; d < DIM; ++d) {
                double s = 0.0;
                for (int a = 0; a < DIM; ++a) s += A.tab_u[A.o_dNu + (q * A.nbu + b) * DIM + a] * Ji[a][d];
                g[d] = s;
            }
        };
        if (urow) {
            double ga[DIM];
            grad(ra, ga);
            for (int b = 0; b < A.nbu; ++b) {
                double gb[DIM];
                grad(b, gb);
                double dot = 0.0;
                for (int k = 0; k < DIM; ++k) dot += ga[k] * gb[k];
                for (int d = 0; d < DIM; ++d) {
                    // dev3d(sym(e_c (x) g_a)) : dev3d(sym(e_d (x) g_b)) = (delta_cd g_a.g_b + g_a[d] g_b[c]) / 2 - g_a[c] g_b[d] / 3
                    const double v = 0.5 * ((rc == d ? dot : 0.0) + ga[d] * gb[rc]) - ga[rc] * gb[d] / 3.0;
                    Ke[A.ou + b * DIM + d] += 2.0 * A.G * v * dO;
                }
            }
            for (int i = 0; i < A.nbp; ++i) Ke[A.op + i] -= A.tab_p[A.o_Np + q * A.nbp + i] * ga[rc] * dO;   // K_up = K_pu'
        } else {
            const double psi = A.tab_p[A.o_Np + q * A.nbp + ra];
            for (int b = 0; b < A.nbu; ++b) {
                double gb[DIM];
                grad(b, gb);
                for (int d = 0; d < DIM; ++d) Ke[A.ou + b * DIM + d] -= psi * gb[d] * dO;
            }
            for (int j = 0; j < A.nbp; ++j) Ke[A.op + j] -= A.invK * psi * A.tab_p[A.o_Np + q * A.nbp + j] * dO;
        }
    }
    bool missing = false;
    for (int j = 0; j < A.n; ++j) {
        const double v = Ke[j];
        if (v == 0.0) continue;
        const unsigned off = A.map[(size_t)(j * A.n + row) * A.ncells_pad + cell];
        if (off == 0xFFFFu) { missing = true; continue; }
        atomicAdd(A.nzval + A.colptr[A.cell_dofs[(size_t)j * A.ncells_pad + cell]] + off, v);
    }
    if (missing && atomicCAS(&A.errflag[0], 0, FB2_ERR_MISSING_PATTERN_ENTRY) == 0) A.errflag[1] = (int)cell;
}

int ensure_tables(fb2_cv* cv, int device) {
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
    cv->tables_device = device;
    cv->tables_count = h.size();
    return FB2_OK;
}

}  // namespace

// `a` = fb2_assembler_create(dh, pattern, NULL) of a two-field DofHandler (the map covers all local dofs); cv_u / cv_p =
// CellValues of the displacement (vdim == dim) and pressure (scalar) interpolations on the SAME quadrature rule and geometric
// interpolation = the two members of MultiFieldCellValues(qr, (u = ip_u, p = ip_p)).  field_u / field_p: their field indices
// in the DofHandler (add! order).  K (the bulk modulus) may be infinite: pass inv_bulk = 0.
extern "C" int fb2_assemble_mixed_up(fb2_assembler* a, fb2_cv* cv_u, fb2_cv* cv_p, int field_u, int field_p, double shear_G, double inv_bulk,
                                     double* nzval_dev, double* f_dev, const fb2_asm_opts* opts) {
    FB2_CHECK(a && cv_u && cv_p && nzval_dev, FB2_ERR_BAD_ARG, "fb2_assemble_mixed_up: null argument");
    fb2_dh* dh = a->dh;
    fb2_grid* g = dh->grid;
    fb2_ctx* ctx = g->ctx;
    FB2_NEED_DEVICE(ctx);
    const int nf = (int)dh->fields.size();
    FB2_CHECK(field_u >= 0 && field_u < nf && field_p >= 0 && field_p < nf && field_u != field_p, FB2_ERR_BAD_ARG, "fb2_assemble_mixed_up: bad field indices");
    const int dim = g->sdim;
    FB2_CHECK(dim == 2 || dim == 3, FB2_ERR_UNSUPPORTED, "fb2_assemble_mixed_up: 2-D and 3-D only");
    FB2_CHECK(cv_u->celltype == g->celltype && cv_p->celltype == g->celltype && cv_u->rdim == dim, FB2_ERR_BAD_ARG, "fb2_assemble_mixed_up: CellValues do not match the grid");
    FB2_CHECK(cv_u->vdim == dim && cv_p->vdim == 1, FB2_ERR_BAD_ARG, "fb2_assemble_mixed_up: u must have dim components, p one");
    FB2_CHECK(cv_u->nb * dim == dh->ips[field_u].nbase * dh->fields[field_u].vdim && cv_p->nb == dh->ips[field_p].nbase * dh->fields[field_p].vdim,
              FB2_ERR_BAD_ARG, "fb2_assemble_mixed_up: the CellValues do not match the fields of the DofHandler");
    // MultiFieldCellValues: one quadrature rule and one geometric mapping for all fields (src/FEValues/CellValues.jl:229-298)
    FB2_CHECK(cv_u->nq == cv_p->nq && cv_u->ngeo == cv_p->ngeo && cv_u->w == cv_p->w && cv_u->dM == cv_p->dM, FB2_ERR_BAD_ARG,
              "fb2_assemble_mixed_up: both CellValues must share the quadrature rule and the geometric interpolation");
    FB2_CHECK(a->n == dh->ndpc && dh->ndpc <= 40 && cv_u->ngeo <= 8, FB2_ERR_UNSUPPORTED, "fb2_assemble_mixed_up: at most 40 dofs and 8 geometric nodes per cell");
    FB2_CHECK(cv_u->ngeo == g->nnpc, FB2_ERR_BAD_ARG, "fb2_assemble_mixed_up: geometric interpolation has %d nodes, cells have %d", cv_u->ngeo, g->nnpc);
    FB2_CUDA(cudaSetDevice(ctx->device));
    FB2_TRY(ensure_tables(cv_u, ctx->device));
    FB2_TRY(ensure_tables(cv_p, ctx->device));
    fb2_asm_opts o = {1, FB2_SCATTER_ATOMIC, 0, 0};
    if (opts) o = *opts;
    FB2_CHECK(o.scatter_mode == FB2_SCATTER_ATOMIC, FB2_ERR_UNSUPPORTED, "fb2_assemble_mixed_up: atomic scatter only");
    if (o.fillzero) {
        FB2_CUDA(cudaMemsetAsync(nzval_dev, 0, (size_t)a->pat->nnz * sizeof(double), ctx->stream));
        if (f_dev) FB2_CUDA(cudaMemsetAsync(f_dev, 0, (size_t)dh->ndofs * sizeof(double), ctx->stream));
    }
    MixedArgs A;
    memset(&A, 0, sizeof(A));
    A.conn = g->d_conn; A.xyz = g->d_xyz; A.cell_dofs = dh->d_cell_dofs; A.colptr = a->pat->d_colptr; A.map = a->d_map;
    A.ncells = g->ncells; A.ncells_pad = g->ncells_pad;
    A.tab_u = cv_u->d_tables; A.tab_p = cv_p->d_tables;
    A.nq = cv_u->nq; A.nbu = cv_u->nb; A.nbp = cv_p->nb; A.ngeo = cv_u->ngeo; A.xstride = g->xstride; A.n = dh->ndpc;
    A.ou = dh->field_offset(field_u); A.op = dh->field_offset(field_p);
    A.o_dNu = cv_u->nq + cv_u->nq * cv_u->nb;
    A.o_dMu = A.o_dNu + cv_u->nq * cv_u->nb * dim + cv_u->nq * cv_u->ngeo;
    A.o_Np = cv_p->nq;
    A.G = shear_G; A.invK = inv_bulk;
    A.nzval = nzval_dev; A.errflag = ctx->d_errflag;
    const int64_t total = g->ncells * (int64_t)dh->ndpc;
    if (total == 0) return FB2_OK;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (dim == 2) k_cell_mixed_up<2><<<grid, 128, 0, ctx->stream>>>(A);
    else k_cell_mixed_up<3><<<grid, 128, 0, ctx->stream>>>(A);
    ctx->launches++;
    FB2_CUDA(cudaGetLastError());
    return fb2_check_device_error(ctx);
}
