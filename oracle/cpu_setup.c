/*
 * CPU restatement of the reference's SET-UP for structured hexahedral Q1 problems in plain C (oracle; TEST INFRASTRUCTURE
 * ONLY): lets bench.py's `--impl reference` arm and `cpu_baseline` time the reference algorithm on the FULL 200^3
 * configuration -- the numpy oracle needs minutes to set that size up.  Never linked into or called from the product library.
 *
 * Follows generate_grid(Hexahedron, nel, left, right) (src/Grid/grid_generators.jl:159-203; nodes :550-559: xi = 2 (idx-1) /
 * (nn-1) - 1, x = sum_c M_c(xi) corner_c, first index fastest), the deterministic interior-node perturbation of the
 * benchmark inputs (oracle/grid.py perturb_grid, same splitmix64 hash), close!(dh) for one first-order Lagrange field of
 * `vdim` components (src/Dofs/DofHandler.jl:493-794: cells ascending, vertices in local order, a new vertex takes vdim
 * consecutive dofs) and allocate_matrix(dh) (src/Dofs/sparsity_pattern.jl:1136-1204: column j = ascending unique rows coupled
 * through some cell).  Checked against the numpy oracle at small sizes in tests/test_oracle_goldens.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double hash01(uint64_t id, uint64_t salt) {
    uint64_t x = id * 4u + salt;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

void oracle_hex_q1_sizes(int64_t nx, int64_t ny, int64_t nz, int vdim, int64_t* nnodes, int64_t* ncells, int64_t* ndofs, int64_t* nnz) {
    *nnodes = (nx + 1) * (ny + 1) * (nz + 1);
    *ncells = nx * ny * nz;
    *ndofs = *nnodes * vdim;
    *nnz = (int64_t)vdim * vdim * (3 * nx + 1) * (3 * ny + 1) * (3 * nz + 1);
}

static int cmp_i64(const void* a, const void* b) {
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* cells: 8 x ncells (1-based, column-major), xyz: 3 x nnodes, cell_dofs: 8*vdim x ncells (1-based), colptr: ndofs + 1,
 * rowval: nnz (both 1-based).  Returns 0, or -1 if the pattern did not come out with the expected nnz. */
int oracle_hex_q1_setup(int64_t nx, int64_t ny, int64_t nz, const double* left, const double* right, double amplitude, int vdim,
                        int64_t* cells, double* xyz, int64_t* cell_dofs, int64_t* colptr, int64_t* rowval) {
    const int64_t nnx = nx + 1, nny = ny + 1, nnz_ = nz + 1, nn = nnx * nny * nnz_, nc = nx * ny * nz;
    /* corners in the vertex order of the linear hexahedron (_extrema_to_corners, :565-578) */
    static const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    double corner[8][3];
    for (int c = 0; c < 8; ++c) {
        const int s[3] = {sx[c], sy[c], sz[c]};
        for (int d = 0; d < 3; ++d) corner[c][d] = left[d] + ((right[d] - left[d]) / 2.0) * ((double)s[d] - (-1.0));
    }
    const double h[3] = {(right[0] - left[0]) / nx, (right[1] - left[1]) / ny, (right[2] - left[2]) / nz};
#pragma omp parallel for
    for (int64_t n = 0; n < nn; ++n) {
        const int64_t i = n % nnx, j = (n / nnx) % nny, k = n / (nnx * nny);
        const double xi[3] = {2.0 * (double)i / (double)(nnx - 1) - 1.0, 2.0 * (double)j / (double)(nny - 1) - 1.0, 2.0 * (double)k / (double)(nnz_ - 1) - 1.0};
        double x[3] = {0, 0, 0};
        for (int c = 0; c < 8; ++c) {
            double M = 1.0;
            M = M * (sx[c] > 0 ? 1 + xi[0] : 1 - xi[0]);
            M = M * (sy[c] > 0 ? 1 + xi[1] : 1 - xi[1]);
            M = M * (sz[c] > 0 ? 1 + xi[2] : 1 - xi[2]);
            M = M / 8.0;
            for (int d = 0; d < 3; ++d) x[d] += M * corner[c][d];
        }
        const int interior = i > 0 && i < nnx - 1 && j > 0 && j < nny - 1 && k > 0 && k < nnz_ - 1;
        for (int d = 0; d < 3; ++d) xyz[3 * n + d] = x[d] + (interior ? amplitude * h[d] * (hash01((uint64_t)(n + 1), (uint64_t)d) - 0.5) : 0.0);
    }
    /* cells pushed for k, j, i with local nodes (i,j,k),(i+1,j,k),(i+1,j+1,k),(i,j+1,k), then the same at k+1 (:170-178) */
    static const int ox[8] = {0, 1, 1, 0, 0, 1, 1, 0}, oy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, oz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma omp parallel for
    for (int64_t c = 0; c < nc; ++c) {
        const int64_t i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
        for (int v = 0; v < 8; ++v) cells[8 * c + v] = 1 + (i + ox[v]) + nnx * ((j + oy[v]) + nny * (k + oz[v]));
    }
    /* close!: first visit of a vertex (cells ascending, local order) hands out vdim consecutive dofs */
    int64_t* node_dof = (int64_t*)calloc((size_t)nn, sizeof(int64_t));
    if (!node_dof) return -2;
    int64_t next = 1;
    for (int64_t c = 0; c < nc; ++c)
        for (int v = 0; v < 8; ++v) {
            const int64_t n = cells[8 * c + v] - 1;
            if (node_dof[n] == 0) { node_dof[n] = next; next += vdim; }
            for (int d = 0; d < vdim; ++d) cell_dofs[(size_t)(8 * vdim) * c + v * vdim + d] = node_dof[n] + d;
        }
    /* allocate_matrix: the dofs of node n couple with the dofs of its <= 27 neighbour nodes; column = sorted rows.  Two
     * passes over the nodes in DOF order (node_of[first dof / vdim]) so that colptr is cumulative. */
    int64_t* node_of = (int64_t*)malloc((size_t)nn * sizeof(int64_t));
    if (!node_of) { free(node_dof); return -2; }
    for (int64_t n = 0; n < nn; ++n) node_of[(node_dof[n] - 1) / vdim] = n;
    colptr[0] = 1;
    for (int64_t q = 0; q < nn; ++q) {
        const int64_t n = node_of[q], i = n % nnx, j = (n / nnx) % nny, k = n / (nnx * nny);
        const int64_t cnt = ((i > 0) + 1 + (i < nnx - 1)) * ((j > 0) + 1 + (j < nny - 1)) * ((k > 0) + 1 + (k < nnz_ - 1)) * vdim;
        for (int d = 0; d < vdim; ++d) colptr[q * vdim + d + 1] = colptr[q * vdim + d] + cnt;
    }
#pragma omp parallel for
    for (int64_t q = 0; q < nn; ++q) {
        const int64_t n = node_of[q], i = n % nnx, j = (n / nnx) % nny, k = n / (nnx * nny);
        int64_t rows[81];
        int m = 0;
        for (int dk = -1; dk <= 1; ++dk)
            for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                    const int64_t a = i + di, b = j + dj, c = k + dk;
                    if (a < 0 || a >= nnx || b < 0 || b >= nny || c < 0 || c >= nnz_) continue;
                    const int64_t nd = node_dof[a + nnx * (b + nny * c)];
                    for (int d = 0; d < vdim; ++d) rows[m++] = nd + d;
                }
        qsort(rows, (size_t)m, sizeof(int64_t), cmp_i64);
        for (int d = 0; d < vdim; ++d) memcpy(rowval + (colptr[q * vdim + d] - 1), rows, (size_t)m * sizeof(int64_t));
    }
    const int64_t total = colptr[nn * vdim] - 1;
    free(node_of);
    free(node_dof);
    return total == (int64_t)vdim * vdim * (3 * nx + 1) * (3 * ny + 1) * (3 * nz + 1) ? 0 : -1;
}
