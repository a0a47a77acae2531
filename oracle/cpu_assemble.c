/*
 * CPU restatement of the reference's assembly loop in plain C (oracle; TEST INFRASTRUCTURE ONLY).
 *
 * Used (a) to cross-check the numpy oracle and (b) as the timed CPU baseline of bench.py
 * (`cpu_baseline`, `--impl reference`): the reference itself is Julia and cannot run in this image.
 * Never linked into or called from the product library.
 *
 * Follows, per cell (docs/src/literate-howto/threaded_assembly.jl:321-349, the reference's "atomic"
 * threaded loop): reinit!(cc, i) gather (src/iterators.jl:72-87), reinit!(cv, cc) per quadrature point
 * (src/FEValues/CellValues.jl:122-140: J = sum x_j (x) dMdxi, det > 0, dOmega = det*w, dNdx = dNdxi . inv(J)),
 * the tutorial element routine with the full i,j double loop
 * (heat_equation.jl:143-164 / threaded_assembly.jl:105-119), then assemble! (src/assembler.jl:322-331,347-457):
 * f[dofs] += fe, sort the dofs with a permutation, and for each sorted column merge-walk the column's
 * rowval segment against the sorted row dofs, nzval[k] += Ke[perm[r], perm[c]] (atomic add when threaded,
 * src/arrayutils.jl:57-74).
 *
 * Layouts: cells nnpc x ncells (1-based, column-major), xyz sdim x nnodes, cell_dofs n x ncells (1-based),
 * colptr/rowval 1-based Int64, tables q-major: N[q][i], dN[q][i][d], dM[q][j][d], w[q].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 96   /* max dofs per cell */
#define MAXG 8    /* max geometric nodes */

static double det_inv(int dim, const double* J, double* Ji) {
    if (dim == 2) {
        double det = J[0] * J[3] - J[1] * J[2];
        Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
        return det;
    }
    double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
    Ji[0] = (J[4] * J[8] - J[5] * J[7]) / det;
    Ji[1] = -(J[1] * J[8] - J[2] * J[7]) / det;
    Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
    Ji[3] = -(J[3] * J[8] - J[5] * J[6]) / det;
    Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det;
    Ji[5] = -(J[0] * J[5] - J[2] * J[3]) / det;
    Ji[6] = (J[3] * J[7] - J[4] * J[6]) / det;
    Ji[7] = -(J[0] * J[7] - J[1] * J[6]) / det;
    Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
    return det;
}

/* insertion sort of the dofs with permutation (sortperm2!, src/assembler.jl:513-576) */
static void sortperm(int n, const int64_t* dofs, int64_t* sorted, int* perm) {
    for (int i = 0; i < n; ++i) { sorted[i] = dofs[i]; perm[i] = i; }
    for (int i = 1; i < n; ++i) {
        int64_t v = sorted[i];
        int p = perm[i], j = i - 1;
        while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; perm[j + 1] = perm[j]; --j; }
        sorted[j + 1] = v; perm[j + 1] = p;
    }
}

/* element: 1 = heat (params: k, source), 3 = elasticity (params: lambda, mu, b[3]) ; returns 0 or the 1-based id of a bad cell */
/* Coloured variant (docs/src/literate-howto/threaded_assembly.jl:330-377: colours one after the other, the cells of a colour in
   parallel, plain adds): ncolors > 0, color_ptr[ncolors + 1] (0-based offsets) into color_cells (0-based cell ids).
   ncolors == 0: all cells in one parallel sweep with atomic adds (the how-to's other scheme, :308-310). */
int64_t oracle_assemble_colored(int element, int dim, int ngeo, int nbs, int vdim, int nq, int64_t ncells, const int64_t* cells,
                                const double* xyz, const int64_t* cell_dofs, const int64_t* colptr, const int64_t* rowval,
                                const double* N, const double* dN, const double* dM, const double* w, const double* params,
                                const double* u, double* nzval, double* f, int nthreads, int ncolors, const int64_t* color_ptr,
                                const int64_t* color_cells) {
    const int n = nbs * vdim;
    int64_t bad = 0;
    if (n > MAXN || ngeo > MAXG) return -1;
    const int use_atomic = ncolors <= 0;
    const int nphases = use_atomic ? 1 : ncolors;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double* Ke = (double*)malloc(sizeof(double) * n * n);
        double fe[MAXN], g[MAXN * 3], x[MAXG * 3];
        int64_t dofs[MAXN], sorted[MAXN];
        int perm[MAXN];
        for (int phase = 0; phase < nphases; ++phase) {
        const int64_t i0 = use_atomic ? 0 : color_ptr[phase], i1 = use_atomic ? ncells : color_ptr[phase + 1];
#pragma omp for schedule(static)
        for (int64_t idx = i0; idx < i1; ++idx) {
            const int64_t ci = use_atomic ? idx : color_cells[idx];
            /* reinit!(cc, i) */
            for (int j = 0; j < ngeo; ++j) {
                int64_t node = cells[ci * ngeo + j] - 1;
                for (int d = 0; d < dim; ++d) x[j * dim + d] = xyz[node * dim + d];
            }
            for (int i = 0; i < n; ++i) dofs[i] = cell_dofs[ci * n + i];
            memset(Ke, 0, sizeof(double) * n * n);
            memset(fe, 0, sizeof(double) * n);
            for (int q = 0; q < nq; ++q) {
                /* reinit!(cv, cc) at this quadrature point */
                double J[9] = {0}, Ji[9];
                for (int j = 0; j < ngeo; ++j)
                    for (int a = 0; a < dim; ++a)
                        for (int b = 0; b < dim; ++b) J[a * dim + b] += x[j * dim + a] * dM[(q * ngeo + j) * dim + b];
                double det = det_inv(dim, J, Ji);
                if (!(det > 0.0)) {
#pragma omp critical
                    if (!bad) bad = ci + 1;
                }
                double dO = det * w[q];
                for (int i = 0; i < nbs; ++i)
                    for (int b = 0; b < dim; ++b) {
                        double s = 0;
                        for (int a = 0; a < dim; ++a) s += dN[(q * nbs + i) * dim + a] * Ji[a * dim + b];
                        g[i * dim + b] = s;
                    }
                /* element routine */
                if (element == 1) {
                    for (int i = 0; i < n; ++i) {
                        fe[i] += params[1] * N[q * nbs + i] * dO;
                        for (int j = 0; j < n; ++j) {
                            double s = 0;
                            for (int b = 0; b < dim; ++b) s += g[i * dim + b] * g[j * dim + b];
                            Ke[j * n + i] += params[0] * s * dO;
                        }
                    }
                } else if (element == 4) {
                    /* Neo-Hooke tangent + residual, docs/src/literate-tutorials/hyperelasticity.jl:162-176,241-276:
                       Psi = mu/2 (Ic - 3 - 2 ln J) + lam/2 (J - 1)^2, S and dS/dC in closed form (dim == vdim == 3) */
                    const double lam = params[0], mu = params[1];
                    double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Cm[9], Ci[9], S[9], P[9], A[81];
                    for (int a = 0; a < nbs; ++a)
                        for (int c = 0; c < 3; ++c) {
                            double uc = u[dofs[a * 3 + c] - 1];
                            for (int b = 0; b < 3; ++b) F[c * 3 + b] += uc * g[a * 3 + b];
                        }
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) Cm[i * 3 + j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
                    double detC = det_inv(3, Cm, Ci);
                    if (!(detC > 0.0)) {
#pragma omp critical
                        if (!bad) bad = ci + 1;
                    }
                    double Jd = sqrt(detC), cS = lam * Jd * (Jd - 1.0);
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) S[i * 3 + j] = mu * ((i == j) - Ci[i * 3 + j]) + cS * Ci[i * 3 + j];
                    double c1 = (mu - cS) * 0.5, c2 = lam * (2.0 * Jd - 1.0) * (Jd * 0.5);
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) P[i * 3 + j] = F[i * 3] * S[j] + F[i * 3 + 1] * S[3 + j] + F[i * 3 + 2] * S[6 + j];
                    /* dP_ij/dF_mn = delta_im S_jn + 2 F_ia dSdC_ajkn F_mk,
                       dSdC_ajkn = c1 (Ci_ak Ci_nj + Ci_an Ci_kj) + c2 Ci_aj Ci_kn */
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j)
                            for (int m = 0; m < 3; ++m)
                                for (int nn = 0; nn < 3; ++nn) {
                                    double t = 0;
                                    for (int a = 0; a < 3; ++a)
                                        for (int k = 0; k < 3; ++k) {
                                            double d4 = c1 * (Ci[a * 3 + k] * Ci[nn * 3 + j] + Ci[a * 3 + nn] * Ci[k * 3 + j]) + c2 * Ci[a * 3 + j] * Ci[k * 3 + nn];
                                            t += F[i * 3 + a] * d4 * F[m * 3 + k];
                                        }
                                    A[((i * 3 + j) * 3 + m) * 3 + nn] = (i == m ? S[j * 3 + nn] : 0.0) + 2.0 * t;
                                }
                    for (int a = 0; a < nbs; ++a)
                        for (int c = 0; c < 3; ++c) {
                            int I = a * 3 + c;
                            double s = 0;
                            for (int j = 0; j < 3; ++j) s += g[a * 3 + j] * P[c * 3 + j];
                            fe[I] += (s - N[q * nbs + a] * params[2 + c]) * dO;
                            double h[9];   /* hoisted: grad(du_i) : dP/dF */
                            for (int d = 0; d < 3; ++d)
                                for (int nn = 0; nn < 3; ++nn) {
                                    double t = 0;
                                    for (int j = 0; j < 3; ++j) t += g[a * 3 + j] * A[((c * 3 + j) * 3 + d) * 3 + nn];
                                    h[d * 3 + nn] = t;
                                }
                            for (int b = 0; b < nbs; ++b)
                                for (int d = 0; d < 3; ++d) {
                                    double t = 0;
                                    for (int nn = 0; nn < 3; ++nn) t += h[d * 3 + nn] * g[b * 3 + nn];
                                    Ke[(b * 3 + d) * n + I] += t * dO;
                                }
                        }
                } else {
                    const double lam = params[0], mu = params[1];
                    for (int a = 0; a < nbs; ++a)
                        for (int c = 0; c < vdim; ++c) {
                            int I = a * vdim + c;
                            fe[I] += N[q * nbs + a] * params[2 + c] * dO;
                            for (int b = 0; b < nbs; ++b) {
                                double dot = 0;
                                for (int k = 0; k < dim; ++k) dot += g[a * dim + k] * g[b * dim + k];
                                for (int d = 0; d < vdim; ++d) {
                                    int Jx = b * vdim + d;
                                    double v = lam * g[a * dim + c] * g[b * dim + d] + mu * g[a * dim + d] * g[b * dim + c];
                                    if (c == d) v += mu * dot;
                                    Ke[Jx * n + I] += v * dO;
                                }
                            }
                        }
                }
            }
            /* assemble!(assembler, dofs, Ke, fe) */
            if (f)
                for (int i = 0; i < n; ++i) {
                    if (use_atomic) {
#pragma omp atomic
                        f[dofs[i] - 1] += fe[i];
                    } else {
                        f[dofs[i] - 1] += fe[i];
                    }
                }
            sortperm(n, dofs, sorted, perm);
            for (int c = 0; c < n; ++c) {
                int64_t col = sorted[c];
                int64_t k = colptr[col - 1] - 1, kend = colptr[col] - 1;
                int r = 0;
                while (r < n && k < kend) {   /* merge walk */
                    int64_t row = rowval[k];
                    if (row == sorted[r]) {
                        double v = Ke[perm[c] * n + perm[r]];
                        if (v != 0.0) {
                            if (use_atomic) {
#pragma omp atomic
                                nzval[k] += v;
                            } else {
                                nzval[k] += v;
                            }
                        }
                        ++r;
                        if (r < n && sorted[r] == sorted[r - 1]) continue;  /* duplicate dof: same k again */
                        ++k;
                    } else if (row < sorted[r]) {
                        ++k;
                    } else {
                        ++r;  /* missing entry: the reference errors for non-zero values; the test grids never hit this */
                    }
                }
            }
        }
        }   /* phase: the implicit barrier of the omp for separates the colours */
        free(Ke);
    }
    return bad;
}

int64_t oracle_assemble(int element, int dim, int ngeo, int nbs, int vdim, int nq, int64_t ncells, const int64_t* cells,
                        const double* xyz, const int64_t* cell_dofs, const int64_t* colptr, const int64_t* rowval,
                        const double* N, const double* dN, const double* dM, const double* w, const double* params,
                        const double* u, double* nzval, double* f, int nthreads) {
    return oracle_assemble_colored(element, dim, ngeo, nbs, vdim, nq, ncells, cells, xyz, cell_dofs, colptr, rowval, N, dN, dM, w,
                                   params, u, nzval, f, nthreads, 0, 0, 0);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
