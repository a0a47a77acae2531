// Fused per-cell assembly kernels (sm_100a).
//
// One launch replaces the reference's whole cell loop (docs/src/literate-tutorials/heat_equation.jl:
// 181-204): CellCache gather reinit!(cc, i) (src/iterators.jl:72-87), reinit!(cv, cc)
// (src/FEValues/CellValues.jl:122-140: J = sum_j x_j (x) dM_j/dxi GeometryMapping.jl:124-130, detJ > 0
// check CellValues.jl:110-115, dNdx = dNdxi . inv(J) FunctionValues.jl:186-192), the element routine
// (kernel menu, see ferrite_b200.h) and assemble! (src/assembler.jl:322-331, 347-457).  The scatter goes
// through the precomputed uint16 offset map: nzval[colptr[dof_j] + map[i,j]] += Ke[i,j], either with FP64
// RED atomics (any cell order) or with plain read-modify-write inside one colour.
//
// Two kernel families:
//   k_cell_scalar  -- one thread per cell, Ke (upper triangle) and fe in FP64 registers, quadrature loop
//                     fully unrolled, reference tables read as constant-bank operands.  Scalar fields with
//                     few dofs per cell (Q1 quad/hex, P1/P2 simplices).
//   k_cell_blocks  -- one CTA per batch of cells; phase A (thread per (cell, qp)) stages physical shape
//                     gradients and dOmega in shared memory (cell index fastest => conflict free), phase B
//                     (thread per (cell, row node a, tile of column nodes b)) integrates vdim x vdim node
//                     blocks in registers and scatters them.  Vector fields and higher order.
#pragma once
#include "common.h"

#define FB2_TAB_MAX 7168
__constant__ double c_tab[FB2_TAB_MAX];  // this header is included by assemble.cu only

struct AsmArgs {
    const int32_t* conn;
    const double* xyz;
    const int32_t* cell_dofs;
    const int64_t* colptr;
    const uint16_t* map;
    const uint4* map8;     // packed offsets for k_cell_scalar
    const int32_t* cells;  // optional indirection (colour / partition subset)
    int64_t ncount;        // number of cells handled by this launch
    int64_t ncells_pad;
    double* nzval;
    double* f;
    const double* u;
    int* errflag;
    double p[6];  // element parameters
    int nq;
    // table offsets into c_tab (in doubles)
    int o_w, o_N, o_dN, o_M, o_dM;
};

__device__ __forceinline__ void fb2_flag_error(int* errflag, int code, int64_t cell) {
    if (atomicCAS(&errflag[0], 0, code) == 0) errflag[1] = (int)cell;
}

template <bool ATOMIC>
__device__ __forceinline__ void fb2_add(double* p, double v) {
    if (ATOMIC) atomicAdd(p, v);  // result unused -> RED.E.ADD.F64
    else *p += v;
}

// Volatile read-only loads: ptxas keeps them in program order ahead of the REDs that follow, so a batch of
// index loads is in flight together instead of one memory round trip per scattered entry.
__device__ __forceinline__ int fb2_ldv_s32(const int32_t* p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned fb2_ldv_u16(const uint16_t* p) {
    unsigned short v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t fb2_ldv_s64(const int64_t* p) {
    int64_t v;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

template <int DIM>
__device__ __forceinline__ void fb2_load_x(const double* __restrict__ xyz, int node, double* x) {
    if (DIM == 3) {
        const double2* p = reinterpret_cast<const double2*>(xyz + 4 * (size_t)node);
        double2 a = __ldg(p), b = __ldg(p + 1);
        x[0] = a.x; x[1] = a.y; x[2] = b.x;
    } else if (DIM == 2) {
        double2 a = __ldg(reinterpret_cast<const double2*>(xyz + 2 * (size_t)node));
        x[0] = a.x; x[1] = a.y;
    } else {
        x[0] = __ldg(xyz + node);
    }
}

// J (DIM x DIM) -> det, inverse (Tensors.jl closed forms: Sarrus / cofactors)
template <int DIM>
__device__ __forceinline__ double fb2_det_inv(const double (&J)[DIM][DIM], double (&Ji)[DIM][DIM]) {
    if (DIM == 2) {
        double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        double r = 1.0 / det;
        Ji[0][0] = J[1][1] * r; Ji[0][1] = -J[0][1] * r;
        Ji[1][0] = -J[1][0] * r; Ji[1][1] = J[0][0] * r;
        return det;
    } else {
        double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        double c01 = J[1][0] * J[2][2] - J[1][2] * J[2][0];
        double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        double det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02;
        double r = 1.0 / det;
        Ji[0][0] = c00 * r;
        Ji[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]) * r;
        Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
        Ji[1][0] = -c01 * r;
        Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
        Ji[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]) * r;
        Ji[2][0] = c02 * r;
        Ji[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]) * r;
        Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
        return det;
    }
}

// Element integration shared by the thread-per-cell kernels: per quadrature point J = sum_j x_j (x) dM_j/dxi, det > 0,
// dOmega = det*w, dNdx = dNdxi . inv(J); upper triangle of Ke (packed: (i,j), i <= j at j(j+1)/2 + i) and fe in registers.
// Returns true if some det(J) was not positive.
template <int DIM, int NGEO, int NB, int NQ, int ELEM>
__device__ __forceinline__ bool fb2_scalar_element(const AsmArgs& A, const double (&x)[NGEO][DIM],
                                                   double (&Ke)[NB * (NB + 1) / 2], double (&fe)[NB]) {
    constexpr int NSYM = NB * (NB + 1) / 2;
#pragma unroll
    for (int i = 0; i < NSYM; ++i) Ke[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) fe[i] = 0.0;

    const double* tw = c_tab + A.o_w;
    const double* tN = c_tab + A.o_N;
    const double* tdN = c_tab + A.o_dN;
    const double* tdM = c_tab + A.o_dM;
    bool bad = false;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double J[DIM][DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = 0; b < DIM; ++b) J[a][b] = 0.0;
#pragma unroll
        for (int j = 0; j < NGEO; ++j)
#pragma unroll
            for (int a = 0; a < DIM; ++a)
#pragma unroll
                for (int b = 0; b < DIM; ++b) J[a][b] = fma(x[j][a], tdM[(q * NGEO + j) * DIM + b], J[a][b]);
        double Ji[DIM][DIM];
        double det = fb2_det_inv<DIM>(J, Ji);
        bad |= !(det > 0.0);
        const double dO = det * tw[q];
        if (ELEM == FB2_ELEM_HEAT) {
            double g[NB][DIM];
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int b = 0; b < DIM; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) s = fma(tdN[(q * NB + i) * DIM + a], Ji[a][b], s);
                    g[i][b] = s;
                }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                double gs[DIM];
#pragma unroll
                for (int b = 0; b < DIM; ++b) gs[b] = g[j][b] * dO;
#pragma unroll
                for (int i = 0; i <= j; ++i) {
                    double s = Ke[j * (j + 1) / 2 + i];
#pragma unroll
                    for (int b = 0; b < DIM; ++b) s = fma(g[i][b], gs[b], s);
                    Ke[j * (j + 1) / 2 + i] = s;
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) fe[i] = fma(tN[q * NB + i], dO, fe[i]);
        } else {  // mass
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                double ns = tN[q * NB + j] * dO;
#pragma unroll
                for (int i = 0; i <= j; ++i) Ke[j * (j + 1) / 2 + i] = fma(tN[q * NB + i], ns, Ke[j * (j + 1) / 2 + i]);
            }
        }
    }
    return bad;
}

// ------------------------------------------------------------------------------------------------
// k_cell_scalar: thread per cell, scalar field.  ELEM: FB2_ELEM_HEAT or FB2_ELEM_MASS.
// ------------------------------------------------------------------------------------------------
template <int DIM, int NGEO, int NB, int NQ, int ELEM, bool ATOMIC>
__global__ void __launch_bounds__(128) k_cell_scalar(const AsmArgs A) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.ncount) return;
    const int64_t cell = A.cells ? (int64_t)A.cells[idx] : idx;
    const int64_t np = A.ncells_pad;

    int node[NGEO];
#pragma unroll
    for (int j = 0; j < NGEO; ++j) node[j] = __ldg(A.conn + (size_t)j * np + cell);
    int dof[NB];   // loaded early: the dependent colptr loads can then be issued right after the quadrature loop
#pragma unroll
    for (int i = 0; i < NB; ++i) dof[i] = __ldg(A.cell_dofs + (size_t)i * np + cell);
    double x[NGEO][DIM];
#pragma unroll
    for (int j = 0; j < NGEO; ++j) fb2_load_x<DIM>(A.xyz, node[j], x[j]);
    // Stage the scatter indices (packed uint16 offsets, column bases) into shared memory with cp.async now;
    // they land while the quadrature loop runs.  Plain loads placed here are sunk by ptxas next to the
    // REDs (one exposed memory round trip per 8 entries, profiles/r1 notes); cp.async cannot be sunk.
    constexpr int NCH = (NB * NB + 7) / 8;
    __shared__ uint4 s_map[NCH][128];
    __shared__ int64_t s_base[NB][128];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_map[k][threadIdx.x]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(A.map8 + (size_t)k * np + cell) : "memory");
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_base[j][threadIdx.x]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(A.colptr + dof[j]) : "memory");
    }
    constexpr int NSYM = NB * (NB + 1) / 2;
    double Ke[NSYM];
    double fe[NB];
    const bool bad = fb2_scalar_element<DIM, NGEO, NB, NQ, ELEM>(A, x, Ke, fe);
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (bad) {
        fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
        return;
    }
    const double kscale = A.p[0];  // heat: conductivity k; mass: rho
    const double fscale = A.p[1];  // heat: source
    // Scatter.  All index loads (column bases, packed offsets) are issued as one batch before the first RED:
    // interleaving them with the atomics serialises 64 memory round trips per cell (profiles/r1 notes).
    // They were staged into shared memory with cp.async before the quadrature loop (see above).
    asm volatile("cp.async.wait_all;" ::: "memory");
    int64_t base[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) base[j] = s_base[j][threadIdx.x];
    uint4 mp[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) mp[k] = s_map[k][threadIdx.x];
    bool missing = false;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double v = kscale * (i <= j ? Ke[j * (j + 1) / 2 + i] : Ke[i * (i + 1) / 2 + j]);
            const int e = j * NB + i;
            const unsigned w32 = (e & 7) < 2 ? mp[e >> 3].x : ((e & 7) < 4 ? mp[e >> 3].y : ((e & 7) < 6 ? mp[e >> 3].z : mp[e >> 3].w));
            const unsigned off = (e & 1) ? (w32 >> 16) : (w32 & 0xFFFFu);
            if (v != 0.0) {
                if (off == 0xFFFFu) missing = true;
                else fb2_add<ATOMIC>(A.nzval + base[j] + off, v);
            }
        }
    }
    if (A.f != nullptr && ELEM == FB2_ELEM_HEAT) {
#pragma unroll
        for (int i = 0; i < NB; ++i) fb2_add<ATOMIC>(A.f + dof[i], fscale * fe[i]);
    }
    if (missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
}

// ------------------------------------------------------------------------------------------------
// k_tile_scalar: CTA per tile of TC cells (tiles.cu).  Phase 1: every thread integrates TC/128 cells and parks
// Ke (upper triangle) and fe in shared memory, slot-major with the cell index fastest (conflict-free).  Phase 2:
// thread per distinct output entry of the tile: sum the scheduled shared-memory slots in fixed order, write the
// entry once -- plain store when the column is complete inside the tile, one RED otherwise.  Consecutive entries
// of a column are consecutive in CSC, so the writes of a warp coalesce.
// ------------------------------------------------------------------------------------------------
struct TileArgs {
    const int32_t* conn;
    const int32_t* ncells;
    const int32_t* cell_ids;
    const int64_t* col_ptr;
    const int32_t* col_dof;
    const int64_t* ent_ptr;
    const uint32_t* ent_rec;
    const uint16_t* ent_srcend;
    const int64_t* src_ptr;
    const uint16_t* src;
    int max_cols;
    int accumulate;  // 1: add onto existing values (fillzero = false)
};

template <int DIM, int NGEO, int NB, int NQ, int ELEM, int TC>
__global__ void __launch_bounds__(128, 2) k_tile_scalar(const AsmArgs A, const TileArgs T) {
    extern __shared__ double sm[];                      // [NSYM + NB][TC] element values
    constexpr int NSYM = NB * (NB + 1) / 2;
    int64_t* s_colbase = reinterpret_cast<int64_t*>(sm + (size_t)(NSYM + NB) * TC);  // [max_cols]
    int32_t* s_coldof = reinterpret_cast<int32_t*>(s_colbase + T.max_cols);           // [max_cols]
    const int64_t tile = blockIdx.x;
    const int ncell = __ldg(T.ncells + tile);
    const int64_t c0 = __ldg(T.col_ptr + tile);
    const int ncols = (int)(__ldg(T.col_ptr + tile + 1) - c0);
    for (int c = threadIdx.x; c < ncols; c += 128) {
        const int32_t d = __ldg(T.col_dof + c0 + c);
        s_coldof[c] = d;
        s_colbase[c] = __ldg(A.colptr + (d & 0x7fffffff));
    }
    const double kscale = A.p[0], fscale = A.p[1];
#pragma unroll 1
    for (int cl = threadIdx.x; cl < TC; cl += 128) {
        if (cl < ncell) {
            double x[NGEO][DIM];
#pragma unroll
            for (int j = 0; j < NGEO; ++j) {
                const int node = __ldg(T.conn + ((size_t)tile * NGEO + j) * TC + cl);
                fb2_load_x<DIM>(A.xyz, node, x[j]);
            }
            double Ke[NSYM];
            double fe[NB];
            const bool bad = fb2_scalar_element<DIM, NGEO, NB, NQ, ELEM>(A, x, Ke, fe);
            if (bad) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, __ldg(T.cell_ids + (size_t)tile * TC + cl));
#pragma unroll
            for (int s = 0; s < NSYM; ++s) sm[s * TC + cl] = kscale * Ke[s];
#pragma unroll
            for (int i = 0; i < NB; ++i) sm[(NSYM + i) * TC + cl] = fscale * fe[i];
        }
    }
    __syncthreads();
    const int64_t e0 = __ldg(T.ent_ptr + tile);
    const int ne = (int)(__ldg(T.ent_ptr + tile + 1) - e0);
    const uint16_t* src = T.src + __ldg(T.src_ptr + tile);
    const bool want_f = A.f != nullptr && ELEM == FB2_ELEM_HEAT;
    for (int e = threadIdx.x; e < ne; e += 128) {
        const uint32_t rec = __ldg(T.ent_rec + e0 + e);
        const int send = __ldg(T.ent_srcend + e0 + e);
        const int sbeg = e ? (int)__ldg(T.ent_srcend + e0 + e - 1) : 0;
        double sum = 0.0;
        for (int s = sbeg; s < send; ++s) sum += sm[__ldg(src + s)];
        const int col = rec >> 16;
        const unsigned k = rec & 0xFFFFu;
        const int32_t d = s_coldof[col];
        const bool complete = d < 0;
        double* dst;
        if (k == 0xFFFFu) {
            if (!want_f) continue;
            dst = A.f + (d & 0x7fffffff);
        } else {
            dst = A.nzval + s_colbase[col] + k;
        }
        if (complete) {
            if (T.accumulate) *dst += sum;   // nobody else touches a complete column
            else *dst = sum;
        } else if (sum != 0.0) {
            atomicAdd(dst, sum);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_cell_blocks: CTA per batch of CELLS cells.
// ------------------------------------------------------------------------------------------------
template <int NBS>
struct TileOf {
    static constexpr int TB = (NBS <= 8) ? NBS : (NBS == 10 ? 10 : 9);
};

// doubles of shared memory per cell
template <int DIM, int NBS, int ELEM>
__host__ __device__ constexpr int fb2_blocks_smem_per_cell(int nq) {
    return nq * (NBS * DIM + 1) + (ELEM == FB2_ELEM_NEOHOOKE ? nq * 90 : 0);
}

template <int DIM, int NGEO, int NBS, int VDIM, int ELEM, bool ATOMIC>
__global__ void __launch_bounds__(256) k_cell_blocks(const AsmArgs A, const int CELLS) {
    extern __shared__ double sm[];
    constexpr int TB = TileOf<NBS>::TB;
    constexpr int NT = (NBS + TB - 1) / TB;
    constexpr int N = NBS * VDIM;
    const int NQ = A.nq;
    const int64_t np = A.ncells_pad;
    double* s_g = sm;                                   // [NQ][NBS][DIM][CELLS]
    double* s_dO = s_g + (size_t)NQ * NBS * DIM * CELLS;  // [NQ][CELLS]
    double* s_A = s_dO + (size_t)NQ * CELLS;              // neohooke: [NQ][81][CELLS] dP/dF * dO
    double* s_P = s_A + (ELEM == FB2_ELEM_NEOHOOKE ? (size_t)NQ * 81 * CELLS : 0);  // [NQ][9][CELLS] P * dO
    const int64_t cell0 = (int64_t)blockIdx.x * CELLS;

    const double* tw = c_tab + A.o_w;
    const double* tN = c_tab + A.o_N;
    const double* tdN = c_tab + A.o_dN;
    const double* tdM = c_tab + A.o_dM;

    // ---- phase A: geometry (+ material state) per (qp, cell) --------------------------------------
    for (int item = threadIdx.x; item < NQ * CELLS; item += blockDim.x) {
        const int q = item / CELLS, cl = item - q * CELLS;
        const int64_t idx = cell0 + cl;
        if (idx >= A.ncount) continue;
        const int64_t cell = A.cells ? (int64_t)A.cells[idx] : idx;
        double J[DIM][DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = 0; b < DIM; ++b) J[a][b] = 0.0;
#pragma unroll
        for (int j = 0; j < NGEO; ++j) {
            int node = __ldg(A.conn + (size_t)j * np + cell);
            double xj[DIM];
            fb2_load_x<DIM>(A.xyz, node, xj);
#pragma unroll
            for (int a = 0; a < DIM; ++a)
#pragma unroll
                for (int b = 0; b < DIM; ++b) J[a][b] = fma(xj[a], tdM[(q * NGEO + j) * DIM + b], J[a][b]);
        }
        double Ji[DIM][DIM];
        const double det = fb2_det_inv<DIM>(J, Ji);
        if (!(det > 0.0)) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
        const double dO = det * tw[q];
        s_dO[q * CELLS + cl] = dO;
        double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < NBS; ++i) {
            double g[DIM];
#pragma unroll
            for (int b = 0; b < DIM; ++b) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < DIM; ++a) s = fma(tdN[(q * NBS + i) * DIM + a], Ji[a][b], s);
                g[b] = s;
                s_g[((size_t)(q * NBS + i) * DIM + b) * CELLS + cl] = s;
            }
            if (ELEM == FB2_ELEM_NEOHOOKE) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double uc = __ldg(A.u + __ldg(A.cell_dofs + (size_t)(i * 3 + c) * np + cell));
#pragma unroll
                    for (int b = 0; b < DIM; ++b) F[c][b] = fma(uc, g[b], F[c][b]);
                }
            }
        }
        if (ELEM == FB2_ELEM_NEOHOOKE) {
            const double lam = A.p[0], mu = A.p[1];
            double C[3][3], Ci[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) C[i][j] = F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j];
            const double detC = fb2_det_inv<3>(C, Ci);
            if (!(detC > 0.0)) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
            const double Jd = sqrt(detC);
            const double cS = lam * Jd * (Jd - 1.0);
            double S[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) S[i][j] = mu * ((i == j ? 1.0 : 0.0) - Ci[i][j]) + cS * Ci[i][j];
            const double c1 = (mu - cS) * 0.5, c2 = lam * (2.0 * Jd - 1.0) * (Jd * 0.5);
            // P = F S
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double pij = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
                    s_P[((size_t)q * 9 + i * 3 + j) * CELLS + cl] = pij * dO;
                }
            // dP_ij/dF_mn = delta_im S_jn + 2 F_ia dSdC_ajkn F_mk,
            // dSdC_ajkn = c1 (Ci_ak Ci_nj + Ci_an Ci_kj) + c2 Ci_aj Ci_kn
            // => F_ia dSdC_ajkn F_mk = c1 (G_ik' ... ) ; evaluate with FCi = F Ci (3x3): FCi_ik = F_ia Ci_ak
            double FCi[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) FCi[i][k] = F[i][0] * Ci[0][k] + F[i][1] * Ci[1][k] + F[i][2] * Ci[2][k];
            // W_im = F_ia Ci_ak F_mk = (FCi F^T)_im
            double W[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int m = 0; m < 3; ++m) W[i][m] = FCi[i][0] * F[m][0] + FCi[i][1] * F[m][1] + FCi[i][2] * F[m][2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int m = 0; m < 3; ++m)
#pragma unroll
                        for (int n = 0; n < 3; ++n) {
                            // F_ia Ci_ak Ci_nj F_mk = W_im Ci_nj ; F_ia Ci_an Ci_kj F_mk = FCi_in FCi_mj ; F_ia Ci_aj Ci_kn F_mk = FCi_ij FCi_mn
                            double t = c1 * (W[i][m] * Ci[n][j] + FCi[i][n] * FCi[m][j]) + c2 * FCi[i][j] * FCi[m][n];
                            double v = (i == m ? S[j][n] : 0.0) + 2.0 * t;
                            s_A[((size_t)q * 81 + ((i * 3 + j) * 3 + m) * 3 + n) * CELLS + cl] = v * dO;
                        }
        }
    }
    __syncthreads();

    // ---- phase B: integrate node blocks and scatter ------------------------------------------------
    const int nitems = NBS * NT * CELLS;
    for (int item = threadIdx.x; item < nitems; item += blockDim.x) {
        const int cl = item % CELLS;
        const int r = item / CELLS;
        const int bt = r % NT, a = r / NT;
        const int64_t idx = cell0 + cl;
        if (idx >= A.ncount) continue;
        const int64_t cell = A.cells ? (int64_t)A.cells[idx] : idx;
        const int b0 = bt * TB;
        double acc[TB][VDIM][VDIM];
#pragma unroll
        for (int t = 0; t < TB; ++t)
#pragma unroll
            for (int c = 0; c < VDIM; ++c)
#pragma unroll
                for (int d = 0; d < VDIM; ++d) acc[t][c][d] = 0.0;
        double fa[VDIM];
#pragma unroll
        for (int c = 0; c < VDIM; ++c) fa[c] = 0.0;

        for (int q = 0; q < NQ; ++q) {
            const double dO = s_dO[q * CELLS + cl];
            double ga[DIM];
#pragma unroll
            for (int b = 0; b < DIM; ++b) ga[b] = s_g[((size_t)(q * NBS + a) * DIM + b) * CELLS + cl];
            const double Na = tN[q * NBS + a];
            if (ELEM == FB2_ELEM_HEAT) {
                fa[0] = fma(Na, dO, fa[0]);
#pragma unroll
                for (int b = 0; b < DIM; ++b) ga[b] *= dO;
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
                        double s = acc[t][0][0];
#pragma unroll
                        for (int b = 0; b < DIM; ++b) s = fma(ga[b], s_g[((size_t)(q * NBS + b0 + t) * DIM + b) * CELLS + cl], s);
                        acc[t][0][0] = s;
                    }
                }
            } else if (ELEM == FB2_ELEM_MASS) {
                const double ns = Na * dO;
#pragma unroll
                for (int t = 0; t < TB; ++t)
                    if (b0 + t < NBS) acc[t][0][0] = fma(ns, tN[q * NBS + b0 + t], acc[t][0][0]);
            } else if (ELEM == FB2_ELEM_ELASTICITY) {
                const double lam = A.p[0] * dO, mu = A.p[1] * dO;
#pragma unroll
                for (int c = 0; c < VDIM; ++c) fa[c] = fma(Na * dO, A.p[2 + c], fa[c]);
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
                        double gb[DIM];
#pragma unroll
                        for (int b = 0; b < DIM; ++b) gb[b] = s_g[((size_t)(q * NBS + b0 + t) * DIM + b) * CELLS + cl];
                        double dot = 0.0;
#pragma unroll
                        for (int b = 0; b < DIM; ++b) dot = fma(ga[b], gb[b], dot);
                        const double mdot = mu * dot;
#pragma unroll
                        for (int c = 0; c < VDIM; ++c)
#pragma unroll
                            for (int d = 0; d < VDIM; ++d) {
                                double v = fma(lam * ga[c], gb[d], mu * ga[d] * gb[c]);
                                if (c == d) v += mdot;
                                acc[t][c][d] += v;
                            }
                    }
                }
            } else {  // neo-hooke: block(a,b)[c][d] = sum_{j,n} ga[j] A[c][j][d][n] gb[n]
                double h[3][3][3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int n = 0; n < 3; ++n) {
                            double s = 0.0;
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                s = fma(ga[j], s_A[((size_t)q * 81 + ((c * 3 + j) * 3 + d) * 3 + n) * CELLS + cl], s);
                            h[c][d][n] = s;
                        }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double s = fa[c];
#pragma unroll
                    for (int j = 0; j < 3; ++j) s = fma(ga[j], s_P[((size_t)q * 9 + c * 3 + j) * CELLS + cl], s);
                    fa[c] = fma(-Na * dO, A.p[2 + c], s);
                }
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
                        double gb[3];
#pragma unroll
                        for (int b = 0; b < 3; ++b) gb[b] = s_g[((size_t)(q * NBS + b0 + t) * 3 + b) * CELLS + cl];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                double s = acc[t][c][d];
#pragma unroll
                                for (int n = 0; n < 3; ++n) s = fma(h[c][d][n], gb[n], s);
                                acc[t][c][d] = s;
                            }
                    }
                }
            }
        }
        // scatter the row block (a, :) x tile columns
        const double kscale = (ELEM == FB2_ELEM_HEAT || ELEM == FB2_ELEM_MASS) ? A.p[0] : 1.0;
        bool missing = false;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            if (b0 + t < NBS) {
                // batch the index loads of this node block before its REDs (loads interleaved with atomics
                // are serialised by the compiler: one memory round trip per entry)
                int64_t base[VDIM];
                unsigned off[VDIM][VDIM];
                int dj[VDIM];
#pragma unroll
                for (int d = 0; d < VDIM; ++d) {
                    const int jl = (b0 + t) * VDIM + d;
                    dj[d] = fb2_ldv_s32(A.cell_dofs + (size_t)jl * np + cell);
#pragma unroll
                    for (int c = 0; c < VDIM; ++c) off[d][c] = fb2_ldv_u16(A.map + (size_t)(jl * N + a * VDIM + c) * np + cell);
                }
#pragma unroll
                for (int d = 0; d < VDIM; ++d) base[d] = fb2_ldv_s64(A.colptr + dj[d]);
#pragma unroll
                for (int d = 0; d < VDIM; ++d) {
#pragma unroll
                    for (int c = 0; c < VDIM; ++c) {
                        const double v = kscale * acc[t][c][d];
                        if (v != 0.0) {
                            if (off[d][c] == 0xFFFFu) missing = true;
                            else fb2_add<ATOMIC>(A.nzval + base[d] + off[d][c], v);
                        }
                    }
                }
            }
        }
        if (missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
        if (bt == 0 && A.f != nullptr && ELEM != FB2_ELEM_MASS) {
            const double fscale = ELEM == FB2_ELEM_HEAT ? A.p[1] : 1.0;
#pragma unroll
            for (int c = 0; c < VDIM; ++c) {
                const int di = __ldg(A.cell_dofs + (size_t)(a * VDIM + c) * np + cell);
                fb2_add<ATOMIC>(A.f + di, fscale * fa[c]);
            }
        }
    }
}
