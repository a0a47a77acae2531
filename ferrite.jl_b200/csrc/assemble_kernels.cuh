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
// Kernels (DESIGN.md section 4 has the measurements):
//   k_cell_scalar     -- one thread per cell, Ke (upper triangle) and fe in FP64 registers, reference tables as
//                        constant-bank operands, scatter indices staged with cp.async, x-face merge through warp
//                        shuffles and lane-parity sector pairing of the REDs.  Scalar fields with few dofs per cell
//                        (Q1 quad/hex, P1/P2 simplices).  The default for heat / mass.
//   k_cell_syrk       -- isotropic elasticity on the FP64 tensor cores: G = X^T diag(dOmega) X per cell with DMMA
//                        (warp per cell for small elements, an 8-warp CTA per Q2 hexahedron), Ke from G in the scatter.
//   k_cell_blocks     -- one CTA per batch of cells; phase A stages physical shape gradients and dOmega (Neo-Hooke: P and
//                        dP/dF) in shared memory, phase B (thread per (cell, row node, tile of column nodes)) integrates
//                        vdim x vdim node blocks in registers and scatters them.  Neo-Hooke, Q2 scalar fields, variant 1.
//   k_cell_scalar_ws  -- warp-specialised k_cell_scalar (variant 8), k_tile_scalar -- tile aggregation (variant 5):
//                        measured alternatives, not defaults.
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
    const uint16_t* mapc;  // cell-major offsets for k_cell_blocks: [cell][mapstride], entry jl * n + il
    const int32_t* dofc;   // cell-major dofs for the CTA kernels: [cell][n]
    const int64_t* basec;  //   and colptr[dof]: [cell][n]
    const int32_t* cells;  // optional indirection (colour subset)
    int64_t cell_first;    // without a list the launch covers cells [cell_first, cell_first + ncount)
    const int32_t* wfirst; // optional warp list of k_cell_scalar: first cell of every warp (4 consecutive warps = 4 grid rows)
    const uint8_t* wcount; //   and its number of cells (<= 32)
    int64_t ncount;        // number of cells handled by this launch
    int64_t ncells_pad;
    double* nzval;
    double* f;
    const double* u;
    int* errflag;
    double p[6];  // element parameters
    int nq;
    const double* cmat;    // FB2_ELEM_ELASTICITY_GENERAL: the 81 entries of C on the device
    int zero_pending;      // host only: start_assemble's zero fill of nzval is still owed (a kernel that fuses it clears this)
    const double* tab;     // the c_tab contents in global memory (per-lane indexed reads)
    bool const_ok;         // the tables fit (and are) in c_tab
    // table offsets into c_tab / tab (in doubles)
    int o_w, o_N, o_dN, o_M, o_dM;
};

__device__ __forceinline__ int64_t fb2_cell_of(const AsmArgs& A, int64_t i) {
    return A.cells ? (int64_t)A.cells[i] : A.cell_first + i;
}

__device__ __forceinline__ void fb2_flag_error(int* errflag, int code, int64_t cell) {
    if (atomicCAS(&errflag[0], 0, code) == 0) errflag[1] = (int)cell;
}

__device__ __forceinline__ void fb2_cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

template <bool ATOMIC>
__device__ __forceinline__ void fb2_add(double* p, double v) {
    if (ATOMIC) atomicAdd(p, v);  // result unused -> RED.E.ADD.F64
    else *p += v;
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
// xs != nullptr: the coordinates are read from shared memory (xs[(j * DIM + a) * 128], the thread's own column) at
// every quadrature point instead of living in 2 * NGEO * DIM registers.
template <int DIM, int NGEO, int NB, int NQ, int ELEM, bool ROLLQ = false, bool XSMEM = false>
__device__ __forceinline__ bool fb2_scalar_element(const AsmArgs& A, const double (&x)[NGEO][DIM],
                                                   double (&Ke)[NB * (NB + 1) / 2], double (&fe)[NB], const double* xs = nullptr) {
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
    // ROLLQ keeps the quadrature loop rolled: 1/NQ of the code size (the unrolled Q1-hex body is ~70 KB of SASS, more
    // than the 32 KB instruction cache level shared by the SM's warps)
#pragma unroll(ROLLQ ? 1 : NQ)
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
                for (int b = 0; b < DIM; ++b) J[a][b] = fma(XSMEM ? xs[(j * DIM + a) * 128] : x[j][a], tdM[(q * NGEO + j) * DIM + b], J[a][b]);
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
// DEBUGMODE (measurement only, wrong results): 1 = integrate but never scatter, 2 = scatter without integrating
template <int DIM, int NGEO, int NB, int NQ, int ELEM, bool ATOMIC, int MB = 1, bool ROLLQ = false, bool CHECK = true, bool PAIRX = true, bool XSMEM = false, int DEBUGMODE = 0>
__global__ void __launch_bounds__(128, MB) k_cell_scalar(const AsmArgs A) {
    // lanes past the end stay alive (they redo a valid cell and write nothing): the face merges below shuffle across
    // the whole warp and synchronise the CTA
    bool active;
    int64_t cell;
    if (A.wfirst) {   // warp list: warp w covers cells [wfirst[w], wfirst[w] + wcount[w])
        const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31, cnt = A.wcount[w];
        active = lane < cnt;
        cell = (int64_t)A.wfirst[w] + (active ? lane : 0);
    } else {
        const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        active = idx < A.ncount;
        const int64_t idc = active ? idx : A.ncount - 1;
        cell = fb2_cell_of(A, idc);
    }
    const int64_t np = A.ncells_pad;

    int node[NGEO];
#pragma unroll
    for (int j = 0; j < NGEO; ++j) node[j] = __ldg(A.conn + (size_t)j * np + cell);
    int dof[NB];   // loaded early: the dependent colptr loads can then be issued right after the quadrature loop
#pragma unroll
    for (int i = 0; i < NB; ++i) dof[i] = __ldg(A.cell_dofs + (size_t)i * np + cell);
    double x[NGEO][DIM];
    __shared__ double s_x[XSMEM ? NGEO * DIM : 1][128];
#pragma unroll
    for (int j = 0; j < NGEO; ++j) {
        fb2_load_x<DIM>(A.xyz, node[j], x[j]);
        if (XSMEM) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) s_x[j * DIM + a][threadIdx.x] = x[j][a];
        }
    }
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
    bool bad = false;
    if (DEBUGMODE == 2) {
#pragma unroll
        for (int e = 0; e < NSYM; ++e) Ke[e] = x[e % NGEO][0] + (double)e;
#pragma unroll
        for (int i = 0; i < NB; ++i) fe[i] = x[i % NGEO][DIM - 1];
    } else {
        bad = fb2_scalar_element<DIM, NGEO, NB, NQ, ELEM, ROLLQ, XSMEM>(A, x, Ke, fe, &s_x[0][threadIdx.x]);
    }
    if (DEBUGMODE == 1 && !(Ke[0] == 123.456)) active = false;   // never true: the integration stays, the scatter goes
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (bad && active) {
        fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
        active = false;
    }
    const double kscale = A.p[0];  // heat: conductivity k; mass: rho
    const double fscale = A.p[1];  // heat: source
    // Face merges (Q1 quadrilateral / hexahedron, atomic mode).  A cell hands the block of Ke / fe that belongs to a
    // face shared with a neighbouring cell to that neighbour and zeroes it (zeros are skipped by the scatter), so the
    // neighbour adds both contributions with ONE RED per entry.  Faces are matched by dof id, so any mesh ordering
    // qualifies; on grids stored in rows the previous lane is the x-neighbour and, with the warp list (4 consecutive
    // grid rows per CTA), the previous warp holds the y-neighbours.  The kernel is bound by L2 atomic sector
    // throughput (profiles/r01_prof_c2_r1c.txt): x-merge alone removes 25 % of the REDs (2.86 -> 2.41 ms on C2).
    constexpr bool MERGE = ATOMIC && CHECK && ((DIM == 3 && NB == 8 && NGEO == 8) || (DIM == 2 && NB == 4 && NGEO == 4));
    constexpr int NF = DIM == 3 ? 4 : 2;                       // nodes per face
    constexpr int RF[4] = {1, 2, 5, 6}, LF[4] = {0, 3, 4, 7};   // right face of the left cell <-> left face of this cell
    constexpr int TF[4] = {3, 2, 7, 6}, BF[4] = {0, 1, 4, 5};   // top face of the lower cell  <-> bottom face of this cell
    constexpr int NPAIR = NF * (NF + 1) / 2;
    if (MERGE && !XSMEM && A.wfirst != nullptr) {   // ---- y-merge through shared memory (uniform branch) ----
        constexpr int YW = XSMEM ? 1 : 128;             // (no static shared memory for it in the XSMEM instantiations)
        __shared__ double s_T[NPAIR + NF][YW];
        __shared__ int s_Tdof[NF][YW];
        __shared__ unsigned char s_act[YW], s_taken[YW];
        const int t = threadIdx.x;
        {
            int pi = 0;
#pragma unroll
            for (int q = 0; q < NF; ++q)
#pragma unroll
                for (int p = 0; p <= q; ++p) {
                    const int i0 = TF[p] < TF[q] ? TF[p] : TF[q], j0 = TF[p] < TF[q] ? TF[q] : TF[p];
                    s_T[pi++][t] = Ke[j0 * (j0 + 1) / 2 + i0];
                }
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                s_T[NPAIR + k][t] = fe[TF[k]];
                s_Tdof[k][t] = dof[TF[k]];
            }
            s_act[t] = active;
            s_taken[t] = 0;
        }
        __syncthreads();
        if (t >= 32) {
            const int lo = t - 32;
            bool ym = active && s_act[lo];
#pragma unroll
            for (int k = 0; k < NF; ++k) ym = ym & (s_Tdof[k][lo] == dof[BF[k]]);
            if (ym) {
                int pi = 0;
#pragma unroll
                for (int q = 0; q < NF; ++q)
#pragma unroll
                    for (int p = 0; p <= q; ++p) {
                        const int i0 = BF[p] < BF[q] ? BF[p] : BF[q], j0 = BF[p] < BF[q] ? BF[q] : BF[p];
                        Ke[j0 * (j0 + 1) / 2 + i0] += s_T[pi++][lo];
                    }
#pragma unroll
                for (int k = 0; k < NF; ++k) fe[BF[k]] += s_T[NPAIR + k][lo];
                s_taken[lo] = 1;
            }
        }
        __syncthreads();
        if (s_taken[t]) {   // the cell above took my top-face block
#pragma unroll
            for (int q = 0; q < NF; ++q)
#pragma unroll
                for (int p = 0; p <= q; ++p) {
                    const int i0 = TF[p] < TF[q] ? TF[p] : TF[q], j0 = TF[p] < TF[q] ? TF[q] : TF[p];
                    Ke[j0 * (j0 + 1) / 2 + i0] = 0.0;
                }
#pragma unroll
            for (int k = 0; k < NF; ++k) fe[TF[k]] = 0.0;
        }
    }
    bool next_takes_x = false;
    if (MERGE) {   // ---- x-merge through warp shuffles ----
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        bool match = lane > 0;
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            const int pd = __shfl_up_sync(full, dof[RF[k]], 1);   // unconditional: every lane must take part
            match = match & (pd == dof[LF[k]]);
        }
        const bool prev_active = __shfl_up_sync(full, (int)active, 1) != 0;
        match = match & prev_active & active;
        const bool nm = __shfl_down_sync(full, (int)match, 1) != 0;
        const bool next_takes = nm & (lane < 31);   // the next lane consumes my right-face block
        next_takes_x = next_takes;
#pragma unroll
        for (int q = 0; q < NF; ++q)
#pragma unroll
            for (int p = 0; p <= q; ++p) {
                const double mine = Ke[RF[q] * (RF[q] + 1) / 2 + RF[p]];
                const double recv = __shfl_up_sync(full, mine, 1);
                if (next_takes) Ke[RF[q] * (RF[q] + 1) / 2 + RF[p]] = 0.0;
                if (match) Ke[LF[q] * (LF[q] + 1) / 2 + LF[p]] += recv;
            }
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            const double mine = fe[RF[k]];
            const double recv = __shfl_up_sync(full, mine, 1);
            if (next_takes) fe[RF[k]] = 0.0;
            if (match) fe[LF[k]] += recv;
        }
    }
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
    // Sector pairing (Q1 quadrilateral / hexahedron): local nodes j and j^1 are x-mates, so column j of an even lane
    // and column j^1 of the next (odd) lane are the same matrix column when j lies on the shared face, and their rows
    // i are x-neighbours = adjacent entries of that CSC column.  Odd lanes therefore walk the columns in the order
    // j^1: the two REDs of a lane pair fall into one 32-byte sector three times out of four.  The values stay in
    // registers (compile-time indices), only selected by lane parity.
    constexpr bool PAIR = MERGE && PAIRX;
    const bool odd = PAIR && (threadIdx.x & 1);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int jo = PAIR ? (j ^ 1) : j;
            const double ve = i <= j ? Ke[j * (j + 1) / 2 + i] : Ke[i * (i + 1) / 2 + j];
            const double vo = i <= jo ? Ke[jo * (jo + 1) / 2 + i] : Ke[i * (i + 1) / 2 + jo];
            const double v = kscale * (odd ? vo : ve);
            const int e = j * NB + i, eo = jo * NB + i;   // e and eo differ by NB = 8: same slot of the packed word
            const unsigned we = (e & 7) < 2 ? mp[e >> 3].x : ((e & 7) < 4 ? mp[e >> 3].y : ((e & 7) < 6 ? mp[e >> 3].z : mp[e >> 3].w));
            const unsigned wo = (eo & 7) < 2 ? mp[eo >> 3].x : ((eo & 7) < 4 ? mp[eo >> 3].y : ((eo & 7) < 6 ? mp[eo >> 3].z : mp[eo >> 3].w));
            const unsigned w32 = odd ? wo : we;
            const unsigned off = (e & 1) ? (w32 >> 16) : (w32 & 0xFFFFu);
            const int64_t bj = odd ? base[jo] : base[j];
            if (!active) continue;
            if (CHECK) {   // zero values are skipped; a non-zero aimed at a missing entry is an error
                if (v != 0.0) {
                    if (off == 0xFFFFu) missing = true;
                    else fb2_add<ATOMIC>(A.nzval + bj + off, v);
                }
            } else {       // complete map: straight-line code; only the blocks handed to the x-neighbour are skipped
                const bool iR = (i & 1) != ((i >> 1) & 1);
                const bool jR = odd ? ((jo & 1) != ((jo >> 1) & 1)) : ((j & 1) != ((j >> 1) & 1));
                if (!(MERGE && iR && jR && next_takes_x)) fb2_add<ATOMIC>(A.nzval + bj + off, v);
            }
        }
    }
    if (A.f != nullptr && ELEM == FB2_ELEM_HEAT && active) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (!MERGE || fe[i] != 0.0) fb2_add<ATOMIC>(A.f + dof[i], fscale * fe[i]);
    }
    if (missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
}

// ------------------------------------------------------------------------------------------------
// k_cell_scalar_ws: warp-specialised version of k_cell_scalar for Q1 quadrilaterals / hexahedra (atomic mode).
//
// k_cell_scalar needs 216 registers per thread for the element integration, i.e. two warps per scheduler, and every
// warp alternates between an FP64-bound phase (integration) and an L2-bound phase (scatter): the FP64 pipe idles while
// both warps of a scheduler scatter (52 % pipe utilisation, profiles/r01_prof_c2_r1h.txt).  Here a CTA has two warp
// groups with their own register budgets (setmaxnreg): the COMPUTE group (216 registers) gathers and integrates batches
// of 128 cells and parks Ke (upper triangle) + fe in a double-buffered shared-memory stage; the SCATTER group
// (40 registers) stages the scatter indices of the same cells with cp.async while the batch is integrated, then does
// the face merge (the x-neighbour's block is read from the stage, no shuffles), the sector-paired REDs and the load
// vector.  Named barriers FULL[s] / EMPTY[s] hand the stages back and forth.  CTAs are persistent over batches.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fb2_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fb2_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NB>
struct WsSmem {
    static constexpr int NSYM = NB * (NB + 1) / 2, NV = NSYM + NB, NCH = (NB * NB + 7) / 8;
    static constexpr size_t val = 0;                                              // double [2][NV][128]
    static constexpr size_t map = val + sizeof(double) * 2 * NV * 128;            // uint4  [NCH][128]
    static constexpr size_t base = map + sizeof(uint4) * NCH * 128;               // int64  [NB][128]
    static constexpr size_t total = base + sizeof(int64_t) * NB * 128;
};

template <int DIM, int NGEO, int NB, int NQ, int ELEM, bool ROLLQ, bool CHECK>
__global__ void __launch_bounds__(256, 2) k_cell_scalar_ws(const AsmArgs A, const int64_t nbatch) {
    using SM = WsSmem<NB>;
    constexpr int NSYM = SM::NSYM, NV = SM::NV, NCH = SM::NCH;
    constexpr bool MERGE = (DIM == 3 && NB == 8 && NGEO == 8) || (DIM == 2 && NB == 4 && NGEO == 4);
    constexpr int NF = DIM == 3 ? 4 : 2;
    constexpr int RF[4] = {1, 2, 5, 6}, LF[4] = {0, 3, 4, 7};   // x = 1 face / x = 0 face; x-mates are j ^ 1
    extern __shared__ __align__(16) unsigned char smraw[];
    double* s_val = reinterpret_cast<double*>(smraw + SM::val);
    uint4* s_map = reinterpret_cast<uint4*>(smraw + SM::map);
    int64_t* s_base = reinterpret_cast<int64_t*>(smraw + SM::base);
    const int wg = threadIdx.x >> 7, t = threadIdx.x & 127;
    const int64_t np = A.ncells_pad;
    enum { BAR_FULL = 1, BAR_EMPTY = 3 };

    if (wg == 0) {
        // ================= compute group =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const double kscale = A.p[0], fscale = A.p[1];
        int k = 0;
        for (int64_t b = blockIdx.x; b < nbatch; b += gridDim.x, ++k) {
            const int s = k & 1;
            const int64_t idx = b * 128 + t;
            const bool active = idx < A.ncount;
            const int64_t cell = fb2_cell_of(A, active ? idx : A.ncount - 1);
            double x[NGEO][DIM];
#pragma unroll
            for (int j = 0; j < NGEO; ++j) {
                const int node = __ldg(A.conn + (size_t)j * np + cell);
                fb2_load_x<DIM>(A.xyz, node, x[j]);
            }
            double Ke[NSYM];
            double fe[NB];
            const bool bad = fb2_scalar_element<DIM, NGEO, NB, NQ, ELEM, ROLLQ>(A, x, Ke, fe);
            if (bad && active) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
            const bool drop = bad || !active;   // zeros are skipped by the scatter group
            if (k >= 2) fb2_bar_sync(BAR_EMPTY + s, 256);     // the scatter group has released this stage
            double* st = s_val + (size_t)s * NV * 128 + t;
#pragma unroll
            for (int e = 0; e < NSYM; ++e) st[e * 128] = drop ? 0.0 : kscale * Ke[e];
#pragma unroll
            for (int i = 0; i < NB; ++i) st[(NSYM + i) * 128] = (ELEM == FB2_ELEM_HEAT && !drop) ? fscale * fe[i] : 0.0;
            fb2_bar_arrive(BAR_FULL + s, 256);
        }
    } else {
        // ================= scatter group =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int lane = t & 31;
        const bool odd = MERGE && (t & 1);
        int k = 0;
        for (int64_t b = blockIdx.x; b < nbatch; b += gridDim.x, ++k) {
            const int s = k & 1;
            const int64_t idx = b * 128 + t;
            const bool active = idx < A.ncount;
            const int64_t cell = fb2_cell_of(A, active ? idx : A.ncount - 1);
            int dof[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) dof[i] = __ldg(A.cell_dofs + (size_t)i * np + cell);
#pragma unroll
            for (int c = 0; c < NCH; ++c) fb2_cp_async16(&s_map[c * 128 + t], A.map8 + (size_t)c * np + cell);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_base[j * 128 + t]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(A.colptr + dof[j]) : "memory");
            }
            // face match with the previous / next lane (dof ids; any mesh ordering qualifies)
            bool match = false, next_takes = false;
            if (MERGE) {
                const unsigned full = 0xffffffffu;
                match = lane > 0;
#pragma unroll
                for (int q = 0; q < NF; ++q) {
                    const int pd = __shfl_up_sync(full, dof[RF[q]], 1);
                    match = match & (pd == dof[LF[q]]);
                }
                const bool prev_active = __shfl_up_sync(full, (int)active, 1) != 0;
                match = match & prev_active & active;
                next_takes = (__shfl_down_sync(full, (int)match, 1) != 0) & (lane < 31);
            }
            fb2_bar_sync(BAR_FULL + s, 256);                  // the batch's Ke / fe are in stage s
            asm volatile("cp.async.wait_all;" ::: "memory");
            const double* st = s_val + (size_t)s * NV * 128 + t;
            if (active) {
#pragma unroll
                for (int j = 0; j < NB; ++j) {
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        // sector pairing: odd lanes walk the columns in the order j ^ 1 (see k_cell_scalar)
                        const int jo = MERGE ? (j ^ 1) : j;
                        const int se = i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j;
                        const int so = i <= jo ? jo * (jo + 1) / 2 + i : i * (i + 1) / 2 + jo;
                        const int im = i ^ 1, jme = j ^ 1, jmo = jo ^ 1;   // x-mates: the neighbour's local ids of the same nodes
                        const int sme = im <= jme ? jme * (jme + 1) / 2 + im : im * (im + 1) / 2 + jme;
                        const int smo = im <= jmo ? jmo * (jmo + 1) / 2 + im : im * (im + 1) / 2 + jmo;
                        const bool iL = (i & 1) == ((i >> 1) & 1);          // local nodes 0,3,4,7 lie on x = 0
                        const bool jLe = (j & 1) == ((j >> 1) & 1), jLo = (jo & 1) == ((jo >> 1) & 1);
                        const bool jL = odd ? jLo : jLe;
                        double v = st[(odd ? so : se) * 128];
                        if (MERGE) {
                            if (iL && jL && match) v += st[(odd ? smo : sme) * 128 - 1];   // left neighbour's x = 1 face block
                            if (!iL && !jL && next_takes) v = 0.0;                            // the right neighbour adds mine
                        }
                        const int e = (odd ? jo : j) * NB + i;
                        const unsigned off = reinterpret_cast<const uint16_t*>(s_map)[((e >> 3) * 128 + t) * 8 + (e & 7)];
                        double* dst = A.nzval + s_base[(odd ? jo : j) * 128 + t] + off;
                        if (CHECK) {   // zero values are skipped; a non-zero aimed at a missing entry is an error
                            if (v != 0.0) {
                                if (off == 0xFFFFu) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
                                else atomicAdd(dst, v);
                            }
                        } else {       // complete map: straight-line code, only the blocks handed to the neighbour are skipped
                            if (!(MERGE && !iL && !jL && next_takes)) atomicAdd(dst, v);
                        }
                    }
                }
                if (A.f != nullptr && ELEM == FB2_ELEM_HEAT) {
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        const bool iL = (i & 1) == ((i >> 1) & 1);
                        double v = st[(NSYM + i) * 128];
                        if (MERGE) {
                            if (iL && match) v += st[(NSYM + (i ^ 1)) * 128 - 1];
                            if (!iL && next_takes) v = 0.0;
                        }
                        if (v != 0.0) atomicAdd(A.f + dof[i], v);
                    }
                }
            }
            if (b + 2 * (int64_t)gridDim.x < nbatch) fb2_bar_arrive(BAR_EMPTY + s, 256);   // the compute group waits for it at k + 2
        }
    }
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
    const int64_t* tile_base;
    const int64_t* ent_ptr;
    const uint2* rec;
    const int64_t* src_ptr;
    const uint16_t* src;
    int max_ent, max_src;
    int accumulate;  // 1: add onto existing values (fillzero = false)
};

template <int DIM, int NGEO, int NB, int NQ, int ELEM, int TC, int MB = 2>
__global__ void __launch_bounds__(TC, MB) k_tile_scalar(const AsmArgs A, const TileArgs T) {
    extern __shared__ double sm[];                      // [NSYM + NB][TC] element values, then the staged schedule
    constexpr int NSYM = NB * (NB + 1) / 2;
    uint2* s_rec = reinterpret_cast<uint2*>(sm + (size_t)(NSYM + NB) * TC);   // [max_ent]
    uint16_t* s_src = reinterpret_cast<uint16_t*>(s_rec + T.max_ent);          // [max_src]
    const int64_t tile = blockIdx.x;
    const int ncell = __ldg(T.ncells + tile);
    const int64_t e0 = __ldg(T.ent_ptr + tile);
    const int ne = (int)(__ldg(T.ent_ptr + tile + 1) - e0);
    const int64_t s0 = __ldg(T.src_ptr + tile);
    const int ns = (int)(__ldg(T.src_ptr + tile + 1) - s0);
    // stage the tile's schedule into shared memory; it lands while the element matrices are integrated
    // (segments are padded to multiples of 8 entries, so every chunk is 16-byte aligned)
    for (int i = threadIdx.x; i < ne / 2; i += TC) fb2_cp_async16(s_rec + 2 * i, T.rec + e0 + 2 * i);
    for (int i = threadIdx.x; i < ns / 8; i += TC) fb2_cp_async16(s_src + 8 * i, T.src + s0 + 8 * i);
    const double kscale = A.p[0], fscale = A.p[1];
    {
        const int cl = threadIdx.x;
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
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const bool want_f = A.f != nullptr && ELEM == FB2_ELEM_HEAT;
    double* const nzbase = A.nzval + __ldg(T.tile_base + tile);
    // two entries per thread and iteration: two independent shared-memory gather chains in flight
    for (int eb = threadIdx.x; eb < ne; eb += 2 * TC) {
        const uint2 r0 = s_rec[eb];
        const bool has1 = eb + TC < ne;
        const uint2 r1 = has1 ? s_rec[eb + TC] : make_uint2(0xFFFFFFFFu, 0u);
        const int b0 = r0.y & 0xFFFFu, n0 = r0.y >> 16;
        const int b1 = r1.y & 0xFFFFu, n1 = r1.y >> 16;
        double sum0 = 0.0, sum1 = 0.0;
        const int nmin = min(n0, n1);
        int s = 0;
        for (; s < nmin; ++s) {
            const double v0 = sm[s_src[b0 + s]];
            const double v1 = sm[s_src[b1 + s]];
            sum0 += v0;
            sum1 += v1;
        }
        for (int q = s; q < n0; ++q) sum0 += sm[s_src[b0 + q]];
        for (int q = s; q < n1; ++q) sum1 += sm[s_src[b1 + q]];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const unsigned x = u ? r1.x : r0.x;
            const double sum = u ? sum1 : sum0;
            if (x == 0xFFFFFFFFu) continue;
            double* dst;
            if (x & 0x40000000u) {
                if (!want_f) continue;
                dst = A.f + (x & 0x3fffffffu);
            } else {
                dst = nzbase + (x & 0x3fffffffu);
            }
            if (x & 0x80000000u) {
                if (T.accumulate) *dst += sum;   // nobody else touches a complete column
                else *dst = sum;
            } else if (sum != 0.0) {
                atomicAdd(dst, sum);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_cell_blocks: CTA per batch of CELLS cells (vector fields, higher order, Neo-Hooke).
//   phase A : thread per (cell, qp): J, det > 0, J^-1, dOmega and the physical gradients of all scalar basis functions
//             go to shared memory ([q][cell][a][d]); Neo-Hooke also stages P*dOmega and dP/dF*dOmega.
//   staging : the cell's scatter indices are brought on chip once: its block of the cell-major uint16 offset map with
//             cp.async (contiguous bytes), column bases colptr[dof_j] and the dofs themselves.
//   phase B : thread per (cell, local row r = (a, c), tile of TB column nodes): integrates its TB*VDIM entries of row
//             r in registers (low register count => several CTAs per SM hide the latencies) and scatters them.  The
//             lanes of a warp hold CONSECUTIVE rows of one cell; rows (a, 0..VDIM-1) are adjacent in a CSC column, so
//             the REDs of one instruction share 32-byte sectors (about 2.4x fewer L2 atomic sectors than one RED per
//             (a, b) block and lane).
// ------------------------------------------------------------------------------------------------
template <int NBS, int VDIM>
struct TileOf {
    // column nodes per phase-B thread: keeps the accumulators (TB * VDIM^2 doubles) around 30-40 registers pairs
    static constexpr int TB = VDIM == 1 ? ((NBS <= 8) ? NBS : (NBS == 10 ? 10 : 9))
                            : VDIM == 2 ? NBS
                                        : (NBS == 27 ? 3 : (NBS == 10 ? 5 : 4));
    // Optional transposed scatter (off): the element matrices of the CTA's cells go through shared memory so that
    // the lanes of a warp add CONSECUTIVE rows of one column (rows (a, 0..VDIM-1) are adjacent in a CSC column and
    // share 32-byte sectors).  Measured on Q1^3 hex elasticity (profiles/r01_prof_c5_r1f.txt): RED sectors 1100 M ->
    // 522 M, lts throughput 49 % -> 27 %, but +40 % instructions and an extra barrier: 7.2 ms -> 8.1 ms, so it stays off.
    static constexpr bool TRANSPOSE = false;
};

struct BlockSmem {
    size_t g, dO, Ji, A, P, base, dof, map, K, total;   // byte offsets
    int mapstride;                                   // uint16 entries per cell in the staged map (multiple of 8)
};

template <int DIM, int NBS, int VDIM, int ELEM>
__host__ __device__ inline BlockSmem fb2_blocks_smem(int nq, int cells) {
    constexpr int N = NBS * VDIM;
    BlockSmem L;
    size_t o = 0;
    L.g = o; o += sizeof(double) * (size_t)nq * cells * NBS * DIM;
    L.dO = o; o += sizeof(double) * (size_t)nq * cells;
    L.Ji = o; o += sizeof(double) * (size_t)nq * cells * DIM * DIM;
    L.A = o; o += (ELEM == FB2_ELEM_NEOHOOKE ? sizeof(double) * (size_t)nq * cells * 81 : 0);
    L.P = o; o += (ELEM == FB2_ELEM_NEOHOOKE ? sizeof(double) * (size_t)nq * cells * 9 : 0);
    L.base = o; o += sizeof(int64_t) * (size_t)cells * N;
    L.dof = o; o += sizeof(int32_t) * (size_t)cells * N;
    o = (o + 15) / 16 * 16;
    L.mapstride = (N * N + 7) / 8 * 8;
    L.map = o; o += sizeof(uint16_t) * (size_t)cells * L.mapstride;
    o = (o + 15) / 16 * 16;
    L.K = o; o += (TileOf<NBS, VDIM>::TRANSPOSE ? sizeof(double) * (size_t)cells * N * N : 0);
    L.total = o;
    return L;
}

// Neo-Hooke keeps 27 + TB * 9 doubles live in phase B: under a 128-register cap it spills (the C4 profile showed 3.1 G L2
// sectors of local-memory traffic, long-scoreboard stalls on every DFMA), so it runs with at most 192 threads per CTA = a cap of 168 registers.
template <int DIM, int NGEO, int NBS, int VDIM, int ELEM, bool ATOMIC>
__global__ void __launch_bounds__(ELEM == FB2_ELEM_NEOHOOKE ? 192 : 256, 2) k_cell_blocks(const AsmArgs A, const int CELLS) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int TB = TileOf<NBS, VDIM>::TB;
    constexpr int NT = (NBS + TB - 1) / TB;
    constexpr int N = NBS * VDIM;
    const int NQ = A.nq;
    const int64_t np = A.ncells_pad;
    const BlockSmem L = fb2_blocks_smem<DIM, NBS, VDIM, ELEM>(NQ, CELLS);
    double* s_g = reinterpret_cast<double*>(smraw + L.g);        // [NQ][CELLS][NBS][DIM]
    double* s_dO = reinterpret_cast<double*>(smraw + L.dO);      // [NQ][CELLS]
    double* s_Ji = reinterpret_cast<double*>(smraw + L.Ji);      // [NQ][CELLS][DIM][DIM]
    double* s_A = reinterpret_cast<double*>(smraw + L.A);        // neo-hooke [NQ][CELLS][81]
    double* s_P = reinterpret_cast<double*>(smraw + L.P);        // neo-hooke [NQ][CELLS][9]
    int64_t* s_base = reinterpret_cast<int64_t*>(smraw + L.base);  // [CELLS][N]
    int32_t* s_dof = reinterpret_cast<int32_t*>(smraw + L.dof);    // [CELLS][N]
    uint16_t* s_map = reinterpret_cast<uint16_t*>(smraw + L.map);  // [CELLS][mapstride]
    double* s_K = reinterpret_cast<double*>(smraw + L.K);          // TRANSPOSE: [CELLS][N * N], entry jl * N + il
    constexpr bool TRANSPOSE = TileOf<NBS, VDIM>::TRANSPOSE;
    const int64_t cell0 = (int64_t)blockIdx.x * CELLS;
    const int ncl = (int)min((int64_t)CELLS, A.ncount - cell0);

    // tables from global memory (L1-resident): the indices below differ per lane, which the constant bank would serialise
    const double* __restrict__ tw = A.tab + A.o_w;
    const double* __restrict__ tN = A.tab + A.o_N;
    const double* __restrict__ tdN = A.tab + A.o_dN;
    const double* __restrict__ tdM = A.tab + A.o_dM;

    // ---- staging of the scatter indices (asynchronous; consumed by the scatter at the end of phase B) --------------
    for (int i = threadIdx.x; i < ncl * (L.mapstride / 8); i += blockDim.x) {
        const int cl = i / (L.mapstride / 8), ch = i - cl * (L.mapstride / 8);
        const int64_t cell = fb2_cell_of(A, cell0 + cl);
        fb2_cp_async16(s_map + (size_t)cl * L.mapstride + ch * 8, A.mapc + (size_t)cell * L.mapstride + ch * 8);
    }
    for (int i = threadIdx.x; i < ncl * N; i += blockDim.x) {
        const int cl = i / N, jl = i - cl * N;
        const int64_t cell = fb2_cell_of(A, cell0 + cl);
        s_dof[i] = __ldg(A.dofc + (size_t)cell * N + jl);   // cell-major: one contiguous run per cell, no dependent gather
        s_base[i] = __ldg(A.basec + (size_t)cell * N + jl);
    }

    // ---- phase A1: per (qp, cell): J, det > 0, J^-1, dOmega ---------------------------------------------------------
    for (int item = threadIdx.x; item < NQ * ncl; item += blockDim.x) {
        const int q = item / ncl, cl = item - q * ncl;
        const int64_t cell = fb2_cell_of(A, cell0 + cl);
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
        s_dO[q * CELLS + cl] = det * tw[q];
        double* Jq = s_Ji + ((size_t)q * CELLS + cl) * DIM * DIM;
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = 0; b < DIM; ++b) Jq[a * DIM + b] = Ji[a][b];
    }
    __syncthreads();
    // ---- phase A2: per (qp, cell, basis function): dNdx = dNdxi . J^-1 ------------------------------------------------
    for (int item = threadIdx.x; item < NQ * ncl * NBS; item += blockDim.x) {
        const int i = item % NBS;
        const int rest = item / NBS;
        const int cl = rest % ncl, q = rest / ncl;
        const double* Jq = s_Ji + ((size_t)q * CELLS + cl) * DIM * DIM;
        double* gq = s_g + (((size_t)q * CELLS + cl) * NBS + i) * DIM;
#pragma unroll
        for (int b = 0; b < DIM; ++b) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) s = fma(tdN[(q * NBS + i) * DIM + a], Jq[a * DIM + b], s);
            gq[b] = s;
        }
    }
    if (ELEM == FB2_ELEM_NEOHOOKE) {
        __syncthreads();
        // ---- phase A3 (Neo-Hooke): per (qp, cell): F = I + sum_a u_a (x) g_a, S, dS/dC -> P dOmega, dP/dF dOmega --------
        for (int item = threadIdx.x; item < NQ * ncl; item += blockDim.x) {
            const int q = item / ncl, cl = item - q * ncl;
            const int64_t cell = fb2_cell_of(A, cell0 + cl);
            const double dO = s_dO[q * CELLS + cl];
            const double* gq = s_g + ((size_t)q * CELLS + cl) * NBS * DIM;
            double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            for (int i = 0; i < NBS; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double uc = __ldg(A.u + s_dof[cl * N + i * 3 + c]);
#pragma unroll
                    for (int b = 0; b < 3; ++b) F[c][b] = fma(uc, gq[i * 3 + b], F[c][b]);
                }
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
            double* Pq = s_P + ((size_t)q * CELLS + cl) * 9;
            double* Aq = s_A + ((size_t)q * CELLS + cl) * 81;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) Pq[i * 3 + j] = (F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j]) * dO;
            // dP_ij/dF_mn = delta_im S_jn + 2 F_ia dSdC_ajkn F_mk with
            // dSdC_ajkn = c1 (Ci_ak Ci_nj + Ci_an Ci_kj) + c2 Ci_aj Ci_kn; with FCi = F Ci and W = FCi F^T:
            // F_ia Ci_ak Ci_nj F_mk = W_im Ci_nj, F_ia Ci_an Ci_kj F_mk = FCi_in FCi_mj, F_ia Ci_aj Ci_kn F_mk = FCi_ij FCi_mn
            double FCi[3][3], W[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) FCi[i][k] = F[i][0] * Ci[0][k] + F[i][1] * Ci[1][k] + F[i][2] * Ci[2][k];
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
                            double t = c1 * (W[i][m] * Ci[n][j] + FCi[i][n] * FCi[m][j]) + c2 * FCi[i][j] * FCi[m][n];
                            Aq[((i * 3 + j) * 3 + m) * 3 + n] = ((i == m ? S[j][n] : 0.0) + 2.0 * t) * dO;
                        }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // ---- phase B: thread per (cell, row node a, tile of TB column nodes) ---------------------------------------------
    // Elasticity: only G_ab = sum_q dOmega g_a (x) g_b is accumulated (9 FMA per qp and node pair); the block is
    // K_ab = lam G_ab + mu G_ab^T + mu tr(G_ab) I  (from lam g_a[c] g_b[d] + mu (g_a[d] g_b[c] + delta_cd g_a.g_b)).
    const int nitems = ncl * NT * NBS;
    for (int item = threadIdx.x; item < nitems; item += blockDim.x) {
        const int a = item % NBS;            // row node fastest
        const int rest = item / NBS;
        const int bt = rest % NT, cl = rest / NT;
        const int64_t cell = fb2_cell_of(A, cell0 + cl);
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

#pragma unroll 2
        for (int q = 0; q < NQ; ++q) {
            const double dO = s_dO[q * CELLS + cl];
            const double* gq = s_g + ((size_t)q * CELLS + cl) * NBS * DIM;
            double ga[DIM];
#pragma unroll
            for (int b = 0; b < DIM; ++b) ga[b] = gq[a * DIM + b] * dO;      // dOmega folded into g_a
            const double Na = tN[q * NBS + a] * dO;
            if (ELEM == FB2_ELEM_HEAT) {
                fa[0] += Na;
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
                        double s = acc[t][0][0];
#pragma unroll
                        for (int b = 0; b < DIM; ++b) s = fma(ga[b], gq[(b0 + t) * DIM + b], s);
                        acc[t][0][0] = s;
                    }
                }
            } else if (ELEM == FB2_ELEM_MASS) {
#pragma unroll
                for (int t = 0; t < TB; ++t)
                    if (b0 + t < NBS) acc[t][0][0] = fma(Na, tN[q * NBS + b0 + t], acc[t][0][0]);
            } else if (ELEM == FB2_ELEM_ELASTICITY) {
#pragma unroll
                for (int c = 0; c < VDIM; ++c) fa[c] = fma(Na, A.p[2 + c], fa[c]);
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
                        double gb[DIM];
#pragma unroll
                        for (int b = 0; b < DIM; ++b) gb[b] = gq[(b0 + t) * DIM + b];
#pragma unroll
                        for (int c = 0; c < VDIM; ++c)
#pragma unroll
                            for (int d = 0; d < VDIM; ++d) acc[t][c][d] = fma(ga[c], gb[d], acc[t][c][d]);
                    }
                }
            } else if (ELEM == FB2_ELEM_ELASTICITY_GENERAL) {
                // K_ab[c][d] = sum_{j,n} g_a[j] C[c][j][d][n] g_b[n] dOmega for any stiffness tensor (the isotropic special case
                // is k_cell_syrk); C is read through the L1 (81 doubles shared by all threads)
#pragma unroll
                for (int c = 0; c < VDIM; ++c) fa[c] = fma(Na, A.p[2 + c], fa[c]);
                double h[VDIM][VDIM][DIM];
#pragma unroll
                for (int c = 0; c < VDIM; ++c)
#pragma unroll
                    for (int d = 0; d < VDIM; ++d)
#pragma unroll
                        for (int n = 0; n < DIM; ++n) {
                            double s = 0.0;
#pragma unroll
                            for (int j = 0; j < DIM; ++j) s = fma(ga[j], __ldg(A.cmat + ((c * 3 + j) * 3 + d) * 3 + n), s);
                            h[c][d][n] = s;
                        }
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
#pragma unroll
                        for (int c = 0; c < VDIM; ++c)
#pragma unroll
                            for (int d = 0; d < VDIM; ++d) {
                                double v = acc[t][c][d];
#pragma unroll
                                for (int n = 0; n < DIM; ++n) v = fma(h[c][d][n], gq[(b0 + t) * DIM + n], v);
                                acc[t][c][d] = v;
                            }
                    }
                }
            } else {  // neo-hooke: K_ab[c][d] = sum_{j,n} ga[j] A[c][j][d][n] gb[n]   (dOmega is inside A and P)
                const double* Aq = s_A + ((size_t)q * CELLS + cl) * 81;
                const double* Pq = s_P + ((size_t)q * CELLS + cl) * 9;
                double gr[3];
#pragma unroll
                for (int b = 0; b < 3; ++b) gr[b] = gq[a * 3 + b];
                double h[3][3][3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int n = 0; n < 3; ++n) {
                            double s = 0.0;
#pragma unroll
                            for (int j = 0; j < 3; ++j) s = fma(gr[j], Aq[((c * 3 + j) * 3 + d) * 3 + n], s);
                            h[c][d][n] = s;
                        }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double s = fa[c];
#pragma unroll
                    for (int j = 0; j < 3; ++j) s = fma(gr[j], Pq[c * 3 + j], s);
                    fa[c] = fma(-Na, A.p[2 + c], s);
                }
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    if (b0 + t < NBS) {
                        double gb[3];
#pragma unroll
                        for (int b = 0; b < 3; ++b) gb[b] = gq[(b0 + t) * 3 + b];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                double v = acc[t][c][d];
#pragma unroll
                                for (int n = 0; n < 3; ++n) v = fma(h[c][d][n], gb[n], v);
                                acc[t][c][d] = v;
                            }
                    }
                }
            }
        }
        // scatter the node blocks (a, b0..b0+TB-1): column bases and offsets come from shared memory
        const double kscale = (ELEM == FB2_ELEM_HEAT || ELEM == FB2_ELEM_MASS) ? A.p[0] : 1.0;
        const uint16_t* mcell = s_map + (size_t)cl * L.mapstride;
        const int64_t* bcell = s_base + (size_t)cl * N;
        bool missing = false;
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            if (b0 + t < NBS) {
                double tr = 0.0;
                if (ELEM == FB2_ELEM_ELASTICITY) {
#pragma unroll
                    for (int c = 0; c < VDIM; ++c) tr += acc[t][c][c];
                }
#pragma unroll
                for (int d = 0; d < VDIM; ++d) {
                    const int jl = (b0 + t) * VDIM + d;
                    const int64_t base = *reinterpret_cast<const volatile int64_t*>(&bcell[jl]);
                    // volatile: read before the zero tests instead of one exposed shared-memory round trip per RED
                    unsigned offs[VDIM];
#pragma unroll
                    for (int c = 0; c < VDIM; ++c) offs[c] = *reinterpret_cast<const volatile uint16_t*>(&mcell[jl * N + a * VDIM + c]);
#pragma unroll
                    for (int c = 0; c < VDIM; ++c) {
                        double v;
                        if (ELEM == FB2_ELEM_ELASTICITY) v = A.p[0] * acc[t][c][d] + A.p[1] * (acc[t][d][c] + (c == d ? tr : 0.0));
                        else v = kscale * acc[t][c][d];
                        if (TRANSPOSE) {
                            s_K[(size_t)cl * N * N + jl * N + a * VDIM + c] = v;
                        } else {
                            const unsigned off = offs[c];
                            if (v != 0.0) {
                                if (off == 0xFFFFu) missing = true;
                                else fb2_add<ATOMIC>(A.nzval + base + off, v);
                            }
                        }
                    }
                }
            }
        }
        if (missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
        if (bt == 0 && A.f != nullptr && ELEM != FB2_ELEM_MASS) {
            const double fscale = ELEM == FB2_ELEM_HEAT ? A.p[1] : 1.0;
#pragma unroll
            for (int c = 0; c < VDIM; ++c) fb2_add<ATOMIC>(A.f + s_dof[cl * N + a * VDIM + c], fscale * fa[c]);
        }
    }
    if (TRANSPOSE) {
        __syncthreads();
        for (int e2 = threadIdx.x; e2 < ncl * N * N; e2 += blockDim.x) {
            const int cl = e2 / (N * N), e = e2 - cl * (N * N);   // e = jl * N + il: lanes walk down the rows of a column
            const double v = s_K[e2];
            if (v != 0.0) {
                const unsigned off = s_map[(size_t)cl * L.mapstride + e];
                if (off == 0xFFFFu) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, fb2_cell_of(A, cell0 + cl));
                else fb2_add<ATOMIC>(A.nzval + s_base[cl * N + e / N] + off, v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_cell_syrk: isotropic elasticity (VDIM == DIM) on the FP64 tensor cores.
//
// With X[q][(a,c)] = d phi_a / d x_c at quadrature point q (a row per point, a column per local dof) the node-pair sums
// of k_cell_blocks are one symmetric rank-NQ update per cell,
//     G = X^T diag(dOmega) X          (N x N, N = NBS * DIM),
// and Ke[(a,c),(b,d)] = lam G[(a,c),(b,d)] + mu G[(a,d),(b,c)] + delta_cd mu sum_k G[(a,k),(b,k)].
// G is computed in 8 x 8 tiles with mma.sync.m8n8k4.f64 (DMMA): one warp instruction is 256 FMAs, so the instruction
// stream of the contraction shrinks 8x against the DFMA version, which was issue/latency bound (19 % of the FP64 peak on
// Q2 hexahedra).  Only the upper-triangular tiles are computed, accumulators stay in registers, and the result goes to
// shared memory (mirrored) so that the scatter walks down the rows of a column with consecutive lanes: consecutive rows
// of a CSC column share 32-byte sectors, which is what the L2 atomic units are bound by.
//
// Work split (WPC = warps per cell): elements with N < 48 are handled by ONE warp per cell (four independent warps per
// CTA, __syncwarp only: no CTA barrier, the phases of different warps interleave and hide each other's latency); Q2
// hexahedra (N = 81, 66 tiles) use one CTA of 8 warps per cell, tile t going to warp t mod 8.  The tile loop is unrolled
// at compile time, so tile coordinates and accumulator slots are immediates.
template <int NBS, int DIM>
struct SyrkOf {
    static constexpr int N = NBS * DIM;
    static constexpr int NTI = (N + 7) / 8;                             // 8-wide tile rows / columns
    static constexpr int NP = (NTI * 8) % 16 == 0 ? NTI * 8 + 8 : NTI * 8;  // row stride of X (doubles), = 8 mod 16: the four
                                                                        // quadrature rows of a fragment load hit disjoint banks
    static constexpr int NG = N % 2 == 0 ? N + 1 : N;                   // row stride of G: odd -> column walks are conflict-free
    static constexpr int NTU = NTI * (NTI + 1) / 2;                     // upper-triangular tiles per cell
    static constexpr int WPC = N >= 48 ? 8 : 1;                         // warps per cell
    static constexpr int NTHR = WPC == 1 ? 128 : 32 * WPC;              // threads per CTA
    static constexpr int CELLS = WPC == 1 ? 4 : 1;                      // cells per CTA
    static constexpr int TPW = (NTU + WPC - 1) / WPC;                   // tiles (accumulator pairs) per warp
};

struct SyrkSmem {
    size_t X, dO, G, base, dof, map, cell;  // byte offsets inside a cell's region, and its size; J^-1 of phase A aliases G
    int nqp, mapstride;
};

template <int NBS, int DIM>
__host__ __device__ inline SyrkSmem fb2_syrk_smem(int nq) {
    using S = SyrkOf<NBS, DIM>;
    SyrkSmem L;
    L.nqp = (nq + 3) / 4 * 4;
    size_t o = 0;
    L.X = o; o += sizeof(double) * (size_t)L.nqp * S::NP;
    L.dO = o; o += sizeof(double) * (size_t)L.nqp;
    const size_t gbytes = sizeof(double) * (size_t)S::N * S::NG;
    const size_t jbytes = sizeof(double) * (size_t)nq * DIM * DIM;
    L.G = o; o += gbytes > jbytes ? gbytes : jbytes;
    L.base = o; o += sizeof(int64_t) * (size_t)S::N;
    L.dof = o; o += sizeof(int32_t) * (size_t)S::N;
    o = (o + 15) / 16 * 16;
    L.mapstride = (S::N * S::N + 7) / 8 * 8;
    // one CTA per cell (WPC > 1): the staged scatter indices reuse the X buffer after the contraction, which brings
    // Q2 hexahedra from 86 KB to 73 KB per CTA = three CTAs per SM
    if (S::WPC > 1 && sizeof(uint16_t) * (size_t)L.mapstride <= sizeof(double) * (size_t)L.nqp * S::NP) L.map = L.X;
    else { L.map = o; o += sizeof(uint16_t) * (size_t)L.mapstride; }
    L.cell = (o + 15) / 16 * 16;
    return L;
}

__device__ __forceinline__ void fb2_dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int DIM, int NGEO, int NBS, bool ATOMIC>
__global__ void __launch_bounds__(SyrkOf<NBS, DIM>::NTHR) k_cell_syrk(const AsmArgs A) {
    using S = SyrkOf<NBS, DIM>;
    constexpr int N = S::N, NP = S::NP, NG = S::NG, NTI = S::NTI, TPW = S::TPW, WPC = S::WPC;
    constexpr int GS = 32 * WPC;             // threads working on one cell
    extern __shared__ __align__(16) unsigned char smraw[];
    const int NQ = A.nq;
    const int64_t np = A.ncells_pad;
    const SyrkSmem L = fb2_syrk_smem<NBS, DIM>(NQ);
    const int NQP = L.nqp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gt = WPC == 1 ? lane : (int)threadIdx.x;   // thread index inside the cell's group
    const int64_t ci = WPC == 1 ? (int64_t)blockIdx.x * S::CELLS + warp : (int64_t)blockIdx.x;
    if (ci >= A.ncount) return;              // a whole group leaves together
    const int64_t cell = fb2_cell_of(A, ci);
    unsigned char* smc = smraw + (WPC == 1 ? (size_t)warp * L.cell : 0);
    double* s_X = reinterpret_cast<double*>(smc + L.X);          // [NQP][NP]
    double* s_dO = reinterpret_cast<double*>(smc + L.dO);        // [NQP]
    double* s_G = reinterpret_cast<double*>(smc + L.G);          // [N][NG]
    double* s_Ji = s_G;                                          // [NQ][DIM*DIM] (phase A only)
    int64_t* s_base = reinterpret_cast<int64_t*>(smc + L.base);  // [N]
    int32_t* s_dof = reinterpret_cast<int32_t*>(smc + L.dof);    // [N]
    uint16_t* s_map = reinterpret_cast<uint16_t*>(smc + L.map);  // [mapstride]
    // tables from global memory (L1-resident): the indices below differ per lane, which the constant bank would serialise
    const double* __restrict__ tw = A.tab + A.o_w;
    const double* __restrict__ tN = A.tab + A.o_N;
    const double* __restrict__ tdN = A.tab + A.o_dN;
    const double* __restrict__ tdM = A.tab + A.o_dM;
#define FB2_GROUP_SYNC() do { if (WPC == 1) __syncwarp(); else __syncthreads(); } while (0)

    // scatter indices: asynchronous staging, consumed after the contraction.  When they share the X buffer they are
    // only pulled into L2 here and staged once X is dead.
    const bool late_map = L.map == L.X;
    if (late_map) {
        for (int i = gt; i < L.mapstride / 64; i += GS)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.mapc + (size_t)cell * L.mapstride + i * 64));
    } else {
        for (int i = gt; i < L.mapstride / 8; i += GS) fb2_cp_async16(s_map + i * 8, A.mapc + (size_t)cell * L.mapstride + i * 8);
    }
    for (int i = gt; i < N; i += GS) {
        s_dof[i] = __ldg(A.dofc + (size_t)cell * N + i);   // cell-major: one contiguous run per cell, no dependent gather
        s_base[i] = __ldg(A.basec + (size_t)cell * N + i);
    }
    // quadrature rows NQ..NQP-1 are padding of the k dimension: zero.  Padding columns are left
    // uninitialised: they only reach G entries with a row or column >= N, which are never read
    for (int i = gt; i < (NQP - NQ) * NP; i += GS) s_X[NQ * NP + i] = 0.0;
    for (int i = gt; i < NQP - NQ; i += GS) s_dO[NQ + i] = 0.0;
    // node coordinates (the X buffer is free until phase A2)
    double* s_x = s_X;                                           // [NGEO][DIM]
    for (int j = gt; j < NGEO; j += GS) {
        const int node = __ldg(A.conn + (size_t)j * np + cell);
        double xj[DIM];
        fb2_load_x<DIM>(A.xyz, node, xj);
#pragma unroll
        for (int a = 0; a < DIM; ++a) s_x[j * DIM + a] = xj[a];
    }
    FB2_GROUP_SYNC();
    // ---- phase A1: thread per (qp, a, b): J_ab = sum_j x_j[a] dM_j/dxi_b ------------------------------------------------
    for (int item = gt; item < NQ * DIM * DIM; item += GS) {
        const int q = item / (DIM * DIM), ab = item - q * DIM * DIM;
        const int a = ab / DIM, b = ab - a * DIM;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NGEO; ++j) s = fma(s_x[j * DIM + a], tdM[(q * NGEO + j) * DIM + b], s);
        s_Ji[item] = s;
    }
    FB2_GROUP_SYNC();
    // ---- thread per qp: det > 0, J^-1 (in place), dOmega ------------------------------------------------------------------
    for (int q = gt; q < NQ; q += GS) {
        double J[DIM][DIM], Ji[DIM][DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = 0; b < DIM; ++b) J[a][b] = s_Ji[q * DIM * DIM + a * DIM + b];
        const double det = fb2_det_inv<DIM>(J, Ji);
        if (!(det > 0.0)) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
        s_dO[q] = det * tw[q];
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = 0; b < DIM; ++b) s_Ji[q * DIM * DIM + a * DIM + b] = Ji[a][b];
    }
    FB2_GROUP_SYNC();
    // ---- phase A2: X[q][(i, b)] = dN_i/dxi . J^-1 ---------------------------------------------------------------------
    for (int item = gt; item < NQ * NBS; item += GS) {
        const int q = item / NBS, i = item - q * NBS;
        const double* Jq = s_Ji + q * DIM * DIM;
#pragma unroll
        for (int b = 0; b < DIM; ++b) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) s = fma(tdN[(q * NBS + i) * DIM + a], Jq[a * DIM + b], s);
            s_X[q * NP + i * DIM + b] = s;
        }
    }
    FB2_GROUP_SYNC();

    // ---- contraction: G tiles on the tensor cores -------------------------------------------------------------------------
    const int kr = lane & 3, mc = lane >> 2;   // fragment coordinates: A[mc][kr], B[kr][mc], C[mc][2 kr + {0,1}]
    double acc[TPW][2];
#pragma unroll
    for (int s = 0; s < TPW; ++s) acc[s][0] = acc[s][1] = 0.0;
    // one specialised copy of the loop per warp of the group (W is a compile-time constant inside): a warp issues only
    // its own tiles' fragment loads and DMMAs, and the branch is warp-uniform
#pragma unroll
    for (int W = 0; W < WPC; ++W) {
        if (WPC > 1 && warp != W) continue;
        for (int ks = 0; ks < NQP / 4; ++ks) {
            const double w = s_dO[4 * ks + kr];
            const double* xrow = s_X + (4 * ks + kr) * NP + mc;
            double xb[NTI], xw[NTI];
#pragma unroll
            for (int t = 0; t < NTI; ++t) { xb[t] = xrow[8 * t]; xw[t] = xb[t] * w; }
            int t = 0;
#pragma unroll
            for (int ti = 0; ti < NTI; ++ti)
#pragma unroll
                for (int tj = ti; tj < NTI; ++tj, ++t)
                    if (t % WPC == W) fb2_dmma884(acc[t / WPC], xw[ti], xb[tj]);
        }
        // G (and its mirror image) to shared memory; J^-1 aliased this buffer, the group is past phase A2
        int t = 0;
#pragma unroll
        for (int ti = 0; ti < NTI; ++ti)
#pragma unroll
            for (int tj = ti; tj < NTI; ++tj, ++t)
                if (t % WPC == W) {
                    const int r = 8 * ti + mc, c0 = 8 * tj + 2 * kr;
                    const double v0 = acc[t / WPC][0], v1 = acc[t / WPC][1];
                    if (r < N) {
                        if (c0 < N) s_G[r * NG + c0] = v0;
                        if (c0 + 1 < N) s_G[r * NG + c0 + 1] = v1;
                        if (ti != tj) {
                            if (c0 < N) s_G[c0 * NG + r] = v0;
                            if (c0 + 1 < N) s_G[(c0 + 1) * NG + r] = v1;
                        }
                    }
                }
    }
    if (late_map) {
        FB2_GROUP_SYNC();   // every warp is done reading X
        for (int i = gt; i < L.mapstride / 8; i += GS) fb2_cp_async16(s_map + i * 8, A.mapc + (size_t)cell * L.mapstride + i * 8);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    FB2_GROUP_SYNC();

    // ---- scatter: a thread owns row il = (a, c) and walks over the columns (b, d); the groups of N consecutive threads
    // take every (GS / N)-th column node.  G is symmetric, so G[(b,.)][(a,.)] is read: consecutive lanes, consecutive words.
    constexpr int NGRP = GS / N;
    const int grp = gt / N, il = gt - grp * N;
    if (grp < NGRP) {
        const double lam = A.p[0], mu = A.p[1];
        const int a = il / DIM, c = il - a * DIM;
        int missing = 0;
        for (int b = grp; b < NBS; b += NGRP) {
            const double* Gb = s_G + (b * DIM) * NG;       // rows (b, 0..DIM-1)
            // the offsets and column bases of the DIM entries are read up front through volatile pointers: left to the
            // compiler they sink behind the zero test of each value, one exposed shared-memory round trip per RED
            unsigned offs[DIM];
            int64_t bases[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                offs[d] = *reinterpret_cast<const volatile uint16_t*>(&s_map[(b * DIM + d) * N + il]);
                bases[d] = *reinterpret_cast<const volatile int64_t*>(&s_base[b * DIM + d]);
            }
            double tr = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) tr += Gb[k * NG + a * DIM + k];
            const double* Gbc = Gb + c * NG + a * DIM;     // G[(b,c)][(a,.)]
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double v = lam * Gb[d * NG + il] + mu * Gbc[d];
                if (c == d) v = fma(mu, tr, v);
                if (v != 0.0) {
                    if (offs[d] == 0xFFFFu) missing = 1;
                    else fb2_add<ATOMIC>(A.nzval + bases[d] + offs[d], v);
                }
            }
        }
        if (missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
        if (A.f != nullptr && grp == 0) {
            double s = 0.0;
            for (int q = 0; q < NQ; ++q) s = fma(tN[q * NBS + a], s_dO[q], s);
            fb2_add<ATOMIC>(A.f + s_dof[il], s * A.p[2 + c]);
        }
    }
#undef FB2_GROUP_SYNC
}
