// Marching-tile assembly kernel for isotropic linear elasticity on trilinear hexahedra (three dofs per grid node), sm_100a.
// Included by assemble.cu after march_kernels.cuh, whose window layout, parity pads and bulk (TMA) flush it reuses.
//
// The warp-per-cell kernel (k_cell_syrk) ends in 576 FP64 REDs per cell; 243 of them are distinct matrix entries, and the
// chip retires scattered REDs at 120-320 G/s (profiles/r02_red_micro.txt): 2.1 M cells x 576 = 1.2 G REDs = 4-7 ms on
// BASELINE.json configs[4]'s element.  As in k_march_hex the contributions are therefore summed on chip and finished
// matrix columns leave through the TMA engine:
//
//   * a CTA of four warps owns a tile of 4 x 4 cells in (x, y) and marches through the layers of its z-chunk; shared memory
//     holds a two-plane window of the matrix, one copy of the three CSC columns of every one of the 5 x 5 tile nodes of node
//     planes z and z + 1, laid out like the global columns (the assemble! offset map indexes it: src/assembler.jl:347-457);
//     2 x 48 KB, two CTAs per SM;
//   * a WARP integrates one cell (reinit! + the element routine of docs/src/literate-tutorials/linear_elasticity.jl:266-281
//     for C = lambda 1x1 + 2 mu I_sym): lane (m, k) = (lane >> 2, lane & 3) evaluates the gradient of shape function m at the
//     quadrature points k and k + 4 from the analytic trilinear Jacobian (the same closed form as fb2_hex8_heat), and these
//     are exactly its A / B fragments of the FP64 tensor-core contraction H_cd[a][b] = sum_q dOmega g_ac g_bd
//     (mma.m8n8k4.f64, one 8 x 8 tile per component pair (c, d), k = quadrature points): no staging of gradients or of G in
//     shared memory.  After 18 DMMAs the lane holds the complete 3 x 3 blocks of the node pairs (a, 2k) and (a, 2k + 1), so
//     Ke = lambda H + mu H^T + mu tr(H) I is lane-local: 18 entries per lane = 576 per cell;
//   * warp w takes the cells x = w of the tile, in sub-step s the cell y = (s + 2 w) mod 4: cells integrated at the same time
//     never share a node, so the window is updated with plain shared-memory read-modify-writes (no shared atomics), one CTA
//     barrier per sub-step;
//   * after layer z the columns of node plane z are final inside the tile: the nine columns of the three tile-interior nodes
//     of a tile row are one contiguous 5.8 KB piece of nzval -> one cp.async.bulk store; columns of nodes on the tile faces
//     and on the first / last plane of a chunk are shared with other CTAs -> cp.reduce.async.bulk.add.f64.  15 bulk
//     operations per layer and CTA replace 9 216 REDs.  A bulk operation takes its operands from uniform registers, so the
//     pieces of the lanes of one warp are issued one after the other: the 15 pieces sit on lanes 0..3 of the four warps;
//   * the set-up of node plane z + 2 (dofs, colptr extents, scan of the column copies) is spread over the sub-steps of
//     layer z, one step per barrier interval, with the loads issued one interval before their use; the offset-map words of a
//     cell are requested one sub-step ahead;
//   * the window layout is padded against shared-memory bank conflicts (fb2_mvec_cap).
// Measured on one B200 (128^3 cells, profiles/r02_prof_c5_march_*.txt): 3.8 ms against 5.05 ms for k_cell_syrk; what is left
// is latency: two CTAs (eight warps) per SM is all that 2 x 104 KB of window and 232 registers allow, no unit is busier than
// 42 % (shared-memory wavefronts), FP64 pipe 23 %, DRAM 23 %.  A flush by coalesced per-thread stores / REDs with fused
// clearing of the window (FB2_MVEC_FLUSH=thread) costs 26 % more instructions and runs at 4.5 ms.
#pragma once

constexpr int MV_PN = 25;   // nodes of a tile plane (5 x 5)
constexpr int MV_NC = 75;   // matrix columns of a tile plane
constexpr int MV_CS = 76;   // padded

// Doubles per window plane: the 75 column copies, up to 76 parity pads and up to 4 x 14 row pads (fb2_mvec_scan), rounded up
// to 12 mod 16.  The residues matter: in the window updates of the kernel the 16 lanes of a half-warp hit the rows of four
// nodes (offsets {0, 3, 9, 12} + const inside a column) in the columns of four nodes that lie 0 / R +- 243 / P / P + R +- 243
// doubles apart (R = distance of two tile rows, P = of the two planes).  The four groups of four 8-byte banks overlap at
// most pairwise iff 7 + R and 5 + P (mod 16) are in {1, 2, 5, 8, 11, 14, 15}; the dense layout (R = 1215, P = 6152) gave
// three- to four-fold conflicts on every access (profiles/r02_prof_c5_march_d.txt: 85 conflicts per cell; grids with an even
// number of nodes per row ran 25 % slower than those with an odd one).
__host__ __device__ inline int fb2_mvec_cap(int max_col_len) { return (MV_NC * max_col_len + MV_CS + 56 + 3) / 16 * 16 + 12; }
// doubles: [2][cap] matrix window | [2][CS] load vector | [3][PN][4] node coordinates; int64 [3][CS] colptr[dof]; int [3][CS] dof;
// unsigned [3][4] adjacency bits; uint16 [3][CS] start of the column copy; uint8 [3][CS] its length.  106 KB for 81-entry
// columns: two CTAs per SM.
__host__ __device__ inline size_t fb2_mvec_smem(int cap) {
    return sizeof(double) * (2 * (size_t)cap + 2 * MV_CS + 3 * MV_PN * 4) + sizeof(int64_t) * 3 * MV_CS + sizeof(int) * (3 * MV_CS + 12) +
           sizeof(uint16_t) * 3 * MV_CS + sizeof(uint8_t) * 3 * MV_CS + 16;
}

// Lane-major byte map of the kernel: word k (0..4) of lane l of cell c at mapv[(c * 5 + k) * 32 + l] holds bytes 4k .. 4k + 3 of
// the lane's 18 offsets; byte t = (bs * 3 + d) * 3 + c is the offset of row (a, c) inside column (b, d), a = l >> 2,
// b = 2 (l & 3) + bs; 0xFF = no such pattern entry.
__global__ void k_pack_map_vec(const uint16_t* __restrict__ map, int64_t ncells, int64_t ncells_pad, uint32_t* __restrict__ mapv) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * 160) return;
    const int64_t cell = t % ncells;     // cells fastest: the reads of the SoA map are coalesced
    const int r = (int)(t / ncells), k = r >> 5, lane = r & 31;
    const int a = lane >> 2, kr = lane & 3;
    unsigned word = 0;
    for (int bb = 0; bb < 4; ++bb) {
        const int tt = 4 * k + bb;
        unsigned byte = 0xFFu;
        if (tt < 18) {
            const int bs = tt / 9, d = (tt % 9) / 3, c = tt % 3;
            const int j = 3 * (2 * kr + bs) + d, i = 3 * a + c;
            const unsigned v = map[(size_t)(j * 24 + i) * ncells_pad + cell];
            byte = v >= 0xFFu ? 0xFFu : v;
        }
        word |= byte << (8 * bb);
    }
    mapv[((size_t)cell * 5 + k) * 32 + lane] = word;
}

__device__ __forceinline__ int fb2_warp_iscan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Position of the 75 column copies of a node plane (warp 0; lane l takes columns l, l + 32, l + 64): exclusive scan of the
// lengths plus parity pads (see fb2_march_plane_finish), and one bit per column "the next column follows it in nzval".
__device__ __forceinline__ void fb2_mvec_scan(const int64_t* s_gb, const uint8_t* s_len, uint16_t* s_cs, unsigned* s_adj, int lane) {
    const unsigned full = 0xffffffffu;
    int len[3];
    int64_t gb[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int k = lane + 32 * s;
        len[s] = k < MV_NC ? (int)s_len[k] : 0;
        gb[s] = k < MV_NC ? s_gb[k] : 0;
    }
    int carry = 0, carry_p = 0, prev_t = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int k = lane + 32 * s;
        const int inc = fb2_warp_iscan(len[s], lane);
        const int csu = carry + inc - len[s];
        carry += __shfl_sync(full, inc, 31);
        const int t = (csu + (int)gb[s]) & 1;     // parity mismatch of the unpadded layout; pad where it changes
        int pv = __shfl_up_sync(full, t, 1);
        if (lane == 0) pv = prev_t;
        const int incp = fb2_warp_iscan(t ^ pv, lane);
        if (k < MV_NC) s_cs[k] = (uint16_t)(csu + carry_p + incp);
        carry_p += __shfl_sync(full, incp, 31);
        prev_t = __shfl_sync(full, t, 31);
        int64_t nb = __shfl_down_sync(full, gb[s], 1);
        const int64_t nb0 = __shfl_sync(full, gb[s < 2 ? s + 1 : 2], 0);
        if (lane == 31) nb = s < 2 ? nb0 : -1;
        const unsigned m = __ballot_sync(full, len[s] > 0 && k + 1 < MV_NC && nb == gb[s] + len[s]);
        if (lane == 0) s_adj[s] = m;
    }
    if (lane == 0) s_adj[3] = 0u;
    __syncwarp();
    // row pads (see fb2_mvec_cap): shift tile row b >= 1 by an even amount such that its distance to row b - 1 becomes
    // 1, 7, 11 (odd distances) or 4, 8, 10, 14 (even ones) mod 16; every lane derives the four shifts
    int start[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) start[b] = s_cs[15 * b];
    int shift[5];
    shift[0] = 0;
#pragma unroll
    for (int b = 1; b < 5; ++b) {
        const int d = (start[b] - start[b - 1]) & 15;
        // smallest even x with (d + x) mod 16 in the allowed set of d's parity
        int x;
        if (d & 1) x = d <= 1 ? 1 - d : (d <= 7 ? 7 - d : (d <= 11 ? 11 - d : 17 - d));
        else x = d <= 4 ? 4 - d : (d <= 8 ? 8 - d : (d <= 10 ? 10 - d : (d <= 14 ? 14 - d : 20 - d)));
        shift[b] = shift[b - 1] + x;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int k = lane + 32 * s;
        if (k < MV_NC) {
            const int b = k / 15;
            const int sh = b == 0 ? 0 : (b == 1 ? shift[1] : (b == 2 ? shift[2] : (b == 3 ? shift[3] : shift[4])));
            s_cs[k] = (uint16_t)(s_cs[k] + sh);
        }
    }
}

// One run of a finished plane (adjacent in nzval and in the window) -> global memory, by one warp: plain stores (STORE) or
// REDs, two entries per lane and instruction after peeling a leading / trailing entry where the run starts / ends on an odd
// entry (window copy and global column start at the same parity).  The window is cleared on the way.
template <bool STORE>
__device__ __forceinline__ void fb2_mvec_run(double* __restrict__ nzval, double* acc, int c0, int total, int64_t g0, int lane) {
    if (total <= 0) return;
    double* g = nzval + g0;
    double* a = acc + c0;
    int i0 = 0;
    if (g0 & 1) {
        if (lane == 0) {
            const double v = a[0];
            a[0] = 0.0;
            if (STORE) g[0] = v; else atomicAdd(g, v);
        }
        i0 = 1;
    }
    const int body = (total - i0) & ~1;
    for (int i = i0 + 2 * lane; i < i0 + body; i += 64) {
        const double2 v = *reinterpret_cast<const double2*>(a + i);
        *reinterpret_cast<double2*>(a + i) = make_double2(0.0, 0.0);
        if (STORE) *reinterpret_cast<double2*>(g + i) = v;
        else { atomicAdd(g + i, v.x); atomicAdd(g + i + 1, v.y); }
    }
    if (((total - i0) & 1) && lane == 0) {
        const double v = a[i0 + body];
        a[i0 + body] = 0.0;
        if (STORE) g[i0 + body] = v; else atomicAdd(g + i0 + body, v);
    }
}

// Write a finished node plane out and clear its window (after the CTA barrier behind the last update).  The unit of work is
// a tile node (its three columns: one run of nzval in the usual numberings, else column by column): warp w takes the units
// w, w + 4, ...; units 0..8 are the tile-interior nodes -- their columns have received every contribution they will ever get
// and are written with plain coalesced stores unless `redall` --, units 9..24 the nodes on the tile faces, which are shared
// with the neighbouring tiles and go out as coalesced REDs (fire and forget: unlike a TMA reduce they do not hold the window
// until the L2 has taken them, 2300 cycles per layer in profiles/r02_prof_c5_march_a.txt).
__device__ __forceinline__ void fb2_mvec_flush(const AsmArgs& A, double* acc, double* sf, const uint16_t* s_cs, const uint8_t* s_len,
                                               const int64_t* s_gb, int* s_dof, const unsigned* s_adj, int tid, bool redall, bool with_f) {
    const int lane = tid & 31, warp = tid >> 5;
    for (int u = warp; u < MV_PN; u += 4) {
        int n;
        if (u < 9) n = 5 * (1 + u / 3) + 1 + u % 3;
        else {
            const int j = u - 9;
            n = j < 5 ? j : (j < 10 ? 15 + j : (j < 13 ? 5 * (j - 9) : 5 * (j - 12) + 4));
        }
        const bool store = !redall && u < 9;
        const int k0 = 3 * n;
        const unsigned bits = (unsigned)((((unsigned long long)s_adj[k0 >> 5] | ((unsigned long long)s_adj[(k0 >> 5) + 1] << 32)) >> (k0 & 31)) & 3ull);
        if (bits == 3u) {
            const int c0 = s_cs[k0], total = (int)s_cs[k0 + 2] + (int)s_len[k0 + 2] - c0;
            if (store) fb2_mvec_run<true>(A.nzval, acc, c0, total, s_gb[k0], lane);
            else fb2_mvec_run<false>(A.nzval, acc, c0, total, s_gb[k0], lane);
        } else {
            for (int k = k0; k < k0 + 3; ++k) {
                if (store) fb2_mvec_run<true>(A.nzval, acc, s_cs[k], s_len[k], s_gb[k], lane);
                else fb2_mvec_run<false>(A.nzval, acc, s_cs[k], s_len[k], s_gb[k], lane);
            }
        }
    }
    if (tid < MV_NC) {
        const int d = s_dof[tid];
        s_dof[tid] = -1;   // the slot is set up for another node plane two layers from now
        if (with_f && d >= 0) {
            const double v = sf[tid];
            sf[tid] = 0.0;
            const int n = tid / 3, a = n % 5, b = n / 5;
            if (!redall && a >= 1 && a <= 3 && b >= 1 && b <= 3) A.f[d] = v;
            else if (v != 0.0) atomicAdd(A.f + d, v);
        }
    }
}

// The same through the TMA engine (template flag TMAF of the kernel; after fence.proxy.async + CTA barrier): one thread
// takes piece p < 15: tile row b = p / 3, part 0 / 2 = the columns of the face nodes a = 0 / 4 (bulk reduce-add), part 1 = the nine
// columns of the interior nodes a = 1..3 (bulk store unless `redall` or the row is a tile face).  A piece whose columns are not
// contiguous in nzval is written column by column.  The window is NOT cleared: the bulk operations read it asynchronously,
// fb2_bulk_wait_read + a zero fill precede its next use.
__device__ __forceinline__ void fb2_mvec_flush_tma(const AsmArgs& A, const double* acc, double* sf, const uint16_t* s_cs, const uint8_t* s_len,
                                                   const int64_t* s_gb, int* s_dof, const unsigned* s_adj, int tid, bool redall, bool with_f, int dbg) {
    // piece p = 4 * lane + warp on lanes 0..3 of every warp: a bulk operation takes its operands from uniform registers, so a
    // warp issues the pieces of its lanes one after the other -- four per warp instead of fifteen on one warp.  A tile row
    // that is reduce-added as a whole and contiguous in nzval is one piece (part 0).
    const int p = 4 * (tid & 31) + (tid >> 5);
    const bool mine = (tid & 31) < 4 && p < 15 && dbg != 2;
    const int b = p / 3, part = p - 3 * b;
    const bool rowred = redall || b == 0 || b == 4;
    auto contiguous = [&](int k0, int n) -> bool {   // columns k0 .. k0 + n - 1 follow each other in nzval
        const int wd = k0 >> 5, sh = k0 & 31;
        const unsigned long long bits = ((unsigned long long)s_adj[wd] | ((unsigned long long)s_adj[wd + 1] << 32)) >> sh;
        const unsigned long long need = (1ull << (n - 1)) - 1ull;
        return (bits & need) == need;
    };
    int n0 = 15 * b, cnt = 0;
    bool percol = false;
    if (mine) {
        if (rowred && dbg != 1 && contiguous(15 * b, 15)) cnt = part == 0 ? 15 : 0;
        else {
            n0 = 15 * b + (part == 0 ? 0 : (part == 1 ? 3 : 12));
            cnt = part == 1 ? 9 : 3;
            if (!contiguous(n0, cnt)) percol = true;
        }
    }
    const bool store = dbg == 1 || (dbg != 3 && !rowred && part == 1);
    {
        int total = 0, c0 = 0;
        int64_t g0 = 0;
        if (cnt > 0 && !percol) {
            c0 = s_cs[n0];
            total = (int)s_cs[n0 + cnt - 1] + (int)s_len[n0 + cnt - 1] - c0;
            g0 = s_gb[n0];
        }
        fb2_march_emit_piece(A.nzval + g0, acc + c0, total, store);
    }
    if (percol) {   // irregular numbering
        for (int k = 0; k < cnt; ++k) {
            if (store) fb2_march_emit<true>(A.nzval, acc, s_cs, s_len, s_gb, n0 + k, 1);
            else fb2_march_emit<false>(A.nzval, acc, s_cs, s_len, s_gb, n0 + k, 1);
        }
    }
    if (mine) fb2_bulk_commit();
    const int col = 127 - tid;   // the per-column chores sit on the upper warps
    if (col < MV_NC) {
        const int d = s_dof[col];
        s_dof[col] = -1;   // the slot is set up for another node plane two layers from now
        if (with_f && d >= 0) {
            const double v = sf[col];
            sf[col] = 0.0;
            const int n = col / 3, a = n % 5, b = n / 5;
            if (!redall && a >= 1 && a <= 3 && b >= 1 && b <= 3) A.f[d] = v;
            else if (v != 0.0) atomicAdd(A.f + d, v);
        }
    }
}

// A.p: [0] lambda, [1] mu, [2..4] body force, [5] w / 8 (the common weight of the 2 x 2 x 2 Gauss rule; the host checks that
// the CellValues holds the tables of QuadratureRule{RefHexahedron}(2) + Lagrange{RefHexahedron,1}).
// GENERAL: any stiffness tensor with the minor symmetries (FB2_ELEM_ELASTICITY_GENERAL, linear_elasticity.jl:266-281):
// Ke[(a,c),(b,d)] = sum_qs C[c][q][d][s] H_ab[q][s], still lane-local; C travels as a kernel argument (constant bank operands).
struct MarchCmat { double c[81]; };

template <bool CHECK, bool TMAF, bool GENERAL>
__global__ void __launch_bounds__(128, 2) k_march_vec(const AsmArgs A, const MarchArgs M, const MarchCmat CM) {
    constexpr int CS = MV_CS;
    constexpr double NA = fb2_q1n(0, 0), NB = fb2_q1n(1, 0);   // 1-D shape function of the near / far node of a Gauss point
    extern __shared__ __align__(16) unsigned char smraw[];
    const int cap = M.cap;
    double* s_acc = reinterpret_cast<double*>(smraw);                      // [2][cap] matrix window
    double* s_f = s_acc + 2 * (size_t)cap;                                 // [2][CS] load vector window
    double* s_xw = s_f + 2 * CS;                                           // [3][PN][4] node coordinates of three node planes
    int64_t* s_gb = reinterpret_cast<int64_t*>(s_xw + 3 * MV_PN * 4);      // [3][CS] colptr[dof] of the plane's columns
    int* s_dof = reinterpret_cast<int*>(s_gb + 3 * CS);                    // [3][CS] dof of (tile node, component), -1 = none
    unsigned* s_adj = reinterpret_cast<unsigned*>(s_dof + 3 * CS);         // [3][4] column n + 1 follows column n in nzval
    uint16_t* s_cs = reinterpret_cast<uint16_t*>(s_adj + 12);              // [3][CS] start of the column copy in its window plane
    uint8_t* s_len = reinterpret_cast<uint8_t*>(s_cs + 3 * CS);            // [3][CS] its length

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bid = M.ctalist ? __ldg(M.ctalist + blockIdx.x) : (int)blockIdx.x;
    const int tx = bid % M.tiles_x;
    bid /= M.tiles_x;
    const int ty = bid % M.tiles_y, ch = bid / M.tiles_y;
    const int zb = M.z0 + ch * M.lz, ze = min(M.z1, zb + M.lz);
    if (zb >= ze) return;
    const int64_t lay = (int64_t)M.nx * M.ny;
    const int64_t np = A.ncells_pad;

    // ---- cells: lanes 0..3 of warp w look after the cells (x, y) = (w, lane) of the tile ------------------------------------------
    const int cxw = tx * 4 + warp, cyl = ty * 4 + (lane & 3);
    const bool cinside = lane < 4 && cxw < M.nx && cyl < M.ny;
    const int64_t cxy = (int64_t)min(cxw, M.nx - 1) + (int64_t)M.nx * min(cyl, M.ny - 1);
    auto raw_cell = [&](int z) -> int64_t {   // -1 = no cell / not a layer of this chunk
        if (!cinside || z >= ze) return -1;
        return M.cellmap ? (int64_t)__ldg(M.cellmap + cxy + lay * z) : cxy + lay * z;
    };
    auto own = [&](int64_t c) -> int64_t { return (c >= M.cell_lo && c < M.cell_hi) ? c : -1; };
    // the 12 dofs of the lane's cell column in node plane P: from the cell below the plane (its nodes 4..7) if there is one,
    // else from the cell above it (nodes 0..3)
    auto plane_dofs = [&](int64_t cbelow, int64_t cabove, int (&d)[12]) -> bool {
        const int64_t c = cbelow >= 0 ? cbelow : cabove;
        const int base = cbelow >= 0 ? 12 : 0;
        if (c < 0) return false;
#pragma unroll
        for (int i = 0; i < 12; ++i) d[i] = __ldg(A.cell_dofs + (size_t)(base + i) * np + c);
        return true;
    };
    auto publish = [&](int* sd, bool pub, const int (&d)[12]) {   // local nodes 0, 1, 2, 3 of the plane = tile nodes tn, tn+1, tn+6, tn+5
        if (!pub) return;
        const int tn = (lane & 3) * 5 + warp;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sd[tn * 3 + c] = d[c];
            sd[(tn + 1) * 3 + c] = d[3 + c];
            sd[(tn + 6) * 3 + c] = d[6 + c];
            sd[(tn + 5) * 3 + c] = d[9 + c];
        }
    };

    if (M.cellmap) {   // a tile of a partition-local box may hold halo cells only
        bool any = false;
        for (int z = zb; z < ze; z += 8) {
            int64_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = raw_cell(z + k);
#pragma unroll
            for (int k = 0; k < 8; ++k) any |= own(v[k]) >= 0;
        }
        if (!__syncthreads_or(any)) return;
    }

    // ---- node coordinates: thread t < 25 fetches tile node t of a plane (cp.async, 32 bytes) -----------------------------------------
    const int64_t nlay = (int64_t)(M.nx + 1) * (M.ny + 1);
    int64_t mynode = -1;
    if (tid < MV_PN) {
        const int gx = tx * 4 + tid % 5, gy = ty * 4 + tid / 5;
        if (gx <= M.nx && gy <= M.ny) mynode = gx + (int64_t)(M.nx + 1) * gy;
    }
    auto fetch_plane_xyz = [&](int zp) {   // node plane zp -> coordinate slot (zp - zb) % 3
        if (mynode >= 0) {
            double* dst = s_xw + ((size_t)((zp - zb) % 3) * MV_PN + tid) * 4;
            const double* src = A.xyz + 4 * (mynode + nlay * zp);
            fb2_cp_async16(dst, src);
            fb2_cp_async16(dst + 2, src + 2);
        }
    };

    // ---- per-lane constants of the analytic element: lane = (m, k), shape function m, Gauss points k and k + 4 --------------------
    const int mc = lane >> 2, kr = lane & 3;
    const int qx = kr & 1, qy = kr >> 1;
    const double nx0 = qx ? NB : NA, nx1 = qx ? NA : NB, ny0 = qy ? NB : NA, ny1 = qy ? NA : NB;   // n(s, q): s = 0, 1
    const double w00 = nx0 * ny0, w10 = nx1 * ny0, w01 = nx0 * ny1, w11 = nx1 * ny1;               // w[sx][sy]
    const int msx = ((mc & 3) == 1 || (mc & 3) == 2) ? 1 : 0, msy = (mc & 3) >= 2 ? 1 : 0, msz = mc >> 2;
    const double ax = msx ? nx1 : nx0, ay = msy ? ny1 : ny0;
    const double px = msx ? ay : -ay, py = msy ? ax : -ax, pz = msz ? ax * ay : -(ax * ay);      // 2 dN_m/dxi = (px nz, py nz, pz)
    const double nzm0 = msz ? NB : NA, nzm1 = msz ? NA : NB;                                     // n(msz, qz) at qz = 0, 1
    const double lam = A.p[0], mu = A.p[1], w8 = A.p[5];
    const bool with_f = A.f != nullptr;
    const double bforce = kr < 3 ? A.p[2 + kr] : 0.0;

    for (int i = tid; i < 2 * cap + 2 * CS; i += 128) s_acc[i] = 0.0;
    fetch_plane_xyz(zb);
    fetch_plane_xyz(zb + 1);
    asm volatile("cp.async.commit_group;" ::: "memory");

    int64_t c0 = own(raw_cell(zb)), c1 = own(raw_cell(zb + 1)), c2 = own(raw_cell(zb + 2)), c3 = own(raw_cell(zb + 3));
    int64_t r4 = raw_cell(zb + 4);
    int dnext[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) dnext[i] = 0;
    bool pubnext = false;
    // the four steps of a plane set-up; inside the loop one step per barrier interval, here back to back
    int64_t cbeg = 0, cend = 0;
    auto step_colptr = [&](int slot) {
        cbeg = 0; cend = 0;
        if (127 - tid < MV_NC) {
            const int d = s_dof[slot * CS + 127 - tid];
            if (d >= 0) { cbeg = __ldg(A.colptr + d); cend = __ldg(A.colptr + d + 1); }
        }
    };
    auto step_store = [&](int slot) {
        asm volatile("" : "+l"(cbeg), "+l"(cend));   // keeps the consumers of the two loads behind the integration
        if (127 - tid < MV_NC) {
            s_gb[slot * CS + 127 - tid] = cbeg;
            s_len[slot * CS + 127 - tid] = (uint8_t)(cend - cbeg);
        }
    };
    {   // node planes zb (cells above only: the layer below belongs to another chunk) and zb + 1
        for (int i = tid; i < 3 * CS; i += 128) s_dof[i] = -1;
        __syncthreads();
        int d[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) d[i] = 0;
        bool pub = plane_dofs(-1, c0, d);
        publish(s_dof, pub, d);
        pub = plane_dofs(c0, c1, d);
        publish(s_dof + CS, pub, d);
        pubnext = plane_dofs(c1, c2, dnext);   // node plane zb + 2, published in the first iteration
        __syncthreads();
        for (int slot = 0; slot < 2; ++slot) {
            step_colptr(slot);
            step_store(slot);
            __syncthreads();
            if (warp == 0) fb2_mvec_scan(s_gb + slot * CS, s_len + slot * CS, s_cs + slot * CS, s_adj + slot * 4, lane);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    unsigned mpn[5] = {0u, 0u, 0u, 0u, 0u};   // offset words of the cell of the next sub-step
    {
        const int64_t celln = __shfl_sync(full, c0, (2 * warp) & 3);
        if (celln >= 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) mpn[k] = __ldg(M.mapv + ((size_t)celln * 5 + k) * 32 + lane);
        }
    }
    for (int z = zb; z < ze; ++z) {
        const int rel = z - zb;
        const int pb = rel & 1, pt = pb ^ 1;                                   // window planes of node planes z and z + 1
        const int sb = rel % 3, st = (rel + 1) % 3, s2 = (rel + 2) % 3;        // set-up / coordinate slots of planes z, z + 1, z + 2
        const int64_t r5 = raw_cell(z + 5);
        if (z + 2 <= ze) fetch_plane_xyz(z + 2);                               // slot s2 held node plane z - 1
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* xb = s_xw + (size_t)sb * MV_PN * 4;
        const double* xt = s_xw + (size_t)st * MV_PN * 4;
#pragma unroll 1
        for (int ss = 0; ss < 4; ++ss) {
            const int y = (ss + 2 * warp) & 3;
            const int64_t cell = __shfl_sync(full, c0, y);
            const bool have = cell >= 0;
            // set-up of node plane z + 2, the part of this interval that has to precede the integration
            if (ss == 1) step_colptr(s2);
            if (ss == 3 && warp == 1) fb2_mvec_scan(s_gb + s2 * CS, s_len + s2 * CS, s_cs + s2 * CS, s_adj + s2 * 4, lane);
            // offset words of this sub-step's cell (requested one sub-step ago) and request of the next cell's
            unsigned mp[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) mp[k] = mpn[k];
            {
                const int64_t celln = ss < 3 ? __shfl_sync(full, c0, (ss + 1 + 2 * warp) & 3) : __shfl_sync(full, c1, (2 * warp) & 3);
                if (celln >= 0) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) mpn[k] = __ldg(M.mapv + ((size_t)celln * 5 + k) * 32 + lane);
                }
            }
            double acc[3][3][2];
            double fpart = 0.0;
            bool bad = false;
            if (have) {
                // ---- geometry: X[sx][sy][sz][c] -------------------------------------------------------------------------------------
                const int tn0 = y * 5 + warp;
                double X[2][2][2][3];
#pragma unroll
                for (int sz = 0; sz < 2; ++sz)
#pragma unroll
                    for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                        for (int sx = 0; sx < 2; ++sx) {
                            const double* p = (sz ? xt : xb) + 4 * (tn0 + sy * 5 + sx);
                            const double2 v = *reinterpret_cast<const double2*>(p);
                            X[sx][sy][sz][0] = v.x; X[sx][sy][sz][1] = v.y; X[sx][sy][sz][2] = p[2];
                        }
                // Jt = 2 J at the two points (qx, qy, qz = 0 / 1): the zeta column is common, the xi / eta columns are linear in zeta
                double Ex[2][3], Ey[2][3], Jz[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
#pragma unroll
                    for (int sz = 0; sz < 2; ++sz) {
                        Ex[sz][c] = fma(ny1, X[1][1][sz][c] - X[0][1][sz][c], ny0 * (X[1][0][sz][c] - X[0][0][sz][c]));
                        Ey[sz][c] = fma(nx1, X[1][1][sz][c] - X[1][0][sz][c], nx0 * (X[0][1][sz][c] - X[0][0][sz][c]));
                    }
                    Jz[c] = fma(w11, X[1][1][1][c] - X[1][1][0][c],
                                fma(w01, X[0][1][1][c] - X[0][1][0][c], fma(w10, X[1][0][1][c] - X[1][0][0][c], w00 * (X[0][0][1][c] - X[0][0][0][c]))));
                }
                double g[2][3], gw[2][3];
#pragma unroll
                for (int qz = 0; qz < 2; ++qz) {
                    const double nz0 = qz ? NB : NA, nz1 = qz ? NA : NB;
                    double J[3][3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        J[c][0] = fma(nz1, Ex[1][c], nz0 * Ex[0][c]);
                        J[c][1] = fma(nz1, Ey[1][c], nz0 * Ey[0][c]);
                        J[c][2] = Jz[c];
                    }
                    double Aj[3][3];   // adjugate of Jt (= det * inverse)
                    Aj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
                    Aj[1][0] = -(J[1][0] * J[2][2] - J[1][2] * J[2][0]);
                    Aj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
                    const double det = J[0][0] * Aj[0][0] + J[0][1] * Aj[1][0] + J[0][2] * Aj[2][0];
                    Aj[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]);
                    Aj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
                    Aj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
                    Aj[1][2] = -(J[0][0] * J[1][2] - J[0][2] * J[1][0]);
                    Aj[2][1] = -(J[0][0] * J[2][1] - J[0][1] * J[2][0]);
                    Aj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
                    bad |= !(det > 0.0);
                    double rc;   // 1 / det(Jt): two Newton steps on the hardware approximation
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(det));
                    rc = fma(fma(-det, rc, 1.0), rc, rc);
                    rc = fma(fma(-det, rc, 1.0), rc, rc);
                    const double dO = det * w8;
                    const double nzm = qz ? nzm1 : nzm0;
                    const double d0 = px * nzm * rc, d1 = py * nzm * rc, d2 = pz * rc;
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        g[qz][b] = fma(d0, Aj[0][b], fma(d1, Aj[1][b], d2 * Aj[2][b]));
                        gw[qz][b] = g[qz][b] * dO;
                    }
                    fpart = fma(ax * ay * nzm, dO, fpart);
                }
                // ---- contraction on the FP64 tensor cores: acc[c][d][e] = H_cd[a = mc][b = 2 kr + e] ---------------------------------
                // (the two k-steps of a tile depend on each other through the accumulator: nine independent DMMAs, then nine more)
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        acc[c][d][0] = 0.0; acc[c][d][1] = 0.0;
                        fb2_dmma884(acc[c][d], gw[0][c], g[0][d]);
                    }
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int d = 0; d < 3; ++d) fb2_dmma884(acc[c][d], gw[1][c], g[1][d]);
                if (with_f) {   // fe[(a, c)] = b_c sum_q N_a dOmega: the four lanes of a row hold the four pairs of points
                    fpart += __shfl_xor_sync(full, fpart, 1);
                    fpart += __shfl_xor_sync(full, fpart, 2);
                }
                bad = __any_sync(full, bad);
                if (bad && lane == 0) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
            }
            // set-up of node plane z + 2, the parts that may follow the integration
            if (ss == 2) step_store(s2);
            if (ss == 0) {
                // the flush of node plane z - 1 (which cleared the window plane that now becomes node plane z + 1 and reset the
                // dof table of its set-up slot) is behind every warp; the dofs of node plane z + 2 go into that slot
                if (TMAF) {   // the bulk engine has read the window plane: clear it
                    if (lane < 4) fb2_bulk_wait_read();
                    __syncthreads();
                    double2* zp = reinterpret_cast<double2*>(s_acc + (size_t)pt * cap);
                    for (int i = tid; i < cap / 2; i += 128) zp[i] = make_double2(0.0, 0.0);
                }
                __syncthreads();
                publish(s_dof + s2 * CS, pubnext, dnext);
                pubnext = plane_dofs(c2, c3, dnext);   // node plane z + 3
            }
            // ---- Ke = lambda H + mu H^T + mu tr(H) I into the window ---------------------------------------------------------------
            if (have && !bad) {
                const int tn0 = y * 5 + warp;
                const int pl = kr >> 1;                                    // the lane's column nodes 2 kr, 2 kr + 1 lie in plane z + pl
                const int wbase = (pl ? pt : pb) * cap;
                const uint16_t* cs = s_cs + (pl ? st : sb) * CS;
                int cb[2][3];
#pragma unroll
                for (int bs = 0; bs < 2; ++bs) {
                    const int tnb = tn0 + (kr & 1) * 5 + (bs ^ (kr & 1));
#pragma unroll
                    for (int d = 0; d < 3; ++d) cb[bs][d] = wbase + (int)cs[tnb * 3 + d];
                }
                double v[18], t[18];
                int sl[18];
                bool missing = false;
#pragma unroll
                for (int bs = 0; bs < 2; ++bs) {
                    const double mtr = mu * (acc[0][0][bs] + acc[1][1][bs] + acc[2][2][bs]);
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int tt = (bs * 3 + d) * 3 + c;
                            double val;
                            if constexpr (GENERAL) {
                                val = 0.0;
#pragma unroll
                                for (int q = 0; q < 3; ++q)
#pragma unroll
                                    for (int r = 0; r < 3; ++r) val = fma(CM.c[((c * 3 + q) * 3 + d) * 3 + r], acc[q][r][bs], val);
                            } else {
                                val = fma(lam, acc[c][d][bs], mu * acc[d][c][bs]);
                                if (c == d) val += mtr;
                            }
                            v[tt] = val;
                            const unsigned off = (mp[tt >> 2] >> (8 * (tt & 3))) & 0xFFu;
                            sl[tt] = cb[bs][d] + (int)off;
                            if (CHECK && off == 0xFFu) {   // a non-zero aimed at a missing pattern entry is an error (src/assembler.jl:459-467)
                                if (val != 0.0) missing = true;
                                sl[tt] = -1;
                            }
                        }
                }
#pragma unroll
                for (int tt = 0; tt < 18; ++tt) t[tt] = (!CHECK || sl[tt] >= 0) ? s_acc[sl[tt]] : 0.0;
#pragma unroll
                for (int tt = 0; tt < 18; ++tt)
                    if (!CHECK || sl[tt] >= 0) s_acc[sl[tt]] = t[tt] + v[tt];
                if (CHECK && missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
                if (with_f && kr < 3) s_f[(msz ? pt : pb) * CS + (tn0 + msy * 5 + msx) * 3 + kr] += bforce * fpart;
            }
            if (ss == 3) {
                if (TMAF) fb2_fence_async_smem();   // the read-modify-writes (generic proxy) -> visible to the bulk engine
                asm volatile("cp.async.wait_group 0;" ::: "memory");   // coordinates of node plane z + 2
            }
            __syncthreads();
        }
        if (TMAF)
            fb2_mvec_flush_tma(A, s_acc + (size_t)pb * cap, s_f + pb * CS, s_cs + sb * CS, s_len + sb * CS, s_gb + sb * CS, s_dof + sb * CS, s_adj + sb * 4,
                               tid, z == zb || !M.overwrite, with_f, M.dbg);
        else
            fb2_mvec_flush(A, s_acc + (size_t)pb * cap, s_f + pb * CS, s_cs + sb * CS, s_len + sb * CS, s_gb + sb * CS, s_dof + sb * CS, s_adj + sb * 4,
                           tid, z == zb || !M.overwrite, with_f);
        c0 = c1; c1 = c2; c2 = c3; c3 = own(r4); r4 = r5;
    }
    {   // the top plane of the chunk is shared with the chunk above
        const int rel = ze - zb, pl = rel & 1, sl = rel % 3;
        if (TMAF)
            fb2_mvec_flush_tma(A, s_acc + (size_t)pl * cap, s_f + pl * CS, s_cs + sl * CS, s_len + sl * CS, s_gb + sl * CS, s_dof + sl * CS, s_adj + sl * 4,
                               tid, true, with_f, M.dbg);
        else
            fb2_mvec_flush(A, s_acc + (size_t)pl * cap, s_f + pl * CS, s_cs + sl * CS, s_len + sl * CS, s_gb + sl * CS, s_dof + sl * CS, s_adj + sl * 4, tid,
                           true, with_f);
    }
    if (TMAF && lane < 4) fb2_bulk_wait_read();   // shared memory must outlive the bulk reads
}
