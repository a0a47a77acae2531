// Marching-tile assembly kernels for structured hexahedral grids (sm_100a).  Included by assemble.cu after
// assemble_kernels.cuh.
//
// What limits the thread-per-cell kernels on B200 (scripts/micro/red_micro.cu, profiles/r02_red_micro.txt): FP64 REDs whose
// 32 lanes land in 32 different sectors run at 120-160 G elements/s chip-wide, coalesced plain stores at 760 G/s, and one
// Q1 hexahedron needs 50 such REDs after the x-face merge -- 8 M cells x 50 = 2.5 ms of RED time against 1.2 ms of FP64
// work.  The way out is to sum the 8 cell contributions of an entry ON CHIP and to write finished matrix columns in bulk:
//
//   * a warp owns a tile of 8 x 4 cells in (x, y) and marches through the layers z = zb .. ze-1 of its chunk, one cell per
//     lane and layer (the same reinit! + element routine as k_cell_scalar: src/FEValues/CellValues.jl:122-140,
//     heat_equation.jl:143-164);
//   * shared memory holds a two-plane window of the matrix: for each of the 9 x 5 tile nodes of the node planes z and
//     z+1 a copy of that node's CSC column, laid out exactly like the global column (the cell-local -> nz offset map of
//     assemble! indexes it, so any dof numbering works, src/assembler.jl:347-457).  Lanes add their Ke entries with plain
//     shared-memory read-modify-writes, ordered in batches such that no two lanes of one instruction hit the same entry
//     (local pairs (i, j) with the same node offset alias across cells, pairs with different offsets never do);
//   * after layer z the columns of node plane z are final inside the tile.  They leave the SM through the TMA engine:
//     the seven tile-interior columns of a tile row are one contiguous piece of nzval (1.5 KB) that has received every
//     contribution it will ever get -- ONE bulk copy shared -> global (cp.async.bulk, no zero fill and no read-for-ownership
//     needed); the columns of the tile faces (shared with the neighbouring tiles) and the first / last plane of a chunk go
//     out as bulk reduce-adds (cp.reduce.async.bulk.add.f64).  15 bulk operations per layer replace 1600 per-lane REDs; the
//     copies of the columns are placed at the parity of their global position so that both ends are 16-byte aligned;
//   * node coordinates and dofs of the next node plane are prefetched (cp.async) while a layer is integrated;
//   * no CTA barrier anywhere: one warp per CTA, so the FP64 phase of one warp overlaps the shared-memory and store
//     phases of the others.
#pragma once

struct MarchArgs {
    int nx, ny;            // cells per grid row / rows per layer (generate_grid order: x fastest, then y, then z)
    int z0, z1;            // layers [z0, z1) of this launch
    int tiles_x, tiles_y;  // tiles of 8 x 4 cells per layer
    int lz;                // layers per chunk (one warp marches through one chunk of one tile)
    int nfull, lt;         // k_march_hex: the first nfull chunks have lz layers, the chunks after them lt <= lz (short chunks at the
                           // end of the launch keep its tail short: CTAs are dispatched in chunk order)
    int cap;               // accumulator doubles per node plane: 45 x (longest matrix column) + 48 parity pads, even
    int overwrite;         // 1: nzval / f were zero-filled for this launch and nobody else adds to tile-interior columns
                           //    => they are written with plain (bulk) stores; 0: everything is added
    const uint8_t* mapb;   // byte-packed offset map (fb2_map_build_bytes)
    const uint32_t* mapv;  // lane-major byte map of k_march_vec (fb2_map_build_vec)
    const int32_t* ctalist; // optional: CTA blockIdx.x works on tile x chunk number ctalist[blockIdx.x] (launch over a subset)
    int dbg;               // measurement only (FB2_MVEC_DBG, results are wrong): 1 every piece as a bulk store, 2 no flush, 3 every piece as a reduce-add
    // cell of box position (x, y, z): x + nx (y + ny z) for grids in generate_grid order (cellmap == nullptr), else
    // cellmap[that] (-1 = no cell there).  Only cells with cell_lo <= id < cell_hi are assembled by this launch (the own cells
    // of a partition, a slab of the streamed host path, ...); the others count as absent.
    const int32_t* cellmap;
    int64_t cell_lo, cell_hi;
};

constexpr int MARCH_PN = 45;   // nodes of a tile plane (9 x 5)
constexpr int MARCH_PS = 48;   // padded

__host__ __device__ inline int fb2_march_cap(int max_col_len) { return (MARCH_PN * max_col_len + MARCH_PS + 1) / 2 * 2; }
// doubles: [2][cap] matrix window | [2][PS] load vector | [32] dummies | [2][PS][4] node coordinates; then colptr[dof] (int64),
// dof (int), column start (uint16) and length (uint8) per window plane and tile node; then the staged offset map (4 x 16 bytes
// per lane).  27.2 KB for Q1 hexahedra: eight single-warp CTAs per SM.
__host__ __device__ inline size_t fb2_march_smem(int cap) {
    return sizeof(double) * (2 * (size_t)cap + 2 * MARCH_PS + 32 + 2 * MARCH_PS * 4) + sizeof(int64_t) * 2 * MARCH_PS +
           sizeof(int) * (2 * MARCH_PS + 8) + sizeof(uint16_t) * 2 * MARCH_PS + sizeof(uint8_t) * 2 * MARCH_PS + sizeof(uint4) * 4 * 32;
}

// local node positions of the 8-node hexahedron (src/Grid/grid_generators.jl:170-178)
__host__ __device__ constexpr int fb2_hx(int j) { return (j == 1 || j == 2 || j == 5 || j == 6) ? 1 : 0; }
__host__ __device__ constexpr int fb2_hy(int j) { return (j == 2 || j == 3 || j == 6 || j == 7) ? 1 : 0; }
__host__ __device__ constexpr int fb2_hz(int j) { return j >= 4 ? 1 : 0; }
// Entry e = j * 8 + i (row node i, column node j) of cell c is the matrix entry (node(c) + p_i, node(c) + p_j): two entries
// of DIFFERENT cells coincide iff they have the same offset p_j - p_i.  Batch = rank of the entry among the entries with
// its offset, so the entries of one batch are pairwise distinct over all cells of the tile.
struct MarchBatches { int b[64]; int nbatch; };
__host__ __device__ constexpr int fb2_march_dclass(int e) {
    return (fb2_hz(e >> 3) - fb2_hz(e & 7) + 1) * 9 + (fb2_hy(e >> 3) - fb2_hy(e & 7) + 1) * 3 + (fb2_hx(e >> 3) - fb2_hx(e & 7) + 1);
}
__host__ __device__ constexpr MarchBatches fb2_march_batches() {
    MarchBatches r{};
    r.nbatch = 0;
    for (int e = 0; e < 64; ++e) {
        int n = 0;
        for (int k = 0; k < e; ++k)
            if (fb2_march_dclass(k) == fb2_march_dclass(e)) ++n;
        r.b[e] = n;
        if (n + 1 > r.nbatch) r.nbatch = n + 1;
    }
    return r;
}

// ---- TMA bulk operations (1-D, shared::cta -> global), tracked by the issuing thread's bulk async-group -------------------
__device__ __forceinline__ void fb2_bulk_store(double* gdst, const double* ssrc, int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fb2_bulk_red_add(double* gdst, const double* ssrc, int bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fb2_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fb2_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fb2_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Set-up of a node plane of the tile, first half: every lane publishes the dofs of the four nodes of its cell that lie in
// the plane and the column extents of tile nodes `lane` and `lane + 32` are requested.
__device__ __forceinline__ void fb2_march_plane_issue(int* s_dof, const int64_t* __restrict__ colptr, int lane, int tn0, bool inside,
                                                      int d0, int d1, int d2, int d3, int64_t& bA, int64_t& bB, int64_t& eA, int64_t& eB) {
    s_dof[lane] = -1;
    if (lane < MARCH_PS - 32) s_dof[lane + 32] = -1;
    __syncwarp();
    if (inside) {   // neighbouring lanes write the same value to shared nodes
        s_dof[tn0] = d0;
        s_dof[tn0 + 1] = d1;
        s_dof[tn0 + 10] = d2;
        s_dof[tn0 + 9] = d3;
    }
    __syncwarp();
    const int dA = s_dof[lane], dB = lane < MARCH_PS - 32 ? s_dof[lane + 32] : -1;
    bA = 0; bB = 0; eA = 0; eB = 0;
    if (dA >= 0) { bA = __ldg(colptr + dA); eA = __ldg(colptr + dA + 1); }
    if (dB >= 0) { bB = __ldg(colptr + dB); eB = __ldg(colptr + dB + 1); }
}

// Second half: position of every column copy inside the plane's accumulator = exclusive scan of the column lengths plus
// parity pads, such that copy and global column start at the same parity (entry index mod 2): with 16-byte aligned bases
// the bulk operations of the flush then see 16-byte aligned addresses on both sides after peeling at most one entry.  Columns
// that are adjacent in nzval stay adjacent in the accumulator (no pad between them).  *rowok: bit b = the seven interior
// columns of tile row b are one contiguous piece of nzval.
__device__ __forceinline__ void fb2_march_plane_finish(uint16_t* s_cs, uint8_t* s_len, int64_t* s_gb, int* s_rowok, int lane, int64_t bA, int64_t bB,
                                                       int64_t eA, int64_t eB) {
    const unsigned full = 0xffffffffu;
    // the four colptr values were requested before the integration; this empty statement keeps every consumer (starting
    // with the subtractions below) behind it, so the warp does not stall on them before it has work to hide the latency
    asm volatile("" : "+l"(bA), "+l"(bB), "+l"(eA), "+l"(eB));
    const int lenA = (int)(eA - bA), lenB = (int)(eB - bB);
    int sA = lenA, sB = lenB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, sA, o), u = __shfl_up_sync(full, sB, o);
        if (lane >= o) { sA += t; sB += u; }
    }
    const int totA = __shfl_sync(full, sA, 31);
    const int csA = sA - lenA, csB = totA + sB - lenB;
    // parity mismatch of the unpadded layout, pad where it changes
    const int tA = (csA + (int)bA) & 1, tB = (csB + (int)bB) & 1;
    const int tA31 = __shfl_sync(full, tA, 31);
    int pvA = __shfl_up_sync(full, tA, 1), pvB = __shfl_up_sync(full, tB, 1);
    if (lane == 0) { pvA = 0; pvB = tA31; }
    int pA = tA ^ pvA, pB = tB ^ pvB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, pA, o), u = __shfl_up_sync(full, pB, o);
        if (lane >= o) { pA += t; pB += u; }
    }
    const int totP = __shfl_sync(full, pA, 31);
    s_cs[lane] = (uint16_t)(csA + pA);
    s_len[lane] = (uint8_t)lenA;
    s_gb[lane] = bA;
    if (lane < MARCH_PS - 32) {
        s_cs[lane + 32] = (uint16_t)(csB + totP + pB);
        s_len[lane + 32] = (uint8_t)lenB;
        s_gb[lane + 32] = bB;
    }
    // contiguity of neighbouring columns: column n + 1 starts where column n ends (bits 0..4 of *rowok: the seven interior
    // columns of tile row b are contiguous; bits 8..12: all nine are)
    int64_t nbA = __shfl_down_sync(full, bA, 1);
    const int64_t nbB = __shfl_down_sync(full, bB, 1), bB0 = __shfl_sync(full, bB, 0);
    if (lane == 31) nbA = bB0;
    const unsigned mA = __ballot_sync(full, lenA > 0 && nbA == bA + lenA), mB = __ballot_sync(full, lenB > 0 && nbB == bB + lenB);
    if (lane == 0) {
        const unsigned long long m = (unsigned long long)mA | ((unsigned long long)mB << 32);
        int ok = 0;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            if (((m >> (9 * b + 1)) & 0x3Full) == 0x3Full) ok |= 1 << b;
            if (((m >> (9 * b)) & 0xFFull) == 0xFFull) ok |= 256 << b;   // all nine columns of the row are one piece of nzval
        }
        *s_rowok = ok;
    }
    __syncwarp();
}

// One piece of a finished plane: columns n0 .. n0 + cnt - 1 (adjacent in nzval and in the accumulator) -> global memory,
// as a bulk copy (STORE) or a bulk reduce-add, after peeling a leading / trailing entry where the piece starts / ends on
// an odd entry.  Called by ONE lane per piece.
template <bool STORE>
__device__ __forceinline__ void fb2_march_emit(double* __restrict__ nzval, const double* acc, const uint16_t* s_cs, const uint8_t* s_len, const int64_t* s_gb,
                                               int n0, int cnt) {
    const int c0 = s_cs[n0];
    int total = (int)s_cs[n0 + cnt - 1] + (int)s_len[n0 + cnt - 1] - c0;
    if (total <= 0) return;
    const int64_t g0 = s_gb[n0];
    double* g = nzval + g0;
    const double* a = acc + c0;
    if (g0 & 1) {
        if (STORE) g[0] = a[0]; else atomicAdd(g, a[0]);
        ++g; ++a; --total;
    }
    const int body = total & ~1;
    if (body) {
        if (STORE) fb2_bulk_store(g, a, body * 8); else fb2_bulk_red_add(g, a, body * 8);
    }
    if (total & 1) {
        if (STORE) g[body] = a[body]; else atomicAdd(g + body, a[body]);
    }
}

// The same for the pieces of all lanes of a warp at once (total = 0: the lane has none).  One call site per kind of bulk
// operation: a bulk operation takes its operands from uniform registers, so ptxas wraps every site in a loop that elects
// one lane at a time -- with the fifteen pieces of a plane spread over the eleven inlined sites of fb2_march_emit's branches
// the flush was 11 % of the kernel's instructions and 20 % of its stall samples (profiles/r02_prof_c2_final.txt).
__device__ __forceinline__ void fb2_march_emit_piece(double* g, const double* a, int total, bool store) {
    const int head = (total > 0 && (reinterpret_cast<uintptr_t>(g) & 8)) ? 1 : 0;   // nzval is 16-byte aligned: odd entry index
    const int rem = total - head;
    const int body = rem & ~1;
    if (head) {
        if (store) g[0] = a[0]; else atomicAdd(g, a[0]);
    }
    if (rem & 1) {
        if (store) g[head + body] = a[head + body]; else atomicAdd(g + head + body, a[head + body]);
    }
    if (body > 0 && store) fb2_bulk_store(g + head, a + head, body * 8);
    if (body > 0 && !store) fb2_bulk_red_add(g + head, a + head, body * 8);
}

// Write a finished node plane out.  Lane s < 15 takes piece s of tile row b = s / 3: part 0 / 2 = the face columns a = 0 / 8
// (reduce-add: shared with the neighbouring tiles), part 1 = the interior columns a = 1..7 (bulk store unless `redall` or the
// row is a tile face, b = 0 / 4); a row that is reduce-added as a whole and contiguous in nzval is ONE piece (part 0).
// Rows whose interior columns are not contiguous in nzval (irregular numbering) are written column by column by lanes
// 16..22.  The accumulator is NOT zeroed here: the bulk operations read it asynchronously; fb2_bulk_wait_read + a zero fill
// precede its next use.
__device__ __forceinline__ void fb2_march_flush(const AsmArgs& A, const double* acc, double* sf, const uint16_t* s_cs, const uint8_t* s_len,
                                                const int64_t* s_gb, const int* s_dof, int rowok, int lane, bool redall, bool with_f) {
    fb2_fence_async_smem();   // the read-modify-writes of this warp (generic proxy) -> visible to the bulk engine (async proxy)
    __syncwarp();
    {
        const int b = lane / 3, part = lane - 3 * b, r0 = 9 * b;
        const bool rowred = redall || b == 0 || b == 4;
        const bool whole = rowred && ((rowok >> (8 + b)) & 1);
        int n0 = r0, cnt = 0;
        if (lane < 15) {
            if (whole) cnt = part == 0 ? 9 : 0;
            else if (part == 0) cnt = 1;
            else if (part == 2) { n0 = r0 + 8; cnt = 1; }
            else if ((rowok >> b) & 1) { n0 = r0 + 1; cnt = 7; }
        }
        int total = 0, c0 = 0;
        int64_t g0 = 0;
        if (cnt > 0) {
            c0 = s_cs[n0];
            total = (int)s_cs[n0 + cnt - 1] + (int)s_len[n0 + cnt - 1] - c0;
            g0 = s_gb[n0];
        }
        fb2_march_emit_piece(A.nzval + g0, acc + c0, total, !rowred && part == 1);
        if (lane < 15) fb2_bulk_commit();
    }
    if (lane >= 16 && lane < 23 && (rowok & 31) != 31) {
        for (int b = 0; b < 5; ++b) {
            if ((rowok >> b) & 1) continue;
            const int n = 9 * b + 1 + (lane - 16);
            if (!redall && b >= 1 && b <= 3) fb2_march_emit<true>(A.nzval, acc, s_cs, s_len, s_gb, n, 1);
            else fb2_march_emit<false>(A.nzval, acc, s_cs, s_len, s_gb, n, 1);
        }
        fb2_bulk_commit();
    }
    if (with_f) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int n = lane + 32 * k;
            const int d = n < MARCH_PN ? s_dof[n] : -1;
            if (d >= 0) {
                const double v = sf[n];
                sf[n] = 0.0;
                const int a = n % 9, b = n / 9;
                if (!redall && a >= 1 && a <= 7 && b >= 1 && b <= 3) A.f[d] = v;
                else if (v != 0.0) atomicAdd(A.f + d, v);
            }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------------------------
// Heat element of the trilinear hexahedron with the 2 x 2 x 2 Gauss rule, written out analytically: the same numbers as
// reinit! + the element loop of heat_equation.jl:143-164 produce from the tables of
// CellValues(QuadratureRule{RefHexahedron}(2), Lagrange{RefHexahedron,1}()) (the host checks that the CellValues really
// holds those tables before this path is taken), with 35 % fewer FP64 instructions than the table-driven loop:
//   * J: the columns of the Jacobian of a trilinear map are bilinear in the other two coordinates, so they come from
//     face / edge interpolations shared between quadrature points (276 instead of 576 operations per cell);
//   * grad N_i = dN_i/dxi . adj(J) / det: the division by det is folded into the weight (dOmega / det^2);
//   * constant functions are in the kernel of the operator, so Ke_ii = -sum_{j != i} Ke_ij: only the 28 strict upper
//     entries are accumulated;
//   * the shape-function values are two constants (no table loads).
// Point q = qx + 2 qy + 4 qz sits at ((2 qx - 1) g, (2 qy - 1) g, (2 qz - 1) g), g = 1/sqrt(3) (src/Quadrature/
// quadrature.jl:96-104: first coordinate fastest, points ascending); the common weight comes from the CellValues.
// ------------------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr double fb2_q1n(int s, int p) {   // 1-D linear shape function of node s at Gauss point p
    return s == p ? 0.5 * (1.0 + 0.5773502691896257) : 0.5 * (1.0 - 0.5773502691896257);
}
__host__ __device__ constexpr int fb2_hexnode(int sx, int sy, int sz) { return sz * 4 + (sy ? (sx ? 2 : 3) : (sx ? 1 : 0)); }

// The loop over (qy, qz) is ROLLED (four iterations of a body that handles the two points qx = 0, 1): the fully unrolled
// body is 35 KB of SASS, and with eight independent warps per SM at different places of it the instruction fetch stalls
// cost more (29 % of all stall samples, profiles/r02_prof_c2_march_c.txt) than the shared sub-expressions save.
// The node coordinates are read from the shared-memory window inside the loop (xs[j] = the lane's node j, 32 bytes apart
// in z-pairs) instead of living in 48 registers next to the 72 of Ke / fe.
__device__ __forceinline__ bool fb2_hex8_heat(const double* const (&xs)[8], const double w8, double (&Ke)[36], double (&fe)[8]) {
    constexpr double NA = fb2_q1n(0, 0), NB = fb2_q1n(1, 0);   // shape function of the near / far node of a Gauss point
#pragma unroll
    for (int e = 0; e < 36; ++e) Ke[e] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) fe[i] = 0.0;
    bool bad = false;
    // All factors 1/2 of the linear shape-function derivatives are dropped: with Jt = 2 J one has adj(Jt) = 4 adj(J) and
    // det(Jt) = 8 det(J), so grad N_i = (sgn_x ny nz, nx sgn_y nz, nx ny sgn_z) . adj(Jt) / det(Jt) and dOmega = det(Jt) w / 8.
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int qy = it & 1, qz = it >> 1;
        const double ny[2] = {qy ? NB : NA, qy ? NA : NB}, nz[2] = {qz ? NB : NA, qz ? NA : NB};
        double w[2][2];       // ny nz
#pragma unroll
        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
            for (int sz = 0; sz < 2; ++sz) w[sy][sz] = ny[sy] * nz[sz];
        // the x-line through the two points: Jt[:][0] = difference of its end points; 2 d x / d eta, 2 d x / d zeta at the ends
        double J[3][3], Dy[2][3], Dz[2][3];
        {
            double x[8][3];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#ifdef FB2_HEX8_XSMEM   // volatile: re-read in every iteration, not hoisted back into 48 registers
                const unsigned sa = (unsigned)__cvta_generic_to_shared(xs[j]);
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[j][0]), "=d"(x[j][1]) : "r"(sa));
                asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(x[j][2]) : "r"(sa));
#else
                const double2 v = *reinterpret_cast<const double2*>(xs[j]);
                x[j][0] = v.x; x[j][1] = v.y; x[j][2] = xs[j][2];
#endif
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int sx = 0; sx < 2; ++sx) {
                    const double x00 = x[fb2_hexnode(sx, 0, 0)][c], x10 = x[fb2_hexnode(sx, 1, 0)][c];
                    const double x01 = x[fb2_hexnode(sx, 0, 1)][c], x11 = x[fb2_hexnode(sx, 1, 1)][c];
                    Dy[sx][c] = fma(nz[1], x11 - x01, nz[0] * (x10 - x00));
                    Dz[sx][c] = fma(ny[1], x11 - x10, ny[0] * (x01 - x00));
                }
                J[c][0] = fma(w[1][1], x[fb2_hexnode(1, 1, 1)][c] - x[fb2_hexnode(0, 1, 1)][c],
                              fma(w[0][1], x[fb2_hexnode(1, 0, 1)][c] - x[fb2_hexnode(0, 0, 1)][c],
                                  fma(w[1][0], x[fb2_hexnode(1, 1, 0)][c] - x[fb2_hexnode(0, 1, 0)][c],
                                      w[0][0] * (x[fb2_hexnode(1, 0, 0)][c] - x[fb2_hexnode(0, 0, 0)][c]))));
            }
        }
#pragma unroll
        for (int qx = 0; qx < 2; ++qx) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                J[c][1] = fma(Dy[1][c], fb2_q1n(1, qx), Dy[0][c] * fb2_q1n(0, qx));
                J[c][2] = fma(Dz[1][c], fb2_q1n(1, qx), Dz[0][c] * fb2_q1n(0, qx));
            }
            // adjugate (= det * inverse) and determinant of Jt
            double Aj[3][3];
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
            const double dO = det * w8;     // w8 = w_q / 8; the eight weights of the 2 x 2 x 2 rule are equal (checked on the host)
            // dOmega / det(Jt)^2 (the gradients below are det(Jt) * grad N): reciprocal by two Newton steps on the hardware
            // approximation (2^-20 -> 2^-40 -> below rounding; det is a cell volume, far from the subnormal range)
            double rc;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(det));
            rc = fma(fma(-det, rc, 1.0), rc, rc);
            rc = fma(fma(-det, rc, 1.0), rc, rc);
            const double sc = w8 * rc;
            //   g_i = sgn_x U[sy][sz] + nx(sx) V[sy][sz],  U = ny nz Aj[0][:],  V = sgn_y nz Aj[1][:] + sgn_z ny Aj[2][:]
            double A1[2][3], A2[2][3];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int b = 0; b < 3; ++b) { A1[s][b] = nz[s] * Aj[1][b]; A2[s][b] = ny[s] * Aj[2][b]; }
            double g[8][3], gs[8][3];
#pragma unroll
            for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                for (int sz = 0; sz < 2; ++sz) {
                    const int i0 = fb2_hexnode(0, sy, sz), i1 = fb2_hexnode(1, sy, sz);
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const double U = w[sy][sz] * Aj[0][b];
                        const double V = (sy ? A1[sz][b] : -A1[sz][b]) + (sz ? A2[sy][b] : -A2[sy][b]);
                        g[i0][b] = fma(fb2_q1n(0, qx), V, -U);
                        g[i1][b] = fma(fb2_q1n(1, qx), V, U);
                        gs[i0][b] = g[i0][b] * sc;
                        gs[i1][b] = g[i1][b] * sc;
                    }
                    const double wd = w[sy][sz] * dO;
                    fe[i0] = fma(fb2_q1n(0, qx), wd, fe[i0]);
                    fe[i1] = fma(fb2_q1n(1, qx), wd, fe[i1]);
                }
#pragma unroll
            for (int j = 1; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < j; ++i) {
                    double s = Ke[j * (j + 1) / 2 + i];
#pragma unroll
                    for (int b = 0; b < 3; ++b) s = fma(g[i][b], gs[j][b], s);
                    Ke[j * (j + 1) / 2 + i] = s;
                }
        }
    }
    // diagonal from the zero row sums
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j != i) s += (i < j ? Ke[j * (j + 1) / 2 + i] : Ke[i * (i + 1) / 2 + j]);
        Ke[i * (i + 1) / 2 + i] = -s;
    }
    return bad;
}

template <int ELEM, bool CHECK, bool ANALYTIC>
__global__ void __launch_bounds__(32, 8) k_march_hex(const AsmArgs A, const MarchArgs M) {
    constexpr int PS = MARCH_PS;
    constexpr MarchBatches MB = fb2_march_batches();
    static_assert(fb2_march_batches().nbatch == 8, "one load-vector entry rides along with every batch");
    extern __shared__ __align__(16) unsigned char smraw[];
    const int cap = M.cap;
    double* s_acc = reinterpret_cast<double*>(smraw);                   // [2][cap] matrix window | [2][PS] load vector | [32] dummies
    const int o_f = 2 * cap, o_dummy = o_f + 2 * PS;
    double* s_xw = s_acc + o_dummy + 32;                                // [2][PS][4] node coordinates of two node planes
    int64_t* s_gb = reinterpret_cast<int64_t*>(s_xw + 2 * PS * 4);      // [2][PS] colptr[dof] of the tile nodes
    int* s_dof = reinterpret_cast<int*>(s_gb + 2 * PS);                 // [2][PS] dof of the tile node, -1 = no such node
    int* s_rowok = s_dof + 2 * PS;                                      // [2] (+6 pad) contiguity bits of the tile rows
    uint16_t* s_cs = reinterpret_cast<uint16_t*>(s_rowok + 8);          // [2][PS] start of the node's column copy
    uint8_t* s_len = reinterpret_cast<uint8_t*>(s_cs + 2 * PS);         // [2][PS] its length
    uint4* s_map = reinterpret_cast<uint4*>(s_len + 2 * PS);            // [4][32] byte-packed offset map of the lane's cell

    const int lane = threadIdx.x;
    const int lx = lane & 7, ly = lane >> 3;
    int bid = M.ctalist ? __ldg(M.ctalist + blockIdx.x) : (int)blockIdx.x;
    const int tx = bid % M.tiles_x;
    bid /= M.tiles_x;
    const int ty = bid % M.tiles_y, ch = bid / M.tiles_y;
    const int zb = ch < M.nfull ? M.z0 + ch * M.lz : M.z0 + M.nfull * M.lz + (ch - M.nfull) * M.lt;
    const int ze = min(M.z1, zb + (ch < M.nfull ? M.lz : M.lt));
    if (zb >= ze) return;
    const int cx = tx * 8 + lx, cy = ty * 4 + ly;
    const bool inside = cx < M.nx && cy < M.ny;     // lanes outside the grid redo a valid cell and add nothing
    const int64_t cxy = (int64_t)min(cx, M.nx - 1) + (int64_t)M.nx * min(cy, M.ny - 1);
    const int64_t lay = (int64_t)M.nx * M.ny;
    const int64_t np = A.ncells_pad;
    // id of the lane's cell in layer z: the raw look-up (-1 = no cell / no such layer) and, separately, the test whether it
    // belongs to this launch -- the test is applied when the id is USED, layers after the look-up was issued
    auto raw_cell = [&](int z) -> int64_t {
        if (!inside || z >= ze) return -1;
        return M.cellmap ? (int64_t)__ldg(M.cellmap + cxy + lay * z) : cxy + lay * z;
    };
    auto own = [&](int64_t c) -> int64_t { return (c >= M.cell_lo && c < M.cell_hi) ? c : -1; };
    // dofs of the lane's four corners of node plane zp + 1: from the cell below it (layer zp) if present, else from the
    // cell above it -- a node plane between an assembled and an absent layer still needs its column copies
    auto plane_dofs = [&](int64_t cbelow, int64_t cabove, int (&d)[4]) -> bool {
        const int64_t c = cbelow >= 0 ? cbelow : cabove;
        const int base = cbelow >= 0 ? 4 : 0;
        if (c < 0) return false;
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = __ldg(A.cell_dofs + (size_t)(base + i) * np + c);
        return true;
    };
    const int tn0 = ly * 9 + lx;                    // tile node of the cell's corner (0, 0)
    const bool with_f = A.f != nullptr && ELEM == FB2_ELEM_HEAT;
    const double kscale = A.p[0], fscale = A.p[1];

    // grid nodes of tile nodes `lane` and `lane + 32` (generate_grid numbers the nodes x fastest, then y, then z:
    // src/Grid/grid_generators.jl:550-559); -1 = outside the grid
    const int64_t nlay = (int64_t)(M.nx + 1) * (M.ny + 1);
    int64_t nodeA, nodeB;
    {
        const int aA = lane % 9, bA_ = lane / 9, nB = lane + 32, aB = nB % 9, bB_ = nB / 9;
        const int gxA = tx * 8 + aA, gyA = ty * 4 + bA_, gxB = tx * 8 + aB, gyB = ty * 4 + bB_;
        nodeA = (gxA <= M.nx && gyA <= M.ny) ? gxA + (int64_t)(M.nx + 1) * gyA : -1;
        nodeB = (nB < MARCH_PN && gxB <= M.nx && gyB <= M.ny) ? gxB + (int64_t)(M.nx + 1) * gyB : -1;
    }
    auto fetch_plane_xyz = [&](int zp, int slot) {   // node plane zp -> window slot (cp.async, 32 bytes per node)
        double* dst = s_xw + (size_t)slot * PS * 4;
        if (nodeA >= 0) {
            const double* src = A.xyz + 4 * (nodeA + nlay * zp);
            fb2_cp_async16(dst + 4 * lane, src);
            fb2_cp_async16(dst + 4 * lane + 2, src + 2);
        }
        if (nodeB >= 0) {
            const double* src = A.xyz + 4 * (nodeB + nlay * zp);
            fb2_cp_async16(dst + 4 * (lane + 32), src);
            fb2_cp_async16(dst + 4 * (lane + 32) + 2, src + 2);
        }
    };

    if (M.cellmap) {   // a tile of a partition-local box may hold halo cells only: nothing to do for this launch
        bool any = false;
        for (int z = zb; z < ze; z += 8) {   // eight independent look-ups per round trip
            int64_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = raw_cell(z + k);
#pragma unroll
            for (int k = 0; k < 8; ++k) any |= own(v[k]) >= 0;
        }
        if (!__any_sync(0xffffffffu, any)) return;
    }
    for (int i = lane; i < o_dummy + 32; i += 32) s_acc[i] = 0.0;
    // cp.async groups, in issue order: [coordinates of the first two node planes], then per layer [offset map of the
    // layer's cell] before and [coordinates of node plane z + 2] after the integration; "wait_group 1" = everything but the
    // newest group has landed
    fetch_plane_xyz(zb, 0);
    fetch_plane_xyz(zb + 1, 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    // cells of the lane in layers z, z + 1, z + 2 (looked up ahead of their use), dofs of node plane z + 1 (prefetched one
    // layer ahead) and whether the lane has any
    int64_t c0 = own(raw_cell(zb)), c1 = own(raw_cell(zb + 1)), c2 = own(raw_cell(zb + 2));
    int64_t r3 = raw_cell(zb + 3);   // raw: its test waits until it becomes c2
    int dnext[4] = {0, 0, 0, 0};
    bool pubnext;
    {   // bottom plane of the first layer (the cells below it belong to another chunk)
        int d[4] = {0, 0, 0, 0};
        const bool pub = plane_dofs(-1, c0, d);
        pubnext = plane_dofs(c0, c1, dnext);
        int64_t bA, bB, eA, eB;
        fb2_march_plane_issue(s_dof, A.colptr, lane, tn0, pub, d[0], d[1], d[2], d[3], bA, bB, eA, eB);
        fb2_march_plane_finish(s_cs, s_len, s_gb, s_rowok, lane, bA, bB, eA, eB);
    }
    const int64_t csafe = M.cell_lo;   // a valid cell for the loads of lanes without one

    for (int z = zb; z < ze; ++z) {
        const int pb = (z - zb) & 1, pt = pb ^ 1;   // window planes holding the node planes z and z + 1
        const bool have = c0 >= 0;
        const int64_t cell = have ? c0 : csafe;
        const int64_t r4 = raw_cell(z + 4);
        // offset map of this layer's cell and set-up of the top plane: requested now, needed after the integration
#pragma unroll
        for (int k = 0; k < 4; ++k) fb2_cp_async16(&s_map[k * 32 + lane], M.mapb + ((size_t)k * np + cell) * 16);
        asm volatile("cp.async.commit_group;" ::: "memory");
        int64_t bA, bB, eA, eB;
        fb2_march_plane_issue(s_dof + pt * PS, A.colptr, lane, tn0, pubnext, dnext[0], dnext[1], dnext[2], dnext[3], bA, bB, eA, eB);
        pubnext = plane_dofs(c1, c2, dnext);   // node plane z + 2, used by the next iteration
        // node coordinates: both planes were requested at least one layer ago
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the offset map just requested
        __syncwarp();
        const double* xs[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) xs[j] = s_xw + ((size_t)(fb2_hz(j) ? pt : pb) * PS + tn0 + fb2_hy(j) * 9 + fb2_hx(j)) * 4;
        double Ke[36], fe[8];
        bool bad;
        if constexpr (ANALYTIC && ELEM == FB2_ELEM_HEAT) {
            bad = fb2_hex8_heat(xs, A.p[2], Ke, fe);
        } else {
            double x[8][3];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double2 v = *reinterpret_cast<const double2*>(xs[j]);
                x[j][0] = v.x; x[j][1] = v.y; x[j][2] = xs[j][2];
            }
            bad = fb2_scalar_element<3, 8, 8, 8, ELEM, true, false>(A, x, Ke, fe);
        }
        if (bad && have) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
        const bool act = have && !bad;
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // the offset map
        __syncwarp();
        if (z + 1 < ze) fetch_plane_xyz(z + 2, pb);   // the slot of node plane z is free now; lands during the flush
        asm volatile("cp.async.commit_group;" ::: "memory");
        fb2_march_plane_finish(s_cs + pt * PS, s_len + pt * PS, s_gb + pt * PS, s_rowok + pt, lane, bA, bB, eA, eB);
        // the window plane that now becomes the top plane was flushed one layer ago: wait until the bulk engine has read
        // it, then clear it
        fb2_bulk_wait_read();
        __syncwarp();
        {
            double2* zp = reinterpret_cast<double2*>(s_acc + (size_t)pt * cap);
            for (int i = lane; i < cap / 2; i += 32) zp[i] = make_double2(0.0, 0.0);
        }
        uint4 mp[4];   // mp[k] = offsets of entries 16 k .. 16 k + 15 (columns 2 k and 2 k + 1), one byte each
#pragma unroll
        for (int k = 0; k < 4; ++k) mp[k] = act ? s_map[k * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
        int cb[8];     // start of the copy of column j inside s_acc; inactive lanes add into their private dummy
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = fb2_hz(j) ? pt : pb;
            cb[j] = act ? p * cap + s_cs[p * PS + tn0 + fb2_hy(j) * 9 + fb2_hx(j)] : o_dummy + lane;
        }
#pragma unroll
        for (int e = 0; e < 36; ++e) Ke[e] *= kscale;
        bool missing = false;
#pragma unroll
        for (int b = 0; b < MB.nbatch; ++b) {
            double t[64];
            double tf = 0.0;
            int fidx = 0;
#pragma unroll
            for (int e = 0; e < 64; ++e) {
                if (MB.b[e] != b) continue;
                const int j = e >> 3;
                const unsigned w32 = ((e >> 2) & 3) == 0 ? mp[e >> 4].x : (((e >> 2) & 3) == 1 ? mp[e >> 4].y : (((e >> 2) & 3) == 2 ? mp[e >> 4].z : mp[e >> 4].w));
                const unsigned off = (w32 >> (8 * (e & 3))) & 0xFFu;
                const int sl = (CHECK && off == 0xFFu) ? o_dummy + lane : cb[j] + (int)off;
                t[e] = s_acc[sl];
            }
            if (with_f) {   // load-vector entry of local node b rides along (entries of one local node never alias)
                const int p = fb2_hz(b) ? pt : pb;
                fidx = act ? o_f + p * PS + tn0 + fb2_hy(b) * 9 + fb2_hx(b) : o_dummy + lane;
                tf = s_acc[fidx];
            }
#pragma unroll
            for (int e = 0; e < 64; ++e) {
                if (MB.b[e] != b) continue;
                const int i = e & 7, j = e >> 3;
                const unsigned w32 = ((e >> 2) & 3) == 0 ? mp[e >> 4].x : (((e >> 2) & 3) == 1 ? mp[e >> 4].y : (((e >> 2) & 3) == 2 ? mp[e >> 4].z : mp[e >> 4].w));
                const unsigned off = (w32 >> (8 * (e & 3))) & 0xFFu;
                const double v = i <= j ? Ke[j * (j + 1) / 2 + i] : Ke[i * (i + 1) / 2 + j];
                if (CHECK && off == 0xFFu) {   // a non-zero aimed at a missing pattern entry is an error (src/assembler.jl:459-467)
                    if (v != 0.0 && act) missing = true;
                } else {
                    s_acc[cb[j] + (int)off] = t[e] + v;
                }
            }
            if (with_f) s_acc[fidx] = tf + fscale * fe[b];
            __syncwarp();
        }
        if (CHECK && missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
        fb2_march_flush(A, s_acc + (size_t)pb * cap, s_acc + o_f + pb * PS, s_cs + pb * PS, s_len + pb * PS, s_gb + pb * PS, s_dof + pb * PS,
                        s_rowok[pb], lane, z == zb || !M.overwrite, with_f);
        c0 = c1; c1 = c2; c2 = own(r3); r3 = r4;
    }
    const int pl = (ze - zb) & 1;   // the top plane of the chunk is shared with the chunk above
    fb2_march_flush(A, s_acc + (size_t)pl * cap, s_acc + o_f + pl * PS, s_cs + pl * PS, s_len + pl * PS, s_gb + pl * PS, s_dof + pl * PS,
                    s_rowok[pl], lane, true, with_f);
    fb2_bulk_wait_read();   // shared memory must outlive the bulk reads
}

// ---- selective zero fill -----------------------------------------------------------------------------------------------------
// With `overwrite` the marching kernel writes the columns of tile-interior nodes with plain stores, so start_assemble's
// zero fill (src/assembler.jl:287-291) is only needed for the columns that receive reduce-adds: nodes on tile faces
// (x % 8 == 0 or y % 4 == 0) and the node planes where two chunks meet -- 38 % of nzval on C2 instead of all of it.
// k_march_mark flags those columns (thread per grid node, same predicate as fb2_march_flush); the flagged dofs are
// compacted into a sorted list once per assembler; k_zero_columns (warp per listed column) runs in front of every launch.
__global__ void k_march_mark(const int32_t* __restrict__ cell_dofs, int64_t np, int nx, int ny, int nz, int lz, int nfull, int lt, int tx, int ty,
                             int vdim, uint8_t* __restrict__ flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nn = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
    if (t >= nn) return;
    const int x = (int)(t % (nx + 1)), y = (int)((t / (nx + 1)) % (ny + 1)), p = (int)(t / ((int64_t)(nx + 1) * (ny + 1)));
    // tiles of tx x ty cells; chunks of lz layers, behind the first nfull of them chunks of lt layers
    const int pt = p - nfull * lz;
    const bool need = (x % tx == 0) || (y % ty == 0) || (pt <= 0 ? p % lz == 0 : pt % lt == 0) || p == nz;
    if (!need) return;
    // the node is corner (sx, sy, sz) of the cell below / left of it (clamped to the grid)
    const int cx = min(x, nx - 1), cy = min(y, ny - 1), cz = min(p, nz - 1);
    const int ln = fb2_hexnode(x - cx, y - cy, p - cz);
    const int64_t cell = cx + (int64_t)nx * (cy + (int64_t)ny * cz);
    for (int c = 0; c < vdim; ++c) flag[cell_dofs[(size_t)(ln * vdim + c) * np + cell]] = 1;
}

// A warp takes 32 listed columns: one lane per column fetches its extent (two dependent loads for 32 columns instead of two
// per column); listed columns that follow each other in nzval (the node rows of tile faces in y, the node planes between
// chunks) are merged into runs, and the whole warp zeroes one run after the other with 16-byte stores.
__global__ void k_zero_columns(const int32_t* __restrict__ cols, int64_t n, const int64_t* __restrict__ colptr, double* __restrict__ nzval, int64_t nnz) {
    const unsigned full = 0xffffffffu;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t i = w * 32 + lane;
    if (w * 32 >= n) return;
    int64_t b = 0, e = 0;
    if (i < n) {
        const int c = __ldg(cols + i);
        b = __ldg(colptr + c);
        e = __ldg(colptr + c + 1);
    }
    const int64_t pe = __shfl_up_sync(full, e, 1);
    unsigned heads = __ballot_sync(full, i < n && (lane == 0 || b != pe));
    while (heads) {
        const int s = __ffs(heads) - 1;
        heads &= heads - 1;
        const int t = (heads ? __ffs(heads) - 1 : (int)min((int64_t)32, n - w * 32)) - 1;   // last column of the run
        int64_t rb = __shfl_sync(full, b, s);
        int64_t re = __shfl_sync(full, e, t);
        if (rb >= re) continue;
        // Whole 32-byte sectors (nzval is 16-byte aligned and comes from cudaMalloc / a torch tensor: 32-byte aligned in
        // practice, else the 16-byte stores below still work): a partly written sector costs the L2 a read from DRAM before
        // it can be written back.  The few entries of neighbouring columns this touches are either listed themselves or
        // belong to columns the marching kernel writes completely afterwards (that is what makes them unlisted).
        rb &= ~(int64_t)3;
        re = min((re + 3) & ~(int64_t)3, nnz);
        const int64_t body = (re - rb) & ~(int64_t)1;
        for (int64_t p = rb + 2 * lane; p < rb + body; p += 64) *reinterpret_cast<double2*>(nzval + p) = make_double2(0.0, 0.0);
        if (((re - rb) & 1) && lane == 0) nzval[re - 1] = 0.0;
    }
}
