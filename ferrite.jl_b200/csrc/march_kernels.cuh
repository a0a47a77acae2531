// Marching-tile assembly kernels for structured hexahedral grids (sm_100a).  Included by assemble.cu after
// assemble_kernels.cuh.
//
// What limits the thread-per-cell kernels on B200 (scripts/micro/red_micro.cu, profiles/r02_red_micro.txt): FP64 REDs whose
// 32 lanes land in 32 different sectors run at 120-160 G elements/s chip-wide, coalesced plain stores at 760 G/s, and one
// Q1 hexahedron needs 50 such REDs after the x-face merge -- 8 M cells x 50 = 2.5 ms of RED time against 1.2 ms of FP64
// work.  The way out is to sum the 8 cell contributions of an entry ON CHIP and to write finished matrix columns with
// coalesced stores:
//
//   * a warp owns a tile of 8 x 4 cells in (x, y) and marches through the layers z = zb .. ze-1 of its chunk, one cell per
//     lane and layer (the same reinit! + element routine as k_cell_scalar: src/FEValues/CellValues.jl:122-140,
//     heat_equation.jl:143-164);
//   * shared memory holds a two-plane window of the matrix: for each of the 9 x 5 tile nodes of the node planes z and
//     z+1 a copy of that node's CSC column, laid out exactly like the global column (the cell-local -> nz offset map of
//     assemble! indexes it, so any dof numbering works, src/assembler.jl:347-457).  Lanes add their Ke entries with plain
//     shared-memory read-modify-writes, ordered in batches such that no two lanes of one instruction hit the same entry
//     (local pairs (i, j) with the same node offset alias across cells, pairs with different offsets never do);
//   * after layer z the columns of node plane z are final inside the tile: columns of tile-interior nodes have received
//     every contribution they will ever get and are written once with coalesced plain stores (216 contiguous bytes per
//     column; no zero fill and no read-for-ownership needed for them); columns on the tile faces, and the first and last
//     plane of a chunk, are shared with neighbouring warps and go out as REDs (12.75 per cell instead of 50, and those
//     are coalesced along the column as well);
//   * no CTA barrier anywhere: one warp per CTA, so the FP64 phase of one warp overlaps the shared-memory and store
//     phases of the others.
#pragma once

struct MarchArgs {
    int nx, ny;            // cells per grid row / rows per layer (generate_grid order: x fastest, then y, then z)
    int z0, z1;            // layers [z0, z1) of this launch
    int tiles_x, tiles_y;  // tiles of 8 x 4 cells per layer
    int lz;                // layers per chunk (one warp marches through one chunk of one tile)
    int cap;               // accumulator doubles per node plane: 45 x (longest matrix column), even
    int overwrite;         // 1: nzval / f were zero-filled for this launch and nobody else adds to tile-interior columns
                           //    => they are written with plain stores; 0: everything is added with REDs
};

constexpr int MARCH_PN = 45;   // nodes of a tile plane (9 x 5)
constexpr int MARCH_PS = 48;   // padded

__host__ __device__ inline size_t fb2_march_smem(int cap) {
    return sizeof(double) * (2 * (size_t)cap + 2 * MARCH_PS + 32) + sizeof(int64_t) * 2 * MARCH_PS + sizeof(int) * 4 * MARCH_PS +
           sizeof(uint4) * 8 * 32;
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

// Set-up of a node plane of the tile, first half: every lane publishes the dofs of the four nodes of its cell that lie in
// the plane and the column extents of tile nodes `lane` and `lane + 32` are requested.
__device__ __forceinline__ void fb2_march_plane_issue(int* s_dof, const int64_t* __restrict__ colptr, int lane, int tn0, bool inside,
                                                      int d0, int d1, int d2, int d3, int64_t& bA, int64_t& bB, int& lenA, int& lenB) {
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
    bA = 0; bB = 0; lenA = 0; lenB = 0;
    if (dA >= 0) { bA = __ldg(colptr + dA); lenA = (int)(__ldg(colptr + dA + 1) - bA); }
    if (dB >= 0) { bB = __ldg(colptr + dB); lenB = (int)(__ldg(colptr + dB + 1) - bB); }
}

// second half: exclusive scan of the column lengths = start of every column copy inside the plane's accumulator
__device__ __forceinline__ void fb2_march_plane_finish(int* s_cs, int64_t* s_gb, int lane, int64_t bA, int64_t bB, int lenA, int lenB) {
    const unsigned full = 0xffffffffu;
    int sA = lenA, sB = lenB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(full, sA, o), u = __shfl_up_sync(full, sB, o);
        if (lane >= o) { sA += t; sB += u; }
    }
    const int totA = __shfl_sync(full, sA, 31);
    s_cs[lane] = sA - lenA;
    s_gb[lane] = bA;
    if (lane < MARCH_PS - 32) {   // entries 45..47 have length 0, so s_cs[45] is the total
        s_cs[lane + 32] = totA + sB - lenB;
        s_gb[lane + 32] = bB;
    }
    __syncwarp();
}

// FP64 RED that skips exact zeros without a branch (entries of a face column this tile never touched stay untouched)
__device__ __forceinline__ void fb2_red_nz(double* p, double v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.neu.f64 p, %1, 0d0000000000000000;\n\t@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v) : "memory");
}

// entries [c0, c1) of a plane accumulator -> nzval[g0 + (e - c0)], plain stores or REDs; the accumulator is zeroed for reuse
template <bool STORE>
__device__ __forceinline__ void fb2_march_flush_run(double* __restrict__ acc, double* __restrict__ nzval, int c0, int c1, int64_t g0, int lane) {
    double* g = nzval + (g0 - c0);
    for (int e = c0 + lane; e < c1; e += 32) {
        const double v = acc[e];
        acc[e] = 0.0;
        if (STORE) g[e] = v;
        else fb2_red_nz(g + e, v);
    }
}

// Write a finished node plane out (and zero its accumulator for reuse).  Tile-interior columns (a = 1..7, b = 1..3) have
// every contribution they will ever get: plain stores unless `redall`; the columns of the tile faces are shared with
// the neighbouring tiles: REDs.  The columns of the seven interior nodes of a tile row are one contiguous run of nzval
// whenever their dofs are consecutive (the reference's numbering away from the grid boundary): six fully coalesced
// requests instead of seven column-sized ones; other numberings take the column-by-column path.
__device__ __forceinline__ void fb2_march_flush(const AsmArgs& A, double* acc, double* sf, const int* s_cs, const int64_t* s_gb,
                                                const int* s_dof, int lane, bool redall, bool with_f) {
    const unsigned full = 0xffffffffu;
#pragma unroll 1
    for (int b = 0; b < 5; ++b) {
        const int r0 = b * 9;
        const bool st = !redall && b >= 1 && b <= 3;
        bool okc = true;
        if (lane < 6) {
            const int n = r0 + 1 + lane;
            okc = s_gb[n + 1] == s_gb[n] + (s_cs[n + 1] - s_cs[n]);
        }
        if (__all_sync(full, okc)) {
            fb2_march_flush_run<false>(acc, A.nzval, s_cs[r0], s_cs[r0 + 1], s_gb[r0], lane);
            if (st) fb2_march_flush_run<true>(acc, A.nzval, s_cs[r0 + 1], s_cs[r0 + 8], s_gb[r0 + 1], lane);
            else fb2_march_flush_run<false>(acc, A.nzval, s_cs[r0 + 1], s_cs[r0 + 8], s_gb[r0 + 1], lane);
            fb2_march_flush_run<false>(acc, A.nzval, s_cs[r0 + 8], s_cs[r0 + 9], s_gb[r0 + 8], lane);
        } else {
#pragma unroll 1
            for (int a = 0; a < 9; ++a) {
                if (st && a >= 1 && a <= 7) fb2_march_flush_run<true>(acc, A.nzval, s_cs[r0 + a], s_cs[r0 + a + 1], s_gb[r0 + a], lane);
                else fb2_march_flush_run<false>(acc, A.nzval, s_cs[r0 + a], s_cs[r0 + a + 1], s_gb[r0 + a], lane);
            }
        }
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
                else fb2_red_nz(A.f + d, v);
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
//   * all shape-function values are compile-time constants (no table loads).
// Point q = qx + 2 qy + 4 qz sits at ((2 qx - 1) g, (2 qy - 1) g, (2 qz - 1) g), g = 1/sqrt(3) (src/Quadrature/
// quadrature.jl:96-104: first coordinate fastest, points ascending); w[q] comes from the CellValues.
// ------------------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr double fb2_q1n(int s, int p) {   // 1-D linear shape function of node s at Gauss point p
    return s == p ? 0.5 * (1.0 + 0.5773502691896257) : 0.5 * (1.0 - 0.5773502691896257);
}
__host__ __device__ constexpr int fb2_hexnode(int sx, int sy, int sz) { return sz * 4 + (sy ? (sx ? 2 : 3) : (sx ? 1 : 0)); }

__device__ __forceinline__ bool fb2_hex8_heat(const double (&x)[8][3], const double* __restrict__ tw, double (&Ke)[36], double (&fe)[8]) {
#pragma unroll
    for (int e = 0; e < 36; ++e) Ke[e] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) fe[i] = 0.0;
    bool bad = false;
    double dz[2][2][3];   // (x(sx,sy,1) - x(sx,sy,0)) / 2
#pragma unroll
    for (int sx = 0; sx < 2; ++sx)
#pragma unroll
        for (int sy = 0; sy < 2; ++sy)
#pragma unroll
            for (int c = 0; c < 3; ++c) dz[sx][sy][c] = 0.5 * (x[fb2_hexnode(sx, sy, 1)][c] - x[fb2_hexnode(sx, sy, 0)][c]);
#pragma unroll
    for (int qz = 0; qz < 2; ++qz) {
        double xz[2][2][3];   // position interpolated along z
#pragma unroll
        for (int sx = 0; sx < 2; ++sx)
#pragma unroll
            for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    xz[sx][sy][c] = fma(x[fb2_hexnode(sx, sy, 1)][c], fb2_q1n(1, qz), x[fb2_hexnode(sx, sy, 0)][c] * fb2_q1n(0, qz));
#pragma unroll
        for (int qy = 0; qy < 2; ++qy) {
            double J[3][3];       // J[a][b] = d x_a / d xi_b
            double dy[2][3], dzy[2][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double p0 = fma(xz[0][1][c], fb2_q1n(1, qy), xz[0][0][c] * fb2_q1n(0, qy));
                const double p1 = fma(xz[1][1][c], fb2_q1n(1, qy), xz[1][0][c] * fb2_q1n(0, qy));
                J[c][0] = 0.5 * (p1 - p0);
#pragma unroll
                for (int sx = 0; sx < 2; ++sx) {
                    dy[sx][c] = xz[sx][1][c] - xz[sx][0][c];
                    dzy[sx][c] = fma(dz[sx][1][c], fb2_q1n(1, qy), dz[sx][0][c] * fb2_q1n(0, qy));
                }
            }
#pragma unroll
            for (int qx = 0; qx < 2; ++qx) {
                const int q = qx + 2 * qy + 4 * qz;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    J[c][1] = fma(dy[1][c], 0.5 * fb2_q1n(1, qx), dy[0][c] * (0.5 * fb2_q1n(0, qx)));
                    J[c][2] = fma(dzy[1][c], fb2_q1n(1, qx), dzy[0][c] * fb2_q1n(0, qx));
                }
                // adjugate (= det * inverse) and determinant
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
                const double w = tw[q];
                const double dO = det * w;
                const double sc = w / det;     // dOmega / det^2: the gradients below are det * grad N
                double g[8][3], gs[8][3];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int sx = fb2_hx(i), sy = fb2_hy(i), sz = fb2_hz(i);
                    const double d0 = (sx ? 0.5 : -0.5) * fb2_q1n(sy, qy) * fb2_q1n(sz, qz);
                    const double d1 = fb2_q1n(sx, qx) * (sy ? 0.5 : -0.5) * fb2_q1n(sz, qz);
                    const double d2 = fb2_q1n(sx, qx) * fb2_q1n(sy, qy) * (sz ? 0.5 : -0.5);
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        g[i][b] = fma(d2, Aj[2][b], fma(d1, Aj[1][b], d0 * Aj[0][b]));
                        gs[i][b] = g[i][b] * sc;
                    }
                    fe[i] = fma(fb2_q1n(sx, qx) * fb2_q1n(sy, qy) * fb2_q1n(sz, qz), dO, fe[i]);
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
    extern __shared__ __align__(16) unsigned char smraw[];
    const int cap = M.cap;
    double* s_acc = reinterpret_cast<double*>(smraw);                   // [2][cap] matrix window | [2][PS] load vector | [32] dummies
    const int o_f = 2 * cap, o_dummy = o_f + 2 * PS;
    int64_t* s_gb = reinterpret_cast<int64_t*>(s_acc + o_dummy + 32);   // [2][PS] colptr[dof] of the tile nodes
    int* s_cs = reinterpret_cast<int*>(s_gb + 2 * PS);                  // [2][PS] start of the node's column copy
    int* s_dof = s_cs + 2 * PS;                                         // [2][PS] dof of the tile node, -1 = no such node
    uint4* s_map = reinterpret_cast<uint4*>(s_dof + 2 * PS);            // [8][32] packed offset map of the lane's cell

    const int lane = threadIdx.x;
    const int lx = lane & 7, ly = lane >> 3;
    int bid = blockIdx.x;
    const int tx = bid % M.tiles_x;
    bid /= M.tiles_x;
    const int ty = bid % M.tiles_y, ch = bid / M.tiles_y;
    const int zb = M.z0 + ch * M.lz, ze = min(M.z1, zb + M.lz);
    if (zb >= ze) return;
    const int cx = tx * 8 + lx, cy = ty * 4 + ly;
    const bool inside = cx < M.nx && cy < M.ny;     // lanes outside the grid redo a valid cell and add nothing
    const int64_t cxy = (int64_t)min(cx, M.nx - 1) + (int64_t)M.nx * min(cy, M.ny - 1);
    const int64_t lay = (int64_t)M.nx * M.ny;
    const int64_t np = A.ncells_pad;
    const int tn0 = ly * 9 + lx;                    // tile node of the cell's corner (0, 0)
    const bool with_f = A.f != nullptr && ELEM == FB2_ELEM_HEAT;
    const double kscale = A.p[0], fscale = A.p[1];

    for (int i = lane; i < o_dummy + 32; i += 32) s_acc[i] = 0.0;
    {   // bottom plane of the first layer
        const int64_t cell = cxy + lay * zb;
        int d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = __ldg(A.cell_dofs + (size_t)i * np + cell);
        int64_t bA, bB;
        int lenA, lenB;
        fb2_march_plane_issue(s_dof, A.colptr, lane, tn0, inside, d[0], d[1], d[2], d[3], bA, bB, lenA, lenB);
        fb2_march_plane_finish(s_cs, s_gb, lane, bA, bB, lenA, lenB);
    }

    for (int z = zb; z < ze; ++z) {
        const int pb = (z - zb) & 1, pt = pb ^ 1;   // window planes holding the node planes z and z + 1
        const int64_t cell = cxy + lay * z;
        double x[8][3];
        {
            int node[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) node[j] = __ldg(A.conn + (size_t)j * np + cell);
#pragma unroll
            for (int j = 0; j < 8; ++j) fb2_load_x<3>(A.xyz, node[j], x[j]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) fb2_cp_async16(&s_map[k * 32 + lane], A.map8 + (size_t)k * np + cell);
        int64_t bA, bB;
        int lenA, lenB;
        {
            int d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = __ldg(A.cell_dofs + (size_t)(4 + i) * np + cell);
            fb2_march_plane_issue(s_dof + pt * PS, A.colptr, lane, tn0, inside, d[0], d[1], d[2], d[3], bA, bB, lenA, lenB);
        }
        double Ke[36], fe[8];
        bool bad;
        if constexpr (ANALYTIC && ELEM == FB2_ELEM_HEAT) bad = fb2_hex8_heat(x, c_tab + A.o_w, Ke, fe);
        else bad = fb2_scalar_element<3, 8, 8, 8, ELEM, true, false>(A, x, Ke, fe);
        if (bad && inside) fb2_flag_error(A.errflag, FB2_ERR_DETJ_NOT_POSITIVE, cell);
        const bool act = inside && !bad;
        fb2_march_plane_finish(s_cs + pt * PS, s_gb + pt * PS, lane, bA, bB, lenA, lenB);
        asm volatile("cp.async.wait_all;" ::: "memory");
        uint4 mp[8];   // mp[j] = offsets of rows 0..7 inside column j
#pragma unroll
        for (int k = 0; k < 8; ++k) mp[k] = act ? s_map[k * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
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
                const int i = e & 7, j = e >> 3;
                const unsigned w32 = (i >> 1) == 0 ? mp[j].x : ((i >> 1) == 1 ? mp[j].y : ((i >> 1) == 2 ? mp[j].z : mp[j].w));
                const unsigned off = (i & 1) ? (w32 >> 16) : (w32 & 0xFFFFu);
                const int sl = (CHECK && off == 0xFFFFu) ? o_dummy + lane : cb[j] + (int)off;
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
                const unsigned w32 = (i >> 1) == 0 ? mp[j].x : ((i >> 1) == 1 ? mp[j].y : ((i >> 1) == 2 ? mp[j].z : mp[j].w));
                const unsigned off = (i & 1) ? (w32 >> 16) : (w32 & 0xFFFFu);
                const double v = i <= j ? Ke[j * (j + 1) / 2 + i] : Ke[i * (i + 1) / 2 + j];
                if (CHECK && off == 0xFFFFu) {   // a non-zero aimed at a missing pattern entry is an error (src/assembler.jl:459-467)
                    if (v != 0.0 && act) missing = true;
                    s_acc[o_dummy + lane] = 0.0;
                } else {
                    s_acc[cb[j] + (int)off] = t[e] + v;
                }
            }
            if (with_f) s_acc[fidx] = tf + fscale * fe[b];
            __syncwarp();
        }
        if (CHECK && missing) fb2_flag_error(A.errflag, FB2_ERR_MISSING_PATTERN_ENTRY, cell);
        fb2_march_flush(A, s_acc + pb * cap, s_acc + o_f + pb * PS, s_cs + pb * PS, s_gb + pb * PS, s_dof + pb * PS, lane,
                        z == zb || !M.overwrite, with_f);
    }
    const int pl = (ze - zb) & 1;   // the top plane of the chunk is shared with the chunk above
    fb2_march_flush(A, s_acc + pl * cap, s_acc + o_f + pl * PS, s_cs + pl * PS, s_gb + pl * PS, s_dof + pl * PS, lane, true, with_f);
}
