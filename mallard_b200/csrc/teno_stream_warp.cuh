// TENO reconstruction, streaming variant with WARP-PRIVATE RINGS (FAST floating-point mode, FAST_CT = 8) —
// TENOFunctor::operator() (numerics/face_reconstruction.cpp:866-1039).  Same compact tables, same arithmetic and the same
// re-association as teno_stream.cuh (read its header first); what changes is who moves the bytes and who multiplies them:
//
//   * A tile is 8 cells and belongs to ONE WARP.  Every warp streams its own tiles through its own shared-memory ring with
//     1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx → SASS UBLKCP) issued by its lane 0: after the warp has
//     consumed a chunk, lane 0 re-arms the stage's mbarrier and requests the chunk STAGES ahead.  There is no producer
//     warp, no "empty" barrier and no CTA-wide synchronisation at all; a CTA is just four independent warps.  Without
//     the fifth warp two CTAs per SM leave 2 warps per scheduler, i.e. the full 255 registers per thread (the CTA-ring
//     kernel had to live in 168: 2 x 5 warps put three warps on a scheduler, and at 200 registers it silently dropped to
//     one CTA per SM — profiles/r01h).
//   * The four lanes of a cell are (variable pair p, column half h) instead of (variable): a lane multiplies HALF of a
//     row's columns with the right-hand sides of BOTH variables of its pair, then the halves are combined with one
//     shuffle per row and every lane keeps the dof of its own variable 2p + h.  Why: a 128-bit shared-memory load is served
//     per quarter-warp, and with one variable per lane the eight lanes of a quarter read two distinct 16-byte entries
//     (ncu, CTA-ring kernel: shared-memory wavefronts at 76 % of the LSU peak next to 69 % of DRAM peak).  Here h is
//     constant within a quarter-warp, whose eight lanes read four cells' entries (64 contiguous bytes, conflict-free),
//     and every lane issues half as many table loads for the same number of FMAs.
//   * Neighbour states: ubuf[m][cell][4], single-buffered.  The right-hand sides are in registers before the next
//     stencil's states are requested (a __syncwarp() orders the warp's reads before its own cp.async writes).  Writes
//     (lane = 16-byte half p of neighbour 4i + 2h + j) are 128 contiguous bytes per quarter-warp; a lane reads its own
//     variable X = 2p + h and the partner's Y = X ^ 1 of each of its columns, so the row loop needs no lane-dependent
//     selects: it accumulates (x, y), ships y to the partner lane and keeps x.
//   * The stencil loop is rolled (four unrolled copies are ~100 kB of SASS and thrash the instruction cache); the rows
//     of a chunk form one basic block so that loads of one row overlap the FMA chains of another; the ids of a stencil
//     are fetched with volatile loads one stencil before they are used.
#pragma once

namespace stream {

constexpr int WARPS = 4;                       // independent warps per CTA
constexpr int WTHREADS = 32 * WARPS;
static_assert(CT == 8, "warp-private rings: one tile = the 8 cells of one warp");

#ifndef MLB_WARP_STAGES
#define MLB_WARP_STAGES 0
#endif
#ifndef MLB_WARP_MINB
#define MLB_WARP_MINB 2                        // CTAs per SM the register allocation is tuned for
#endif
// Ring depth: the stream runs at the memory system's ceiling with 3 stages already (A/B on B200, 3 / 4 / 5 stages: 2.05 / 2.07 /
// 2.09 ms per launch, and 1.89 / 1.95 / 1.96 ms with the arithmetic compiled out); the shared memory a shallower ring leaves
// unused goes to L1, where the neighbour-state gathers hit.
constexpr int warp_stages(int /*order*/) { return MLB_WARP_STAGES ? MLB_WARP_STAGES : 3; }

template <int ORDER> struct WSmem {
    using C = Cfg<ORDER>;
    static constexpr int STAGES = warp_stages(ORDER);
    static constexpr int UROW = CT * 4;                                       // doubles per neighbour row
    static constexpr size_t RING = (size_t)STAGES * C::CHUNK_BYTES;
    static constexpr size_t UBUF = (size_t)C::MC * UROW * 8;
    static constexpr size_t FXBUF = (size_t)2 * FX_ROWS * CT * 8;             // per tile parity
    static constexpr size_t ZERO = 16;                                         // a 0.0 the "no column" lanes multiply with
    static constexpr size_t PER_WARP = (RING + UBUF + FXBUF + ZERO + 127) / 128 * 128;
    static constexpr size_t TOTAL = PER_WARP * WARPS;
};

template <int ORDER, bool MONO>
__global__ void __launch_bounds__(WTHREADS, MLB_WARP_MINB) teno_stream_warp_kernel(const __grid_constant__ ReconStreamArgs a) {
    using C = Cfg<ORDER>;
    using SM = WSmem<ORDER>;
    constexpr int K = C::K, KR = C::KR, MC = C::MC, NP = C::NP, Q = C::Q, S = FAST_S, NF = FAST_S - 1, STAGES = SM::STAGES;
    constexpr int CPT = S * C::NCH;                            // chunks per tile
    constexpr int NSLOT = NP / 2 + 1;                          // column-pair slots per lane (the last: a pair, the single column, or nothing)
    constexpr int NI = (MC + 3) / 4;                           // neighbour-fetch rounds: lane (half p, parity h) requests m = 4 i + 2 h + j
    constexpr int UROW = SM::UROW;
    constexpr int BASIS = MONO ? MLB_BASIS_MONOMIAL : MLB_BASIS_LEGENDRE;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bars[WARPS][STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char * ring = smem + (size_t)warp * SM::PER_WARP;
    double * ubuf = reinterpret_cast<double *>(ring + SM::RING);
    double * fxbuf = reinterpret_cast<double *>(ring + SM::RING + SM::UBUF);
    double * zero = reinterpret_cast<double *>(ring + SM::RING + SM::UBUF + SM::FXBUF);
    uint64_t * full_bar = full_bars[warp];

    const uint32_t n_tiles = a.n_tiles;
    const uint32_t n_warps = gridDim.x * WARPS;
    const uint32_t gw = a.tile_begin + blockIdx.x * WARPS + warp;   // the warp's first tile: neighbouring warps stream neighbouring tiles
    if (gw >= n_tiles) return;
    const uint32_t n_chunks = ((n_tiles - gw + n_warps - 1) / n_warps) * CPT;   // this warp's chunks

    if (lane == 0) {
        for (int i = 0; i < STAGES; i++) mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        zero[0] = 0.0; zero[1] = 0.0;
    }
    __syncwarp();

    // ---- the warp's own producer (lane 0): chunk g of the warp lives in stage g % STAGES
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    uint32_t p_g = 0, p_st = 0, p_ch = 0, p_tile = gw;         // next chunk to request: index, stage, chunk within tile, tile
    auto issue = [&]() {                                       // lane 0 only
        const unsigned char * src = reinterpret_cast<const unsigned char *>(a.mat) + ((size_t)p_tile * CPT + p_ch) * C::CHUNK_BYTES;
        mbar_expect_tx(&full_bar[p_st], C::CHUNK_BYTES);
        bulk_g2s(ring + (size_t)p_st * C::CHUNK_BYTES, src, C::CHUNK_BYTES, &full_bar[p_st], policy);
        p_g++;
        if (++p_st == STAGES) p_st = 0;
        if (++p_ch == CPT) { p_ch = 0; p_tile += n_warps; }
    };
    if (lane == 0) {
        for (int i = 0; i < STAGES && p_g < n_chunks; i++) issue();
    }

    // ---- lane roles
    const int h = (lane >> 3) & 1;                                    // column half (constant within a quarter-warp)
    const int cl = (lane >> 4) * 4 + ((lane & 7) >> 1);               // cell of the tile
    const int p = lane & 1;                                           // variable pair (2p, 2p + 1)
    const int var = 2 * p + h;                                        // X: the variable this lane owns; Y = 2p + 1 - h goes to the partner lane
    const int sub = p + 2 * h;                                        // 0..3: share of the per-cell prefetches
    const uint32_t Np = a.g.Npad;

    // the lane's last column slot: a column pair, the single column (MC - 1), or nothing.  Its two table entries are read
    // with two 64-bit loads whose addresses point at the pair, at (single, zero word) or at (zero word, zero word), so
    // the row loop has no lane-dependent selects and a neighbouring cell's entry is never multiplied (not even by zero).
    const int pi_last = 2 * (NSLOT - 1) + h;
    const int kind = pi_last < NP ? 0 : (pi_last == NP ? 1 : 2);
    const uint32_t off_reg = (uint32_t)(h * CT * 16 + cl * 16);       // + i * 2 CT 16
    const uint32_t zero_off = (uint32_t)(reinterpret_cast<unsigned char *>(zero) - ring);
    const uint32_t off_lx = kind == 0 ? (uint32_t)(pi_last * CT * 16 + cl * 16) : (kind == 1 ? (uint32_t)(NP * CT * 16 + cl * 8) : 0u);
    const uint32_t off_ly = off_lx + 8;
    const bool lx_row = kind != 2, ly_row = kind == 0;                // does the entry live in the table row (else: the zero word)?
    const int lcol0 = kind == 0 ? 2 * pi_last : (kind == 1 ? MC - 1 : 0);
    const int lcol1 = kind == 0 ? lcol0 + 1 : lcol0;

    // Nothing a lane needs from global memory is waited for: it is requested with cp.async (LDGSTS) one stencil (neighbour
    // states) or one tile (geometry, own state) ahead; the ids of a stencil are loaded one stencil before they are used.
    auto prefetch_tile = [&](uint32_t t, uint32_t parity) {
        const uint32_t cell = t * CT + cl;
        if (cell >= a.g.N_recon) return;
        double * dst = fxbuf + (size_t)parity * FX_ROWS * CT + cl;
#pragma unroll
        for (int i = 0; i < 3; i++) {                                 // value index sub + 4 i of the 12 (slot j, component)
            const int v = sub + 4 * i;
            cp_async8(dst + v * CT, a.g.slot_fx + (size_t)v * Np + cell);
        }
        if (sub == 0) cp_async8(dst + 12 * CT, a.area0 + cell);
        cp_async8(dst + (13 + sub) * CT, a.Uin + 4 * (size_t)cell + sub);
    };
    uint32_t id[2 * NI], id0;                                         // ids this lane fetches for the NEXT stencil; the stencil's first id
    auto load_ids = [&](uint32_t t, uint32_t s) {
        const uint32_t * __restrict__ q = a.ids + ((size_t)t * S + s) * (MC * CT) + cl;
        id0 = ld_id(q);
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int m = 4 * i + 2 * h + j;
                id[2 * i + j] = ld_id(q + (m < MC ? m : MC - 1) * CT);   // the surplus slot repeats the last id and is never requested
            }
    };
    auto request_states = [&]() {
        double * dst = ubuf + cl * 4 + 2 * p;                         // this lane moves half p of each 32-byte state
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int m = 4 * i + 2 * h + j;
                if (m < MC) cp_async16(dst + m * UROW, a.Uin + 4 * (size_t)id[2 * i + j] + 2 * p);
            }
    };

    uint32_t tile = gw;
    bool empty_cur;                                                   // is the stencil whose states are in flight empty?
    load_ids(tile, 0);
    prefetch_tile(tile, 0);
    empty_cur = id0 == tile * CT + cl;                                // an empty stencil lists the cell itself (b = 0)
    request_states();
    cp_async_commit();
    load_ids(tile, 1);

    uint32_t c_st = 0, c_par = 0;                                     // stage and phase parity of the chunk being consumed
    uint32_t ready = 0;
    for (uint32_t it = 0; tile < n_tiles; it++) {
        const uint32_t next = tile + n_warps;
        const bool has_next = next < n_tiles;
        const uint32_t cell = tile * CT + cl;
        const bool live = cell < a.g.N_recon;
        const double * fx = fxbuf + (size_t)(it & 1) * FX_ROWS * CT + cl;
        double uX = 0.0, uY = 0.0;
        double dof[S][KR];
        double w[S];
        // The stencil loop is NOT unrolled (the body is ~10 kB of SASS; four copies of it thrash the instruction cache,
        // ncu r01h: 13 % of the stall samples were "no instruction"); its results are filed into dof[s][] / w[s] by
        // warp-uniform branches at the end of the body.
#pragma unroll 1
        for (uint32_t s = 0; s < (uint32_t)S; s++) {
            // 1. right-hand sides b[m] = U[nbr m] - U[cell] of the lane's columns, for X and Y
            cp_async_wait_all();
            __syncwarp();
            if (s == 0) {
                uX = live ? fx[(13 + var) * CT] : 0.0;
                uY = live ? fx[(13 + (var ^ 1)) * CT] : 0.0;          // 2p + 1 - h = (2p + h) ^ 1
            }
            const bool empty = empty_cur;                             // empty stencil (:896-899) or padding cell
            double bX[NSLOT][2], bY[NSLOT][2];
            {
                const double * ubx = ubuf + cl * 4 + var, * uby = ubuf + cl * 4 + (var ^ 1);
#pragma unroll
                for (int i = 0; i < NSLOT - 1; i++) {
                    bX[i][0] = ubx[(4 * i + 2 * h) * UROW] - uX;     bY[i][0] = uby[(4 * i + 2 * h) * UROW] - uY;
                    bX[i][1] = ubx[(4 * i + 2 * h + 1) * UROW] - uX; bY[i][1] = uby[(4 * i + 2 * h + 1) * UROW] - uY;
                }
                const double x0 = ubx[lcol0 * UROW] - uX, y0 = uby[lcol0 * UROW] - uY;
                const double x1 = ubx[lcol1 * UROW] - uX, y1 = uby[lcol1 * UROW] - uY;
                bX[NSLOT - 1][0] = kind == 2 ? 0.0 : x0; bY[NSLOT - 1][0] = kind == 2 ? 0.0 : y0;
                bX[NSLOT - 1][1] = kind == 0 ? x1 : 0.0; bY[NSLOT - 1][1] = kind == 0 ? y1 : 0.0;
            }
            __syncwarp();                                             // the warp's reads of ubuf precede its next writes
            // 2. requests for what comes next (the next tile's first stencil after the last one of this tile)
            empty_cur = id0 == (s + 1 < S ? cell : next * CT + cl);
            if (s + 1 < S || has_next) request_states();
            if (s + 1 == S && has_next) prefetch_tile(next, (it + 1) & 1);
            cp_async_commit();
            // 3. ids of the stencil after that (consumed a stencil later)
            {
                const bool in_tile = s + 2 < S;
                if (in_tile || has_next) load_ids(in_tile ? tile : next, in_tile ? s + 2 : s + 2 - S);
            }
            // 4. dofs a_k = sum_m A'[k][m] b[m]: each lane sums its half of the columns for X and Y; the partner's half of
            //    X arrives with one shuffle per row
            double d[KR];
#pragma unroll
            for (int ch = 0; ch < C::NCH; ch++) {
                if (!ready) mbar_wait(&full_bar[c_st], c_par);
                const uint32_t n_st = c_st + 1 == STAGES ? 0u : c_st + 1, n_par = c_st + 1 == STAGES ? c_par ^ 1u : c_par;
                ready = mbar_test(&full_bar[n_st], n_par);            // probe the next chunk now, use the answer after this chunk's arithmetic
                const unsigned char * base = ring + (size_t)c_st * C::CHUNK_BYTES;
                // the RC rows of a chunk form one basic block (all dot products first, then the shuffles), so that the
                // scheduler overlaps the shared-memory loads of one row with the FMA chains of another
                double sx[C::RC], sy[C::RC];
#pragma unroll
                for (int r = 0; r < C::RC; r++) {
                    const unsigned char * row = base + (size_t)r * C::ROW_DOUBLES * 8;
                    double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
#ifndef MLB_WARP_NOMATH   /* bench-only probe (A/B build): data movement without the dot products */
#pragma unroll
                    for (int i = 0; i < NSLOT - 1; i++) {
                        const double2 c = *reinterpret_cast<const double2 *>(row + off_reg + i * 2 * CT * 16);
                        x0 = fma(c.x, bX[i][0], x0); x1 = fma(c.y, bX[i][1], x1);
                        y0 = fma(c.x, bY[i][0], y0); y1 = fma(c.y, bY[i][1], y1);
                    }
                    {
                        const double cx = *reinterpret_cast<const double *>((lx_row ? row : ring + zero_off) + (lx_row ? off_lx : 0u));
                        const double cy = *reinterpret_cast<const double *>((ly_row ? row : ring + zero_off) + (ly_row ? off_ly : 0u));
                        x0 = fma(cx, bX[NSLOT - 1][0], x0); x1 = fma(cy, bX[NSLOT - 1][1], x1);
                        y0 = fma(cx, bY[NSLOT - 1][0], y0); y1 = fma(cy, bY[NSLOT - 1][1], y1);
                    }
#endif
                    sx[r] = x0 + x1; sy[r] = y0 + y1;
                }
#pragma unroll
                for (int r = 0; r < C::RC; r++) d[ch * C::RC + r] = sx[r] + __shfl_xor_sync(0xffffffffu, sy[r], 8);
                __syncwarp();                                         // every lane is done with the stage
                if (lane == 0 && p_g < n_chunks) {                    // refill it with the chunk STAGES ahead
#ifdef MLB_WARP_PROXY_FENCE
                    fence_proxy_async();                              // generic-proxy READS followed by an async-proxy write need no
#endif                                                                // proxy fence (the same release/acquire pattern as a TMA pipeline's
                    issue();                                          // consumer_release -> producer_acquire); kept as an A/B switch
                }
                c_st = n_st; c_par = n_par;
            }
            // smoothness indicator a^T OI a (:922-936) with the matrix folded onto its upper triangle: OIs[k][k] = OI[k][k],
            // OIs[k][j>k] = OI[k][j] + OI[j][k]; row/column 0 of OI vanish (derivatives of the constant mode)
            double si = 0.0;
#pragma unroll
            for (int k = 0; k < KR; k++) {
                double t = 0.0;
#pragma unroll
                for (int j = k; j < KR; j++) t = fma(a.OIs[k * KR + j], d[j], t);
                si = fma(d[k], t, si);
            }
            const double x = si + 1.0e-12;                            // 1/(SI+eps)^6 :940-944
            const double x2 = x * x, x3 = x2 * x;
            const double ws = empty ? 0.0 : 1.0 / (x3 * x3);
#pragma unroll
            for (int t = 0; t < S; t++) {
                if (s == (uint32_t)t) {                               // warp-uniform
                    w[t] = ws;
#pragma unroll
                    for (int k = 0; k < KR; k++) dof[t][k] = d[k];
                }
            }
        }

        if (live) {
            const double u_self = uX;
            // non-linear weights :948-981 (reference-faithful unless fixed_weights)
            double sd = 0.0;
#pragma unroll
            for (int s = 1; s < S; s++) sd += w[s];
            // thresholds as products: w/x > c  <=>  w > c x for x > 0, and both are false when x is 0, Inf or NaN
            if (w[0] > 1.0e-7 * (sd + w[0])) {
                w[0] = 1.0;
#pragma unroll
                for (int s = 1; s < S; s++) w[s] = 0.0;
            } else {
#pragma unroll
                for (int s = 1; s < S; s++) {
                    if (w[s] > 1.0e-5 * sd) w[s] = (1.0 / K);
                    else if (a.fixed_weights) w[s] = 0.0;
                }
                sd = 0.0;
#pragma unroll
                for (int s = 1; s < S; s++) sd += w[s];
                const double isd = 1.0 / sd;
#pragma unroll
                for (int s = 1; s < S; s++) w[s] *= isd;
                if (a.fixed_weights) w[0] = 0.0;
            }
            // combined polynomial c_k = sum_s w_s a_sk over the stencils the reference does not skip (:1011)
            double c[KR];
#pragma unroll
            for (int k = 0; k < KR; k++) c[k] = 0.0;
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (w[s] != 0.0) {
#pragma unroll
                    for (int k = 0; k < KR; k++) c[k] = fma(w[s], dof[s][k], c[k]);
                }
            }
            // the tile's geometry was waited for at the top of stencil 0 (and published by the __syncwarp() there)
            const double area0 = fx[12 * CT];
            double cb = 0.0;                                          // sum_k c_k psi_bar_k / area_t[0]  (:1028-1029)
            const double cscale = a.fixed_weights ? -1.0 : 1.0 / area0;
#pragma unroll
            for (int k = 0; k < KR; k++) cb = fma(c[k], a.psi_bar[k + 1] * cscale, cb);
            double * out = a.Fc + (size_t)cell * (NF * Q) * 4 + var;  // Fc[cell][slot * Q + q][var]
#pragma unroll
            for (int j = 0; j < NF; j++) {                            // :985-1034 (triangles: always three faces)
                const double x0 = fx[(4 * j) * CT], y0 = fx[(4 * j + 1) * CT], x1 = fx[(4 * j + 2) * CT], y1 = fx[(4 * j + 3) * CT];
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const double tq = (a.qf_x[q] + 1.0) * 0.5;
                    const double xq = tq * (x1 - x0) + x0, yq = tq * (y1 - y0) + y0;
                    double Px[ORDER + 1], Py[ORDER + 1];
                    basis_values<ORDER>(BASIS, xq, Px);
                    basis_values<ORDER>(BASIS, yq, Py);
                    out[(j * Q + q) * 4] = poly_sum(c, Px, Py, u_self + cb, std::make_integer_sequence<int, KR>{});
                }
            }
        }
        __syncwarp();   // fxbuf[it & 1] is rewritten (by other lanes of this warp) two tiles from now at the earliest
        tile = next;
    }
}

template <int ORDER, bool MONO>
static void launch_stream_w(const ReconStreamArgs & a, cudaStream_t st) {
    const size_t smem = WSmem<ORDER>::TOTAL;
    if (a.n_tiles <= a.tile_begin) return;
    const int ctas = persistent_ctas(reinterpret_cast<const void *>(teno_stream_warp_kernel<ORDER, MONO>), WTHREADS, smem);
    const uint32_t need = (a.n_tiles - a.tile_begin + WARPS - 1) / WARPS;
    const unsigned grid = need < (uint32_t)ctas ? need : (unsigned)ctas;   // persistent: one CTA per resident CTA slot
    teno_stream_warp_kernel<ORDER, MONO><<<grid, WTHREADS, smem, st>>>(a);
}

static void launch_stream(const ReconStreamArgs & a, cudaStream_t st) {
    const bool mono = a.basis == MLB_BASIS_MONOMIAL;
    switch (a.order) {
        case 1: if (mono) launch_stream_w<1, true>(a, st); else launch_stream_w<1, false>(a, st); break;
        case 2: if (mono) launch_stream_w<2, true>(a, st); else launch_stream_w<2, false>(a, st); break;
        case 3: if (mono) launch_stream_w<3, true>(a, st); else launch_stream_w<3, false>(a, st); break;
        case 4: if (mono) launch_stream_w<4, true>(a, st); else launch_stream_w<4, false>(a, st); break;
        default: break;
    }
}

}  // namespace stream
