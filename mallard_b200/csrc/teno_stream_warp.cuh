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
//     (lane = 16-byte half p of neighbour 4i + 2h + j) and reads ((column, variable pair) as one 128-bit load) are both
//     128 contiguous bytes per quarter-warp.
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
constexpr int warp_stages(int /*order*/) { return MLB_WARP_STAGES ? MLB_WARP_STAGES : 5; }

template <int ORDER> struct WSmem {
    using C = Cfg<ORDER>;
    static constexpr int STAGES = warp_stages(ORDER);
    static constexpr int UROW = CT * 4;                                       // doubles per neighbour row
    static constexpr size_t RING = (size_t)STAGES * C::CHUNK_BYTES;
    static constexpr size_t UBUF = (size_t)C::MC * UROW * 8;
    static constexpr size_t FXBUF = (size_t)2 * FX_ROWS * CT * 8;             // per tile parity
    static constexpr size_t PER_WARP = (RING + UBUF + FXBUF + 127) / 128 * 128;
    static constexpr size_t TOTAL = PER_WARP * WARPS;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
    uint64_t * full_bar = full_bars[warp];

    const uint32_t n_tiles = a.n_tiles;
    const uint32_t n_warps = gridDim.x * WARPS;
    const uint32_t gw = blockIdx.x * WARPS + warp;             // neighbouring warps stream neighbouring tiles
    if (gw >= n_tiles) return;
    const uint32_t n_chunks = ((n_tiles - gw + n_warps - 1) / n_warps) * CPT;   // this warp's chunks

    if (lane == 0) {
        for (int i = 0; i < STAGES; i++) mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    const int var = 2 * p + h;                                        // the variable this lane owns after the exchange
    const int sub = p + 2 * h;                                        // 0..3: share of the per-cell prefetches
    const uint32_t Np = a.g.Npad;

    // the lane's last column slot: a column pair, the single column (MC - 1), or nothing
    const int pi_last = 2 * (NSLOT - 1) + h;
    const int kind = pi_last < NP ? 0 : (pi_last == NP ? 1 : 2);
    const uint32_t off_reg = (uint32_t)(h * CT * 16 + cl * 16);       // + i * 2 CT 16
    const uint32_t off_last = kind == 0 ? (uint32_t)(pi_last * CT * 16 + cl * 16) : (kind == 1 ? (uint32_t)(NP * CT * 16 + (cl >> 1) * 16) : (uint32_t)(cl * 16));
    const bool sel_y = kind == 1 && (cl & 1);
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
    auto load_ids = [&](uint32_t t, int s) {
        const uint32_t * __restrict__ q = a.ids + ((size_t)t * S + s) * (MC * CT) + cl;
        id0 = q[0];
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int m = 4 * i + 2 * h + j;
                id[2 * i + j] = m < MC ? q[m * CT] : 0u;
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
        double uA = 0.0, uB = 0.0;
        double dof[S][KR];
        double w[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            // 1. right-hand sides b[m] = U[nbr m] - U[cell] of the lane's columns, both variables of its pair
            cp_async_wait_all();
            __syncwarp();
            if (s == 0) { uA = live ? fx[(13 + 2 * p) * CT] : 0.0; uB = live ? fx[(14 + 2 * p) * CT] : 0.0; }
            const bool empty = empty_cur;                             // empty stencil (:896-899) or padding cell
            double bA[NSLOT][2], bB[NSLOT][2];
            {
                const double * ub = ubuf + cl * 4 + 2 * p;
#pragma unroll
                for (int i = 0; i < NSLOT - 1; i++) {
                    const double2 t0 = *reinterpret_cast<const double2 *>(ub + (4 * i + 2 * h) * UROW);
                    const double2 t1 = *reinterpret_cast<const double2 *>(ub + (4 * i + 2 * h + 1) * UROW);
                    bA[i][0] = t0.x - uA; bB[i][0] = t0.y - uB;
                    bA[i][1] = t1.x - uA; bB[i][1] = t1.y - uB;
                }
                const double2 t0 = *reinterpret_cast<const double2 *>(ub + lcol0 * UROW);
                const double2 t1 = *reinterpret_cast<const double2 *>(ub + lcol1 * UROW);
                bA[NSLOT - 1][0] = kind == 2 ? 0.0 : t0.x - uA; bB[NSLOT - 1][0] = kind == 2 ? 0.0 : t0.y - uB;
                bA[NSLOT - 1][1] = kind == 0 ? t1.x - uA : 0.0; bB[NSLOT - 1][1] = kind == 0 ? t1.y - uB : 0.0;
            }
            __syncwarp();                                             // the warp's reads of ubuf precede its next writes
            // 2. requests for what comes next (the next tile's first stencil after the last one of this tile)
            empty_cur = id0 == (s + 1 < S ? cell : next * CT + cl);
            if (s + 1 < S || has_next) request_states();
            if (s + 1 == S && has_next) prefetch_tile(next, (it + 1) & 1);
            cp_async_commit();
            // 3. ids of the stencil after that (plain loads, consumed a stencil later)
            {
                const bool in_tile = s + 2 < S;
                if (in_tile || has_next) load_ids(in_tile ? tile : next, in_tile ? s + 2 : s + 2 - S);
            }
            // 4. dofs a_k = sum_m A'[k][m] b[m]: each lane sums its half of the columns for two variables; the halves
            //    are exchanged with one shuffle per row
#pragma unroll
            for (int ch = 0; ch < C::NCH; ch++) {
                if (!ready) mbar_wait(&full_bar[c_st], c_par);
                const uint32_t n_st = c_st + 1 == STAGES ? 0u : c_st + 1, n_par = c_st + 1 == STAGES ? c_par ^ 1u : c_par;
                ready = mbar_test(&full_bar[n_st], n_par);            // probe the next chunk now, use the answer after this chunk's arithmetic
                const unsigned char * base = ring + (size_t)c_st * C::CHUNK_BYTES;
#pragma unroll
                for (int r = 0; r < C::RC; r++) {
                    const unsigned char * row = base + (size_t)r * C::ROW_DOUBLES * 8;
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                    for (int i = 0; i < NSLOT - 1; i++) {
                        const double2 c = *reinterpret_cast<const double2 *>(row + off_reg + i * 2 * CT * 16);
                        a0 = fma(c.x, bA[i][0], a0); a1 = fma(c.y, bA[i][1], a1);
                        b0 = fma(c.x, bB[i][0], b0); b1 = fma(c.y, bB[i][1], b1);
                    }
                    {
                        const double2 t = *reinterpret_cast<const double2 *>(row + off_last);
                        const double cx = kind == 2 ? 0.0 : (sel_y ? t.y : t.x);
                        const double cy = kind == 0 ? t.y : 0.0;
                        a0 = fma(cx, bA[NSLOT - 1][0], a0); a1 = fma(cy, bA[NSLOT - 1][1], a1);
                        b0 = fma(cx, bB[NSLOT - 1][0], b0); b1 = fma(cy, bB[NSLOT - 1][1], b1);
                    }
                    const double pa = a0 + a1, pb = b0 + b1;
                    const double send = h ? pa : pb, keep = h ? pb : pa;
                    dof[s][ch * C::RC + r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                __syncwarp();                                         // every lane is done with the stage
                if (lane == 0 && p_g < n_chunks) {                    // refill it with the chunk STAGES ahead
                    fence_proxy_async();
                    issue();
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
                for (int j = k; j < KR; j++) t = fma(a.OIs[k * KR + j], dof[s][j], t);
                si = fma(dof[s][k], t, si);
            }
            const double x = si + 1.0e-12;                            // 1/(SI+eps)^6 :940-944
            const double x2 = x * x, x3 = x2 * x;
            w[s] = empty ? 0.0 : 1.0 / (x3 * x3);
        }

        if (live) {
            const double u_self = h ? uB : uA;
            // non-linear weights :948-981 (reference-faithful unless fixed_weights)
            double sd = 0.0;
#pragma unroll
            for (int s = 1; s < S; s++) sd += w[s];
            if (w[0] / (sd + w[0]) > 1.0e-7) {
                w[0] = 1.0;
#pragma unroll
                for (int s = 1; s < S; s++) w[s] = 0.0;
            } else {
#pragma unroll
                for (int s = 1; s < S; s++) {
                    if (w[s] / sd > 1.0e-5) w[s] = (1.0 / K);
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
    static int ctas = 0;
    if (!ctas) {
        cudaFuncSetAttribute(teno_stream_warp_kernel<ORDER, MONO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, teno_stream_warp_kernel<ORDER, MONO>, WTHREADS, smem);
        ctas = sms * (per_sm > 0 ? per_sm : 1);
    }
    if (!a.n_tiles) return;
    const uint32_t need = (a.n_tiles + WARPS - 1) / WARPS;
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
