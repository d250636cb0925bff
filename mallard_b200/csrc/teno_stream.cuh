// TENO reconstruction, streaming variant (FAST floating-point mode only) — TENOFunctor::operator()
// (numerics/face_reconstruction.cpp:866-1039) restructured around the one thing that bounds it on a B200: the
// pseudo-inverse tables (5.5 kB per cell in the compact layout below) have to cross HBM once per RK stage.
//
//   * persistent CTAs (2 per SM): 4 consumer warps + 1 producer warp.  The producer's elected lane streams the tables
//     of the CTA's cell tiles through a ring of shared-memory stages with 1-D TMA bulk copies
//     (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes → SASS UBLKCP) that complete on "full"
//     mbarriers; consumers release a stage through its "empty" mbarrier.  Up to STAGES x 14.6 kB per CTA are in
//     flight, independent of what the consumers are doing, so HBM latency is covered by the copy engine rather than by
//     resident warps.  The stream carries an L2 evict-first policy: it is read exactly once per stage and must not
//     push the conserved state (67 MB, gathered through L2) out of the cache.
//   * thread = (cell, conserved variable); a tile is 32 cells.  Table rows are interleaved across the tile's cells in
//     16-byte column pairs, so that one LDS.128 of a warp covers 8 cells x 16 B = one conflict-free 128 B wavefront
//     (the four variable-threads of a cell read the same address: broadcast).
//   * compact tables: row 0 and column 0 of every reference matrix are exact zeros (face_reconstruction.cpp:620-675
//     re-embeds the (M-1)x(K-1) pseudo-inverse), so only rows 1..K-1 x columns 1..M-1 are stored, and the transformed
//     areas (:903-910) are folded into the columns on the host: a_k = sum_m (A+[k][m] area[m]) (U_m - U_i).
//   * the reference's loop nest (variable → face → quadrature point → stencil → dof, accumulating in memory) is
//     re-associated: c_k = sum_s w_s a_sk first, then one K-term polynomial per quadrature point.  Stencils whose weight
//     is exactly zero are skipped as in the reference (:1011), so a non-finite dof of an unused stencil cannot leak.
//   Results differ from the reference by rounding only (<= 1e-12 relative per step, asserted in tests/test_gpu_parity.py);
//   the bit-faithful variant is teno_recon_kernel in STRICT mode.
#pragma once
#include <utility>

#include "stream_ptx.cuh"

namespace stream {

constexpr int CT = FAST_CT;            // cells per tile
constexpr int CONSUMERS = 4 * CT;      // threads: one per (cell, variable)
constexpr int THREADS = CONSUMERS + 32;

template <int ORDER> struct Cfg {
    static constexpr int K = (ORDER + 1) * (ORDER + 2) / 2;
    static constexpr int M = 2 * K;
    static constexpr int KR = K - 1;                 // stored rows   k = 1..K-1
    static constexpr int MC = M - 1;                 // stored columns m = 1..M-1 (always odd)
    static constexpr int NP = MC / 2;                // 16-byte column pairs per row, plus one single column
    static constexpr int ROW_DOUBLES = (2 * NP + 1) * CT;
    static constexpr int RC = fast_rows_per_chunk(ORDER);
    static constexpr int NCH = KR / RC;              // chunks per stencil
    static constexpr int CHUNK_BYTES = RC * ROW_DOUBLES * 8;
    static constexpr int Q = (ORDER + 1) / 2;        // default face quadrature (face_reconstruction.cpp:115-116)
    static_assert(KR % RC == 0, "rows per chunk must divide K-1");
    static_assert(CHUNK_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
};

// Sum_k c[k] psi_k(x, y) over the stored dofs k = 1..K-1 with the exponents resolved at compile time (a run-time
// dof_ex()/dof_ey() would index Px/Py dynamically and push them into local memory).
template <int... Ks>
__device__ __forceinline__ double poly_sum(const double * c, const double * Px, const double * Py, double out, std::integer_sequence<int, Ks...>) {
    ((out = fma(c[Ks], Px[std::integral_constant<int, dof_ex(Ks + 1)>::value] * Py[std::integral_constant<int, dof_ey(Ks + 1)>::value], out)), ...);
    return out;
}

static bool stream_supported(int order, int M, int Q, int basis, int n_slots) {
    if ((basis != MLB_BASIS_LEGENDRE && basis != MLB_BASIS_MONOMIAL) || n_slots != FAST_S - 1) return false;
    if (order < 1 || order > 4) return false;
    const int K = (order + 1) * (order + 2) / 2;
    return M == 2 * K && Q == (order + 1) / 2;
}

#if MLB_FAST_CT != 32
}  // namespace stream
#include "teno_stream_warp.cuh"
namespace stream {
#else
template <int ORDER> struct Smem {
    using C = Cfg<ORDER>;
    static constexpr int STAGES = fast_stages(ORDER);
    static constexpr size_t RING = (size_t)STAGES * C::CHUNK_BYTES;
    static constexpr size_t UBUF = (size_t)2 * C::MC * CT * 4 * 8;      // neighbour states [m][cell][4], double-buffered over stencils
    static constexpr size_t FXBUF = (size_t)2 * FX_ROWS * CT * 8;      // per tile parity
    static constexpr size_t TOTAL = RING + UBUF + FXBUF;
};

// OWNVAR = false: the four variable-threads of a cell share the fetch of a neighbour list (16-byte copies, thread `var` takes
//                  neighbours var, var + 4, ...); OWNVAR = true: every thread fetches its own variable (8 bytes) of every
//                  neighbour — the four lanes of a cell then hit the same 32-byte sector and the same shared-memory row.
// 2 CTAs of 5 warps per SM put 3 warps on two of the four schedulers: 16384 / 3 / 32 = 170 registers per thread
// (a cap of 200 silently drops the kernel to ONE CTA per SM, profiles/r01h).
#ifndef MLB_STREAM_MAXNREG
#define MLB_STREAM_MAXNREG 168
#endif
template <int ORDER, bool OWNVAR, bool MONO>
__global__ void
#if MLB_STREAM_MAXNREG > 0
__maxnreg__(MLB_STREAM_MAXNREG)
#else
__launch_bounds__(THREADS, 2)
#endif
teno_stream_kernel(const __grid_constant__ ReconStreamArgs a) {
    using C = Cfg<ORDER>;
    using SM = Smem<ORDER>;
    constexpr int K = C::K, KR = C::KR, MC = C::MC, NP = C::NP, Q = C::Q, S = FAST_S, NF = FAST_S - 1, STAGES = SM::STAGES;
    constexpr int CPT = S * C::NCH;                            // chunks per tile
    constexpr int NG = OWNVAR ? MC : (MC + 3) / 4;             // neighbour states a thread requests per stencil
    extern __shared__ __align__(128) unsigned char smem[];     // ring | ubuf | fxbuf
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    unsigned char * ring = smem;
    double * ubuf = reinterpret_cast<double *>(smem + SM::RING);
    double * fxbuf = reinterpret_cast<double *>(smem + SM::RING + SM::UBUF);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t n_tiles = a.n_tiles;
    if (tid >= CONSUMERS) {
        // ---------------- producer warp: one elected lane drives the copy engine ----------------
        if (tid == CONSUMERS) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            uint32_t g = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const unsigned char * src = reinterpret_cast<const unsigned char *>(a.mat) + (size_t)tile * CPT * C::CHUNK_BYTES;
                for (int ch = 0; ch < CPT; ch++, g++) {
                    const uint32_t st = g % STAGES, use = g / STAGES;
                    mbar_wait(&empty_bar[st], (use & 1u) ^ 1u);      // first use of a stage passes immediately
                    mbar_expect_tx(&full_bar[st], C::CHUNK_BYTES);
                    bulk_g2s(ring + (size_t)st * C::CHUNK_BYTES, src + (size_t)ch * C::CHUNK_BYTES, C::CHUNK_BYTES, &full_bar[st], policy);
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    // Nothing a consumer needs from global memory is waited for: it is requested with cp.async (LDGSTS: global → shared
    // memory without register staging) one stencil (neighbour states) or one tile (geometry, own state) ahead.
    //   ubuf[parity of s][m][cell][4]      conserved states of the stencil's cells: one 32-byte sector per neighbour, fetched
    //                                      by the cell's four variable-threads in turn (thread `var` takes m = var, var + 4, ...)
    //   fxbuf[parity of tile][row][cell]   face end points, area_t[0] and the cell's own state
    // The four variable-threads of a cell sit in one warp, so a __syncwarp() after cp.async.wait_all publishes the copies.
    const int cl = tid >> 2, var = tid & 3;
    const uint32_t Np = a.g.Npad;
    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;

    auto prefetch_tile = [&](uint32_t t, uint32_t parity) {
        const uint32_t cell = t * CT + cl;
        if (cell >= a.g.N_recon) return;
        double * dst = fxbuf + (size_t)parity * FX_ROWS * CT + cl;
#pragma unroll
        for (int i = 0; i < 3; i++) {                                 // value index var + 4 i of the 12 (slot j, component)
            const int v = var + 4 * i;
            cp_async8(dst + v * CT, a.g.slot_fx + (size_t)v * Np + cell);
        }
        if (var == 0) cp_async8(dst + 12 * CT, a.area0 + cell);
        cp_async8(dst + (13 + var) * CT, a.Uin + 4 * (size_t)cell + var);
    };
    uint32_t id[NG], id0;                                             // ids this thread fetches for the NEXT stencil; its first id
    auto load_ids = [&](uint32_t t, int s) {
        const uint32_t * __restrict__ p = a.ids + ((size_t)t * S + s) * (MC * CT) + cl;
        if (OWNVAR) {
#pragma unroll
            for (int i = 0; i < NG; i++) id[i] = p[i * CT];
            id0 = id[0];
        } else {
            id0 = p[0];
#pragma unroll
            for (int i = 0; i < NG; i++) id[i] = (var + 4 * i < MC) ? p[(var + 4 * i) * CT] : 0u;
        }
    };
    auto request_states = [&](int parity) {
        if (OWNVAR) {
            double * dst = ubuf + (size_t)parity * MC * CT * 4 + tid;
#pragma unroll
            for (int i = 0; i < NG; i++) cp_async8(dst + (size_t)i * CONSUMERS, a.Uin + 4 * (size_t)id[i] + var);
        } else {
            double * dst = ubuf + ((size_t)parity * MC * CT + cl) * 4;
#pragma unroll
            for (int i = 0; i < NG; i++) {
                if (var + 4 * i < MC) {
                    const double * src = a.Uin + 4 * (size_t)id[i];
                    cp_async16(dst + (size_t)(var + 4 * i) * CT * 4, src);
                    cp_async16(dst + (size_t)(var + 4 * i) * CT * 4 + 2, src + 2);
                }
            }
        }
    };

    bool empty_cur;                                                   // is the stencil whose states are in flight empty?
    load_ids(tile, 0);
    prefetch_tile(tile, 0);
    empty_cur = id0 == tile * CT + cl;                                // an empty stencil lists the cell itself (b = 0)
    request_states(0);
    cp_async_commit();
    load_ids(tile, 1);

    uint32_t ready = 0;
    for (uint32_t it = 0; tile < n_tiles; it++) {
        const uint32_t next = tile + gridDim.x;
        const bool has_next = next < n_tiles;
        const uint32_t cell = tile * CT + cl;
        const bool live = cell < a.g.N_recon;
        const double * fx = fxbuf + (size_t)(it & 1) * FX_ROWS * CT + cl;
        double u_self = 0.0;
        double dof[S][KR];
        double w[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            // 1. right-hand side b[m] = U[nbr m] - U[cell]: requested one stencil (own state: one tile) ago
            cp_async_wait_all();
            __syncwarp();
            if (s == 0) u_self = live ? fx[(13 + var) * CT] : 0.0;
            const bool empty = empty_cur;                             // empty stencil (:896-899) or padding cell
            double b[MC];
            {
                const double * ub = ubuf + (size_t)(s & 1) * MC * CT * 4 + tid;
#pragma unroll
                for (int m = 0; m < MC; m++) b[m] = ub[m * CONSUMERS] - u_self;
            }
            // 2. requests for what comes next (the next tile's first stencil after the last one of this tile); the ids
            //    were loaded one stencil ago
            empty_cur = id0 == (s + 1 < S ? cell : next * CT + cl);
            if (s + 1 < S || has_next) request_states((s + 1) & 1);
            if (s + 1 == S && has_next) prefetch_tile(next, (it + 1) & 1);
            cp_async_commit();
            // 3. ids of the stencil after that (plain coalesced loads, consumed a stencil later)
            {
                const bool in_tile = s + 2 < S;
                if (in_tile || has_next) load_ids(in_tile ? tile : next, in_tile ? s + 2 : s + 2 - S);
            }
            // 4. dofs a_k = sum_m A'[k][m] b[m], rows arriving chunk by chunk through the ring
#pragma unroll
            for (int ch = 0; ch < C::NCH; ch++) {
                const int pos = s * C::NCH + ch;
                uint32_t st, use;
                if (CPT % STAGES == 0) { st = pos % STAGES; use = it * (CPT / STAGES) + pos / STAGES; }
                else { const uint32_t g = it * CPT + pos; st = g % STAGES; use = g / STAGES; }
                if (!ready) mbar_wait(&full_bar[st], use & 1u);
                {   // probe the next chunk's barrier now; the answer is needed only after this chunk's arithmetic
                    uint32_t st2, use2;
                    if (CPT % STAGES == 0) { st2 = (pos + 1) % STAGES; use2 = it * (CPT / STAGES) + (pos + 1) / STAGES; }
                    else { const uint32_t g2 = it * CPT + pos + 1; st2 = g2 % STAGES; use2 = g2 / STAGES; }
                    ready = mbar_test(&full_bar[st2], use2 & 1u);
                }
                const unsigned char * base = ring + (size_t)st * C::CHUNK_BYTES + cl * 16;
#pragma unroll
                for (int r = 0; r < C::RC; r++) {
                    const unsigned char * row = base + (size_t)r * C::ROW_DOUBLES * 8;
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        const double2 c = *reinterpret_cast<const double2 *>(row + p * CT * 16);
                        s0 = fma(c.x, b[2 * p], s0);
                        s1 = fma(c.y, b[2 * p + 1], s1);
                    }
                    const double c1 = *reinterpret_cast<const double *>(row + NP * CT * 16 - cl * 8);
                    s0 = fma(c1, b[MC - 1], s0);
                    dof[s][ch * C::RC + r] = s0 + s1;
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[st]);
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
                    basis_values<ORDER>(MONO ? MLB_BASIS_MONOMIAL : MLB_BASIS_LEGENDRE, xq, Px);
                    basis_values<ORDER>(MONO ? MLB_BASIS_MONOMIAL : MLB_BASIS_LEGENDRE, yq, Py);
                    out[(j * Q + q) * 4] = poly_sum(c, Px, Py, u_self + cb, std::make_integer_sequence<int, KR>{});
                }
            }
        }
        __syncwarp();   // fxbuf[it & 1] is rewritten (by other lanes of this warp) two tiles from now at the earliest
        tile = next;
    }
}

template <int ORDER, bool OWNVAR, bool MONO>
static void launch_stream_v(const ReconStreamArgs & a, cudaStream_t st) {
    const size_t smem = Smem<ORDER>::TOTAL;
    static int ctas = 0;
    if (!ctas) {
        cudaFuncSetAttribute(teno_stream_kernel<ORDER, OWNVAR, MONO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, teno_stream_kernel<ORDER, OWNVAR, MONO>, THREADS, smem);
        ctas = sms * (per_sm > 0 ? per_sm : 1);
    }
    if (!a.n_tiles) return;
    const unsigned grid = a.n_tiles < (uint32_t)ctas ? a.n_tiles : (unsigned)ctas;   // persistent: one CTA slot per resident CTA
    teno_stream_kernel<ORDER, OWNVAR, MONO><<<grid, THREADS, smem, st>>>(a);
}
template <int ORDER>
static void launch_stream_t(const ReconStreamArgs & a, cudaStream_t st) {
    const bool mono = a.basis == MLB_BASIS_MONOMIAL;
    if (a.async_gather == 2) { if (mono) launch_stream_v<ORDER, true, true>(a, st); else launch_stream_v<ORDER, true, false>(a, st); }
    else { if (mono) launch_stream_v<ORDER, false, true>(a, st); else launch_stream_v<ORDER, false, false>(a, st); }
}

static void launch_stream(const ReconStreamArgs & a, cudaStream_t st) {
    switch (a.order) {
        case 1: launch_stream_t<1>(a, st); break;
        case 2: launch_stream_t<2>(a, st); break;
        case 3: launch_stream_t<3>(a, st); break;
        case 4: launch_stream_t<4>(a, st); break;
        default: break;
    }
}
#endif  // MLB_FAST_CT == 32

}  // namespace stream
