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

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

template <int ORDER, int STAGES>
__global__ void __launch_bounds__(THREADS, 2) teno_stream_kernel(const __grid_constant__ ReconStreamArgs a) {
    using C = Cfg<ORDER>;
    constexpr int K = C::K, KR = C::KR, MC = C::MC, NP = C::NP, Q = C::Q, S = FAST_S;
    extern __shared__ __align__(128) unsigned char ring[];   // STAGES x CHUNK_BYTES
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];

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
                const unsigned char * src = reinterpret_cast<const unsigned char *>(a.mat) + (size_t)tile * (S * C::NCH) * C::CHUNK_BYTES;
                for (int ch = 0; ch < S * C::NCH; ch++, g++) {
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
    const int cl = tid >> 2, var = tid & 3;
    const uint32_t Np = a.g.Npad;
    const double * __restrict__ Uv = a.Uin + (size_t)var * Np;
    uint32_t g = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t cell = tile * CT + cl;
        const bool live = cell < a.g.N_recon;
        const double u_self = live ? Uv[cell] : 0.0;
        const uint32_t * __restrict__ ids = a.ids + (size_t)tile * (S * MC * CT) + cl;
        double dof[S][KR];
        double w[S];
        uint32_t id[MC];
#pragma unroll
        for (int m = 0; m < MC; m++) id[m] = ids[m * CT];
#pragma unroll
        for (int s = 0; s < S; s++) {
            const bool empty = !live || id[0] == NO_FACE;            // empty stencil (:896-899) or padding cell
            double b[MC];
#pragma unroll
            for (int m = 0; m < MC; m++) b[m] = empty ? 0.0 : Uv[id[m]] - u_self;
            if (s + 1 < S) {                                          // ids of the next stencil travel while this one computes
#pragma unroll
                for (int m = 0; m < MC; m++) id[m] = ids[((s + 1) * MC + m) * CT];
            }
#pragma unroll
            for (int ch = 0; ch < C::NCH; ch++, g++) {
                const uint32_t st = g % STAGES, use = g / STAGES;
                mbar_wait(&full_bar[st], use & 1u);
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
        if (!live) continue;

        // non-linear weights :948-981 (reference-faithful unless fixed_weights)
        {
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
#pragma unroll
                for (int s = 1; s < S; s++) w[s] /= sd;
                if (a.fixed_weights) w[0] = 0.0;
            }
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
        const double area0 = a.area0[cell];
        double cb = 0.0;                                              // sum_k c_k psi_bar_k / area_t[0]  (:1028-1029)
        const double cscale = a.fixed_weights ? -1.0 : 1.0 / area0;
#pragma unroll
        for (int k = 0; k < KR; k++) cb = fma(c[k], a.psi_bar[k + 1] * cscale, cb);

        const int nf = a.g.nfc[cell];
        for (int j = 0; j < nf; j++) {                                // :985-1034
            const double * fx = a.g.slot_fx + ((size_t)j * 4) * Np + cell;
            const double x0 = fx[0], y0 = fx[Np], x1 = fx[2 * (size_t)Np], y1 = fx[3 * (size_t)Np];
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const double tq = (a.qf_x[q] + 1.0) * 0.5;
                const double xq = tq * (x1 - x0) + x0, yq = tq * (y1 - y0) + y0;
                double Px[ORDER + 1], Py[ORDER + 1];
                legendre_values<ORDER>(xq, Px);
                legendre_values<ORDER>(yq, Py);
                double out = u_self + cb;
#pragma unroll
                for (int k = 0; k < KR; k++) out = fma(c[k], Px[dof_ex(k + 1)] * Py[dof_ey(k + 1)], out);
                a.Fc[((size_t)(j * Q + q) * 4 + var) * Np + cell] = out;
            }
        }
    }
}

template <int ORDER>
static void launch_stream_t(const ReconStreamArgs & a, cudaStream_t st) {
    using C = Cfg<ORDER>;
    constexpr int STAGES = fast_stages(ORDER);
    const size_t smem = (size_t)STAGES * C::CHUNK_BYTES;
    static int ctas = 0;
    if (!ctas) {
        cudaFuncSetAttribute(teno_stream_kernel<ORDER, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, teno_stream_kernel<ORDER, STAGES>, THREADS, smem);
        ctas = sms * (per_sm > 0 ? per_sm : 1);
    }
    if (!a.n_tiles) return;
    const unsigned grid = a.n_tiles < (uint32_t)ctas ? a.n_tiles : (unsigned)ctas;   // persistent: one CTA slot per resident CTA
    teno_stream_kernel<ORDER, STAGES><<<grid, THREADS, smem, st>>>(a);
}

static bool stream_supported(int order, int M, int Q, int basis, int n_slots) {
    if (basis != MLB_BASIS_LEGENDRE || n_slots != FAST_S - 1) return false;
    if (order < 1 || order > 4) return false;
    const int K = (order + 1) * (order + 2) / 2;
    return M == 2 * K && Q == (order + 1) / 2;
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

}  // namespace stream
