// TENO reconstruction, BIT-FAITHFUL streaming variant (STRICT floating-point mode) — TENOFunctor::operator()
// (numerics/face_reconstruction.cpp:866-1039) with the reference's operations in the reference's order (this file is
// compiled with -fmad=false): b[m] = area_t[m] (U[nbr m] - U[i]), k-ascending / m-ascending sums for the dofs, the full
// K x K oscillation matrix, true divisions in the weights, stencil-by-stencil accumulation of the face values.  What it
// shares with the FAST kernel (teno_stream_warp.cuh) is only the data movement: a tile of 8 cells belongs to one warp,
// which streams the tile's reference-layout tables (transformed areas + the full K x Mp matrices, 7.0 kB per cell for
// p = 3) through its own shared-memory ring with TMA bulk copies, fetches neighbour states with cp.async one stencil
// ahead and the tile's geometry one tile ahead.  Lane = (cell, conserved variable) as in teno_recon_kernel, whose
// results it reproduces bit for bit (tests/test_gpu_parity.py runs every STRICT case through it).
#pragma once

#include "stream_ptx.cuh"

namespace sstream {

using namespace stream;

constexpr int SCT = TILE;                      // 8 cells per tile (the STRICT tables are interleaved over TILE cells)
constexpr int SWARPS = 4;
constexpr int STHREADS = 32 * SWARPS;
constexpr int SS = 4;                          // stencils per cell: central + one per face of a triangle

template <int ORDER, int MP> struct SCfg {
    static constexpr int K = (ORDER + 1) * (ORDER + 2) / 2;
    static constexpr int ROWB = (MP / 2) * SCT * 16;                          // bytes of one matrix row of a tile
    static constexpr int AREAB = MP * SCT * 8;                                // bytes of a stencil's transformed areas
    static constexpr int RC = ORDER == 3 ? 5 : 3;                             // rows per chunk (divides K = 3, 6, 10, 15): the rows of a chunk
                                                                              // are independent m-ascending chains - the kernel's only ILP
    static constexpr int NCH = K / RC;
    static constexpr int CPS = 1 + NCH;                                       // chunks per stencil: areas, then the rows
    static constexpr int STAGEB = AREAB > RC * ROWB ? AREAB : RC * ROWB;
    static constexpr int STAGES = 19500 / STAGEB > 8 ? 8 : (19500 / STAGEB < 2 ? 2 : 19500 / STAGEB);
    static constexpr size_t RING = (size_t)STAGES * STAGEB;
    static constexpr size_t UBUF = (size_t)MP * SCT * 4 * 8;                  // neighbour states [m][cell][4]
    static constexpr size_t FXBUF = (size_t)2 * FX_ROWS * SCT * 8;
    static constexpr size_t PER_WARP = (RING + UBUF + FXBUF + 127) / 128 * 128;
    static constexpr size_t TOTAL = PER_WARP * SWARPS;
    static_assert(K % RC == 0, "rows per chunk must divide K");
    static_assert(STAGEB % 16 == 0 && AREAB % 16 == 0, "bulk copies move multiples of 16 bytes");
};

// out += w * dof_k * (psi_k(x, y) + cbar_k) for k = 0..K-1 in order (:1021-1030), exponents resolved at compile time
template <int... Ks>
__device__ __forceinline__ double poly_accumulate_reg(double out, double w, const double * dof, const double * Px, const double * Py,
                                                      const double * cbar, std::integer_sequence<int, Ks...>) {
    ((out += w * dof[Ks] * (Px[std::integral_constant<int, dof_ex(Ks)>::value] * Py[std::integral_constant<int, dof_ey(Ks)>::value] + cbar[Ks])), ...);
    return out;
}

template <int ORDER, int MP>
__global__ void __launch_bounds__(STHREADS, 2) teno_strict_stream_kernel(const __grid_constant__ ReconArgs a) {
    using C = SCfg<ORDER, MP>;
    constexpr int K = C::K, STAGES = C::STAGES, NI = (MP + 3) / 4;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bars[SWARPS][STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char * ring = smem + (size_t)warp * C::PER_WARP;
    double * ubuf = reinterpret_cast<double *>(ring + C::RING);
    double * fxbuf = reinterpret_cast<double *>(ring + C::RING + C::UBUF);
    uint64_t * full_bar = full_bars[warp];

    const uint32_t n_tiles = (a.g.N_recon + SCT - 1) / SCT;
    const uint32_t n_warps = gridDim.x * SWARPS;
    const uint32_t gw = blockIdx.x * SWARPS + warp;
    if (gw >= n_tiles) return;
    const uint32_t n_chunks = ((n_tiles - gw + n_warps - 1) / n_warps) * (SS * C::CPS);

    if (lane == 0) {
        for (int i = 0; i < STAGES; i++) mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // ---- the warp's own producer (lane 0)
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    uint32_t p_g = 0, p_st = 0, p_c = 0, p_s = 0, p_tile = gw;
    auto issue = [&]() {
        const size_t ts = (size_t)p_tile * SS + p_s;
        const void * src;
        uint32_t bytes;
        if (p_c == 0) { src = a.st_area + ts * MP * SCT; bytes = C::AREAB; }
        else { src = a.st_mat + (ts * K + (size_t)(p_c - 1) * C::RC) * (MP / 2) * SCT * 2; bytes = C::RC * C::ROWB; }
        mbar_expect_tx(&full_bar[p_st], bytes);
        bulk_g2s(ring + (size_t)p_st * C::STAGEB, src, bytes, &full_bar[p_st], policy);
        p_g++;
        if (++p_st == STAGES) p_st = 0;
        if (++p_c == C::CPS) { p_c = 0; if (++p_s == SS) { p_s = 0; p_tile += n_warps; } }
    };
    if (lane == 0) {
        for (int i = 0; i < STAGES && p_g < n_chunks; i++) issue();
    }

    const int cl = lane >> 2, var = lane & 3;
    const uint32_t Np = a.g.Npad;

    auto prefetch_tile = [&](uint32_t t, uint32_t parity) {
        const uint32_t cell = t * SCT + cl;
        if (cell >= a.g.N_recon) return;
        double * dst = fxbuf + (size_t)parity * FX_ROWS * SCT + cl;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int v = var + 4 * i;
            cp_async8(dst + v * SCT, a.g.slot_fx + (size_t)v * Np + cell);
        }
        cp_async8(dst + (13 + var) * SCT, a.Uin + 4 * (size_t)cell + var);
    };
    uint32_t id[NI], id0;                                             // neighbours var, var + 4, ... of the NEXT stencil; its first id
    auto load_ids = [&](uint32_t t, uint32_t s) {
        const uint32_t * __restrict__ q = a.st_ids + ((size_t)t * SS + s) * (MP * SCT) + cl;
        id0 = ld_id(q);
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int m = var + 4 * i;
            id[i] = ld_id(q + (m < MP ? m : MP - 1) * SCT);
        }
    };
    auto request_states = [&](uint32_t self) {                        // an empty stencil (ids = NO_FACE) fetches the cell itself; unused
        double * dst = ubuf + cl * 4;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int m = var + 4 * i;
            if (m < MP) {
                const double * src = a.Uin + 4 * (size_t)(id[i] == NO_FACE ? self : id[i]);
                cp_async16(dst + (size_t)m * SCT * 4, src);
                cp_async16(dst + (size_t)m * SCT * 4 + 2, src + 2);
            }
        }
    };

    uint32_t tile = gw;
    bool empty_cur;
    load_ids(tile, 0);
    prefetch_tile(tile, 0);
    empty_cur = id0 == NO_FACE;
    request_states(tile * SCT + cl);
    cp_async_commit();
    load_ids(tile, 1);

    uint32_t c_st = 0, c_par = 0;
    auto release = [&]() {                                            // every lane is done with stage c_st: refill it, move on
        __syncwarp();
        if (lane == 0 && p_g < n_chunks) issue();
        if (++c_st == STAGES) { c_st = 0; c_par ^= 1u; }
    };

    for (uint32_t it = 0; tile < n_tiles; it++) {
        const uint32_t next = tile + n_warps;
        const bool has_next = next < n_tiles;
        const uint32_t cell = tile * SCT + cl;
        const bool live = cell < a.g.N_recon;
        const double * fx = fxbuf + (size_t)(it & 1) * FX_ROWS * SCT + cl;
        double u_self = 0.0, area0 = 0.5;
        double dof[SS][K];
        double w[SS];
#pragma unroll 1
        for (uint32_t s = 0; s < (uint32_t)SS; s++) {
            cp_async_wait_all();
            __syncwarp();
            if (s == 0) u_self = live ? fx[(13 + var) * SCT] : 0.0;
            const bool empty = empty_cur || !live;
            // ---- b[m] = area_t[m] (U[nbr m] - U[i])  :903-910
            double b[MP];
            mbar_wait(&full_bar[c_st], c_par);
            {
                const double * areas = reinterpret_cast<const double *>(ring + (size_t)c_st * C::STAGEB) + cl;
                const double * ub = ubuf + cl * 4 + var;
#pragma unroll
                for (int m = 0; m < MP; m++) b[m] = areas[m * SCT] * (ub[(size_t)m * SCT * 4] - u_self);
                if (s == 0) area0 = areas[0];
            }
            release();
            // ---- requests for the next stencil (the next tile's first one after the last of this tile)
            empty_cur = id0 == NO_FACE;
            if (s + 1 < SS || has_next) request_states(s + 1 < SS ? cell : next * SCT + cl);
            if (s + 1 == SS && has_next) prefetch_tile(next, (it + 1) & 1);
            cp_async_commit();
            {
                const bool in_tile = s + 2 < SS;
                if (in_tile || has_next) load_ids(in_tile ? tile : next, in_tile ? s + 2 : s + 2 - SS);
            }
            // ---- a = A+ b, k-ascending rows, m-ascending sums  :915-918
            double d[K];
#pragma unroll
            for (int ch = 0; ch < C::NCH; ch++) {
                mbar_wait(&full_bar[c_st], c_par);
                const unsigned char * base = ring + (size_t)c_st * C::STAGEB + cl * 16;
#pragma unroll
                for (int r = 0; r < C::RC; r++) {
                    const unsigned char * row = base + (size_t)r * C::ROWB;
                    double sum = 0.0;
#pragma unroll
                    for (int m2 = 0; m2 < MP / 2; m2++) {
                        const double2 c = *reinterpret_cast<const double2 *>(row + m2 * SCT * 16);
                        sum += c.x * b[2 * m2];
                        sum += c.y * b[2 * m2 + 1];
                    }
                    d[ch * C::RC + r] = sum;
                }
                release();
            }
            // ---- SI = a . (OI a), w = 1/(SI + eps)^6  :922-944
            double ws = 0.0;
            if (!empty) {
                double si = 0.0;                                      // (OI a)_k is consumed as soon as it is complete: same sums, same order
#pragma unroll
                for (int k = 0; k < K; k++) {
                    double t = 0.0;
#pragma unroll
                    for (int j = 0; j < K; j++) t += a.OI[k * K + j] * d[j];
                    si += d[k] * t;
                }
                const double x = si + 1.0e-12;
                const double x2 = x * x, x3 = x2 * x;
                ws = 1.0 / (x3 * x3);
            }
#pragma unroll
            for (int t = 0; t < SS; t++) {
                if (s == (uint32_t)t) {                               // warp-uniform
                    w[t] = ws;
#pragma unroll
                    for (int k = 0; k < K; k++) dof[t][k] = d[k];
                }
            }
        }

        if (live) {
            // non-linear weights :948-981 (reference-faithful: the central weight stays raw in the ENO branch)
            double sd = 0.0;
#pragma unroll
            for (int s = 1; s < SS; s++) sd += w[s];
            if (w[0] / (sd + w[0]) > 1.0e-7) {
                w[0] = 1.0;
#pragma unroll
                for (int s = 1; s < SS; s++) w[s] = 0.0;
            } else {
#pragma unroll
                for (int s = 1; s < SS; s++) {
                    if (w[s] / sd > 1.0e-5) w[s] = (1.0 / K);
                    else if (a.fixed_weights) w[s] = 0.0;
                }
                sd = 0.0;
#pragma unroll
                for (int s = 1; s < SS; s++) sd += w[s];
#pragma unroll
                for (int s = 1; s < SS; s++) w[s] /= sd;
                if (a.fixed_weights) w[0] = 0.0;
            }
            double cbar[K];                                           // psi_bar_k / area_t[s][0] :1028-1029
#pragma unroll
            for (int k = 0; k < K; k++) cbar[k] = a.fixed_weights ? -a.psi_bar[k] : a.psi_bar[k] / area0;
            const int nf = a.g.nfc[cell];
            const int Q = a.g.Q;
            for (int j = 0; j < nf; j++) {                            // :985-1034
                const double x0 = fx[(4 * j) * SCT], y0 = fx[(4 * j + 1) * SCT], x1 = fx[(4 * j + 2) * SCT], y1 = fx[(4 * j + 3) * SCT];
                for (int q = 0; q < Q; q++) {
                    const double tq = (a.qf_x[q] + 1.0) * 0.5;
                    const double xq = tq * (x1 - x0) + x0, yq = tq * (y1 - y0) + y0;
                    double Px[ORDER + 1], Py[ORDER + 1];
                    basis_values<ORDER>(a.basis, xq, Px);
                    basis_values<ORDER>(a.basis, yq, Py);
                    double out = u_self;
#pragma unroll
                    for (int s = 0; s < SS; s++) {
                        if (w[s] == 0.0) continue;
                        out = poly_accumulate_reg(out, w[s], dof[s], Px, Py, cbar, std::make_integer_sequence<int, K>{});
                    }
                    a.Fc[((size_t)cell * (a.g.n_slots * Q) + (j * Q + q)) * 4 + var] = out;
                }
            }
        }
        __syncwarp();
        tile = next;
    }
}

template <int ORDER, int MP>
static void launch_strict_stream(const ReconArgs & a, cudaStream_t st) {
    const size_t smem = SCfg<ORDER, MP>::TOTAL;
    const uint32_t n_tiles = (a.g.N_recon + SCT - 1) / SCT;
    if (!n_tiles) return;
    const int ctas = persistent_ctas(reinterpret_cast<const void *>(teno_strict_stream_kernel<ORDER, MP>), STHREADS, smem);
    const uint32_t need = (n_tiles + SWARPS - 1) / SWARPS;
    teno_strict_stream_kernel<ORDER, MP><<<need < (uint32_t)ctas ? need : (unsigned)ctas, STHREADS, smem, st>>>(a);
}

// the streaming variant covers what the reference's TENO supports (triangles: central + 3 directional stencils)
static bool strict_stream_supported(const ReconArgs & a) {
    static const bool off = [] { const char * e = getenv("MLB_STRICT_STREAM"); return e && e[0] == '0'; }();
    return !off && a.S == SS && a.g.n_slots == SS - 1;
}

}  // namespace sstream
