// TENO reconstruction, streaming variant (FAST floating-point mode only) — TENOFunctor::operator()
// (numerics/face_reconstruction.cpp:866-1039) restructured around the one thing that bounds it on a B200: the
// pseudo-inverse tables (5.5 kB per cell in the compact layout below) have to cross HBM once per RK stage.
// This header holds what the streaming kernel and the table builders share (tile geometry, chunking, the polynomial
// evaluation); the kernel itself is teno_stream_warp.cuh (warp-private TMA rings).
//
//   * compact tables: row 0 and column 0 of every reference matrix are exact zeros (face_reconstruction.cpp:620-675
//     re-embeds the (M-1)x(K-1) pseudo-inverse), so only rows 1..K-1 x columns 1..M-1 are stored, and the transformed
//     areas (:903-910) are folded into the columns on the host: a_k = sum_m (A+[k][m] area[m]) (U_m - U_i).
//   * table rows are interleaved across the 8 cells of a tile in 16-byte column pairs; a chunk of RC rows is one contiguous
//     block, moved by one 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx::bytes -> SASS UBLKCP) with an L2
//     evict-first policy: the stream is read exactly once per stage and must not push the conserved state (gathered
//     through L2) out of the cache.
//   * the reference's loop nest (variable -> face -> quadrature point -> stencil -> dof, accumulating in memory) is
//     re-associated: c_k = sum_s w_s a_sk first, then one K-term polynomial per quadrature point.  Stencils whose weight
//     is exactly zero are skipped as in the reference (:1011), so a non-finite dof of an unused stencil cannot leak.
//   Results differ from the reference by rounding only (<= 1e-12 relative per step, asserted in tests/test_gpu_parity.py);
//   the bit-faithful variants are teno_strict_stream_kernel / teno_recon_kernel in STRICT mode.
#pragma once
#include <utility>

#include "stream_ptx.cuh"

namespace stream {

constexpr int CT = FAST_CT;            // cells per tile = the 8 cells of one warp

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

}  // namespace stream
#include "teno_stream_warp.cuh"
