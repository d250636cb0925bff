// TENO reconstruction matrices built on the device — TENO::compute_reconstruction_matrices
// (numerics/face_reconstruction.cpp:477-741): for every (cell, stencil) the (M x K) matrix of basis-function means over
// the stencil's cells (:525-610), its Householder R factor (common_math.h:355-422), the two triangular solves that give
// the pseudo-inverse (:437-493), re-embedded with a zero first row and column (:620-675) — written straight into the
// compact, tile-interleaved table the streaming reconstruction kernel reads (teno_stream.cuh), with the transformed
// areas folded into the columns.  Nothing but node coordinates and stencil ids crosses PCIe; the 5.5 kB per cell of
// matrices never exist on the host.
//
// This file is compiled with -fmad=false and mirrors the host preprocessor (preprocess.cpp, compiled with
// -ffp-contract=off) operation by operation, so the tables are BIT-IDENTICAL to the host-built (= the reference's) ones:
// + - * / and sqrt are correctly rounded on both sides.  tests/test_gpu_parity.py asserts the equality.
//
// Mapping: thread = (cell, stencil); FAST_CT consecutive threads = the cells of one table tile (their stores are FAST_CT x 16
// contiguous bytes).
// The R factor (the only array the O(rows^2 cols^2) Householder sweep touches) lives in shared memory, element-major
// ([element][thread]: conflict-free); the original matrix and the solve workspace are per-thread local arrays.
#include <cuda_runtime.h>

#include "kernel_args.h"

namespace mlb {

namespace {

// exponents of the k-th basis function, graded ordering (TENO::calc_polynomial_indices, face_reconstruction.cpp:182-215)
__host__ __device__ constexpr int tb_ey(int k) { int p = 0; while (k > p) { k -= p + 1; p++; } return k; }
__host__ __device__ constexpr int tb_ex(int k) { int p = 0; while (k > p) { k -= p + 1; p++; } return p - k; }
template <int ORDER> struct TbCfg { static constexpr int THREADS = ORDER <= 3 ? 64 : 32; };   // order 4: 29 x 14 R factors, 118 kB per warp

template <int P> __device__ __forceinline__ double legendre0(int p, double x) {   // preprocess.cpp legendre(0, p, x): same term order
    switch (p) {
        case 0: return 1.0 * 1.0;
        case 1: return 1.0 * (1.0 * x);
        case 2: return 0.5 * (3.0 * x * x + -1.0);
        case 3: return 0.5 * (5.0 * x * x * x + -3.0 * x);
        default: return 0.125 * (35.0 * x * x * x * x + -30.0 * x * x + 3.0);
    }
}

__device__ __forceinline__ void m2_inverse(const double * A, double * Ai) {
    const double det = A[0] * A[3] - A[1] * A[2];
    const double r = 1.0 / det;
    Ai[0] = A[3] * r; Ai[1] = -A[1] * r; Ai[2] = -A[2] * r; Ai[3] = A[0] * r;
}
__device__ __forceinline__ void m2_apply(const double * A, const double * x, double * y) {
    const double y0 = A[0] * x[0] + A[1] * x[1], y1 = A[2] * x[0] + A[3] * x[1];
    y[0] = y0; y[1] = y1;
}

template <int ORDER>
__global__ void __launch_bounds__(TbCfg<ORDER>::THREADS) teno_tables_kernel(const __grid_constant__ TableBuildArgs a) {
    constexpr int K = (ORDER + 1) * (ORDER + 2) / 2, M = 2 * K, KR = K - 1, MC = M - 1, NP = MC / 2;
    constexpr int CT = FAST_CT, S = FAST_S, TB_THREADS = TbCfg<ORDER>::THREADS;
    extern __shared__ double sm[];                     // R[MC * KR][T] | v[MC][T] | colbuf[MC][T]
    const int tid = threadIdx.x;
    double * R = sm + tid;
    double * v = sm + (size_t)MC * KR * TB_THREADS + tid;
    double * cb = v + (size_t)MC * TB_THREADS;
#define R_(i, k) R[((i) * KR + (k)) * TB_THREADS]
#define V_(i) v[(i) * TB_THREADS]
#define C_(i) cb[(i) * TB_THREADS]

    // group g of FAST_CT consecutive threads handles (tile, stencil) = (g / S, g % S)
    static_assert(TB_THREADS % CT == 0, "thread block = whole tiles");
    const uint32_t gw = (blockIdx.x * TB_THREADS + tid) / CT;
    const uint32_t ft = gw / S;
    const int s = gw % S, fl = tid % CT;
    if (ft >= a.n_ftiles) return;
    const uint32_t cell = ft * CT + fl;
    constexpr size_t FROW = (size_t)(2 * NP + 1) * CT;
    double * out = a.fm_mat + ((size_t)ft * S + s) * KR * FROW;
    const uint32_t * ids = a.fm_ids + ((size_t)ft * S + s) * (MC * CT) + fl;
    const bool empty = cell >= a.n_recon || ids[0] == cell;   // boundary-face directional stencil / padding cell
    if (empty) {
        for (int k = 0; k < KR; k++) {
            double * row = out + (size_t)k * FROW;
            for (int p = 0; p < NP; p++) reinterpret_cast<double2 *>(row)[(size_t)p * CT + fl] = make_double2(0.0, 0.0);
            row[(size_t)2 * NP * CT + fl] = 0.0;
        }
        if (s == 0 && cell >= a.n_recon) a.fm_area0[cell] = 0.5;
        return;
    }

    // ---- rows of the matrix before the mean is removed (preprocess.cpp integrate_basis_rows), row 0 = the cell itself
    const double * T0 = a.tri_xy + 6 * (size_t)cell;
    const double o[2] = {T0[0], T0[1]};
    double J[4] = {T0[2] - o[0], T0[4] - o[0], T0[3] - o[1], T0[5] - o[1]}, Ji[4];
    m2_inverse(J, Ji);
    double B[MC * KR];                                  // rows 1..M-1 x columns 1..K-1, mean removed
    double at[M];
    int col0_zero = 1;
#pragma unroll 1
    for (int i = 0; i < M; i++) {
        const double * Tn = (i == 0) ? T0 : a.tri_xy + 6 * (size_t)ids[(size_t)(i - 1) * CT];
        const double v0[2] = {Tn[0], Tn[1]}, v1[2] = {Tn[2], Tn[3]}, v2[2] = {Tn[4], Tn[5]};
        const double Jn[4] = {v1[0] - v0[0], v2[0] - v0[0], v1[1] - v0[1], v2[1] - v0[1]};
        double pa[2] = {v0[0] - o[0], v0[1] - o[1]}, pb[2] = {v1[0] - o[0], v1[1] - o[1]}, pc[2] = {v2[0] - o[0], v2[1] - o[1]};
        m2_apply(Ji, pa, pa); m2_apply(Ji, pb, pb); m2_apply(Ji, pc, pc);
        const double area = 0.5 * fabs(pa[0] * (pb[1] - pc[1]) + pb[0] * (pc[1] - pa[1]) + pc[0] * (pa[1] - pb[1]));
        at[i] = area;
        double px[ORDER + 1][7], py[ORDER + 1][7];
        for (int q = 0; q < a.nq; q++) {
            double x[2] = {a.qc_xy[2 * q], a.qc_xy[2 * q + 1]};
            m2_apply(Jn, x, x);
            x[0] += v0[0]; x[1] += v0[1];
            x[0] -= o[0]; x[1] -= o[1];
            m2_apply(Ji, x, x);
#pragma unroll
            for (int d = 0; d <= ORDER; d++) { px[d][q] = legendre0<ORDER>(d, x[0]); py[d][q] = legendre0<ORDER>(d, x[1]); }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            double sum = 0.0;
            for (int q = 0; q < a.nq; q++) sum += a.qc_w[q] * (px[tb_ex(k)][q] * py[tb_ey(k)][q]);
            double val = sum * area;
            val -= area * a.psi_bar[k];
            if (k == 0) { if (val > 1.0e-12) col0_zero = 0; }
            else if (i > 0) B[(i - 1) * KR + (k - 1)] = val;
        }
    }
    if (!col0_zero) { atomicExch(a.err_flag, 1); return; }   // dense path of :676-706 has no compact table: the host reports it
    if (s == 0) a.fm_area0[cell] = at[0];

    // ---- R factor by Householder reflections (preprocess.cpp householder_R)
#pragma unroll 1
    for (int e = 0; e < MC * KR; e++) R[e * TB_THREADS] = B[e];
#pragma unroll 1
    for (int j = 0; j < KR; j++) {
        double nrm = 0.0;
        for (int i = j; i < MC; i++) { const double r = R_(i, j); nrm += r * r; }
        nrm = sqrt(nrm);
        if (nrm < 1.0e-15) continue;
        const double sgn = (R_(j, j) >= 0.0) ? 1.0 : -1.0;
        const double alpha = -sgn * nrm;
        const int len = MC - j;
        double nu = 0.0;
        for (int k = 0; k < len; k++) {
            double x = R_(j + k, j);
            if (k == 0) x -= alpha;
            V_(k) = x;
            nu += x * x;
        }
        nu = sqrt(nu);
        for (int k = 0; k < len; k++) V_(k) = V_(k) / nu;
#pragma unroll 1
        for (int c = j; c < KR; c++) {
            for (int i = 0; i < len; i++) {
                const double tv = 2.0 * V_(i);
                double sacc = 0.0;
                for (int k = 0; k < len; k++) {
                    const double q = ((i == k) ? 1.0 : 0.0) - tv * V_(k);
                    sacc += q * R_(j + k, c);
                }
                C_(i) = sacc;
            }
            for (int i = 0; i < len; i++) R_(j + i, c) = C_(i);
        }
    }

    // ---- R^T Y = B^T (forward), R X = Y (backward, in place over Y): X = pseudo-inverse, KR x MC (pseudo_inverse_from_R)
    double Y[KR * MC];
#pragma unroll 1
    for (int i = 0; i < KR; i++) {
        const double rii = R_(i, i);
        for (int j = 0; j < MC; j++) {
            double sacc = 0.0;
            for (int k = 0; k < i; k++) sacc += R_(k, i) * Y[k * MC + j];
            Y[i * MC + j] = (B[j * KR + i] - sacc) / rii;
        }
    }
#pragma unroll 1
    for (int i = KR - 1; i >= 0; i--) {
        const double rii = R_(i, i);
        for (int j = 0; j < MC; j++) {
            double sacc = 0.0;
            for (int k = i + 1; k < KR; k++) sacc += R_(i, k) * Y[k * MC + j];
            Y[i * MC + j] = (Y[i * MC + j] - sacc) / rii;
        }
    }

    // ---- compact rows: A+[k][m] * area_t[m] for k = 1..K-1, m = 1..M-1
#pragma unroll 1
    for (int k = 0; k < KR; k++) {
        double * row = out + (size_t)k * FROW;
        for (int p = 0; p < NP; p++)
            reinterpret_cast<double2 *>(row)[(size_t)p * CT + fl] = make_double2(Y[k * MC + 2 * p] * at[2 * p + 1], Y[k * MC + 2 * p + 1] * at[2 * p + 2]);
        row[(size_t)2 * NP * CT + fl] = Y[k * MC + MC - 1] * at[MC];
    }
#undef R_
#undef V_
#undef C_
}

template <int ORDER>
void launch_t(const TableBuildArgs & a, cudaStream_t st) {
    constexpr int K = (ORDER + 1) * (ORDER + 2) / 2, M = 2 * K, KR = K - 1, MC = M - 1, TB_THREADS = TbCfg<ORDER>::THREADS;
    const size_t smem = ((size_t)MC * KR + 2 * MC) * TB_THREADS * sizeof(double);
    cudaFuncSetAttribute(teno_tables_kernel<ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t groups = (uint64_t)a.n_ftiles * FAST_S;
    const unsigned grid = (unsigned)((groups * FAST_CT + TB_THREADS - 1) / TB_THREADS);
    if (grid) teno_tables_kernel<ORDER><<<grid, TB_THREADS, smem, st>>>(a);
}

}  // namespace

bool teno_tables_device_supported(int order, int basis, int nq) { return basis == MLB_BASIS_LEGENDRE && order >= 1 && order <= 4 && nq <= 7; }

void launch_teno_tables(const TableBuildArgs & a, cudaStream_t st) {
    switch (a.order) {
        case 1: launch_t<1>(a, st); break;
        case 2: launch_t<2>(a, st); break;
        case 3: launch_t<3>(a, st); break;
        case 4: launch_t<4>(a, st); break;
        default: break;
    }
}

}  // namespace mlb
