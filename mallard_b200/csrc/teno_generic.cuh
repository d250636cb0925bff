// TENO reconstruction, generic fallback — TENOFunctor::operator() (numerics/face_reconstruction.cpp:866-1039) for every
// configuration the reference accepts that has no specialised kernel: basis_order 5..9, max_stencil_size_factor != 2 (any
// stencil size M), cells with four faces (five stencils).  Same tables, same thread mapping (thread = (cell, conserved
// variable), a 32-thread block = one 8-cell table tile) and the same operation and summation order as teno_recon_kernel, with
// everything whose size depends on (K, M, S) in shared memory instead of registers: the right-hand side b[M] and the dofs
// a[S][K] of a thread.  Basis functions come from the coefficient table of basis.h:81-176 evaluated in the reference's
// expression shape (coefficient, then repeated multiplication by x, terms added left to right).  Not a fast path: it exists
// so that no reference-valid TOML is refused; included in both floating-point namespaces.
#pragma once

namespace generic {

constexpr int GTHREADS = 32;
constexpr int GMAX_ORDER = 9;

struct LegTerm { double c; int e; };
struct LegPoly { double scale; int nt; LegTerm t[5]; };
__constant__ LegPoly LEG[GMAX_ORDER + 1] = {
    {1.0, 1, {{1.0, 0}}},
    {1.0, 1, {{1.0, 1}}},
    {0.5, 2, {{3.0, 2}, {-1.0, 0}}},
    {0.5, 2, {{5.0, 3}, {-3.0, 1}}},
    {0.125, 3, {{35.0, 4}, {-30.0, 2}, {3.0, 0}}},
    {0.125, 3, {{63.0, 5}, {-70.0, 3}, {15.0, 1}}},
    {0.0625, 4, {{231.0, 6}, {-315.0, 4}, {105.0, 2}, {-5.0, 0}}},
    {0.0625, 4, {{429.0, 7}, {-693.0, 5}, {315.0, 3}, {-35.0, 1}}},
    {0.0078125, 5, {{6435.0, 8}, {-12012.0, 6}, {6930.0, 4}, {-1260.0, 2}, {35.0, 0}}},
    {0.0078125, 5, {{12155.0, 9}, {-25740.0, 7}, {18018.0, 5}, {-4620.0, 3}, {315.0, 1}}},
};

__device__ __forceinline__ double basis_1d(int basis, int p, double x) {
    if (basis != MLB_BASIS_LEGENDRE) {
#ifdef MLB_STREAM_KERNELS
        double r = 1.0;
        for (int d = 0; d < p; d++) r *= x;
        return r;
#else
        return pow(x, (double)p);          // Kokkos::pow(x, p), basis.h:66-70
#endif
    }
    const LegPoly & P = LEG[p];
    double acc = 0.0;
    for (int j = 0; j < P.nt; j++) {
        double term = P.t[j].c;
        for (int k = 0; k < P.t[j].e; k++) term *= x;
        acc = j == 0 ? term : acc + term;
    }
    return P.scale * acc;
}

// dynamic shared memory: b[Mp][32] | dof[S][K][32]
__global__ void __launch_bounds__(GTHREADS) teno_generic_kernel(const __grid_constant__ ReconArgs a) {
    MLB_DYNAMIC_SMEM(double, gsm);
    const int tid = threadIdx.x;
    const uint32_t cell = blockIdx.x * (GTHREADS / 4) + (tid >> 2);
    const int var = tid & 3;
    if (cell >= a.g.N_recon) return;
    const int K = a.K, M = a.M, MP = a.Mp, S = a.S, order = a.order;
    double * b = gsm + tid;                                   // b[m * 32]
    double * dof = gsm + (size_t)MP * GTHREADS + tid;          // dof[(s * K + k) * 32]
    const uint32_t Np = a.g.Npad;
    const size_t tile = cell / TILE;
    const int lane = cell % TILE;
    const double * Uv = a.Uin + var;
    const double u_self = Uv[4 * (size_t)cell];
    const double * OI = a.OI_dev;

    double w[1 + MAX_SLOTS];
    double area0 = 0.5;
    for (int s = 0; s < S; s++) {
        const size_t sbase = (tile * S + s) * MP;
        const uint32_t * ids = a.st_ids + sbase * TILE + lane;
        if (ids[0] == NO_FACE) { w[s] = 0.0; continue; }                 // empty stencil :896-899
        const double * areas = a.st_area + sbase * TILE + lane;
        for (int m = 0; m < MP; m++) b[m * GTHREADS] = areas[m * TILE] * (Uv[4 * (size_t)ids[m * TILE]] - u_self);   // :903-910
        if (s == 0) area0 = areas[0];
        const double2 * mat = reinterpret_cast<const double2 *>(a.st_mat) + ((tile * S + s) * K * (MP / 2)) * TILE + lane;
        for (int k = 0; k < K; k++) {                                    // a = A+ b, sums in ascending m :915-918
            double sum = 0.0;
            for (int m2 = 0; m2 < MP / 2; m2++) {
                const double2 c = mat[((size_t)k * (MP / 2) + m2) * TILE];
                sum += c.x * b[(2 * m2) * GTHREADS];
                sum += c.y * b[(2 * m2 + 1) * GTHREADS];
            }
            dof[(size_t)(s * K + k) * GTHREADS] = sum;
        }
        double si = 0.0;                                                 // SI = a . (OI a) :922-936
        for (int k = 0; k < K; k++) {
            double t = 0.0;
            for (int j = 0; j < K; j++) t += OI[k * K + j] * dof[(size_t)(s * K + j) * GTHREADS];
            b[k * GTHREADS] = t;                                         // (OI a)_k parked in the right-hand side's storage (K <= M)
        }
        for (int k = 0; k < K; k++) si += dof[(size_t)(s * K + k) * GTHREADS] * b[k * GTHREADS];
        const double x = si + 1.0e-12;                                   // 1/(SI+eps)^6 :940-944
        const double x2 = x * x, x3 = x2 * x;
        w[s] = 1.0 / (x3 * x3);
    }
    (void)M;

    {   // non-linear weights :948-981 (reference-faithful: the central weight stays raw in the ENO branch)
        double sd = 0.0;
        for (int s = 1; s < S; s++) sd += w[s];
        if (w[0] / (sd + w[0]) > 1.0e-7) {
            w[0] = 1.0;
            for (int s = 1; s < S; s++) w[s] = 0.0;
        } else {
            for (int s = 1; s < S; s++) {
                if (w[s] / sd > 1.0e-5) w[s] = (1.0 / K);
                else if (a.fixed_weights) w[s] = 0.0;
            }
            sd = 0.0;
            for (int s = 1; s < S; s++) sd += w[s];
            for (int s = 1; s < S; s++) w[s] /= sd;
            if (a.fixed_weights) w[0] = 0.0;
        }
    }

    const int nf = a.g.nfc[cell];
    const int Q = a.g.Q;
    for (int j = 0; j < nf; j++) {                                       // :985-1034
        const double * fx = a.g.slot_fx + ((size_t)j * 4) * Np + cell;
        const double x0 = fx[0], y0 = fx[Np], x1 = fx[2 * (size_t)Np], y1 = fx[3 * (size_t)Np];
        for (int q = 0; q < Q; q++) {
            const double tq = (a.qf_x[q] + 1.0) * 0.5;
            const double xq = tq * (x1 - x0) + x0, yq = tq * (y1 - y0) + y0;
            double Px[GMAX_ORDER + 1], Py[GMAX_ORDER + 1];
            for (int d = 0; d <= order; d++) { Px[d] = basis_1d(a.basis, d, xq); Py[d] = basis_1d(a.basis, d, yq); }
            double out = u_self;
            for (int s = 0; s < S; s++) {
                if (w[s] == 0.0) continue;
                for (int k = 0; k < K; k++) {
                    const double pb = a.psi_bar_cell ? a.psi_bar_cell[(size_t)cell * K + k] : a.psi_bar_dev[k];
                    const double cbar = a.fixed_weights ? -pb : pb / area0;   // psi_bar_k / area_t[s][0] :1028-1029
                    out += w[s] * dof[(size_t)(s * K + k) * GTHREADS] * (Px[a.pidx_dev[2 * k]] * Py[a.pidx_dev[2 * k + 1]] + cbar);
                }
            }
            a.Fc[((size_t)cell * (a.g.n_slots * Q) + (j * Q + q)) * 4 + var] = out;
        }
    }
}

#ifndef MLB_HOST_EMULATION
static void launch_generic(const ReconArgs & a, cudaStream_t st) {
    const size_t smem = ((size_t)a.Mp + (size_t)a.S * a.K) * GTHREADS * sizeof(double);
    ensure_dynamic_smem(reinterpret_cast<const void *>(teno_generic_kernel), smem);
    const unsigned cells_per_block = GTHREADS / 4;
    const unsigned grid = (a.g.N_recon + cells_per_block - 1) / cells_per_block;
    if (grid) teno_generic_kernel<<<grid, GTHREADS, smem, st>>>(a);
}
#endif

static bool generic_supported(int order, int K, int Mp, int S) {
    return order >= 1 && order <= GMAX_ORDER && K <= Mp && S <= 1 + MAX_SLOTS &&
           ((size_t)Mp + (size_t)S * K) * GTHREADS * sizeof(double) <= 200 * 1024;
}

}  // namespace generic
