// Layout / bookkeeping kernels that do no floating-point arithmetic of the path (or a single IEEE operation), shared by
// both floating-point modes: AoS(reference numbering) <-> SoA(library numbering) conversion, halo pack/unpack, dt.
#include <cfloat>
#include <map>
#include <mutex>

#include "kernel_args.h"

namespace mlb {

// ---- per-(kernel, device) launch configuration (see kernel_args.h)
namespace {
std::mutex cfg_mutex;
std::map<std::pair<const void *, int>, int> cfg_ctas;      // persistent grid size
std::map<std::pair<const void *, int>, size_t> cfg_smem;   // largest dynamic shared-memory size opted into
void cfg_check(cudaError_t e, const char * what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}
void opt_in_locked(const void * kernel, int dev, size_t smem) {
    size_t & have = cfg_smem[{kernel, dev}];
    if (smem <= have) return;
    cfg_check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    have = smem;
}
}  // namespace

void ensure_dynamic_smem(const void * kernel, size_t smem) {
    int dev = 0;
    cfg_check(cudaGetDevice(&dev), "cudaGetDevice");
    std::lock_guard<std::mutex> lock(cfg_mutex);
    opt_in_locked(kernel, dev, smem);
}

int persistent_ctas(const void * kernel, int threads, size_t smem) {
    int dev = 0;
    cfg_check(cudaGetDevice(&dev), "cudaGetDevice");
    std::lock_guard<std::mutex> lock(cfg_mutex);
    const auto key = std::make_pair(kernel, dev);
    const auto it = cfg_ctas.find(key);
    if (it != cfg_ctas.end()) return it->second;
    opt_in_locked(kernel, dev, smem);
    int sms = 0, per_sm = 0;
    cfg_check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute(MultiProcessorCount)");
    cfg_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (sms < 1 || per_sm < 1) throw std::runtime_error("streaming kernel does not fit on this device (shared memory / registers)");
    return cfg_ctas[key] = sms * per_sm;
}

namespace {

// dst[i][v] (nv == 4: states and residuals are AoS on the device) or dst[v][i] (primitives: SoA) = src[perm[i]][v];
// perm == nullptr: identity
__global__ void import_kernel(const double * __restrict__ aos, const uint32_t * __restrict__ perm, uint32_t n, uint32_t npad, int nv,
                              double * __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t src = (size_t)(perm ? perm[i] : i) * nv;
    if (nv == 4) for (int v = 0; v < 4; v++) dst[4 * (size_t)i + v] = aos[src + v];
    else for (int v = 0; v < nv; v++) dst[(size_t)v * npad + i] = aos[src + v];
}

// the inverse
__global__ void export_kernel(const double * __restrict__ srcdev, const uint32_t * __restrict__ perm, uint32_t n, uint32_t npad, int nv,
                              double * __restrict__ aos) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t dst = (size_t)(perm ? perm[i] : i) * nv;
    if (nv == 4) for (int v = 0; v < 4; v++) aos[dst + v] = srcdev[4 * (size_t)i + v];
    else for (int v = 0; v < nv; v++) aos[dst + v] = srcdev[(size_t)v * npad + i];
}

// prim[5][i] = U[i][0]: the density plane the CFL kernel reads next to the reference's five primitives
__global__ void rho_plane_kernel(const double * __restrict__ U, uint32_t n, uint32_t npad, double * __restrict__ prim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) prim[5 * (size_t)npad + i] = U[4 * (size_t)i];
}

// out[perm[i]] = scal[which] * v[i]   — KokkosBlas::scal(cfl_local, dt, cfl_local), solver/solver.cpp:584
__global__ void export_scaled_kernel(const double * __restrict__ v, const double * __restrict__ scal, int which,
                                     const uint32_t * __restrict__ perm, uint32_t n, double * __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = v[i];                               // 0 until the first calc_dt: the reference's cfl_local is +0 then
    out[perm[i]] = r == 0.0 ? 0.0 : scal[which] * r;
}

// buf[k][v] = soa[v][idx[k]]  /  soa[v][idx[k]] = buf[k][v]   (halo pack / unpack, 4 conserved variables)
__global__ void gather_kernel(const double * __restrict__ soa, const uint32_t * __restrict__ idx, uint32_t n, uint32_t npad,
                              double * __restrict__ buf) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 4u * n) return;
    const uint32_t k = t >> 2, v = t & 3u;
    buf[t] = soa[4 * (size_t)idx[k] + v];
}
__global__ void scatter_kernel(const double * __restrict__ buf, const uint32_t * __restrict__ idx, uint32_t n, uint32_t npad,
                               double * __restrict__ soa) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 4u * n) return;
    const uint32_t k = t >> 2, v = t & 3u;
    soa[4 * (size_t)idx[k] + v] = buf[t];
}

__global__ void apply_dt_kernel(double * scal, long long * max_bits, double cfl, double global_max, int use_global) {
    const double mx = use_global ? global_max : scal[SC_MAX_SR];
    scal[SC_MAX_SR] = mx;
    scal[SC_DT] = cfl / mx;       // Solver::calc_dt solver/solver.cpp:583
    scal[SC_CFL] = cfl;
    (void)max_bits;
}

__global__ void set_scalar_kernel(double * scal, int which, double v) { scal[which] = v; }

// Face values in the reference layout F[f_ref][q][side][v] from the cell-centred storage (or from U for first order)
__global__ void export_faces_kernel(const double * __restrict__ Fc, const double * __restrict__ U, const uint32_t * __restrict__ slot_face,
                                    const uint32_t * __restrict__ perm_faces, uint32_t n, uint32_t npad, int n_slots, int Q,
                                    double * __restrict__ F) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < n_slots; j++) {
        const uint32_t fcode = slot_face[(size_t)j * npad + i];
        if (fcode == NO_FACE) continue;
        const uint32_t f = perm_faces[fcode & 0x7FFFFFFFu], side = fcode >> 31;
        for (int q = 0; q < Q; q++)
            for (int v = 0; v < 4; v++)
                F[(((size_t)f * Q + q) * 2 + side) * 4 + v] = Fc ? Fc[((size_t)i * (n_slots * Q) + (j * Q + q)) * 4 + v] : U[4 * (size_t)i + v];
    }
}

// Output fields as the reference's writer sees them (solver.cpp:336-350: Data views of conservatives, primitives, cfl_local),
// one plane per requested variable in REFERENCE numbering: out[v][perm[i]].  code 0-3 conserved, 4-8 primitives, 9 CFL.
struct FieldCodes { int32_t n; int32_t code[16]; };
__device__ __forceinline__ double field_value(int code, uint32_t i, uint32_t npad, const double * U, const double * prim, const double * sr,
                                              const double * scal) {
    if (code < 4) return U[4 * (size_t)i + code];
    if (code < 9) return prim[(size_t)(code - 4) * npad + i];
    const double r = sr[i];                              // 0 until the first calc_dt: the reference's cfl_local is +0 then
    return r == 0.0 ? 0.0 : scal[SC_DT] * r;             // KokkosBlas::scal(cfl_local, dt, cfl_local), solver.cpp:584
}
__global__ void export_fields_kernel(FieldCodes fc, const double * __restrict__ U, const double * __restrict__ prim,
                                     const double * __restrict__ sr, const double * __restrict__ scal, const uint32_t * __restrict__ perm,
                                     uint32_t n, uint32_t npad, uint32_t n_ref, double * __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = perm ? perm[i] : i;
    for (int v = 0; v < fc.n; v++) out[(size_t)v * n_ref + r] = field_value(fc.code[v], i, npad, U, prim, sr, scal);
}

// max_array / min_array of Solver::do_checks (solver.cpp:434-437, common_math.h:577-614) and the NaN count of check_fields
// (solver.cpp:470-498) in one pass over the device-resident fields: 9 fields x {min, max}.  A comparison with NaN is false
// on both sides, exactly as `if (a > max)` / `if (a < min)` in the reference's reducers.
__global__ void __launch_bounds__(256) field_ranges_kernel(const double * __restrict__ U, const double * __restrict__ prim, uint32_t n, uint32_t npad,
                                                           double * __restrict__ partial /* [grid][18] */, unsigned long long * __restrict__ nan_count) {
    __shared__ double red[8][18];
    double mn[9], mx[9];
#pragma unroll
    for (int f = 0; f < 9; f++) { mn[f] = DBL_MAX; mx[f] = -DBL_MAX; }
    unsigned nan = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int f = 0; f < 9; f++) {
            const double v = f < 4 ? U[4 * (size_t)i + f] : prim[(size_t)(f - 4) * npad + i];
            if (v < mn[f]) mn[f] = v;
            if (v > mx[f]) mx[f] = v;
            nan += (v != v);
        }
    }
#pragma unroll
    for (int f = 0; f < 9; f++)
        for (int o = 16; o > 0; o >>= 1) {
            const double a = __shfl_xor_sync(0xffffffffu, mn[f], o), b = __shfl_xor_sync(0xffffffffu, mx[f], o);
            if (a < mn[f]) mn[f] = a;
            if (b > mx[f]) mx[f] = b;
        }
    for (int o = 16; o > 0; o >>= 1) nan += __shfl_xor_sync(0xffffffffu, nan, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        for (int f = 0; f < 9; f++) { red[w][f] = mn[f]; red[w][9 + f] = mx[f]; }
        if (nan) atomicAdd(nan_count, (unsigned long long)nan);
    }
    __syncthreads();
    if (threadIdx.x < 18) {
        double v = red[0][threadIdx.x];
        for (int k = 1; k < 8; k++) {
            const double o = red[k][threadIdx.x];
            if (threadIdx.x < 9 ? o < v : o > v) v = o;
        }
        partial[(size_t)blockIdx.x * 18 + threadIdx.x] = v;
    }
}
__global__ void field_ranges_finish_kernel(const double * __restrict__ partial, int n_blocks, double * __restrict__ out /* [18] */) {
    const int t = threadIdx.x;
    if (t >= 18) return;
    double v = partial[t];
    for (int k = 1; k < n_blocks; k++) {
        const double o = partial[(size_t)k * 18 + t];
        if (t < 9 ? o < v : o > v) v = o;
    }
    out[t] = v;
}

inline unsigned blocks(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_export_fields(int n_fields, const int32_t * codes, const double * U, const double * prim, const double * sr, const double * scal,
                          const uint32_t * perm, uint32_t n, uint32_t npad, uint32_t n_ref, double * out, cudaStream_t st) {
    FieldCodes fc{};
    fc.n = n_fields;
    for (int i = 0; i < n_fields && i < 16; i++) fc.code[i] = codes[i];
    if (n) export_fields_kernel<<<blocks(n, 256), 256, 0, st>>>(fc, U, prim, sr, scal, perm, n, npad, n_ref, out);
}
int field_ranges_blocks(uint32_t n) { const unsigned b = blocks(n, 256); return (int)(b < 592u ? (b ? b : 1u) : 592u); }
void launch_field_ranges(const double * U, const double * prim, uint32_t n, uint32_t npad, double * partial, double * out18,
                         unsigned long long * nan_count, cudaStream_t st) {
    const int nb = field_ranges_blocks(n);
    field_ranges_kernel<<<nb, 256, 0, st>>>(U, prim, n, npad, partial, nan_count);
    field_ranges_finish_kernel<<<1, 32, 0, st>>>(partial, nb, out18);
}

void launch_import_state(const double * aos, const uint32_t * perm, uint32_t n, uint32_t npad, int nv, double * soa, cudaStream_t st) {
    if (n) import_kernel<<<blocks(n, 256), 256, 0, st>>>(aos, perm, n, npad, nv, soa);
}
void launch_export_state(const double * soa, const uint32_t * perm, uint32_t n, uint32_t npad, int nv, double * aos, cudaStream_t st) {
    if (n) export_kernel<<<blocks(n, 256), 256, 0, st>>>(soa, perm, n, npad, nv, aos);
}
void launch_rho_plane(const double * U, uint32_t n, uint32_t npad, double * prim, cudaStream_t st) {
    if (n) rho_plane_kernel<<<blocks(n, 256), 256, 0, st>>>(U, n, npad, prim);
}
void launch_export_scaled(const double * v, const double * scal, int which, const uint32_t * perm, uint32_t n, double * out, cudaStream_t st) {
    if (n) export_scaled_kernel<<<blocks(n, 256), 256, 0, st>>>(v, scal, which, perm, n, out);
}
void launch_gather(const double * soa, const uint32_t * idx, uint32_t n, uint32_t npad, double * buf, cudaStream_t st) {
    if (n) gather_kernel<<<blocks(4ull * n, 256), 256, 0, st>>>(soa, idx, n, npad, buf);
}
void launch_scatter(const double * buf, const uint32_t * idx, uint32_t n, uint32_t npad, double * soa, cudaStream_t st) {
    if (n) scatter_kernel<<<blocks(4ull * n, 256), 256, 0, st>>>(buf, idx, n, npad, soa);
}
void launch_apply_dt(double * scal, long long * max_bits, double cfl, double global_max, int use_global, cudaStream_t st) {
    apply_dt_kernel<<<1, 1, 0, st>>>(scal, max_bits, cfl, global_max, use_global);
}
void launch_set_scalar(double * scal, int which, double v, cudaStream_t st) { set_scalar_kernel<<<1, 1, 0, st>>>(scal, which, v); }
void launch_export_faces(const double * Fc, const double * U, const uint32_t * slot_face, const uint32_t * perm_faces, uint32_t n,
                         uint32_t npad, int n_slots, int Q, double * F_aos, cudaStream_t st) {
    if (n) export_faces_kernel<<<blocks(n, 256), 256, 0, st>>>(Fc, U, slot_face, perm_faces, n, npad, n_slots, Q, F_aos);
}

}  // namespace mlb
