// Whole time steps of a SMALL first-order mesh in ONE cooperative kernel (new).  examples/sod (1000 cells) and
// examples/wedge (7500 cells) are launch-bound: CFL kernel + 3 x (face fluxes, residual gather / RK update) = 7 kernels of 1-3 us of
// work each, 48 / 58 us per step even as a replayed CUDA graph.  Here the grid is resident for the whole run (cooperative launch) and
// the phases of a step are separated by grid-wide barriers instead of kernel boundaries; n_steps steps are ONE launch.
//
// The phases call the very bodies the stand-alone kernels are made of (spectral_radius_body, face_flux_body, gather_stage_body:
// kernels_impl.cuh) on the same argument blocks, so the arithmetic - and in STRICT mode every bit of the result - is that of the
// multi-kernel path; what differs is who calls them: a grid-stride loop instead of one thread per face / cell.
//   phase 0          (cfl > 0 only) spectral radius of every owned cell, running maximum (warp shuffle + one atomic per warp)
//   phase 1 + 2 s    face fluxes of stage s (thread 0 first publishes dt = cfl / max of phase 0, when s == 0)
//   phase 2 + 2 s    residual gather + RK update of stage s (the last stage refreshes the primitives, t and the step counter)
// Restrictions (checked by the caller, api.cu): first-order reconstruction, no viscous terms, one GPU, SSPRK3 / RK4 (their stage
// buffers return to the same rotation after a step).  tests/emul runs the phase function on the host against the oracle.
#pragma once

namespace small {

constexpr int SS_THREADS = 256;

template <int RS>
__device__ __forceinline__ void small_step_phase(const SmallStepArgs & p, const int phase, const uint32_t tid, const uint32_t nthreads) {
    if (phase == 0) {
        double m = -1.0;
        for (uint32_t i = tid; i < p.cfl.g.N_owned; i += nthreads) {
            const double sr = spectral_radius_body(p.cfl, i);
            if (sr == sr) m = fmax(m, sr);                               // NaN never wins (Kokkos::Max joins with `<`), as in cfl_kernel
        }
#ifdef MLB_HOST_EMULATION
        if (m > *reinterpret_cast<double *>(p.cfl.max_bits)) *reinterpret_cast<double *>(p.cfl.max_bits) = m;
#else
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) atomicMax(p.cfl.max_bits, __double_as_longlong(m));   // all candidates are >= 0 or exactly -1: integer order = value order
#endif
        return;
    }
    const int s = (phase - 1) >> 1;
    const StageArgs & a = p.st[s];
    if ((phase & 1) != 0) {
        if (s == 0 && tid == 0 && p.cfl.cfl > 0.0) {                     // Solver::calc_dt solver/solver.cpp:583; read by phase 2 (a grid barrier later)
#ifdef MLB_HOST_EMULATION
            const double mx = *reinterpret_cast<double *>(p.cfl.max_bits);
            *reinterpret_cast<double *>(p.cfl.max_bits) = -1.0;
#else
            const double mx = __longlong_as_double((long long)atomicAdd(reinterpret_cast<unsigned long long *>(p.cfl.max_bits), 0ull));   // as cfl_kernel reads it
            *p.cfl.max_bits = __double_as_longlong(-1.0);
#endif
            p.cfl.scal[SC_MAX_SR] = mx;
            p.cfl.scal[SC_DT] = p.cfl.cfl / mx;
            p.cfl.scal[SC_CFL] = p.cfl.cfl;
        }
        // whole warps walk the faces: face_flux_body combines its quadrature lanes with __shfl_sync(full mask) (a width-1 shuffle on this
        // first-order path, but still a warp-synchronous instruction), so a warp stays together as long as its FIRST lane has a face;
        // lanes past the last face recompute it and do not store (the body's `valid`), exactly as in the stand-alone kernel's last block
        for (uint32_t f = tid; f - (tid & 31u) < a.g.NF; f += nthreads) face_flux_body<RS, false, 1, false>(a, f);
    } else {
        for (uint32_t i = tid; i < a.g.N_owned; i += nthreads) gather_stage_body(a, i);
    }
}

#ifndef MLB_HOST_EMULATION
template <int RS>
__global__ void __launch_bounds__(SS_THREADS) small_step_kernel(const __grid_constant__ SmallStepArgs p) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int first = p.cfl.cfl > 0.0 ? 0 : 1, n_phases = 1 + 2 * p.n_stages;
    for (uint32_t step = 0; step < p.n_steps; step++)
        for (int ph = first; ph < n_phases; ph++) {
            small_step_phase<RS>(p, ph, tid, nthreads);
            grid.sync();
        }
}

template <int RS>
static void launch_small_step_rs(const SmallStepArgs & p, int max_blocks, cudaStream_t st) {
    const void * fn = reinterpret_cast<const void *>(small_step_kernel<RS>);
    const int resident = persistent_ctas(fn, SS_THREADS, 0);               // SMs x co-resident blocks per SM on the current device
    const uint32_t work = p.st[0].g.NF > p.st[0].g.N_owned ? p.st[0].g.NF : p.st[0].g.N_owned;
    int blocks = (int)((work + SS_THREADS - 1) / SS_THREADS);
    if (blocks > resident) blocks = resident;
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    void * args[] = {const_cast<SmallStepArgs *>(&p)};
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(SS_THREADS), args, 0, st);
    if (e != cudaSuccess) throw std::runtime_error(std::string("small_step_kernel: cooperative launch failed: ") + cudaGetErrorString(e));
}
static void launch_small_step(const SmallStepArgs & p, int max_blocks, cudaStream_t st) {
    switch (p.st[0].ph.riemann) {
        case MLB_RIEMANN_RUSANOV: launch_small_step_rs<MLB_RIEMANN_RUSANOV>(p, max_blocks, st); break;
        case MLB_RIEMANN_HLL: launch_small_step_rs<MLB_RIEMANN_HLL>(p, max_blocks, st); break;
        default: launch_small_step_rs<MLB_RIEMANN_HLLC>(p, max_blocks, st); break;
    }
}
#endif

}  // namespace small
