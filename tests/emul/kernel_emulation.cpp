// TEST INFRASTRUCTURE ONLY - never linked into, loaded by or called from the product path (mallard_b200/ has no CPU fallback).
//
// Host emulation of the CUDA kernel SOURCE.  mallard_b200/csrc/kernels_impl.cuh - the very file nvcc compiles for sm_100a - is
// compiled here for the host with the CUDA keywords defined away and threadIdx / blockIdx as plain variables, and its kernels
// are executed thread by thread in a loop.  What can be emulated this way is every kernel whose threads do not cooperate:
//   teno_recon_kernel<ORDER, MP>   (specialised TENO reconstruction, per-thread slices of shared memory)
//   generic::teno_generic_kernel   (orders 1-9, any stencil size, up to five stencils: quadrilaterals)
//   visc_grad_kernel               (least-squares gradients)
//   face_flux_kernel<RS, TENO, QT = 0 | 1, VISC>   (QT = 0: run-time quadrature loop, no warp shuffles)
//   gather_stage_kernel            (residual gather + RK update; used here in its "bare residual" mode)
// Not emulated: the PTX streaming kernels (TMA, mbarriers) and the CFL kernel (block reduction, atomics).
// Compiled with -ffp-contract=off this is the STRICT instantiation: on triangles its results must equal the oracle's bit for
// bit wherever pow() is not involved.  Purpose: (i) a check of kernels that have not run on a B200 yet against the oracle and
// against analytic solutions, (ii) a second, independent execution of the kernel source for the ones that have.
// The host side (mesh preprocessor) is the product's own: this file links against libmallard_b200.so for mlb::preprocess.
#define MLB_HOST_EMULATION 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
static EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __constant__ static const
struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
static inline double4 make_double4(double x, double y, double z, double w) { return {x, y, z, w}; }
static inline double __shfl_sync(unsigned, double v, int, int width = 32) { if (width != 1) abort(); return v; }   // single-lane groups only
typedef void * cudaStream_t;
static std::vector<double> emu_smem;
#define MLB_DYNAMIC_SMEM(type, name) type * name = reinterpret_cast<type *>(emu_smem.data())
using std::fmax; using std::fmin; using std::sqrt; using std::pow; using std::fabs; using std::fma;

#define MLB_KNS emu
#include "../../mallard_b200/csrc/kernels_impl.cuh"

using namespace mlb;

namespace {

thread_local std::string emu_err;

template <class K, class A> void run_kernel(K kernel, const A & args, unsigned grid, unsigned block) {
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < block; t++) { blockIdx.x = b; threadIdx.x = t; kernel(args); }
}

struct Emu {
    Prep P;
    mlb_numerics num{};
    GasParams gas{};
    DevPhys phys{};
    DevGeom g{};
    uint32_t nc_ref = 0, nf_ref = 0;
    bool teno = false;
    std::vector<double> U, k0, Fc, AF, G, bnd_s;
    int force_generic = 0;
};

ReconArgs recon_args(Emu & e) {
    ReconArgs r{};
    const TenoTables & T = e.P.teno;
    r.g = e.g; r.Uin = e.U.data(); r.Fc = e.Fc.data(); r.st_ids = T.st_ids.data(); r.st_area = T.st_area.data(); r.st_mat = T.st_mat.data();
    r.order = T.order; r.K = T.K; r.M = T.M; r.Mp = T.Mp; r.S = T.S; r.basis = T.basis; r.fixed_weights = e.num.teno_fixed;
    for (size_t i = 0; i < e.P.qf_x.size(); i++) r.qf_x[i] = e.P.qf_x[i];
    if (T.K <= 15) {
        for (int i = 0; i < T.K; i++) { r.psi_bar[i] = T.psi_bar[i]; r.pidx[2 * i] = T.pidx[2 * i]; r.pidx[2 * i + 1] = T.pidx[2 * i + 1]; }
        for (int i = 0; i < T.K * T.K; i++) r.OI[i] = T.OI[i];
    }
    r.OI_dev = T.OI.data(); r.psi_bar_dev = T.psi_bar.data(); r.pidx_dev = T.pidx.data();
    r.psi_bar_cell = T.mixed ? T.psi_bar_cell.data() : nullptr;
    return r;
}

template <int ORDER, int MP> void run_recon_t(const ReconArgs & a) {
    constexpr int K = (ORDER + 1) * (ORDER + 2) / 2;
    emu_smem.assign((size_t)(1 + MAX_SLOTS) * K * emu::RECON_THREADS, 0.0);
    run_kernel(emu::teno_recon_kernel<ORDER, MP>, a, (a.g.N_recon + emu::RECON_THREADS / 4 - 1) / (emu::RECON_THREADS / 4), emu::RECON_THREADS);
}

void run_recon(Emu & e) {
    const ReconArgs a = recon_args(e);
    const bool spec = !e.force_generic && a.S <= 1 + MAX_SLOTS &&
                      ((a.order == 1 && a.Mp == 6) || (a.order == 2 && a.Mp == 12) || (a.order == 3 && a.Mp == 20) || (a.order == 4 && a.Mp == 30));
    if (spec) {
        if (a.order == 1) run_recon_t<1, 6>(a); else if (a.order == 2) run_recon_t<2, 12>(a); else if (a.order == 3) run_recon_t<3, 20>(a); else run_recon_t<4, 30>(a);
        return;
    }
    if (!emu::generic::generic_supported(a.order, a.K, a.Mp, a.S)) throw std::runtime_error("generic kernel: configuration not supported");
    emu_smem.assign(((size_t)a.Mp + (size_t)a.S * a.K) * emu::generic::GTHREADS, 0.0);
    run_kernel(emu::generic::teno_generic_kernel, a, (a.g.N_recon + 7) / 8, emu::generic::GTHREADS);
}

StageArgs stage_args(Emu & e) {
    StageArgs a{};
    a.g = e.g; a.ph = e.phys; a.Uin = e.U.data(); a.Fc = e.Fc.data(); a.AF = e.AF.data(); a.teno = e.teno ? 1 : 0;
    a.G = e.gas.mu > 0.0 ? e.G.data() : nullptr;
    a.rk.mode = 3; a.rk.k_store = e.k0.data();
    return a;
}

template <int RS> void run_faces(Emu & e, const StageArgs & a) {
    const unsigned grid = (a.g.NF + 127u) / 128u;
    const bool visc = a.G != nullptr;
    if (!e.teno) { if (visc) run_kernel(emu::face_flux_kernel<RS, false, 1, true>, a, grid, 128); else run_kernel(emu::face_flux_kernel<RS, false, 1, false>, a, grid, 128); }
    else { if (visc) run_kernel(emu::face_flux_kernel<RS, true, 0, true>, a, grid, 128); else run_kernel(emu::face_flux_kernel<RS, true, 0, false>, a, grid, 128); }
}

}  // namespace

extern "C" {

const char * emu_last_error() { return emu_err.c_str(); }

void * emu_create(const mlb_mesh * mesh, const mlb_numerics * num, const mlb_physics * phys, const mlb_bc * bcs, int n_bcs, const int32_t * part,
                  int rank, int n_ranks) {
    try {
        std::unique_ptr<Emu> e(new Emu());
        e->num = *num;
        e->teno = num->recon == MLB_RECON_TENO;
        e->gas = make_gas(*phys);
        HostMesh hm;
        host_mesh_from_view(hm, *mesh);
        e->nc_ref = hm.nc; e->nf_ref = hm.nf;
        std::vector<std::string> zones;
        e->phys.gas = e->gas; e->phys.riemann = num->riemann; e->phys.n_bcs = n_bcs;
        for (int b = 0; b < n_bcs; b++) {      // as create_impl (api.cu) binds the [[boundaries]] entries
            zones.push_back(bcs[b].zone_name);
            BcParams & d = e->phys.bcs[b];
            d.type = bcs[b].type;
            for (double & x : d.data) x = 0.0;
            if (d.type == MLB_BC_UPT) {
                const double rho = bcs[b].p / (e->gas.R * bcs[b].T), en = e->gas.cv * bcs[b].T;
                d.data[0] = rho; d.data[1] = bcs[b].u[0]; d.data[2] = bcs[b].u[1]; d.data[3] = bcs[b].p; d.data[4] = bcs[b].T; d.data[5] = en + bcs[b].p / rho;
            } else if (d.type == MLB_BC_P_OUT) d.data[0] = bcs[b].p;
            else if (d.type == MLB_BC_WALL_NOSLIP) { d.data[1] = bcs[b].u[0]; d.data[2] = bcs[b].u[1]; d.data[4] = bcs[b].T; }
        }
        PrepOptions opt;
        opt.renumber = num->renumber;
        opt.viscous = e->gas.mu > 0.0;
        opt.part = part; opt.rank = rank; opt.n_ranks = n_ranks;      // a rank's context of a partitioned mesh (ghosts are filled by emu_set_state)
        preprocess(hm, *num, zones, opt, e->P);
        Prep & P = e->P;
        for (size_t i = 0; i < P.qf_x.size(); i++) { e->phys.qf_x[i] = P.qf_x[i]; e->phys.qf_w[i] = P.qf_w[i]; }
        DevGeom & g = e->g;
        g.N = P.N; g.N_owned = P.N_owned; g.N_recon = P.N_recon; g.Npad = P.Npad; g.NF = P.NF; g.n_slots = P.n_slots; g.Q = P.Q; g.NFpad = P.NFpad;
        g.slot_face = P.slot_face.data(); g.slot_nbr = P.slot_nbr.data(); g.slot_nslot = P.slot_nslot.data();
        g.rhs_order = P.rhs_order.data(); g.nfc = P.n_faces_of_cell.data(); g.cell_vol = P.cell_vol.data(); g.cell_xy = P.cell_xy.data();
        g.face_nx = P.face_nx.data(); g.face_ny = P.face_ny.data(); g.face_area = P.face_area.data();
        g.slot_fx = e->teno ? P.slot_fx.data() : nullptr;
        g.face_cl = P.face_cl.data(); g.face_cr = P.face_cr.data(); g.face_slots = P.face_slots.data();
        if (opt.viscous) { g.slot_d = P.slot_d.data(); g.face_d = P.face_d.data(); e->G.assign(6 * (size_t)P.Npad, 0.0); }
        e->U.assign(4 * (size_t)P.Npad, 0.0); e->k0.assign(4 * (size_t)P.Npad, 0.0);
        e->AF.assign(4 * (size_t)std::max<uint32_t>(P.NFpad, 1), 0.0);
        if (e->teno) e->Fc.assign((size_t)P.n_slots * P.Q * 4 * P.Npad, 0.0);
        return e.release();
    } catch (const std::exception & ex) { emu_err = ex.what(); return nullptr; }
}
void emu_destroy(void * h) { delete static_cast<Emu *>(h); }
void emu_force_generic(void * h, int on) { static_cast<Emu *>(h)->force_generic = on; }
int emu_n_quad(void * h) { return static_cast<Emu *>(h)->P.Q; }

int emu_n_owned(void * h) { return (int)static_cast<Emu *>(h)->P.N_owned; }
int emu_n_held(void * h) { return (int)static_cast<Emu *>(h)->P.N; }

// every held cell - owned and ghost - takes its state from the global array: what a completed halo exchange leaves behind
int emu_set_state(void * h, const double * U_ref) {
    Emu & e = *static_cast<Emu *>(h);
    for (uint32_t i = 0; i < e.P.N; i++)
        for (int v = 0; v < 4; v++) e.U[4 * (size_t)i + v] = U_ref[4 * (size_t)e.P.perm_cells[i] + v];
    return 0;
}

// FaceReconstruction::calc_face_values: F[nf_ref][Q][2][4], zero where undefined (as mlb_calc_face_values)
int emu_face_values(void * h, double * F) {
    Emu & e = *static_cast<Emu *>(h);
    try {
        if (e.teno) run_recon(e);
        const Prep & P = e.P;
        std::fill(F, F + (size_t)e.nf_ref * P.Q * 8, 0.0);
        for (uint32_t i = 0; i < P.N_owned; i++)
            for (int j = 0; j < P.n_slots; j++) {
                const uint32_t fcode = P.slot_face[(size_t)j * P.Npad + i];
                if (fcode == NO_FACE) continue;
                const uint32_t f = P.perm_faces[fcode & 0x7FFFFFFFu], side = fcode >> 31;
                for (int q = 0; q < P.Q; q++)
                    for (int v = 0; v < 4; v++)
                        F[(((size_t)f * P.Q + q) * 2 + side) * 4 + v] = e.teno ? e.Fc[((size_t)i * (P.n_slots * P.Q) + (j * P.Q + q)) * 4 + v] : e.U[4 * (size_t)i + v];
            }
        return 0;
    } catch (const std::exception & ex) { emu_err = ex.what(); return 1; }
}

// Solver::calc_rhs on the current state: rhs[nc_ref][4]
int emu_rhs(void * h, double * rhs) {
    Emu & e = *static_cast<Emu *>(h);
    try {
        if (e.teno) run_recon(e);
        StageArgs a = stage_args(e);
        if (a.G) run_kernel(emu::visc_grad_kernel, a, (a.g.N_recon + 255u) / 256u, 256);
        switch (e.num.riemann) {
            case MLB_RIEMANN_RUSANOV: run_faces<MLB_RIEMANN_RUSANOV>(e, a); break;
            case MLB_RIEMANN_HLL: run_faces<MLB_RIEMANN_HLL>(e, a); break;
            default: run_faces<MLB_RIEMANN_HLLC>(e, a); break;
        }
        run_kernel(emu::gather_stage_kernel, a, (a.g.N_owned + 255u) / 256u, 256);
        for (uint32_t i = 0; i < e.P.N_owned; i++)
            for (int v = 0; v < 4; v++) rhs[4 * (size_t)e.P.perm_cells[i] + v] = e.k0[4 * (size_t)i + v];
        return 0;
    } catch (const std::exception & ex) { emu_err = ex.what(); return 1; }
}

// least-squares gradients [nc_ref][6] (viscous contexts)
int emu_gradients(void * h, double * G) {
    Emu & e = *static_cast<Emu *>(h);
    if (e.G.empty()) { emu_err = "context is inviscid"; return 1; }
    StageArgs a = stage_args(e);
    run_kernel(emu::visc_grad_kernel, a, (a.g.N_recon + 255u) / 256u, 256);
    for (uint32_t i = 0; i < e.P.N_owned; i++)
        for (int v = 0; v < 6; v++) G[6 * (size_t)e.P.perm_cells[i] + v] = e.G[6 * (size_t)i + v];
    return 0;
}

}  // extern "C"
