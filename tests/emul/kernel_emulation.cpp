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
#include "../../mallard_b200/csrc/stage_plan.h"

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
    // time stepping (first-order contexts): the buffers, scalars and schedule of api.cu's mlb_ctx
    std::vector<double> Ub[3], kb[4], prim, sr;
    double scal[SC_COUNT] = {-1.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned long long step_counter = 0;
    double max_val = -1.0;            // stands in for max_bits (the emulation keeps the running maximum as a double)
    unsigned int blocks_done = 0;
    int cur = 0;
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

void run_recon(Emu & e, const double * Uin = nullptr) {
    ReconArgs a = recon_args(e);
    if (Uin) a.Uin = Uin;
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

// ---- whole time steps (first order): api.cu's stage_args / cfl_args on the emulation's buffers
StageArgs step_stage_args(Emu & e, const StagePlan & s) {
    StageArgs a{};
    a.g = e.g; a.ph = e.phys; a.Uin = e.Ub[s.in].data(); a.Fc = e.teno ? e.Fc.data() : nullptr; a.AF = e.AF.data(); a.teno = e.teno ? 1 : 0;
    a.scal = e.scal; a.step_counter = &e.step_counter; a.G = e.gas.mu > 0.0 ? e.G.data() : nullptr;
    RkArgs & rk = a.rk;
    rk.mode = s.mode; rk.n_prev = s.n_prev; rk.last_stage = s.last;
    rk.base = e.Ub[s.base].data(); rk.out = e.Ub[s.out].data();
    rk.k_store = e.kb[s.kstore].data();                                  // keep_stage_rhs = true
    for (int j = 0; j < s.n_prev; j++) { rk.kprev[j] = e.kb[s.kprev[j]].data(); rk.cprev[j] = s.cprev[j]; }
    rk.c0 = s.c0; rk.c1 = s.c1; rk.coef = s.coef;
    rk.prim_out = s.last ? e.prim.data() : nullptr;
    return a;
}
CflArgs step_cfl_args(Emu & e, double cfl) {
    CflArgs a{};
    a.g = e.g; a.gas = e.gas; a.U = e.Ub[e.cur].data(); a.prim = e.prim.data(); a.sr_out = e.sr.data(); a.scal = e.scal;
    a.max_bits = reinterpret_cast<long long *>(&e.max_val); a.blocks_done = &e.blocks_done; a.cfl = cfl;
    return a;
}

void step_faces(Emu & e, const StageArgs & a) {          // run_stage's launches before the gather: gradients (viscous), face fluxes
    if (a.G) run_kernel(emu::visc_grad_kernel, a, (a.g.N_recon + 255u) / 256u, 256);
    switch (e.num.riemann) {
        case MLB_RIEMANN_RUSANOV: run_faces<MLB_RIEMANN_RUSANOV>(e, a); break;
        case MLB_RIEMANN_HLL: run_faces<MLB_RIEMANN_HLL>(e, a); break;
        default: run_faces<MLB_RIEMANN_HLLC>(e, a); break;
    }
}

// Solver::calc_dt: the CFL kernel (cell loop + max; the last block's dt = cfl / max)
double emulated_calc_dt(Emu & e, double cfl) {
    const CflArgs ca = step_cfl_args(e, cfl);
    double mx = -1.0;
    for (uint32_t i = 0; i < e.g.N_owned; i++) { const double sr = emu::spectral_radius_body(ca, i); if (sr == sr) mx = std::fmax(mx, sr); }
    e.scal[SC_MAX_SR] = mx; e.scal[SC_DT] = cfl / mx; e.scal[SC_CFL] = cfl;
    return e.scal[SC_DT];
}

// the multi-kernel sequence of mlb_run: CFL "kernel" (cell loop + max + dt), then per stage face kernel and gather kernel
void steps_kernel_by_kernel(Emu & e, uint32_t n_steps, double cfl) {
    for (uint32_t n = 0; n < n_steps; n++) {
        const auto plan = make_stage_plan(e.cur, e.num.integrator);      // (do_step, api.cu: the plan of every step starts from the current buffer)
        if (cfl > 0.0) emulated_calc_dt(e, cfl);
        for (const StagePlan & sp : plan) {
            const StageArgs a = step_stage_args(e, sp);
            if (e.teno) run_recon(e, a.Uin);
            step_faces(e, a);
            run_kernel(emu::gather_stage_kernel, a, (a.g.N_owned + 255u) / 256u, 256);
        }
        if (e.num.integrator == MLB_INTEGRATOR_FE) e.cur = plan.back().out;   // finish_plan (api.cu): forward Euler's result lives in the other buffer
    }
}

// the cooperative kernel: phases separated by grid barriers = every thread of the grid finishes a phase before the next one starts
template <int RS> void steps_small_rs(const SmallStepArgs & p, unsigned blocks) {
    const uint32_t nthreads = blocks * emu::small::SS_THREADS;
    const int first = p.cfl.cfl > 0.0 ? 0 : 1, n_phases = 1 + 2 * p.n_stages;
    for (uint32_t step = 0; step < p.n_steps; step++)
        for (int ph = first; ph < n_phases; ph++)
            for (uint32_t tid = 0; tid < nthreads; tid++) emu::small::small_step_phase<RS>(p, ph, tid, nthreads);
}
void steps_small(Emu & e, uint32_t n_steps, double cfl, unsigned blocks) {
    const auto plan = make_stage_plan(e.cur, e.num.integrator);
    SmallStepArgs p{};
    for (size_t st = 0; st < plan.size(); st++) p.st[st] = step_stage_args(e, plan[st]);
    p.cfl = step_cfl_args(e, cfl > 0.0 ? cfl : 0.0);
    p.n_stages = (int32_t)plan.size(); p.n_steps = n_steps;
    switch (e.num.riemann) {
        case MLB_RIEMANN_RUSANOV: steps_small_rs<MLB_RIEMANN_RUSANOV>(p, blocks); break;
        case MLB_RIEMANN_HLL: steps_small_rs<MLB_RIEMANN_HLL>(p, blocks); break;
        default: steps_small_rs<MLB_RIEMANN_HLLC>(p, blocks); break;
    }
}

}  // namespace

// mlb_riemann_flux (the body of riemann_kernel, kernels_impl.cuh): the flux function this build's face kernel calls - riemann_flux in the
// STRICT build, the lean re-formulation in the FAST build - on a list of face states, rows (rho, u, v, p, h)
template <int RS>
static void riemann_list(uint64_t n, const double * nunit, const double * L, const double * R, double gamma, double * flux) {
    for (uint64_t i = 0; i < n; i++) {
        const emu::FaceState l = {L[5 * i], L[5 * i + 1], L[5 * i + 2], L[5 * i + 3], L[5 * i + 4]};
        const emu::FaceState r = {R[5 * i], R[5 * i + 1], R[5 * i + 2], R[5 * i + 3], R[5 * i + 4]};
#ifdef MLB_STREAM_KERNELS
        emu::riemann_flux_lean<RS>(&flux[4 * i], nunit[2 * i], nunit[2 * i + 1], emu::face_cons(l), emu::face_cons(r), gamma);
#else
        emu::riemann_flux<RS>(&flux[4 * i], nunit[2 * i], nunit[2 * i + 1], l, r, gamma);
#endif
    }
}
extern "C" {

// n_steps time steps of a first-order context from its current state.  cfl > 0: dt from the CFL condition every step, else the fixed
// dt_fixed.  small_blocks == 0: the multi-kernel sequence of mlb_run; > 0: the cooperative kernel's phases on a grid of that many blocks.
int emu_run(void * h, unsigned n_steps, double cfl, double dt_fixed, unsigned small_blocks, double * t_out, double * dt_out) {
    Emu & e = *static_cast<Emu *>(h);
    try {
        if ((e.teno || e.gas.mu > 0.0) && small_blocks) throw std::runtime_error("emu_run: the cooperative kernel takes first-order inviscid contexts only");
        if (e.P.N != e.P.N_owned) throw std::runtime_error("emu_run: single contexts only (no halo exchange in the emulation)");
        if (e.num.integrator == MLB_INTEGRATOR_FE && small_blocks) throw std::runtime_error("emu_run: the cooperative kernel takes SSPRK3 / RK4 only (as small_step_eligible, api.cu)");
        if (!(cfl > 0.0)) e.scal[SC_DT] = dt_fixed;
        if (small_blocks) steps_small(e, n_steps, cfl, small_blocks); else steps_kernel_by_kernel(e, n_steps, cfl);
        if (t_out) *t_out = e.scal[SC_T];
        if (dt_out) *dt_out = e.scal[SC_DT];
        return 0;
    } catch (const std::exception & ex) { emu_err = ex.what(); return 1; }
}
// U[nc_ref][4] of the owned cells after emu_run (reference numbering)
int emu_get_state(void * h, double * U_ref) {
    Emu & e = *static_cast<Emu *>(h);
    for (uint32_t i = 0; i < e.P.N_owned; i++)
        for (int v = 0; v < 4; v++) U_ref[4 * (size_t)e.P.perm_cells[i] + v] = e.Ub[e.cur][4 * (size_t)i + v];
    return 0;
}
unsigned long long emu_step_count(void * h) { return static_cast<Emu *>(h)->step_counter; }
double emu_time(void * h) { return static_cast<Emu *>(h)->scal[SC_T]; }
// what mlb_get_array / mlb_get_state export after a step, reference numbering: "rhs0".."rhs3" (stage residuals, keep_stage_rhs), "U_temp"
// (the reference's solution_vec[1]: the last intermediate stage state), "cfl_local" (spectral radius x dt, [nc_ref])
int emu_get_array(void * h, const char * name, double * out) {
    Emu & e = *static_cast<Emu *>(h);
    const std::string n = name;
    const auto plan = make_stage_plan(e.cur, e.num.integrator);
    const double * src = nullptr;
    if (n.size() == 4 && n.rfind("rhs", 0) == 0 && n[3] >= '0' && n[3] < '4') src = e.kb[n[3] - '0'].data();
    else if (n == "U_temp") {
        if (plan.size() < 2) { emu_err = "emu_get_array: forward Euler has no intermediate stage state"; return 1; }
        src = e.Ub[plan[plan.size() - 2].out].data();
    }
    else if (n == "cfl_local") {
        for (uint32_t i = 0; i < e.P.N_owned; i++) { const double r = e.sr[i]; out[e.P.perm_cells[i]] = r == 0.0 ? 0.0 : e.scal[SC_DT] * r; }
        return 0;
    } else { emu_err = "emu_get_array: unknown array " + n; return 1; }
    for (uint32_t i = 0; i < e.P.N_owned; i++)
        for (int v = 0; v < 4; v++) out[4 * (size_t)e.P.perm_cells[i] + v] = src[4 * (size_t)i + v];
    return 0;
}
// Solver::calc_dt on the stepping state (sets the dt the next fixed-dt emu_run uses); < 0 as the device reports it
double emu_calc_dt(void * h, double cfl) { return emulated_calc_dt(*static_cast<Emu *>(h), cfl); }
// primitives [nc_ref][5] of the stepping state (refreshed by the last stage, Solver::update_primitives)
int emu_get_primitives(void * h, double * P_ref) {
    Emu & e = *static_cast<Emu *>(h);
    for (uint32_t i = 0; i < e.P.N_owned; i++)
        for (int v = 0; v < 5; v++) P_ref[5 * (size_t)e.P.perm_cells[i] + v] = e.prim[(size_t)v * e.P.Npad + i];
    return 0;
}

// mlb_set_state with prim != NULL (api.cu: import_state): the stepping state's primitives are the caller's (the reference steps from the
// primitives its initial condition defines, which need not be the ones recomputed from U in every bit); the density plane stays U's
int emu_set_primitives(void * h, const double * P_ref) {
    Emu & e = *static_cast<Emu *>(h);
    for (uint32_t i = 0; i < e.P.N; i++)
        for (int v = 0; v < 5; v++) e.prim[(size_t)v * e.P.Npad + i] = P_ref[5 * (size_t)e.P.perm_cells[i] + v];
    return 0;
}

// mlb_compute_primitives (prims_aos_kernel's body): cons_to_prim of the kernel source on a list of conserved states
int emu_primitives(const mlb_physics * phys, unsigned long long n, const double * U, double * P) {
    const GasParams g = make_gas(*phys);
    for (unsigned long long i = 0; i < n; i++) emu::cons_to_prim(g, &U[4 * i], &P[5 * i]);
    return 0;
}
int emu_riemann_flux(int riemann, unsigned long long n, const double * nunit, const double * L, const double * R, double gamma, double * flux) {
    switch (riemann) {
        case MLB_RIEMANN_RUSANOV: riemann_list<MLB_RIEMANN_RUSANOV>(n, nunit, L, R, gamma, flux); break;
        case MLB_RIEMANN_HLL: riemann_list<MLB_RIEMANN_HLL>(n, nunit, L, R, gamma, flux); break;
        default: riemann_list<MLB_RIEMANN_HLLC>(n, nunit, L, R, gamma, flux); break;
    }
    return 0;
}

const char * emu_last_error() { return emu_err.c_str(); }

// cell0_nodes != NULL: `mesh` is a rank-local mesh (mlb_create_local): the six node coordinates of the GLOBAL mesh's cell 0
void * emu_create(const mlb_mesh * mesh, const mlb_numerics * num, const mlb_physics * phys, const mlb_bc * bcs, int n_bcs, const int32_t * part,
                  int rank, int n_ranks, const double * cell0_nodes) {
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
        if (cell0_nodes) { opt.psi_ref_tri = cell0_nodes; opt.keep_ref_tables = false; }      // as create_impl (api.cu) for mlb_create_local
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
        e->bnd_s.assign(P.Npad, 0.0);
        for (uint32_t i = 0; i < P.N; i++) e->bnd_s[i] = 2.0 * std::pow(P.cell_vol[i], 1.0 / 2);     // as create_impl (api.cu), solver.cpp:664-665
        g.bnd_s = e->bnd_s.data();
        for (auto & b : e->Ub) b.assign(4 * (size_t)P.Npad, 0.0);
        for (auto & b : e->kb) b.assign(4 * (size_t)P.Npad, 0.0);
        e->prim.assign(6 * (size_t)P.Npad, 0.0); e->sr.assign(P.Npad, 0.0);
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
    // the stepping state: mlb_set_state with prim == NULL (primitives_soa_kernel: five primitives + the density plane)
    e.Ub[e.cur] = e.U;
    for (uint32_t i = 0; i < e.P.N; i++) {
        double P5[5];
        emu::cons_to_prim(e.gas, &e.U[4 * (size_t)i], P5);
        for (int v = 0; v < 5; v++) e.prim[(size_t)v * e.P.Npad + i] = P5[v];
        e.prim[5 * (size_t)e.P.Npad + i] = e.U[4 * (size_t)i];
    }
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
