// C-ABI layer (include/mallard_b200.h): context lifetime, state movement, the stepping seams and the parity hooks.
// Everything that computes runs in the CUDA kernels of kernels_impl.cuh; there is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <map>
#include <memory>
#include <omp.h>

#include "kernel_args.h"
#include "nccl_dyn.h"
#include "stage_plan.h"

using namespace mlb;

namespace {

thread_local std::string g_err;

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CUDA_OK(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t e_ = (expr);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr);                 \
    } while (0)

template <class T> T * dev_alloc(size_t n) {
    T * p = nullptr;
    CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
template <class T> T * dev_upload(const std::vector<T> & v, cudaStream_t st) {
    T * p = dev_alloc<T>(v.size());
    if (!v.empty()) CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return p;
}

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        throw CudaError(std::string("mallard_b200 needs a CUDA device and has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= n) throw CudaError("mallard_b200: CUDA device ordinal out of range");
    CUDA_OK(cudaSetDevice(device));
}

struct ProfileEntry { double ms = 0.0; uint64_t launches = 0; };

#define NCCL_OK(expr)                                                                                              \
    do {                                                                                                           \
        ncclResult_t r_ = (expr);                                                                                  \
        if (r_ != ncclSuccess)                                                                                     \
            throw std::runtime_error(std::string("NCCL error: ") + nccl().GetErrorString(r_) + " at " #expr);      \
    } while (0)

}  // namespace

struct mlb_ctx {
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    mlb_numerics num{};
    GasParams gas{};
    DevPhys phys{};
    const KernelTable * kt = nullptr;
    Prep prep;                         // host copy (big TENO tables are released after upload)
    uint32_t nc_ref = 0, nf_ref = 0;   // reference mesh sizes
    int n_stages = 1, n_rhs = 1;
    bool teno = false;
    int rank = 0, n_ranks = 1;

    DevGeom g{};
    double * U[3] = {nullptr, nullptr, nullptr};
    double * k[4] = {nullptr, nullptr, nullptr, nullptr};
    double * prim = nullptr, * sr = nullptr, * Fc = nullptr, * AF = nullptr, * scal = nullptr, * k_override = nullptr;
    double * G = nullptr;              // viscous runs: least-squares gradients of (u, v, T), AoS [Npad][6]
    long long * max_bits = nullptr;
    unsigned int * blocks_done = nullptr;
    unsigned long long * step_counter = nullptr;
    uint32_t * d_perm_cells = nullptr, * d_perm_faces = nullptr;
    uint32_t * d_st_ids = nullptr;
    double * d_st_area = nullptr, * d_st_mat = nullptr;
    double * d_OI = nullptr, * d_psi_bar = nullptr, * d_psi_bar_cell = nullptr;   // oscillation-indicator matrix, integral_psi_target, exponents (generic kernel)
    uint8_t * d_pidx = nullptr;
    bool streaming = false;            // FAST mode: compact streaming tables + teno_stream_kernel
    uint32_t * d_fm_ids = nullptr;
    double * d_fm_mat = nullptr, * d_fm_area0 = nullptr;
    uint32_t n_ftiles = 0;
    double * d_stage = nullptr;        // AoS staging, 5*N doubles (and face export)
    size_t d_stage_elems = 0;
    double * h_stage = nullptr;        // pinned host staging
    size_t h_stage_elems = 0;
    std::vector<void *> owned_dev;     // everything to cudaFree
    int cur = 0;                       // U[cur] holds the solution
    int last_temp = 1;                 // buffer that corresponds to the reference's solution_vec[1]
    bool has_override = false;
    size_t device_bytes = 0;
    double table_build_ms = 0.0;       // device-side TENO table construction (teno_tables.cu)

    // halo
    std::vector<int32_t> peers;
    std::vector<uint64_t> send_counts, recv_counts;      // cells per peer
    std::vector<std::vector<uint32_t>> recv_ref_ids;     // per peer: reference ids of ghosts, in recv-buffer order
    uint32_t * d_send_idx = nullptr, * d_recv_idx = nullptr;
    double * d_send_buf = nullptr, * d_recv_buf = nullptr;
    uint64_t n_send = 0, n_recv = 0;
    // The exchange lives on its own stream: pack -> (caller's communicator) -> unpack run on comm_stream while the compute
    // stream reconstructs the interior cells; ev_state orders pack after the producer of the state, ev_halo orders the
    // first reader of ghost data after unpack.
    double * d_ranges = nullptr;       // mlb_field_ranges scratch
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_state = nullptr, ev_halo = nullptr;
    bool halo_pending = false;         // an unpack has been enqueued that the compute stream has not waited for yet
    int stage_begun = -1;              // stage whose interior reconstruction is already enqueued (mlb_stage_begin)
    // rank-local ingest: global id of every cell of the (local) mesh this context was created from; empty = the mesh is global
    std::vector<uint32_t> global_ids;
    // native multi-GPU driver (mlb_comm_init): one communicator for the per-stage neighbour exchange on comm_stream, a second
    // (ncclCommSplit duplicate) for the per-step all-reduce of the spectral radius on the compute stream
    ncclComm_t nccl_p2p = nullptr, nccl_coll = nullptr;

    // measurement
    cudaEvent_t ev[16] = {};
    uint64_t launches = 0;
    uint64_t graph_replays = 0;        // steps of mlb_run executed as CUDA graph replays
    uint64_t small_steps = 0;          // steps of mlb_run executed inside the cooperative small-mesh kernel
    bool profiling = false;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::map<std::string, ProfileEntry> profile;
    std::vector<std::string> profile_names;   // stable storage for mlb_profile_read

    template <class T> T * track(T * p, size_t n) { owned_dev.push_back((void *)p); device_bytes += n * sizeof(T); return p; }
    template <class T> T * alloc(size_t n) { return track(dev_alloc<T>(n), n); }
    template <class T> T * upload(const std::vector<T> & v) { return track(dev_upload(v, stream), v.size()); }

    void flush_profile() {
        for (auto & p : pending) {
            float ms = 0.f;
            cudaEventSynchronize(p.second.second);
            cudaEventElapsedTime(&ms, p.second.first, p.second.second);
            auto & e = profile[p.first];
            e.ms += ms; e.launches++;
            cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second);
        }
        pending.clear();
    }
    template <class F> void launch(const char * name, F && f) {
        if (profiling) {
            cudaEvent_t a, b;
            CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
            CUDA_OK(cudaEventRecord(a, stream));
            f();
            CUDA_OK(cudaEventRecord(b, stream));
            pending.push_back({name, {a, b}});
            if (pending.size() >= 2048) flush_profile();
        } else f();
        launches++;
        CUDA_OK(cudaGetLastError());
    }
    void ensure_stage(size_t elems) {
        if (elems > d_stage_elems) {
            if (d_stage) cudaFree(d_stage);
            d_stage = dev_alloc<double>(elems); d_stage_elems = elems;
        }
        if (elems > h_stage_elems) {
            if (h_stage) cudaFreeHost(h_stage);
            CUDA_OK(cudaMallocHost(&h_stage, elems * sizeof(double))); h_stage_elems = elems;
        }
    }
    ~mlb_ctx() {
        if (stream) cudaStreamSynchronize(stream);
        if (comm_stream) cudaStreamSynchronize(comm_stream);
        try {
            if (nccl_coll) nccl().CommDestroy(nccl_coll);
            if (nccl_p2p) nccl().CommDestroy(nccl_p2p);
        } catch (...) {}
        for (auto & p : pending) { cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second); }
        for (void * p : owned_dev) cudaFree(p);
        if (d_stage) cudaFree(d_stage);
        if (h_stage) cudaFreeHost(h_stage);
        for (auto & e : ev) if (e) cudaEventDestroy(e);
        if (comm_stream) { cudaStreamSynchronize(comm_stream); cudaStreamDestroy(comm_stream); }
        if (ev_state) cudaEventDestroy(ev_state);
        if (ev_halo) cudaEventDestroy(ev_halo);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct mlb_host_mesh { HostMesh m; };

// Ghost cells grouped by owning rank (ascending), in library order within a group: the layout of the receive buffer.
static void halo_recv_lists(const Prep & P, const int32_t * part, std::vector<int32_t> & peers, std::vector<uint64_t> & counts,
                            std::vector<std::vector<uint32_t>> & ref_ids, std::vector<uint32_t> & recv_idx) {
    std::map<int32_t, std::vector<uint32_t>> by_owner;
    for (uint32_t i = P.N_owned; i < P.N; i++) by_owner[part[P.perm_cells[i]]].push_back(i);
    peers.clear(); counts.clear(); ref_ids.clear(); recv_idx.clear();
    for (auto & kv : by_owner) {
        peers.push_back(kv.first);
        counts.push_back(kv.second.size());
        std::vector<uint32_t> ref;
        for (uint32_t i : kv.second) { ref.push_back(P.perm_cells[i]); recv_idx.push_back(i); }
        ref_ids.push_back(std::move(ref));
    }
}
struct mlb_plan { Prep prep; uint32_t nc_ref = 0, nf_ref = 0; std::vector<int32_t> part; int rank = 0; std::string err; };

namespace {

std::vector<StagePlan> make_plan(const mlb_ctx & c) { return make_stage_plan(c.cur, c.num.integrator); }   // stage_plan.h

ReconArgs recon_args(mlb_ctx & c, const double * Uin) {
    ReconArgs r{};
    r.g = c.g; r.Uin = Uin; r.Fc = c.Fc; r.st_ids = c.d_st_ids; r.st_area = c.d_st_area; r.st_mat = c.d_st_mat;
    const TenoTables & T = c.prep.teno;
    r.order = T.order; r.K = T.K; r.M = T.M; r.Mp = T.Mp; r.S = T.S; r.basis = T.basis; r.fixed_weights = c.num.teno_fixed;
    for (size_t i = 0; i < c.prep.qf_x.size(); i++) r.qf_x[i] = c.prep.qf_x[i];
    if (T.K <= 15) {
        for (int i = 0; i < T.K; i++) { r.psi_bar[i] = T.psi_bar[i]; r.pidx[2 * i] = T.pidx[2 * i]; r.pidx[2 * i + 1] = T.pidx[2 * i + 1]; }
        for (int i = 0; i < T.K * T.K; i++) r.OI[i] = T.OI[i];
    }
    r.OI_dev = c.d_OI; r.psi_bar_dev = c.d_psi_bar; r.pidx_dev = c.d_pidx; r.psi_bar_cell = c.d_psi_bar_cell;
    return r;
}

ReconStreamArgs stream_args(mlb_ctx & c, const double * Uin) {
    ReconStreamArgs r{};
    const TenoTables & T = c.prep.teno;
    r.g = c.g; r.Uin = Uin; r.Fc = c.Fc; r.mat = c.d_fm_mat; r.ids = c.d_fm_ids; r.area0 = c.d_fm_area0;
    r.n_tiles = c.n_ftiles; r.order = T.order; r.fixed_weights = c.num.teno_fixed; r.basis = T.basis;
    static const int variant = [] { const char * e = getenv("MLB_STREAM_GATHER"); return e && !strcmp(e, "ownvar") ? 2 : 1; }();
    r.async_gather = variant;   // tuning knob: which lanes fetch which neighbour bytes (see teno_stream.cuh)
    for (size_t i = 0; i < c.prep.qf_x.size() && i < 4; i++) r.qf_x[i] = c.prep.qf_x[i];
    for (int i = 0; i < T.K; i++) r.psi_bar[i] = T.psi_bar[i];
    for (size_t i = 0; i < T.OIs.size(); i++) r.OIs[i] = T.OIs[i];
    return r;
}

// The compute stream's first reader of ghost data waits for the unpack enqueued on the communication stream.
void wait_halo(mlb_ctx & c) {
    if (c.halo_pending) { CUDA_OK(cudaStreamWaitEvent(c.stream, c.ev_halo, 0)); c.halo_pending = false; }
}

// Tiles made of interior cells only (stencils without ghosts): their reconstruction overlaps the halo exchange.
uint32_t interior_tiles(const mlb_ctx & c) {
    return (c.streaming && c.n_ranks > 1) ? c.prep.N_interior / FAST_CT : 0u;
}

// phase 0: everything; 1: interior tiles only (before the halo wait); 2: the remaining tiles
void run_recon(mlb_ctx & c, const double * Uin, int phase = 0) {
    if (c.streaming) {
        ReconStreamArgs r = stream_args(c, Uin);
        const uint32_t ti = interior_tiles(c);
        if (phase == 1) r.n_tiles = ti;
        if (phase == 2) r.tile_begin = ti;
        if (r.n_tiles > r.tile_begin) c.launch("teno_stream", [&] { c.kt->recon_stream(r, c.stream); });
    } else if (phase != 1) {
        ReconArgs r = recon_args(c, Uin);
        c.launch("teno_recon", [&] { c.kt->recon(r, c.stream); });
    }
}

// the argument block of one stage's face / gather kernels
StageArgs stage_args(mlb_ctx & c, const StagePlan & s, bool bare, double * k_out) {
    StageArgs a{};
    a.g = c.g; a.ph = c.phys; a.Uin = c.U[s.in]; a.Fc = c.Fc; a.AF = c.AF; a.teno = c.teno ? 1 : 0;
    a.k_override = c.has_override ? c.k_override : nullptr;
    a.scal = c.scal; a.step_counter = c.step_counter; a.G = c.G;
    RkArgs & rk = a.rk;
    if (bare) { rk.mode = 3; rk.k_store = k_out; }
    else {
        rk.mode = s.mode; rk.n_prev = s.n_prev; rk.last_stage = s.last;
        rk.base = c.U[s.base]; rk.out = c.U[s.out];
        const bool need_k = c.num.keep_stage_rhs || !s.last;
        rk.k_store = need_k ? c.k[s.kstore] : nullptr;
        for (int j = 0; j < s.n_prev; j++) { rk.kprev[j] = c.k[s.kprev[j]]; rk.cprev[j] = s.cprev[j]; }
        rk.c0 = s.c0; rk.c1 = s.c1; rk.coef = s.coef;
        rk.prim_out = s.last ? c.prim : nullptr;
    }
    return a;
}

void run_stage(mlb_ctx & c, const StagePlan & s, bool bare, double * k_out) {
    const double * Uin = c.U[s.in];
    const bool recon = c.teno && !c.has_override;
    if (c.stage_begun >= 0) {              // mlb_stage_begin already enqueued the interior tiles
        c.stage_begun = -1;
        wait_halo(c);
        if (recon) run_recon(c, Uin, 2);
    } else if (recon && c.halo_pending && interior_tiles(c)) {
        run_recon(c, Uin, 1);
        wait_halo(c);
        run_recon(c, Uin, 2);
    } else {
        wait_halo(c);
        if (recon) run_recon(c, Uin);
    }
    const StageArgs a = stage_args(c, s, bare, k_out);
    if (!c.has_override && c.G) c.launch("visc_grad", [&] { c.kt->gradients(a, c.stream); });
    if (!c.has_override) c.launch(c.teno ? "face_flux_teno" : "face_flux_fo", [&] { c.kt->faces(a, c.stream); });
    c.launch("gather_stage", [&] { c.kt->stage(a, c.stream); });
}

void finish_plan(mlb_ctx & c, const std::vector<StagePlan> & plan) {
    if (c.num.integrator == MLB_INTEGRATOR_FE) { c.last_temp = c.cur; c.cur = plan.back().out; }
    else c.last_temp = plan[plan.size() - 2].out;   // reference's solution_vec[1] after the step
}

void do_step(mlb_ctx & c) {
    const auto plan = make_plan(c);
    for (auto & s : plan) run_stage(c, s, false, nullptr);
    finish_plan(c, plan);
}

CflArgs cfl_args(mlb_ctx & c, double cfl) {
    CflArgs a{};
    a.g = c.g; a.gas = c.gas; a.U = c.U[c.cur]; a.prim = c.prim; a.sr_out = c.sr; a.scal = c.scal;
    a.max_bits = c.max_bits; a.blocks_done = c.blocks_done; a.cfl = cfl;
    return a;
}

void do_calc_dt(mlb_ctx & c, double cfl) {
    const CflArgs a = cfl_args(c, cfl);
    wait_halo(c);   // the CFL kernel reads the ghosts' primitives
    c.launch("cfl", [&] { c.kt->cfl(a, c.stream); });
}

// Small first-order meshes: whole steps in one cooperative launch (csrc/small_step.cuh).  Opt-in (MLB_SMALL_STEP=1, read per call)
// until it has been measured on hardware; the result is that of the multi-kernel path (same bodies, same argument blocks).
bool small_step_eligible(const mlb_ctx & c) {
    const char * e = getenv("MLB_SMALL_STEP");
    return e && e[0] == '1' && !c.teno && !c.G && !c.has_override && c.n_ranks == 1 && !c.profiling && !c.halo_pending &&
           c.num.integrator != MLB_INTEGRATOR_FE && c.prep.N_owned <= 262144u;
}
void run_small_steps(mlb_ctx & c, uint32_t n_steps, double cfl) {
    const auto plan = make_plan(c);
    SmallStepArgs p{};
    for (size_t st = 0; st < plan.size(); st++) p.st[st] = stage_args(c, plan[st], false, nullptr);
    p.cfl = cfl_args(c, cfl > 0.0 ? cfl : 0.0);
    p.n_stages = (int32_t)plan.size();
    p.n_steps = n_steps;
    const char * b = getenv("MLB_SMALL_STEP_BLOCKS");      // A/B knob: cap the grid (fewer blocks = cheaper grid barriers)
    const int max_blocks = b ? atoi(b) : 0;
    c.launch("small_step", [&] { c.kt->small_step(p, max_blocks, c.stream); });
    CUDA_OK(cudaGetLastError());
    c.small_steps += n_steps;
    finish_plan(c, plan);
}

void read_scalars(mlb_ctx & c, double * out) {
    CUDA_OK(cudaMemcpyAsync(out, c.scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
}

// Solver::calc_dt's `if (dt < 0) throw` (solver.cpp:587-589).  The stage kernels treat a negative device-resident dt as
// "no step" (the solution, t and the step counter stay as they are), so reporting after the fact loses nothing.
double require_dt(mlb_ctx & c) {
    double sc[SC_COUNT];
    read_scalars(c, sc);
    if (sc[SC_DT] < 0.0) throw std::runtime_error("dt negative: " + std::to_string(sc[SC_DT]) + ".");
    return sc[SC_DT];
}

void import_state(mlb_ctx & c, const double * U_host, double * soa, int nv) {
    const size_t n = (size_t)c.nc_ref * nv;
    c.ensure_stage(std::max<size_t>(n, 1));
    CUDA_OK(cudaMemcpyAsync(c.d_stage, U_host, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    launch_import_state(c.d_stage, c.d_perm_cells, c.prep.N, c.prep.Npad, nv, soa, c.stream);
    c.launches++;
}

void export_state(mlb_ctx & c, const double * soa, int nv, double * out_host) {
    const size_t n = (size_t)c.nc_ref * nv;
    c.ensure_stage(std::max<size_t>(n, 1));
    if (c.n_ranks > 1) CUDA_OK(cudaMemsetAsync(c.d_stage, 0, n * sizeof(double), c.stream));
    launch_export_state(soa, c.d_perm_cells, c.prep.N_owned, c.prep.Npad, nv, c.d_stage, c.stream);
    c.launches++;
    CUDA_OK(cudaMemcpyAsync(out_host, c.d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
}

// the reconstruction matrices of the compact tables are computed in HBM (teno_tables.cu), bit-identical to
// the host preprocessor's; only stencil ids and node coordinates are uploaded.
void build_tables_on_device(mlb_ctx & c) {
    TenoTables & T = c.prep.teno;
    const int KR = T.K - 1, MC = T.M - 1;
    const size_t frow = (size_t)(2 * (MC / 2) + 1) * FAST_CT;
    const size_t n_mat = (size_t)c.n_ftiles * FAST_S * KR * frow;
    c.d_fm_mat = c.alloc<double>(n_mat);
    c.d_fm_area0 = c.alloc<double>((size_t)c.n_ftiles * FAST_CT);
    double * d_tri = dev_upload(T.tri_xy, c.stream);
    int * d_err = dev_alloc<int>(1);
    CUDA_OK(cudaMemsetAsync(d_err, 0, sizeof(int), c.stream));
    TableBuildArgs a{};
    a.n_recon = c.prep.N_recon; a.n_ftiles = c.n_ftiles; a.order = T.order; a.nq = T.nq_cell;
    a.tri_xy = d_tri; a.fm_ids = c.d_fm_ids; a.fm_mat = c.d_fm_mat; a.fm_area0 = c.d_fm_area0; a.err_flag = d_err;
    for (int q = 0; q < T.nq_cell; q++) { a.qc_xy[2 * q] = T.qc_xy[2 * q]; a.qc_xy[2 * q + 1] = T.qc_xy[2 * q + 1]; a.qc_w[q] = T.qc_w[q]; }
    for (int k = 0; k < T.K; k++) a.psi_bar[k] = T.psi_bar[k];
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    CUDA_OK(cudaEventRecord(e0, c.stream));
    launch_teno_tables(a, c.stream);
    c.launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(e1, c.stream));
    int err = 0;
    CUDA_OK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    c.table_build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_tri); cudaFree(d_err);
    if (err) throw std::runtime_error("TENO: reconstruction matrix has a non-zero first column; compact tables unavailable (use fp_mode strict)");
    if (getenv("MLB_PREP_TIMING")) fprintf(stderr, "[mlb] device table build: %u cells, %.2f ms\n", c.prep.N_recon, ms);
}

mlb_ctx * create_impl(const mlb_mesh * mesh, const int32_t * part, const mlb_numerics * numerics, const mlb_physics * physics,
                      const mlb_bc * bcs, int32_t n_bcs, const mlb_parallel * par, const mlb_local_mesh * local = nullptr) {
    if (!mesh || !numerics || !physics) throw std::runtime_error("mlb_create: NULL argument");
    if (n_bcs > MAX_BCS) throw std::runtime_error("mlb_create: too many boundaries");
    std::unique_ptr<mlb_ctx> c(new mlb_ctx());
    c->device = par ? par->device : 0;
    c->rank = par ? par->rank : 0;
    c->n_ranks = par ? std::max(1, par->n_ranks) : 1;
    require_device(c->device);
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (c->n_ranks > 1) {
        CUDA_OK(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_state, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    }
    c->num = *numerics;
    if (c->num.recon != MLB_RECON_FO && c->num.recon != MLB_RECON_TENO) throw std::runtime_error("Unknown face reconstruction type.");
    if (c->num.riemann < 0 || c->num.riemann > 2) throw std::runtime_error("Unknown Riemann solver type.");
    if (c->num.integrator < 0 || c->num.integrator > 2) throw std::runtime_error("Unknown time integrator type.");
    c->teno = c->num.recon == MLB_RECON_TENO;
    c->kt = c->num.fp_mode == MLB_FP_FAST ? kernels_fast() : kernels_strict();
    c->gas = make_gas(*physics);
    c->n_stages = c->num.integrator == MLB_INTEGRATOR_FE ? 1 : (c->num.integrator == MLB_INTEGRATOR_RK4 ? 4 : 3);
    c->n_rhs = c->n_stages;

    HostMesh hm;
    host_mesh_from_view(hm, *mesh);
    c->nc_ref = hm.nc; c->nf_ref = hm.nf;

    std::vector<std::string> bc_zones;
    c->phys.gas = c->gas; c->phys.riemann = c->num.riemann; c->phys.n_bcs = n_bcs;
    for (int b = 0; b < n_bcs; b++) {
        if (!bcs[b].zone_name) throw std::runtime_error("Boundary name not specified.");
        if (bcs[b].type < 0 || bcs[b].type > MLB_BC_WALL_NOSLIP) throw std::runtime_error("Unknown boundary type.");
        bc_zones.push_back(bcs[b].zone_name);
        BcParams & d = c->phys.bcs[b];
        d.type = bcs[b].type;
        for (double & x : d.data) x = 0.0;
        if (d.type == MLB_BC_UPT) {   // boundary_upt.cpp:38-73
            const double rho = bcs[b].p / (c->gas.R * bcs[b].T);
            const double e = c->gas.cv * bcs[b].T;
            d.data[0] = rho; d.data[1] = bcs[b].u[0]; d.data[2] = bcs[b].u[1]; d.data[3] = bcs[b].p; d.data[4] = bcs[b].T;
            d.data[5] = e + bcs[b].p / rho;
        } else if (d.type == MLB_BC_P_OUT) d.data[0] = bcs[b].p;
        else if (d.type == MLB_BC_WALL_NOSLIP) { d.data[1] = bcs[b].u[0]; d.data[2] = bcs[b].u[1]; d.data[4] = bcs[b].T; }
    }

    PrepOptions opt;
    opt.renumber = c->num.renumber;
    opt.keep_ref_tables = c->teno && !part && hm.nc <= 200000;
    if (c->teno && c->num.fp_mode == MLB_FP_FAST && c->kt->stream_supported) {
        int n_slots = 0;
        for (uint32_t cc = 0; cc < hm.nc; cc++) n_slots = std::max(n_slots, hm.nfc(cc));
        const int p = c->num.basis_order, K = (p + 1) * (p + 2) / 2;
        const double factor = c->num.max_stencil_size_factor > 0 ? c->num.max_stencil_size_factor : 2.0;
        const int Q = c->num.quadrature_order_face > 0 ? c->num.quadrature_order_face : (p + 1) / 2;
        c->streaming = c->kt->stream_supported(p, (int)(uint16_t)(factor * K), Q, c->num.basis, n_slots);
    }
    opt.fast_tables = c->streaming;
    {   // the compact tables' matrices are built on the device unless MLB_HOST_TABLES=1 asks for the host path (parity check)
        const char * e = getenv("MLB_HOST_TABLES");
        const int p = c->num.basis_order;
        opt.device_tables = c->streaming && !(e && e[0] == '1') && teno_tables_device_supported(p, c->num.basis, 7 /* Dunavant rules have <= 7 points */);
    }
    opt.part = part; opt.rank = c->rank; opt.n_ranks = c->n_ranks;
    opt.viscous = c->gas.mu > 0.0;
    if (local) {   // rank-local ingest: the mesh holds this rank's cells and enough ghost layers, in the global order
        if (!local->global_cell_ids) throw std::runtime_error("mlb_create_local: global_cell_ids is NULL");
        c->global_ids.assign(local->global_cell_ids, local->global_cell_ids + hm.nc);
        for (uint32_t i = 1; i < hm.nc; i++)
            if (c->global_ids[i] <= c->global_ids[i - 1]) throw std::runtime_error("mlb_create_local: global_cell_ids must be strictly ascending (the local numbering keeps the global order)");
        if (hm.nc && c->global_ids.back() >= local->n_global_cells) throw std::runtime_error("mlb_create_local: global cell id out of range");
        opt.psi_ref_tri = local->cell0_nodes;
        opt.keep_ref_tables = false;
    } else {
        for (size_t f = 0; f < hm.nf; f++)
            if (hm.cof[2 * f + 1] == CUT_FACE) throw std::runtime_error("mlb_create: cells_of_face holds a cut-face marker (-2); rank-local meshes go through mlb_create_local");
    }
    preprocess(hm, c->num, bc_zones, opt, c->prep);
    Prep & P = c->prep;
    for (size_t i = 0; i < P.qf_x.size(); i++) { c->phys.qf_x[i] = P.qf_x[i]; c->phys.qf_w[i] = P.qf_w[i]; }
    if (c->teno && !c->streaming && !c->kt->recon_supported(P.teno.order, P.teno.K, P.teno.Mp, P.teno.S, P.teno.basis))
        throw std::runtime_error("TENO: this (basis, order, stencil size) combination exceeds what the device kernels hold in shared memory "
                                 "(order 1-9; stencils of up to ~500 cells)");

    // ---- upload
    DevGeom & g = c->g;
    g.N = P.N; g.N_owned = P.N_owned; g.N_recon = P.N_recon; g.Npad = P.Npad; g.NF = P.NF; g.n_slots = P.n_slots; g.Q = P.Q;
    g.slot_face = c->upload(P.slot_face); g.slot_nbr = c->upload(P.slot_nbr); g.slot_nslot = c->upload(P.slot_nslot);
    g.rhs_order = c->upload(P.rhs_order); g.nfc = c->upload(P.n_faces_of_cell);
    g.cell_vol = c->upload(P.cell_vol); g.cell_xy = c->upload(P.cell_xy);
    {
        dvec bs(P.Npad, 0.0);
        for (uint32_t i = 0; i < P.N; i++) bs[i] = 2.0 * std::pow(P.cell_vol[i], 1.0 / 2);   // solver.cpp:664-665 (libm pow)
        g.bnd_s = c->upload(bs);
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    g.face_nx = c->upload(P.face_nx); g.face_ny = c->upload(P.face_ny); g.face_area = c->upload(P.face_area);
    g.slot_fx = c->teno ? c->upload(P.slot_fx) : nullptr;
    g.NFpad = P.NFpad;
    if (opt.viscous) {
        g.slot_d = c->upload(P.slot_d); g.face_d = c->upload(P.face_d);
        c->G = c->alloc<double>(6 * (size_t)P.Npad);
        CUDA_OK(cudaMemsetAsync(c->G, 0, 6 * (size_t)P.Npad * sizeof(double), c->stream));
    }
    g.face_cl = c->upload(P.face_cl); g.face_cr = c->upload(P.face_cr); g.face_slots = c->upload(P.face_slots);
    c->AF = c->alloc<double>(4 * (size_t)std::max<uint32_t>(P.NFpad, 1));
    c->d_perm_cells = c->upload(P.perm_cells);
    c->d_perm_faces = c->upload(P.perm_faces);
    const size_t NP = P.Npad;
    for (int i = 0; i < 3; i++) { c->U[i] = c->alloc<double>(4 * NP); CUDA_OK(cudaMemsetAsync(c->U[i], 0, 4 * NP * sizeof(double), c->stream)); }
    for (int i = 0; i < c->n_rhs; i++) { c->k[i] = c->alloc<double>(4 * NP); CUDA_OK(cudaMemsetAsync(c->k[i], 0, 4 * NP * sizeof(double), c->stream)); }
    c->prim = c->alloc<double>(6 * NP); CUDA_OK(cudaMemsetAsync(c->prim, 0, 6 * NP * sizeof(double), c->stream));
    c->sr = c->alloc<double>(NP); CUDA_OK(cudaMemsetAsync(c->sr, 0, NP * sizeof(double), c->stream));
    c->scal = c->alloc<double>(SC_COUNT);
    {
        double init[SC_COUNT] = {0};
        init[SC_DT] = -1.0;   // "no dt yet": mlb_take_step refuses negative dt like Solver::calc_dt (solver.cpp:587-589)
        CUDA_OK(cudaMemcpyAsync(c->scal, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    c->max_bits = c->alloc<long long>(1);
    {
        const double m1 = -1.0; long long bits; std::memcpy(&bits, &m1, 8);
        CUDA_OK(cudaMemcpy(c->max_bits, &bits, 8, cudaMemcpyHostToDevice));
    }
    c->blocks_done = c->alloc<unsigned int>(1); CUDA_OK(cudaMemset(c->blocks_done, 0, 4));
    c->step_counter = c->alloc<unsigned long long>(1); CUDA_OK(cudaMemset(c->step_counter, 0, 8));
    if (c->teno) {
        TenoTables & T = P.teno;
        c->Fc = c->alloc<double>((size_t)P.n_slots * P.Q * 4 * NP);
        CUDA_OK(cudaMemsetAsync(c->Fc, 0, (size_t)P.n_slots * P.Q * 4 * NP * sizeof(double), c->stream));
        if (c->streaming) {
            c->n_ftiles = (uint32_t)((P.N_recon + FAST_CT - 1) / FAST_CT);
            c->d_fm_ids = c->upload(T.fm_ids);
            if (opt.device_tables) build_tables_on_device(*c);
            else { c->d_fm_area0 = c->upload(T.fm_area0); c->d_fm_mat = c->upload(T.fm_mat); }
        } else {
            c->d_st_ids = c->upload(T.st_ids); c->d_st_area = c->upload(T.st_area); c->d_st_mat = c->upload(T.st_mat);
            c->d_OI = c->upload(T.OI); c->d_psi_bar = c->upload(T.psi_bar); c->d_pidx = c->upload(T.pidx);
            if (T.mixed) c->d_psi_bar_cell = c->upload(T.psi_bar_cell);
        }
        CUDA_OK(cudaStreamSynchronize(c->stream));
        uvec().swap(T.st_ids); dvec().swap(T.st_area); dvec().swap(T.st_mat);   // host copies no longer needed
        uvec().swap(T.fm_ids); dvec().swap(T.fm_area0); dvec().swap(T.fm_mat); dvec().swap(T.tri_xy);
    }
    for (auto & e : c->ev) CUDA_OK(cudaEventCreate(&e));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return c.release();
}


// Cell ids crossing the ABI are in the numbering of the mesh the context was created from; contexts created from a rank-local
// mesh (mlb_create_local) exchange GLOBAL ids with their peers.
uint32_t to_global(const mlb_ctx & c, uint32_t mesh_id) { return c.global_ids.empty() ? mesh_id : c.global_ids[mesh_id]; }
uint32_t from_global(const mlb_ctx & c, uint32_t gid) {
    if (c.global_ids.empty()) return gid < c.nc_ref ? gid : NO_FACE;
    const auto it = std::lower_bound(c.global_ids.begin(), c.global_ids.end(), gid);
    return (it != c.global_ids.end() && *it == gid) ? (uint32_t)(it - c.global_ids.begin()) : NO_FACE;
}

// What this rank must send: per peer (ascending), the cells the peer holds as ghosts, in the order of the peer's receive buffer.
// A peer that only receives from this rank (TENO stencils are not symmetric) joins the peer list with an empty receive list,
// which leaves the receive-buffer offsets unchanged.
void set_send_ids(mlb_ctx & c, int32_t n_lists, const int32_t * peer_ranks, const uint64_t * counts, const uint32_t * ids) {
    for (int l = 1; l < n_lists; l++)
        if (peer_ranks[l] <= peer_ranks[l - 1]) throw std::runtime_error("halo: send lists must be in ascending peer order");
    for (int l = 0; l < n_lists; l++) {
        if (!counts[l] || std::find(c.peers.begin(), c.peers.end(), peer_ranks[l]) != c.peers.end()) continue;
        const size_t at = std::lower_bound(c.peers.begin(), c.peers.end(), peer_ranks[l]) - c.peers.begin();
        c.peers.insert(c.peers.begin() + at, peer_ranks[l]);
        c.recv_counts.insert(c.recv_counts.begin() + at, 0);
        c.recv_ref_ids.insert(c.recv_ref_ids.begin() + at, std::vector<uint32_t>());
    }
    std::vector<uint32_t> idx;
    c.send_counts.assign(c.peers.size(), 0);
    size_t off = 0;
    for (int l = 0; l < n_lists; l++) {
        auto it = std::find(c.peers.begin(), c.peers.end(), peer_ranks[l]);
        for (uint64_t k = 0; k < counts[l]; k++) {
            const uint32_t ref = from_global(c, ids[off + k]);
            const uint32_t loc = ref != NO_FACE ? c.prep.iperm_cells[ref] : NO_FACE;
            if (loc == NO_FACE || loc >= c.prep.N_owned) throw std::runtime_error("halo: peer requested a cell this rank does not own");
            idx.push_back(loc);
        }
        if (it != c.peers.end()) c.send_counts[it - c.peers.begin()] = counts[l];
        off += counts[l];
    }
    c.n_send = idx.size();
    c.d_send_idx = c.upload(idx);
    c.d_send_buf = c.alloc<double>(4 * std::max<size_t>(c.n_send, 1));
    CUDA_OK(cudaStreamSynchronize(c.stream));
}

void halo_pack(mlb_ctx & c, int buf) {
    cudaStream_t cs = c.comm_stream ? c.comm_stream : c.stream;
    if (c.comm_stream) {                  // everything enqueued on the compute stream so far produces the state to be sent
        CUDA_OK(cudaEventRecord(c.ev_state, c.stream));
        CUDA_OK(cudaStreamWaitEvent(cs, c.ev_state, 0));
    }
    launch_gather(c.U[buf], c.d_send_idx, (uint32_t)c.n_send, c.prep.Npad, c.d_send_buf, cs);
    c.launches++;
    CUDA_OK(cudaGetLastError());
}

void halo_unpack(mlb_ctx & c, int buf, bool first_stage) {
    cudaStream_t cs = c.comm_stream ? c.comm_stream : c.stream;
    launch_scatter(c.d_recv_buf, c.d_recv_idx, (uint32_t)c.n_recv, c.prep.Npad, c.U[buf], cs);
    c.launches++;
    if (first_stage && c.prep.N > c.prep.N_owned) {   // ghosts' primitives for the CFL kernel (update_primitives of their owners)
        const uint32_t off = c.prep.N_owned & ~31u;    // keep the SoA column alignment
        c.kt->primitives_soa(c.gas, c.prep.N - off, c.prep.Npad, c.U[buf] + 4 * (size_t)off, c.prim + off, cs);
        c.launches++;
    }
    CUDA_OK(cudaGetLastError());
    if (c.comm_stream) { CUDA_OK(cudaEventRecord(c.ev_halo, cs)); c.halo_pending = true; }
}

// ---- native multi-GPU driver: NCCL over NVLink, no host code between the stages of a step -----------------------------------
// MLB_COMM_TRACE=1: one stderr line per phase of the collective set-up and of the first steps (which rank stopped where, should a
// run not come back: the 8-GPU hang of profiles/r02d_8gpu_native_hang.txt left no trace of its own)
void comm_trace(const mlb_ctx & c, const char * what) {
    static const bool on = [] { const char * e = getenv("MLB_COMM_TRACE"); return e && e[0] == '1'; }();
    if (on) { fprintf(stderr, "[mlb comm] rank %d/%d: %s\n", c.rank, c.n_ranks, what); fflush(stderr); }
}
void require_comm(const mlb_ctx & c) {
    if (!c.nccl_p2p) throw std::runtime_error("the context has no communicator: call mlb_comm_init first");
}

// pack -> grouped ncclSend / ncclRecv between the library's device buffers -> unpack, all on the communication stream
void nccl_exchange(mlb_ctx & c, int buf, bool first_stage) {
    const NcclApi & N = nccl();
    halo_pack(c, buf);
    NCCL_OK(N.GroupStart());
    size_t so = 0, ro = 0;
    for (size_t i = 0; i < c.peers.size(); i++) {
        const size_t sc = 4 * (size_t)c.send_counts[i], rc = 4 * (size_t)c.recv_counts[i];
        if (rc) NCCL_OK(N.Recv(c.d_recv_buf + ro, rc, ncclDouble, c.peers[i], c.nccl_p2p, c.comm_stream));
        if (sc) NCCL_OK(N.Send(c.d_send_buf + so, sc, ncclDouble, c.peers[i], c.nccl_p2p, c.comm_stream));
        so += sc; ro += rc;
    }
    NCCL_OK(N.GroupEnd());
    halo_unpack(c, buf, first_stage);
}

// One time step over the partitioned mesh (mallard_b200/parallel.py documents the same schedule):
//   communication stream:  pack -> send/recv -> unpack                                                   (per stage)
//   compute stream:        reconstruction of the INTERIOR tiles | wait(unpack) | rim tiles, fluxes, residual + RK update
//                          stage 0 also: local max spectral radius -> ncclAllReduce(max) -> dt, between the two halves
void do_step_distributed(mlb_ctx & c, double cfl) {
    const NcclApi & N = nccl();
    const auto plan = make_plan(c);
    for (size_t st = 0; st < plan.size(); st++) {
        nccl_exchange(c, plan[st].in, st == 0);
        if (c.teno && !c.has_override && interior_tiles(c)) { run_recon(c, c.U[plan[st].in], 1); c.stage_begun = (int)st; }
        if (st == 0 && cfl > 0.0) {
            do_calc_dt(c, -1.0);           // rank-local maximum; waits for the unpack (the CFL kernel reads the ghosts' primitives)
            NCCL_OK(N.AllReduce(c.scal + SC_MAX_SR, c.scal + SC_MAX_SR, 1, ncclDouble, ncclMax, c.nccl_coll, c.stream));
            launch_apply_dt(c.scal, c.max_bits, cfl, 0.0, 0, c.stream);
            c.launches++;
        }
        run_stage(c, plan[st], false, nullptr);
    }
    finish_plan(c, plan);
}

// Every rank tells every peer which of the peer's cells it holds as ghosts (global / mesh ids, in receive-buffer order): an
// all-gather of the count matrix, then one grouped send/recv of the id lists.  Collective over the communicator.
void exchange_halo_plan(mlb_ctx & c) {
    const NcclApi & N = nccl();
    const int n = c.n_ranks;
    std::vector<uint64_t> want((size_t)n, 0), all((size_t)n * n, 0);
    for (size_t i = 0; i < c.peers.size(); i++) want[c.peers[i]] = c.recv_counts[i];
    uint64_t * d_want = dev_upload(want, c.stream), * d_all = dev_alloc<uint64_t>((size_t)n * n);
    comm_trace(c, "halo plan: all-gather of the count matrix");
    NCCL_OK(N.AllGather(d_want, d_all, (size_t)n, ncclUint64, c.nccl_coll, c.stream));
    CUDA_OK(cudaMemcpyAsync(all.data(), d_all, all.size() * 8, cudaMemcpyDeviceToHost, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
    comm_trace(c, "halo plan: count matrix received, exchanging the id lists");
    std::vector<uint32_t> out_ids;
    for (size_t i = 0; i < c.peers.size(); i++)
        for (uint32_t id : c.recv_ref_ids[i]) out_ids.push_back(to_global(c, id));
    std::vector<int32_t> from;
    std::vector<uint64_t> counts;
    size_t n_in = 0;
    for (int p = 0; p < n; p++) {
        const uint64_t k = all[(size_t)p * n + c.rank];      // what rank p receives from this rank
        if (p != c.rank && k) { from.push_back(p); counts.push_back(k); n_in += k; }
    }
    uint32_t * d_out = dev_upload(out_ids, c.stream), * d_in = dev_alloc<uint32_t>(n_in);
    NCCL_OK(N.GroupStart());
    size_t off = 0;
    for (size_t i = 0; i < c.peers.size(); i++) {
        if (c.recv_counts[i]) NCCL_OK(N.Send(d_out + off, c.recv_counts[i], ncclUint32, c.peers[i], c.nccl_p2p, c.stream));
        off += c.recv_counts[i];
    }
    off = 0;
    for (size_t l = 0; l < from.size(); l++) {
        NCCL_OK(N.Recv(d_in + off, counts[l], ncclUint32, from[l], c.nccl_p2p, c.stream));
        off += counts[l];
    }
    NCCL_OK(N.GroupEnd());
    std::vector<uint32_t> in_ids(n_in);
    if (n_in) CUDA_OK(cudaMemcpyAsync(in_ids.data(), d_in, n_in * 4, cudaMemcpyDeviceToHost, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
    cudaFree(d_want); cudaFree(d_all); cudaFree(d_out); cudaFree(d_in);
    comm_trace(c, "halo plan: id lists received");
    set_send_ids(c, (int32_t)from.size(), from.data(), counts.data(), in_ids.data());
}

}  // namespace

#define API_BEGIN(ctx)  try { if (!(ctx)) throw std::runtime_error("NULL context");
#define API_BEGIN0(ctx) try {
#define API_END(ctx)                                                       \
    return 0; }                                                            \
    catch (const std::exception & e) { if (ctx) (ctx)->err = e.what(); g_err = e.what(); return 1; }

extern "C" {

const char * mlb_version(void) { return "mallard_b200 0.1 (sm_100a)"; }
int mlb_check_abi(int32_t header_abi_version) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (header_abi_version != MLB_ABI_VERSION)
        throw std::runtime_error("ABI mismatch: the caller was built against revision " + std::to_string(header_abi_version) +
                                 " of mallard_b200.h, this library implements revision " + std::to_string(MLB_ABI_VERSION) + " (rebuild the caller)");
    API_END(none)
}
int mlb_set_host_threads(int32_t n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
const char * mlb_last_error(const mlb_ctx * ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int mlb_create(mlb_ctx ** out, const mlb_mesh * mesh, const mlb_numerics * numerics, const mlb_physics * physics,
               const mlb_bc * bcs, int32_t n_bcs, const mlb_parallel * parallel) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out) throw std::runtime_error("mlb_create: out is NULL");
    *out = nullptr;
    *out = create_impl(mesh, nullptr, numerics, physics, bcs, n_bcs, parallel);
    API_END(none)
}

static mlb_ctx * create_with_halo(const mlb_mesh * mesh, const int32_t * part, const mlb_numerics * numerics, const mlb_physics * physics,
                                  const mlb_bc * bcs, int32_t n_bcs, const mlb_parallel * parallel, const mlb_local_mesh * local) {
    mlb_ctx * c = create_impl(mesh, part, numerics, physics, bcs, n_bcs, parallel, local);
    try {
        std::vector<uint32_t> recv_idx;
        halo_recv_lists(c->prep, part, c->peers, c->recv_counts, c->recv_ref_ids, recv_idx);
        c->n_recv = recv_idx.size();
        c->d_recv_idx = c->upload(recv_idx);
        c->d_recv_buf = c->alloc<double>(4 * std::max<size_t>(c->n_recv, 1));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    } catch (...) { delete c; throw; }
    return c;
}

int mlb_create_partitioned(mlb_ctx ** out, const mlb_mesh * mesh, const int32_t * part, const mlb_numerics * numerics,
                           const mlb_physics * physics, const mlb_bc * bcs, int32_t n_bcs, const mlb_parallel * parallel) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out || !part || !parallel) throw std::runtime_error("mlb_create_partitioned: NULL argument");
    *out = nullptr;
    *out = create_with_halo(mesh, part, numerics, physics, bcs, n_bcs, parallel, nullptr);
    API_END(none)
}

int mlb_create_local(mlb_ctx ** out, const mlb_mesh * local_mesh, const int32_t * part_local, const mlb_local_mesh * local,
                     const mlb_numerics * numerics, const mlb_physics * physics, const mlb_bc * bcs, int32_t n_bcs,
                     const mlb_parallel * parallel) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out || !part_local || !parallel || !local) throw std::runtime_error("mlb_create_local: NULL argument");
    *out = nullptr;
    *out = create_with_halo(local_mesh, part_local, numerics, physics, bcs, n_bcs, parallel, local);
    API_END(none)
}

void mlb_destroy(mlb_ctx * ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    delete ctx;
}

int mlb_set_state(mlb_ctx * c, const double * U, const double * prim) {
    API_BEGIN(c)
    if (!U) throw std::runtime_error("mlb_set_state: U is NULL");
    CUDA_OK(cudaSetDevice(c->device));
    import_state(*c, U, c->U[c->cur], 4);
    if (prim) { import_state(*c, prim, c->prim, 5); launch_rho_plane(c->U[c->cur], c->prep.N, c->prep.Npad, c->prim, c->stream); c->launches++; }
    else { c->kt->primitives_soa(c->gas, c->prep.N, c->prep.Npad, c->U[c->cur], c->prim, c->stream); c->launches++; }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}

int mlb_get_state(mlb_ctx * c, double * U, double * prim, double * cfl_local) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (U) export_state(*c, c->U[c->cur], 4, U);
    if (prim) export_state(*c, c->prim, 5, prim);
    if (cfl_local) {
        c->ensure_stage(c->nc_ref);
        if (c->n_ranks > 1) CUDA_OK(cudaMemsetAsync(c->d_stage, 0, (size_t)c->nc_ref * sizeof(double), c->stream));
        launch_export_scaled(c->sr, c->scal, SC_DT, c->d_perm_cells, c->prep.N_owned, c->d_stage, c->stream);
        CUDA_OK(cudaMemcpyAsync(cfl_local, c->d_stage, (size_t)c->nc_ref * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    API_END(c)
}

int mlb_n_face_quadrature_points(const mlb_ctx * c) { return c ? c->prep.Q : 0; }

int mlb_calc_face_values(mlb_ctx * c, double * F_out) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (c->teno) run_recon(*c, c->U[c->cur]);
    if (F_out) {
        const size_t n = (size_t)c->nf_ref * c->prep.Q * 8;
        c->ensure_stage(n);
        CUDA_OK(cudaMemsetAsync(c->d_stage, 0, n * sizeof(double), c->stream));
        launch_export_faces(c->teno ? c->Fc : nullptr, c->U[c->cur], c->g.slot_face, c->d_perm_faces, c->prep.N_owned, c->prep.Npad,
                            c->prep.n_slots, c->prep.Q, c->d_stage, c->stream);
        c->launches++;
        CUDA_OK(cudaMemcpyAsync(F_out, c->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}

int mlb_calc_rhs(mlb_ctx * c, double * rhs_out) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    StagePlan s{}; s.in = c->cur;
    run_stage(*c, s, true, c->k[0]);
    if (rhs_out) export_state(*c, c->k[0], 4, rhs_out);
    else CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}

int mlb_calc_rhs_host(mlb_ctx * c, const double * U_in, double * rhs_out) {
    API_BEGIN(c)
    if (!U_in || !rhs_out) throw std::runtime_error("mlb_calc_rhs_host: NULL buffer");
    CUDA_OK(cudaSetDevice(c->device));
    const int tmp = (c->cur + 1) % 3;
    import_state(*c, U_in, c->U[tmp], 4);
    StagePlan s{}; s.in = tmp;
    run_stage(*c, s, true, c->k[0]);
    export_state(*c, c->k[0], 4, rhs_out);
    API_END(c)
}

int mlb_calc_dt(mlb_ctx * c, double cfl, double * dt_out) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (!(cfl > 0.0)) throw std::runtime_error("mlb_calc_dt: cfl must be positive");
    do_calc_dt(*c, cfl);
    double sc[SC_COUNT];
    read_scalars(*c, sc);
    if (dt_out) *dt_out = sc[SC_DT];
    if (sc[SC_DT] < 0.0) throw std::runtime_error("dt negative: " + std::to_string(sc[SC_DT]) + ".");
    API_END(c)
}

int mlb_set_dt(mlb_ctx * c, double dt) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (dt < 0.0) throw std::runtime_error("dt negative: " + std::to_string(dt) + ".");
    launch_set_scalar(c->scal, SC_DT, dt, c->stream);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}

int mlb_take_step(mlb_ctx * c) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    require_dt(*c);
    do_step(*c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}

int mlb_take_step_host(mlb_ctx * c, double cfl, double * U_inout, double * dt_out) {
    API_BEGIN(c)
    if (!U_inout) throw std::runtime_error("mlb_take_step_host: NULL buffer");
    CUDA_OK(cudaSetDevice(c->device));
    if (!(cfl > 0.0)) require_dt(*c);      // fixed dt: refuse before the resident state is touched
    import_state(*c, U_inout, c->U[c->cur], 4);
    if (cfl > 0.0) {
        c->kt->primitives_soa(c->gas, c->prep.N, c->prep.Npad, c->U[c->cur], c->prim, c->stream); c->launches++;
        do_calc_dt(*c, cfl);
    }
    do_step(*c);
    export_state(*c, c->U[c->cur], 4, U_inout);
    double sc[SC_COUNT];
    read_scalars(*c, sc);                  // 64 bytes behind the state's D2H on the same stream
    if (dt_out) *dt_out = sc[SC_DT];
    if (sc[SC_DT] < 0.0) throw std::runtime_error("dt negative: " + std::to_string(sc[SC_DT]) + ".");   // the step was a no-op
    API_END(c)
}

int mlb_run(mlb_ctx * c, uint32_t n_steps, double cfl, double * t_out, double * dt_last_out) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (c->n_ranks > 1) throw std::runtime_error("mlb_run: partitioned contexts are stepped with mlb_run_distributed or the split-phase API");
    if (!(cfl > 0.0) && n_steps) require_dt(*c);
    auto one_step = [&] {
        if (cfl > 0.0) do_calc_dt(*c, cfl);
        do_step(*c);
    };
    // Small meshes (examples/sod: 1000 cells, examples/wedge: 7500) are launch-bound: 7-10 kernels of a few microseconds per
    // step.  One step is captured into a CUDA graph and replayed; dt, t and the step counter live on the device, and the
    // stage buffers of SSPRK3 / RK4 return to the same rotation after a step, so every replay is the same graph.
    static const bool graphs = [] { const char * e = getenv("MLB_RUN_GRAPH"); return !(e && e[0] == '0'); }();
    uint32_t done = 0;
    if (n_steps && small_step_eligible(*c)) {         // launch-bound first-order meshes: all n_steps in ONE cooperative kernel (opt-in)
        run_small_steps(*c, n_steps, cfl);
        done = n_steps;
    }
    if (done < n_steps && graphs && !c->profiling && n_steps >= 8 && c->num.integrator != MLB_INTEGRATOR_FE) {
        one_step();                                   // eager: function attributes, occupancy queries, error reporting
        done = 1;
        const uint64_t l0 = c->launches;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        bool ok = true;
        try { one_step(); } catch (...) { ok = false; }
        const cudaError_t ec = cudaStreamEndCapture(c->stream, &g);
        const uint64_t per_step = c->launches - l0;
        c->launches = l0;                             // nothing ran during the capture
        if (ok && ec == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
            for (; done < n_steps; done++) { CUDA_OK(cudaGraphLaunch(ge, c->stream)); c->launches += per_step; }
            c->graph_replays += n_steps - 1;
        } else {
            cudaGetLastError();                       // capture not possible here: fall through to plain launches
        }
        if (ge) cudaGraphExecDestroy(ge);
        if (g) cudaGraphDestroy(g);
    }
    for (; done < n_steps; done++) one_step();
    double sc[SC_COUNT];
    read_scalars(*c, sc);
    if (t_out) *t_out = sc[SC_T];
    if (dt_last_out) *dt_last_out = sc[SC_DT];
    if (sc[SC_DT] < 0.0) throw std::runtime_error("dt negative: " + std::to_string(sc[SC_DT]) + ".");
    API_END(c)
}

int mlb_get_time(mlb_ctx * c, double * t, uint64_t * step) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    double sc[SC_COUNT];
    read_scalars(*c, sc);
    if (t) *t = sc[SC_T];
    if (step) { unsigned long long s; CUDA_OK(cudaMemcpy(&s, c->step_counter, 8, cudaMemcpyDeviceToHost)); *step = s; }
    API_END(c)
}

// ---- N3: diagnostics and output fed from the device-resident fields ------------------------------------
// Solver::do_checks' scalar ranges (solver.cpp:434-443) and check_fields' NaN test (solver.cpp:470-498) as one device
// reduction: no per-step copy of the state to the host (the reference does copy_device_to_host() every step, solver.cpp:423).
int mlb_field_ranges(mlb_ctx * c, double * min9, double * max9, uint64_t * n_nan) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    const int nb = field_ranges_blocks(c->prep.N_owned);
    if (!c->d_ranges) {                                   // per-block partials, the 18 results and the NaN counter: allocated once
        c->d_ranges = c->alloc<double>((size_t)nb * 18 + 18 + 1);
    }
    double * d_part = c->d_ranges;
    unsigned long long * d_nan = reinterpret_cast<unsigned long long *>(c->d_ranges + (size_t)nb * 18 + 18);
    CUDA_OK(cudaMemsetAsync(d_nan, 0, 8, c->stream));
    launch_field_ranges(c->U[c->cur], c->prim, c->prep.N_owned, c->prep.Npad, d_part, d_part + (size_t)nb * 18, d_nan, c->stream);
    c->launches += 2;
    CUDA_OK(cudaGetLastError());
    double h[18];
    unsigned long long hn = 0;
    CUDA_OK(cudaMemcpyAsync(h, d_part + (size_t)nb * 18, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(&hn, d_nan, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int f = 0; f < 9; f++) { if (min9) min9[f] = h[f]; if (max9) max9[f] = h[9 + f]; }
    if (n_nan) *n_nan = hn;
    API_END(c)
}

// DataWriter::write_vtu (io/data_writer.cpp:93-244), byte for byte: same XML header, same appended-data blocks (4-byte
// length words, Float64 cell data, 3-component points, UInt32 connectivity / offsets, cell type 7).  The requested
// variables are gathered on the device into one plane each (reference numbering), copied once, and written with one
// fwrite per block instead of one ofstream::write per value.
int mlb_write_vtu(mlb_ctx * c, const mlb_mesh * mesh, const char * prefix, uint32_t step, int32_t n_vars, const char * const * names) {
    API_BEGIN(c)
    if (!mesh || !prefix || !names) throw std::runtime_error("mlb_write_vtu: NULL argument");
    if (n_vars <= 0) throw std::runtime_error("DataWriter: No variables specified.");
    if (n_vars > 16) throw std::runtime_error("mlb_write_vtu: at most 16 variables");
    if (c->n_ranks > 1) throw std::runtime_error("mlb_write_vtu: partitioned contexts write through mlb_get_owned");
    if (mesh->n_cells != c->nc_ref) throw std::runtime_error("mlb_write_vtu: mesh does not match the context");
    static const char * const NAMES[10] = {"RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X", "U_Y", "P", "T", "H", "CFL"};   // common_typedef.h:36-49, solver.cpp:349
    int32_t codes[16];
    for (int v = 0; v < n_vars; v++) {
        codes[v] = -1;
        for (int k = 0; k < 10; k++) if (!strcmp(names[v], NAMES[k])) codes[v] = k;
        if (codes[v] < 0) throw std::runtime_error(std::string("DataWriter: Unknown variable: ") + names[v] + ".");
    }
    CUDA_OK(cudaSetDevice(c->device));
    const uint32_t nc = mesh->n_cells, nn = mesh->n_nodes;
    const size_t n = (size_t)n_vars * nc;
    c->ensure_stage(n);
    launch_export_fields(n_vars, codes, c->U[c->cur], c->prim, c->sr, c->scal, c->d_perm_cells, c->prep.N_owned, c->prep.Npad, nc, c->d_stage, c->stream);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(c->h_stage, c->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));

    char fname[4096];
    snprintf(fname, sizeof(fname), "%s_%06u.vtu", prefix, step);          // LEN_STEP = 6, common_io.h:19
    FILE * out = fopen(fname, "wb");
    if (!out) throw std::runtime_error(std::string("DataWriter::write_vtu: Could not open file: ") + fname + ".");
    const uint64_t len_conn = mesh->offsets_nodes_of_cell[nc] - mesh->offsets_nodes_of_cell[0];
    uint64_t off = 0;
    fprintf(out, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
    fprintf(out, "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"%u\" NumberOfCells=\"%u\">\n", nn, nc);
    fprintf(out, "      <PointData>\n      </PointData>\n      <CellData>\n");
    for (int v = 0; v < n_vars; v++) {
        fprintf(out, "        <DataArray type=\"Float64\" Name=\"%s\" format=\"appended\" offset=\"%llu\">\n        </DataArray>\n", names[v], (unsigned long long)off);
        off += 4 + (uint64_t)nc * 8;
    }
    fprintf(out, "      </CellData>\n      <Points>\n");
    fprintf(out, "        <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"appended\" offset=\"%llu\">\n        </DataArray>\n", (unsigned long long)off);
    off += 4 + (uint64_t)nn * 3 * 8;
    fprintf(out, "      </Points>\n      <Cells>\n");
    fprintf(out, "        <DataArray type=\"UInt32\" Name=\"connectivity\" format=\"appended\" offset=\"%llu\">\n        </DataArray>\n", (unsigned long long)off);
    off += 4 + len_conn * 4;
    fprintf(out, "        <DataArray type=\"UInt32\" Name=\"offsets\" format=\"appended\" offset=\"%llu\">\n        </DataArray>\n", (unsigned long long)off);
    off += 4 + (uint64_t)nc * 4;
    fprintf(out, "        <DataArray type=\"UInt8\" Name=\"types\" format=\"appended\" offset=\"%llu\">\n        </DataArray>\n", (unsigned long long)off);
    fprintf(out, "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n<AppendedData encoding=\"raw\">\n_");
    auto block = [&](uint64_t n_bytes, const void * data) {               // the reference writes the low 4 bytes of a u_int64_t length
        fwrite(&n_bytes, 4, 1, out);
        if (n_bytes) fwrite(data, 1, n_bytes, out);
    };
    for (int v = 0; v < n_vars; v++) block((uint64_t)nc * 8, c->h_stage + (size_t)v * nc);
    {
        std::vector<double> pts((size_t)nn * 3);
        for (uint32_t i = 0; i < nn; i++) { pts[3 * (size_t)i] = mesh->node_coords[2 * (size_t)i]; pts[3 * (size_t)i + 1] = mesh->node_coords[2 * (size_t)i + 1]; pts[3 * (size_t)i + 2] = 0.0; }
        block((uint64_t)nn * 3 * 8, pts.data());
    }
    block(len_conn * 4, mesh->nodes_of_cell + mesh->offsets_nodes_of_cell[0]);
    {
        std::vector<uint32_t> offs(nc);
        for (uint32_t i = 0; i < nc; i++) offs[i] = mesh->offsets_nodes_of_cell[i + 1] - mesh->offsets_nodes_of_cell[0];
        block((uint64_t)nc * 4, offs.data());
    }
    {
        std::vector<uint8_t> types(nc, 7);
        block((uint64_t)nc, types.data());
    }
    fprintf(out, "\n</AppendedData>\n</VTKFile>\n");
    if (fclose(out) != 0) throw std::runtime_error(std::string("DataWriter::write_vtu: Could not write to file: ") + fname + ".");
    API_END(c)
}

int mlb_set_rhs_override(mlb_ctx * c, const double * rhs) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (!rhs) { c->has_override = false; return 0; }
    if (!c->k_override) c->k_override = c->alloc<double>(4 * (size_t)c->prep.Npad);
    import_state(*c, rhs, c->k_override, 4);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->has_override = true;
    API_END(c)
}

int mlb_get_array(mlb_ctx * c, const char * name, void * out, uint64_t * nbytes) {
    API_BEGIN(c)
    if (!name || !nbytes) throw std::runtime_error("mlb_get_array: NULL argument");
    CUDA_OK(cudaSetDevice(c->device));
    const std::string n = name;
    const Prep & P = c->prep;
    const TenoTables & T = P.teno;
    auto host = [&](const void * p, size_t nb) { if (out) std::memcpy(out, p, nb); *nbytes = nb; };
    if (n == "perm_cells") host(P.perm_cells.data(), P.perm_cells.size() * 4);
    else if (n == "perm_faces") host(P.perm_faces.data(), P.perm_faces.size() * 4);
    else if (n.size() == 4 && n.rfind("rhs", 0) == 0) {
        const int r = n[3] - '0';
        if (r < 0 || r >= c->n_rhs) throw std::runtime_error("no such stage residual: " + n);
        *nbytes = (uint64_t)c->nc_ref * 32;
        if (out) export_state(*c, c->k[r], 4, (double *)out);
    } else if (n == "U_temp") {
        *nbytes = (uint64_t)c->nc_ref * 32;
        if (out) export_state(*c, c->U[c->last_temp], 4, (double *)out);
    } else if (n == "cfl_local") {
        *nbytes = (uint64_t)c->nc_ref * 8;
        if (out) { if (mlb_get_state(c, nullptr, nullptr, (double *)out)) throw std::runtime_error(c->err); }
    } else if (n == "dev:fm_mat" || n == "dev:fm_area0" || n == "dev:fm_ids") {   // the device-resident compact tables, as they are
        if (!c->streaming) throw std::runtime_error("context has no streaming TENO tables");
        const int KR = T.K - 1, MC = T.M - 1;
        const size_t frow = (size_t)(2 * (MC / 2) + 1) * FAST_CT;
        const void * src = n == "dev:fm_mat" ? (const void *)c->d_fm_mat : n == "dev:fm_area0" ? (const void *)c->d_fm_area0 : (const void *)c->d_fm_ids;
        *nbytes = n == "dev:fm_mat" ? (uint64_t)c->n_ftiles * FAST_S * KR * frow * 8 : n == "dev:fm_area0" ? (uint64_t)c->n_ftiles * FAST_CT * 8
                                    : (uint64_t)c->n_ftiles * FAST_S * MC * FAST_CT * 4;
        if (out) CUDA_OK(cudaMemcpy(out, src, *nbytes, cudaMemcpyDeviceToHost));
    } else if (n == "stats") {
        double s[13] = {(double)c->launches, P.seconds, (double)c->device_bytes, (double)P.N, (double)P.N_owned, (double)P.NF,
                        (double)P.N_recon, (double)c->n_stages, P.seconds_stencils, P.seconds_matrices, c->table_build_ms * 1e-3, (double)c->graph_replays,
                        (double)c->small_steps};
        host(s, sizeof(s));
    } else if (n.rfind("teno:", 0) == 0) {
        if (!c->teno) throw std::runtime_error("context has no TENO tables");
        if (n == "teno:integral_psi_target") host(T.psi_bar.data(), T.psi_bar.size() * 8);
        else if (n == "teno:oscillation_indicator") host(T.OI.data(), T.OI.size() * 8);
        else if (n == "teno:poly_indices") host(T.pidx.data(), T.pidx.size());
        else {
            if (!T.keep_ref) throw std::runtime_error("reference-layout TENO tables were not kept for this context (mesh too large or partitioned)");
            if (n == "teno:offsets_stencil_groups") host(T.ref_off_groups.data(), T.ref_off_groups.size() * 4);
            else if (n == "teno:offsets_stencils") host(T.ref_off_stencils.data(), T.ref_off_stencils.size() * 4);
            else if (n == "teno:stencils") host(T.ref_stencils.data(), T.ref_stencils.size() * 4);
            else if (n == "teno:offsets_reconstruction_matrices") host(T.ref_off_mats.data(), T.ref_off_mats.size() * 4);
            else if (n == "teno:reconstruction_matrices") host(T.ref_mats.data(), T.ref_mats.size() * 8);
            else if (n == "teno:transformed_areas") host(T.ref_areas.data(), T.ref_areas.size() * 8);
            else throw std::runtime_error("unknown array " + n);
        }
    } else throw std::runtime_error("unknown array " + n);
    API_END(c)
}

int mlb_event_record(mlb_ctx * c, int32_t slot) {
    API_BEGIN(c)
    if (slot < 0 || slot >= 16) throw std::runtime_error("event slot out of range");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaEventRecord(c->ev[slot], c->stream));
    API_END(c)
}
int mlb_event_elapsed_ms(mlb_ctx * c, int32_t a, int32_t b, float * ms) {
    API_BEGIN(c)
    if (a < 0 || a >= 16 || b < 0 || b >= 16 || !ms) throw std::runtime_error("bad event query");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaEventSynchronize(c->ev[b]));
    CUDA_OK(cudaEventElapsedTime(ms, c->ev[a], c->ev[b]));
    API_END(c)
}
int mlb_profile_enable(mlb_ctx * c, int32_t on) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    c->flush_profile();
    if (on && !c->profiling) c->profile.clear();
    c->profiling = on != 0;
    API_END(c)
}
int mlb_profile_read(mlb_ctx * c, int32_t max_entries, const char ** names, double * ms, uint64_t * launches) {
    if (!c) return 0;
    try {
        cudaSetDevice(c->device);
        c->flush_profile();
        c->profile_names.clear();
        for (auto & kv : c->profile) c->profile_names.push_back(kv.first);
        int i = 0;
        for (auto & kv : c->profile) {
            if (i >= max_entries) break;
            if (names) names[i] = c->profile_names[i].c_str();
            if (ms) ms[i] = kv.second.ms;
            if (launches) launches[i] = kv.second.launches;
            i++;
        }
        return i;
    } catch (const std::exception & e) { c->err = e.what(); return -1; }
}
uint64_t mlb_launch_count(const mlb_ctx * c) { return c ? c->launches : 0; }
int mlb_synchronize(mlb_ctx * c) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (c->comm_stream) CUDA_OK(cudaStreamSynchronize(c->comm_stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}
void * mlb_stream(mlb_ctx * c) { return c ? (void *)c->stream : nullptr; }
void * mlb_comm_stream(mlb_ctx * c) { return c ? (void *)(c->comm_stream ? c->comm_stream : c->stream) : nullptr; }

// ---- multi-GPU -------------------------------------------------------------------------------------------------
int mlb_halo_info(mlb_ctx * c, int32_t * n_peers, int32_t * peers, uint64_t * send_counts, uint64_t * recv_counts) {
    API_BEGIN(c)
    if (n_peers) *n_peers = (int32_t)c->peers.size();
    for (size_t i = 0; i < c->peers.size(); i++) {
        if (peers) peers[i] = c->peers[i];
        if (recv_counts) recv_counts[i] = c->recv_counts[i];
        if (send_counts) send_counts[i] = i < c->send_counts.size() ? c->send_counts[i] : 0;
    }
    API_END(c)
}
int mlb_halo_recv_ids(mlb_ctx * c, int32_t peer_index, uint32_t * ref_ids_out) {
    API_BEGIN(c)
    if (peer_index < 0 || peer_index >= (int)c->recv_ref_ids.size()) throw std::runtime_error("bad peer index");
    const auto & v = c->recv_ref_ids[peer_index];
    if (ref_ids_out) for (size_t i = 0; i < v.size(); i++) ref_ids_out[i] = to_global(*c, v[i]);
    API_END(c)
}
int mlb_halo_set_send_ids(mlb_ctx * c, int32_t n_lists, const int32_t * peer_ranks, const uint64_t * counts, const uint32_t * ref_ids) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    set_send_ids(*c, n_lists, peer_ranks, counts, ref_ids);
    API_END(c)
}
int mlb_halo_buffers(mlb_ctx * c, void ** send_dev, void ** recv_dev) {
    API_BEGIN(c)
    if (send_dev) *send_dev = c->d_send_buf;
    if (recv_dev) *recv_dev = c->d_recv_buf;
    API_END(c)
}
static int stage_input_buffer(mlb_ctx * c, int32_t stage) {
    const auto plan = make_plan(*c);
    if (stage < 0 || stage >= (int)plan.size()) throw std::runtime_error("stage out of range");
    return plan[stage].in;
}
int mlb_halo_pack(mlb_ctx * c, int32_t stage) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    halo_pack(*c, stage_input_buffer(c, stage));
    API_END(c)
}
int mlb_halo_unpack(mlb_ctx * c, int32_t stage) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    halo_unpack(*c, stage_input_buffer(c, stage), stage == 0);
    API_END(c)
}
int mlb_n_stages(const mlb_ctx * c) { return c ? c->n_stages : 0; }
int mlb_stage(mlb_ctx * c, int32_t stage) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    const auto plan = make_plan(*c);
    if (stage < 0 || stage >= (int)plan.size()) throw std::runtime_error("stage out of range");
    if (c->stage_begun >= 0 && c->stage_begun != stage) throw std::runtime_error("mlb_stage: a different stage was begun with mlb_stage_begin");
    run_stage(*c, plan[stage], false, nullptr);
    if (stage == (int)plan.size() - 1) finish_plan(*c, plan);
    API_END(c)
}
int mlb_stage_begin(mlb_ctx * c, int32_t stage) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    const auto plan = make_plan(*c);
    if (stage < 0 || stage >= (int)plan.size()) throw std::runtime_error("stage out of range");
    if (c->stage_begun >= 0) throw std::runtime_error("mlb_stage_begin: the previous stage has not been finished with mlb_stage");
    if (c->teno && !c->has_override && interior_tiles(*c)) {
        run_recon(*c, c->U[plan[stage].in], 1);
        c->stage_begun = stage;
    }
    API_END(c)
}
int mlb_local_max_spectral_radius(mlb_ctx * c, double * max_out) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    do_calc_dt(*c, -1.0);
    if (max_out) {   // NULL: stay asynchronous, the value is left in the device scalar block (mlb_scalars_device()[2])
        double sc[SC_COUNT];
        read_scalars(*c, sc);
        *max_out = sc[SC_MAX_SR];
    }
    API_END(c)
}
void * mlb_scalars_device(mlb_ctx * c) { return c ? (void *)c->scal : nullptr; }
int mlb_apply_dt_device(mlb_ctx * c, double cfl) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    launch_apply_dt(c->scal, c->max_bits, cfl, 0.0, 0, c->stream);
    c->launches++;
    API_END(c)
}
int mlb_set_owned(mlb_ctx * c, const double * U_owned) {
    API_BEGIN(c)
    if (!U_owned) throw std::runtime_error("mlb_set_owned: NULL buffer");
    CUDA_OK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->prep.N_owned * 4;
    c->ensure_stage(std::max<size_t>(n, 1));
    CUDA_OK(cudaMemcpyAsync(c->d_stage, U_owned, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    launch_import_state(c->d_stage, nullptr, c->prep.N_owned, c->prep.Npad, 4, c->U[c->cur], c->stream);
    c->launches++;
    // primitives of the owned cells (Solver::update_primitives); ghosts get theirs at the stage-0 unpack
    c->kt->primitives_soa(c->gas, c->prep.N_owned, c->prep.Npad, c->U[c->cur], c->prim, c->stream);
    c->launches++;
    API_END(c)
}
int mlb_get_owned(mlb_ctx * c, double * U_owned) {
    API_BEGIN(c)
    if (!U_owned) throw std::runtime_error("mlb_get_owned: NULL buffer");
    CUDA_OK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->prep.N_owned * 4;
    c->ensure_stage(std::max<size_t>(n, 1));
    launch_export_state(c->U[c->cur], nullptr, c->prep.N_owned, c->prep.Npad, 4, c->d_stage, c->stream);
    c->launches++;
    CUDA_OK(cudaMemcpyAsync(U_owned, c->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    API_END(c)
}
int mlb_apply_dt(mlb_ctx * c, double cfl, double global_max) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    launch_apply_dt(c->scal, c->max_bits, cfl, global_max, 1, c->stream);
    c->launches++;
    API_END(c)
}
int mlb_finish_step(mlb_ctx * c) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    if (c->comm_stream) CUDA_OK(cudaStreamSynchronize(c->comm_stream));
    require_dt(*c);                        // synchronises the compute stream
    API_END(c)
}
int mlb_owned_cells(mlb_ctx * c, uint32_t * n_owned, uint32_t * cells_out) {
    API_BEGIN(c)
    if (n_owned) *n_owned = c->prep.N_owned;
    if (cells_out) std::memcpy(cells_out, c->prep.perm_cells.data(), (size_t)c->prep.N_owned * 4);
    API_END(c)
}

// ---- native multi-GPU driver ------------------------------------------------------------------------------------------------
int mlb_comm_unique_id(void * id_out) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!id_out) throw std::runtime_error("mlb_comm_unique_id: NULL argument");
    static_assert(sizeof(ncclUniqueId) == MLB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_OK(nccl().GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    API_END(none)
}

int mlb_comm_init(mlb_ctx * c, const void * id) {
    API_BEGIN(c)
    if (!id) throw std::runtime_error("mlb_comm_init: NULL id");
    if (c->n_ranks < 2 || !c->comm_stream) throw std::runtime_error("mlb_comm_init: the context is not partitioned");
    if (c->nccl_p2p) throw std::runtime_error("mlb_comm_init: the context already has a communicator");
    CUDA_OK(cudaSetDevice(c->device));
    const NcclApi & N = nccl();
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    comm_trace(*c, "ncclCommInitRank");
    NCCL_OK(N.CommInitRank(&c->nccl_p2p, c->n_ranks, uid, c->rank));
    comm_trace(*c, "ncclCommSplit (communicator of the dt all-reduce)");
    NCCL_OK(N.CommSplit(c->nccl_p2p, 0, c->rank, &c->nccl_coll, nullptr));
    exchange_halo_plan(*c);
    comm_trace(*c, "communicators and halo plan ready");
    API_END(c)
}

int mlb_run_distributed(mlb_ctx * c, uint32_t n_steps, double cfl, double * t_out, double * dt_last_out) {
    API_BEGIN(c)
    CUDA_OK(cudaSetDevice(c->device));
    require_comm(*c);
    if (!(cfl > 0.0) && n_steps) require_dt(*c);
    // As in mlb_run: one eager step (NCCL sets up its connections, kernels get their attributes), then the step is captured -
    // both streams, the grouped send/recv and the all-reduce included - and replayed as one CUDA graph per step.
    static const bool graphs = [] { const char * e = getenv("MLB_RUN_GRAPH"); return !(e && e[0] == '0'); }();
    uint32_t done = 0;
    if (graphs && !c->profiling && n_steps >= 4 && c->num.integrator != MLB_INTEGRATOR_FE) {
        comm_trace(*c, "run: eager step");
        do_step_distributed(*c, cfl);
        done = 1;
        const uint64_t l0 = c->launches;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        comm_trace(*c, "run: capturing a step");
        CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        bool ok = true;
        std::string why;
        try { do_step_distributed(*c, cfl); } catch (const std::exception & e) { ok = false; why = e.what(); }
        const cudaError_t ec = cudaStreamEndCapture(c->stream, &g);
        comm_trace(*c, "run: step captured, replaying");
        const uint64_t per_step = c->launches - l0;
        c->launches = l0;                             // nothing ran during the capture
        if (ok && ec == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
            for (; done < n_steps; done++) { CUDA_OK(cudaGraphLaunch(ge, c->stream)); c->launches += per_step; }
            c->graph_replays += n_steps - 1;
        } else {
            cudaGetLastError();
            c->halo_pending = false; c->stage_begun = -1;
            if (ge) cudaGraphExecDestroy(ge);
            if (g) cudaGraphDestroy(g);
            // every rank must issue the same collectives: a rank that cannot capture cannot silently fall back on its own
            throw std::runtime_error("mlb_run_distributed: the step could not be captured into a CUDA graph (" +
                                     (why.empty() ? std::string(cudaGetErrorString(ec)) : why) + "); set MLB_RUN_GRAPH=0 on every rank");
        }
        if (ge) cudaGraphExecDestroy(ge);
        if (g) cudaGraphDestroy(g);
    }
    for (; done < n_steps; done++) do_step_distributed(*c, cfl);
    comm_trace(*c, "run: all steps enqueued, waiting for the device");
    if (c->comm_stream) CUDA_OK(cudaStreamSynchronize(c->comm_stream));
    double sc[SC_COUNT];
    read_scalars(*c, sc);
    comm_trace(*c, "run: done");
    if (t_out) *t_out = sc[SC_T];
    if (dt_last_out) *dt_last_out = sc[SC_DT];
    if (sc[SC_DT] < 0.0) throw std::runtime_error("dt negative: " + std::to_string(sc[SC_DT]) + ".");
    API_END(c)
}

int mlb_take_step_distributed_host(mlb_ctx * c, double cfl, double * U_owned_inout, double * dt_out) {
    API_BEGIN(c)
    if (!U_owned_inout) throw std::runtime_error("mlb_take_step_distributed_host: NULL buffer");
    CUDA_OK(cudaSetDevice(c->device));
    require_comm(*c);
    if (!(cfl > 0.0)) require_dt(*c);
    const size_t n = (size_t)c->prep.N_owned * 4;
    // owned cells are the first N_owned rows of the AoS state: the host buffer is copied straight into / out of it
    CUDA_OK(cudaMemcpyAsync(c->U[c->cur], U_owned_inout, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->kt->primitives_soa(c->gas, c->prep.N_owned, c->prep.Npad, c->U[c->cur], c->prim, c->stream);
    c->launches++;
    do_step_distributed(*c, cfl);
    CUDA_OK(cudaMemcpyAsync(U_owned_inout, c->U[c->cur], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    double sc[SC_COUNT];
    read_scalars(*c, sc);
    if (dt_out) *dt_out = sc[SC_DT];
    if (sc[SC_DT] < 0.0) throw std::runtime_error("dt negative: " + std::to_string(sc[SC_DT]) + ".");
    API_END(c)
}

// Recursive coordinate bisection on cell centroids (deterministic: the order is (coordinate, cell id), a strict total order,
// so the two halves of every cut are uniquely defined whatever the selection algorithm visits first).
static void rcb(std::vector<std::pair<double, uint32_t>> & key, const double * xy, uint32_t * ids, size_t lo, size_t hi, int32_t p0, int32_t np,
                int32_t * part_out) {
    if (np == 1) { for (size_t i = lo; i < hi; i++) part_out[ids[i]] = p0; return; }
    double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
    for (size_t i = lo; i < hi; i++)
        for (int d = 0; d < 2; d++) { const double x = xy[2 * (size_t)ids[i] + d]; mn[d] = std::min(mn[d], x); mx[d] = std::max(mx[d], x); }
    const int d = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;
    const int32_t npl = np / 2;
    const size_t mid = lo + (size_t)((double)(hi - lo) * npl / np);
    for (size_t i = lo; i < hi; i++) key[i] = {xy[2 * (size_t)ids[i] + d], ids[i]};     // contiguous keys: the selection stays in cache lines
    std::nth_element(key.begin() + lo, key.begin() + mid, key.begin() + hi);
    for (size_t i = lo; i < hi; i++) ids[i] = key[i].second;
    // the two halves touch disjoint ranges of key / ids and disjoint cells of part_out: independent tasks
#pragma omp task shared(key) if (hi - lo > 200000)
    rcb(key, xy, ids, lo, mid, p0, npl, part_out);
    rcb(key, xy, ids, mid, hi, p0 + npl, np - npl, part_out);
#pragma omp taskwait
}

int mlb_partition_coords(uint64_t n_cells, const double * cell_xy, int32_t n_parts, int32_t * part_out) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!cell_xy || !part_out || n_parts < 1 || n_cells > 0xFFFFFFFEull) throw std::runtime_error("mlb_partition_coords: bad argument");
    std::vector<uint32_t> ids(n_cells);
    for (uint64_t i = 0; i < n_cells; i++) ids[i] = (uint32_t)i;
    std::vector<std::pair<double, uint32_t>> key(n_cells);
#pragma omp parallel
#pragma omp single
    rcb(key, cell_xy, ids.data(), 0, n_cells, 0, n_parts, part_out);
    API_END(none)
}

int mlb_partition(const mlb_mesh * mesh, int32_t n_parts, int32_t * part_out) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!mesh || !part_out || n_parts < 1) throw std::runtime_error("mlb_partition: bad argument");
    if (mesh->cell_coords) return mlb_partition_coords(mesh->n_cells, mesh->cell_coords, n_parts, part_out);
    HostMesh hm;
    host_mesh_from_view(hm, *mesh);                    // centroids as Mesh::compute_cell_centroids makes them
    return mlb_partition_coords(hm.nc, hm.cell_xy.data(), n_parts, part_out);
    API_END(none)
}

// Graph partition (partition_graph.cpp): multilevel recursive bisection of the cell-face dual graph, no coordinates involved
int mlb_partition_graph_csr(uint32_t n, const uint64_t * xadj, const uint32_t * adj, int32_t n_parts, int32_t * part_out) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!xadj || (!adj && xadj[n]) || !part_out || n_parts < 1 || n > 0xFFFFFFFEu) throw std::runtime_error("mlb_partition_graph_csr: bad argument");
    for (uint32_t v = 0; v < n; v++) if (xadj[v + 1] < xadj[v]) throw std::runtime_error("mlb_partition_graph_csr: xadj must be non-decreasing");
    graph_partition(n, xadj, adj, n_parts, part_out);
    API_END(none)
}

int mlb_partition_graph(const mlb_mesh * mesh, int32_t n_parts, int32_t * part_out) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!mesh || !mesh->cells_of_face || !part_out || n_parts < 1) throw std::runtime_error("mlb_partition_graph: bad argument");
    std::vector<uint64_t> xadj;
    std::vector<uint32_t> adj;
    dual_graph(mesh->n_cells, mesh->n_faces, mesh->cells_of_face, xadj, adj);
    graph_partition(mesh->n_cells, xadj.data(), adj.data(), n_parts, part_out);
    API_END(none)
}

// ---- stateless kernels -------------------------------------------------------------------------------------------
int mlb_riemann_flux(int32_t device, int32_t riemann, int32_t fp_mode, uint64_t n, const double * n_unit, const double * L,
                     const double * R, double gamma, double * flux) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    require_device(device);
    if (riemann < 0 || riemann > 2) throw std::runtime_error("Unknown Riemann solver type.");
    const KernelTable * kt = fp_mode == MLB_FP_FAST ? kernels_fast() : kernels_strict();
    double * dn = dev_alloc<double>(2 * n), * dl = dev_alloc<double>(5 * n), * dr = dev_alloc<double>(5 * n), * df = dev_alloc<double>(4 * n);
    CUDA_OK(cudaMemcpy(dn, n_unit, 16 * n, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dl, L, 40 * n, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dr, R, 40 * n, cudaMemcpyHostToDevice));
    kt->riemann_flux(riemann, n, dn, dl, dr, gamma, df, 0);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(flux, df, 32 * n, cudaMemcpyDeviceToHost));
    cudaFree(dn); cudaFree(dl); cudaFree(dr); cudaFree(df);
    API_END(none)
}

int mlb_compute_primitives(int32_t device, int32_t fp_mode, const mlb_physics * physics, uint64_t n, const double * U, double * prim,
                           double * R_cp_cv) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!physics) throw std::runtime_error("physics is NULL");
    const GasParams g = make_gas(*physics);
    if (R_cp_cv) { R_cp_cv[0] = g.R; R_cp_cv[1] = g.cp; R_cp_cv[2] = g.cv; }
    if (n) {
        require_device(device);
        const KernelTable * kt = fp_mode == MLB_FP_FAST ? kernels_fast() : kernels_strict();
        double * du = dev_alloc<double>(4 * n), * dp = dev_alloc<double>(5 * n);
        CUDA_OK(cudaMemcpy(du, U, 32 * n, cudaMemcpyHostToDevice));
        kt->primitives(g, n, du, dp, 0);
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpy(prim, dp, 40 * n, cudaMemcpyDeviceToHost));
        cudaFree(du); cudaFree(dp);
    }
    API_END(none)
}

// ---- host-only preprocessing plan (no device needed) ------------------------------------------------------------
static mlb_plan * plan_create(const mlb_mesh * mesh, const mlb_numerics * numerics, const mlb_bc * bcs, int32_t n_bcs,
                              const int32_t * part, const mlb_parallel * parallel, const mlb_local_mesh * local) {
    if (!mesh || !numerics) throw std::runtime_error("mlb_plan_create: NULL argument");
    std::unique_ptr<mlb_plan> p(new mlb_plan());
    HostMesh hm;
    host_mesh_from_view(hm, *mesh);
    p->nc_ref = hm.nc; p->nf_ref = hm.nf;
    std::vector<std::string> zones;
    for (int b = 0; b < n_bcs; b++) zones.push_back(bcs[b].zone_name ? bcs[b].zone_name : "");
    PrepOptions opt;
    opt.renumber = numerics->renumber;
    opt.keep_ref_tables = numerics->recon == MLB_RECON_TENO && !part;
    opt.fast_tables = numerics->recon == MLB_RECON_TENO && numerics->fp_mode == MLB_FP_FAST;
    opt.part = part; opt.rank = parallel ? parallel->rank : 0; opt.n_ranks = parallel ? parallel->n_ranks : 1;
    if (local) opt.psi_ref_tri = local->cell0_nodes;
    p->rank = opt.rank;
    if (part) p->part.assign(part, part + hm.nc);
    preprocess(hm, *numerics, zones, opt, p->prep);
    return p.release();
}
int mlb_plan_create(mlb_plan ** out, const mlb_mesh * mesh, const mlb_numerics * numerics, const mlb_bc * bcs, int32_t n_bcs,
                    const int32_t * part, const mlb_parallel * parallel) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out) throw std::runtime_error("mlb_plan_create: NULL argument");
    *out = nullptr;
    *out = plan_create(mesh, numerics, bcs, n_bcs, part, parallel, nullptr);
    API_END(none)
}
int mlb_plan_create_local(mlb_plan ** out, const mlb_mesh * local_mesh, const mlb_numerics * numerics, const mlb_bc * bcs, int32_t n_bcs,
                          const int32_t * part_local, const mlb_parallel * parallel, const mlb_local_mesh * local) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out || !part_local || !parallel || !local) throw std::runtime_error("mlb_plan_create_local: NULL argument");
    *out = nullptr;
    *out = plan_create(local_mesh, numerics, bcs, n_bcs, part_local, parallel, local);
    API_END(none)
}
int mlb_plan_get(mlb_plan * p, const char * name, void * out, uint64_t * nbytes) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!p || !name || !nbytes) throw std::runtime_error("mlb_plan_get: NULL argument");
    const std::string n = name;
    const Prep & P = p->prep;
    const TenoTables & T = P.teno;
    auto host = [&](const void * q, size_t nb) { if (out) std::memcpy(out, q, nb); *nbytes = nb; };
    if (n == "sizes") {
        uint32_t s[12] = {P.N, P.N_owned, P.N_recon, P.NF, (uint32_t)P.n_slots, (uint32_t)P.Q, (uint32_t)T.K, (uint32_t)T.M, P.Npad,
                          (uint32_t)T.S, (uint32_t)T.Mp, (uint32_t)FAST_CT};
        host(s, sizeof(s));
    }
    else if (n == "seconds") host(&P.seconds, 8);
    else if (n == "n_interior") host(&P.N_interior, 4);
    else if (n == "perm_cells") host(P.perm_cells.data(), P.perm_cells.size() * 4);
    else if (n == "perm_faces") host(P.perm_faces.data(), P.perm_faces.size() * 4);
    else if (n == "slot_face") host(P.slot_face.data(), P.slot_face.size() * 4);
    else if (n == "slot_nbr") host(P.slot_nbr.data(), P.slot_nbr.size() * 4);
    else if (n == "rhs_order") host(P.rhs_order.data(), P.rhs_order.size());
    else if (n == "st_ids") host(T.st_ids.data(), T.st_ids.size() * 4);
    else if (n == "fm_ids") host(T.fm_ids.data(), T.fm_ids.size() * 4);
    else if (n == "fm_mat") host(T.fm_mat.data(), T.fm_mat.size() * 8);
    else if (n == "fm_area0") host(T.fm_area0.data(), T.fm_area0.size() * 8);
    else if (n == "OIs") host(T.OIs.data(), T.OIs.size() * 8);
    else if (n == "halo_peers" || n == "halo_recv_counts" || n == "halo_recv_ids") {
        if (p->part.empty()) throw std::runtime_error("plan is not partitioned");
        std::vector<int32_t> peers; std::vector<uint64_t> counts; std::vector<std::vector<uint32_t>> ids; std::vector<uint32_t> idx;
        halo_recv_lists(P, p->part.data(), peers, counts, ids, idx);
        if (n == "halo_peers") host(peers.data(), peers.size() * 4);
        else if (n == "halo_recv_counts") host(counts.data(), counts.size() * 8);
        else { std::vector<uint32_t> flat; for (auto & v : ids) flat.insert(flat.end(), v.begin(), v.end()); host(flat.data(), flat.size() * 4); }
    }
    else if (n == "ghost_owner") {
        std::vector<int32_t> o;
        for (uint32_t i = P.N_owned; i < P.N; i++) o.push_back(p->part.empty() ? 0 : p->part[P.perm_cells[i]]);
        host(o.data(), o.size() * 4);
    }
    else if (n == "teno:integral_psi_target") host(T.psi_bar.data(), T.psi_bar.size() * 8);
    else if (n == "teno:oscillation_indicator") host(T.OI.data(), T.OI.size() * 8);
    else if (n == "teno:poly_indices") host(T.pidx.data(), T.pidx.size());
    else if (n == "teno:offsets_stencil_groups") host(T.ref_off_groups.data(), T.ref_off_groups.size() * 4);
    else if (n == "teno:offsets_stencils") host(T.ref_off_stencils.data(), T.ref_off_stencils.size() * 4);
    else if (n == "teno:stencils") host(T.ref_stencils.data(), T.ref_stencils.size() * 4);
    else if (n == "teno:offsets_reconstruction_matrices") host(T.ref_off_mats.data(), T.ref_off_mats.size() * 4);
    else if (n == "teno:reconstruction_matrices") host(T.ref_mats.data(), T.ref_mats.size() * 8);
    else if (n == "teno:transformed_areas") host(T.ref_areas.data(), T.ref_areas.size() * 8);
    else throw std::runtime_error("unknown plan array " + n);
    API_END(none)
}
void mlb_plan_destroy(mlb_plan * p) { delete p; }

// ---- host mesh generators ---------------------------------------------------------------------------------------
int mlb_host_mesh_generate(mlb_host_mesh ** out, int32_t type, uint32_t nx, uint32_t ny, double Lx, double Ly) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out) throw std::runtime_error("out is NULL");
    std::unique_ptr<mlb_host_mesh> m(new mlb_host_mesh());
    host_mesh_generate(m->m, type, nx, ny, Lx, Ly);
    *out = m.release();
    API_END(none)
}
int mlb_host_mesh_from_arrays(mlb_host_mesh ** out, const mlb_mesh * mesh) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out || !mesh) throw std::runtime_error("NULL argument");
    std::unique_ptr<mlb_host_mesh> m(new mlb_host_mesh());
    mlb_mesh v = *mesh;
    v.cell_coords = nullptr;   // geometry is always recomputed here (Mesh::compute_cell_centroids/volumes/face_areas/normals, mesh/mesh.cpp:167-261)
    host_mesh_from_view(m->m, v);
    *out = m.release();
    API_END(none)
}
int mlb_host_mesh_read_gmsh(mlb_host_mesh ** out, const char * path) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out || !path) throw std::runtime_error("NULL argument");
    std::unique_ptr<mlb_host_mesh> m(new mlb_host_mesh());
    host_mesh_read_gmsh(m->m, path);
    *out = m.release();
    API_END(none)
}
int mlb_host_mesh_from_cells(mlb_host_mesh ** out, uint32_t n_nodes, const double * node_coords, uint32_t n_cells, const uint32_t * offsets_nodes_of_cell,
                             const uint32_t * nodes_of_cell, uint32_t n_boundary_edges, const uint32_t * edge_nodes, const int32_t * edge_tags,
                             uint32_t n_names, const int32_t * name_tags, const char * const * names) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!out) throw std::runtime_error("NULL argument");
    std::unique_ptr<mlb_host_mesh> m(new mlb_host_mesh());
    host_mesh_from_cells(m->m, n_nodes, node_coords, n_cells, offsets_nodes_of_cell, nodes_of_cell, n_boundary_edges, edge_nodes, edge_tags, n_names,
                         name_tags, names);
    *out = m.release();
    API_END(none)
}
int mlb_host_mesh_write_gmsh(const mlb_mesh * mesh, const char * path) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!mesh || !path) throw std::runtime_error("NULL argument");
    host_mesh_write_gmsh(*mesh, path);
    API_END(none)
}
int mlb_host_mesh_view(const mlb_host_mesh * hm, mlb_mesh * v) {
    mlb_ctx * none = nullptr;
    API_BEGIN0(none)
    if (!hm || !v) throw std::runtime_error("NULL argument");
    HostMesh & m = const_cast<HostMesh &>(hm->m);
    v->n_cells = m.nc; v->n_faces = m.nf; v->n_nodes = m.nn;
    v->node_coords = m.node_xy.data();
    v->offsets_nodes_of_cell = m.onc.data(); v->nodes_of_cell = m.noc.data();
    v->offsets_faces_of_cell = m.ofc.data(); v->faces_of_cell = m.foc.data();
    v->offsets_nodes_of_face = m.onf.data(); v->nodes_of_face = m.nof.data();
    v->cells_of_face = m.cof.data();
    v->cell_coords = m.cell_xy.data(); v->cell_volume = m.cell_vol.data(); v->face_area = m.face_area.data(); v->face_normals = m.face_n.data();
    m.zone_views.clear();
    for (auto & z : m.zones) m.zone_views.push_back({z.name.c_str(), (uint32_t)z.faces.size(), z.faces.data()});
    v->n_zones = (uint32_t)m.zone_views.size();
    v->zones = m.zone_views.data();
    API_END(none)
}
void mlb_host_mesh_free(mlb_host_mesh * m) { delete m; }

}  // extern "C"
