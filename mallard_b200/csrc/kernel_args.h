// Argument blocks passed BY VALUE to the kernels (they live in the constant bank: uniform, broadcast reads) and the
// launcher table each floating-point mode exports.
#pragma once
#ifndef MLB_HOST_EMULATION          // tests/emul compiles the kernel SOURCE for the host (test infrastructure, never a product path)
#include <cuda_runtime.h>
#define MLB_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif

#include "mlb_internal.h"

namespace mlb {

struct DevGeom {
    uint32_t N, N_owned, N_recon, Npad, NF;
    int32_t n_slots, Q;
    const uint32_t * slot_face;   // [n_slots][Npad]
    const int32_t * slot_nbr;     // [n_slots][Npad]
    const uint8_t * slot_nslot;   // [n_slots][Npad]
    const uint8_t * rhs_order;    // [Npad]
    const uint8_t * nfc;          // [Npad]
    const double * cell_vol;      // [Npad]
    const double * cell_xy;       // [2][Npad]
    const double * bnd_s;         // [Npad] 2*pow(V,1/2) (solver.cpp:662-666), computed on the host with libm pow
    const double * face_nx, * face_ny, * face_area;   // [NFpad]
    const double * slot_fx;       // [n_slots][4][Npad]
    const double * slot_d;        // viscous: [n_slots][2][Npad] centroid-to-neighbour-centroid vectors of every reconstructed cell (boundary: mirror image)
    const double * face_d;        // viscous: [4][NFpad] centroid line d of every face (boundary: to the mirror image), cell 0 centroid -> face mid-point r
    uint32_t NFpad;
    const uint32_t * face_cl;     // [NFpad] cell on side 0 (normal points out of it)
    const int32_t * face_cr;      // [NFpad] cell on side 1 ; < 0: -(bc index + 1) ; INT32_MIN: no flux
    const uint8_t * face_slots;   // [NFpad] slot of the face in cell 0 | slot in cell 1 << 4 (TENO face values are cell-centred)
};

struct DevPhys {
    GasParams gas;
    int32_t riemann, n_bcs;
    BcParams bcs[MAX_BCS];
    double qf_x[MAX_Q], qf_w[MAX_Q];
};

// Device-resident scalars (one small block per context)
enum { SC_DT = 0, SC_T = 1, SC_MAX_SR = 2, SC_CFL = 3, SC_COUNT = 8 };

// Fused RK stage update, applied to the residual k of the stage just evaluated (numerics/time_integrator.cpp:57-163):
//   mode 0: out = base + (dt*coef)*k
//   mode 1: out = (c0*base + c1*in) + (dt*coef)*k                  (SSPRK3 stage 2: axpby then axpy)
//   mode 2: out = ((base + (dt*cp0)*kp0) + (dt*cp1)*kp1 [+ ...]) + (dt*coef)*k   (final combination)
//   mode 3: no update (bare residual evaluation)
struct RkArgs {
    int32_t mode, n_prev, last_stage, pad_;
    const double * base;
    double * out;
    double * k_store;             // AoS [Npad][4] or null
    const double * kprev[3];
    double cprev[3];
    double c0, c1, coef;
    double * prim_out;            // SoA [6][Npad] (u, v, p, T, h, rho) written when last_stage (Solver::update_primitives)
};

struct StageArgs {
    DevGeom g;
    DevPhys ph;
    RkArgs rk;
    const double * Uin;           // AoS [Npad][4]: state the residual is evaluated on
    const double * Fc;            // TENO: cell-centred face values AoS [Npad][n_slots * Q][4]
    double * AF;                  // [NFpad][4] area * quadrature-averaged flux per face (written by the face kernel)
    double * G;                   // viscous: least-squares gradients AoS [Npad][6] = d(u, v, T)/d(x, y); null when mu == 0
    const double * k_override;    // AoS [Npad][4] or null: state-independent residual (test hook)
    double * scal;                // device scalars
    unsigned long long * step_counter;
    int32_t teno;
};

struct ReconArgs {
    DevGeom g;
    const double * Uin;
    double * Fc;
    const uint32_t * st_ids;
    const double * st_area;
    const double * st_mat;
    int32_t order, K, M, Mp, S, basis, fixed_weights;
    double qf_x[MAX_Q];
    double psi_bar[15];           // the specialised kernels (order <= 4: K <= 15) take these by value ...
    uint8_t pidx[2 * 15];
    double OI[15 * 15];
    const double * OI_dev;        // ... the generic kernel (any order <= 9: K <= 55) reads them from device memory
    const double * psi_bar_dev;
    const uint8_t * pidx_dev;
    const double * psi_bar_cell;  // [Npad][K] or null: meshes with quadrilaterals carry the basis means per cell
};

struct ReconStreamArgs {       // teno_stream.cuh
    DevGeom g;
    const double * Uin;
    double * Fc;
    const double * mat;
    const uint32_t * ids;
    const double * area0;
    uint32_t n_tiles;             // tiles [tile_begin, n_tiles) are processed by this launch
    uint32_t tile_begin;
    int32_t order, fixed_weights, async_gather, basis;
    double qf_x[4];
    double psi_bar[15];
    double OIs[14 * 14];
};

struct TableBuildArgs {        // teno_tables.cu: reconstruction matrices built on the device
    uint32_t n_recon, n_ftiles;
    int32_t order, nq;
    const double * tri_xy;        // [Npad][6] node coordinates of every held cell (library numbering, nodes_of_cell order)
    const uint32_t * fm_ids;      // compact stencil ids (library numbering)
    double * fm_mat;
    double * fm_area0;
    int * err_flag;
    double qc_xy[14], qc_w[7];    // Dunavant cell quadrature
    double psi_bar[15];
};

struct CflArgs {
    DevGeom g;
    GasParams gas;
    const double * U;             // AoS [Npad][4] (unused: rho comes from prim[5])
    const double * prim;          // SoA [6][Npad]
    double * sr_out;              // [Npad] spectral radius per cell (cfl_local before scaling)
    double * scal;
    long long * max_bits;         // running max (as ordered int) — reset by the finishing block
    unsigned int * blocks_done;
    double cfl;                   // <= 0: only the local max is produced (multi-GPU: host/all-reduce finishes)
};

// small_step.cuh: every stage's argument block (same content run_stage hands the stand-alone kernels) + the CFL kernel's
struct SmallStepArgs {
    StageArgs st[4];
    CflArgs cfl;                  // cfl <= 0: fixed dt (already in scal[SC_DT])
    int32_t n_stages;
    uint32_t n_steps;
};

struct KernelTable {
    const char * name;
    void (*gradients)(const StageArgs &, cudaStream_t);   // viscous runs only
    void (*faces)(const StageArgs &, cudaStream_t);
    void (*stage)(const StageArgs &, cudaStream_t);
    void (*recon)(const ReconArgs &, cudaStream_t);
    void (*cfl)(const CflArgs &, cudaStream_t);
    void (*riemann_flux)(int riemann, uint64_t n, const double * nunit, const double * L, const double * R, double gamma,
                         double * flux, cudaStream_t);
    void (*primitives)(const GasParams &, uint64_t n, const double * U_aos, double * P_aos, cudaStream_t);
    void (*primitives_soa)(const GasParams &, uint32_t n, uint32_t npad, const double * U, double * P, cudaStream_t);
    bool (*recon_supported)(int order, int K, int Mp, int S, int basis);
    // FAST mode only (null in the STRICT table): streaming TENO reconstruction over the compact tables
    void (*recon_stream)(const ReconStreamArgs &, cudaStream_t);
    bool (*stream_supported)(int order, int M, int Q, int basis, int n_slots);
    // whole steps of a small first-order mesh in one cooperative launch (small_step.cuh); max_blocks <= 0: as many as are resident
    void (*small_step)(const SmallStepArgs &, int max_blocks, cudaStream_t);
};

const KernelTable * kernels_strict();
const KernelTable * kernels_fast();

// device-side construction of the compact TENO tables (teno_tables.cu)
bool teno_tables_device_supported(int order, int basis, int nq);
void launch_teno_tables(const TableBuildArgs & a, cudaStream_t st);

// Launch configuration is a property of (kernel, DEVICE): the opt-in to more than 48 kB of dynamic shared memory is per
// device, and so are the SM count and the occupancy a persistent grid is sized from.  Cached per (kernel, device ordinal),
// thread-safe, every CUDA return code checked (throws std::runtime_error).  utils.cu.
//   persistent_ctas: opts `kernel` into `smem` bytes on the CURRENT device and returns SMs x resident CTAs per SM
//   ensure_dynamic_smem: the opt-in alone (kernels launched with a problem-sized grid)
int persistent_ctas(const void * kernel, int threads, size_t smem);
void ensure_dynamic_smem(const void * kernel, size_t smem);

// layout helpers (mode independent, utils.cu)
void launch_import_state(const double * aos, const uint32_t * perm, uint32_t n, uint32_t npad, int nv, double * soa, cudaStream_t);
void launch_export_state(const double * soa, const uint32_t * perm, uint32_t n, uint32_t npad, int nv, double * aos, cudaStream_t);
void launch_rho_plane(const double * U_aos, uint32_t n, uint32_t npad, double * prim, cudaStream_t);
void launch_export_scaled(const double * v, const double * scal, int which, const uint32_t * perm, uint32_t n, double * out, cudaStream_t);
void launch_gather(const double * soa, const uint32_t * idx, uint32_t n, uint32_t npad, double * buf, cudaStream_t);
void launch_scatter(const double * buf, const uint32_t * idx, uint32_t n, uint32_t npad, double * soa, cudaStream_t);
void launch_apply_dt(double * scal, long long * max_bits, double cfl, double global_max, int use_global, cudaStream_t);
void launch_set_scalar(double * scal, int which, double v, cudaStream_t);
// N3: output fields as variable planes in reference numbering, and the do_checks / check_fields reductions
void launch_export_fields(int n_fields, const int32_t * codes, const double * U, const double * prim, const double * sr, const double * scal,
                          const uint32_t * perm, uint32_t n, uint32_t npad, uint32_t n_ref, double * out, cudaStream_t);
int field_ranges_blocks(uint32_t n);
void launch_field_ranges(const double * U, const double * prim, uint32_t n, uint32_t npad, double * partial /* [blocks][18] */,
                         double * out18, unsigned long long * nan_count, cudaStream_t);
void launch_export_faces(const double * Fc, const double * U, const uint32_t * slot_face, const uint32_t * perm_faces, uint32_t n,
                         uint32_t npad, int n_slots, int Q, double * F_aos /* [nf_ref][Q][2][4] */, cudaStream_t);

}  // namespace mlb
