/*
 * mallard_b200 — C ABI of the B200-native (sm_100a) replacement for Mallard's explicit residual hot path.
 *
 * The reference (MatthewBonanni/Mallard, C++20/Kokkos) has no FFI of its own; its extension points are C++
 * virtual bases and one std::function seam.  Every entry point below replaces one of those seams and cites it
 * (paths relative to the reference's src/).  Conventions:
 *   - return 0 on success, non-zero on error; the message is available from mlb_last_error();
 *     no C++ exception crosses this boundary;
 *   - all arrays passed in or out are caller-owned HOST buffers in the REFERENCE's numbering and layout
 *     (row-major / Kokkos LayoutRight: U[nc][4], prim[nc][5], face values [nf][Q][2][4]); the library owns all
 *     device memory and applies/undoes its own cell/face renumbering internally;
 *   - one host thread drives one context; calls are synchronous at the ABI unless stated otherwise;
 *   - there is NO CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef MALLARD_B200_H
#define MALLARD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlb_ctx mlb_ctx;
typedef struct mlb_host_mesh mlb_host_mesh;
typedef struct mlb_plan mlb_plan;

/* numerics/face_reconstruction.h:30-38 */
enum { MLB_RECON_FO = 0, MLB_RECON_TENO = 1 };
/* numerics/riemann_solver.h:27-37 */
enum { MLB_RIEMANN_RUSANOV = 0, MLB_RIEMANN_HLL = 1, MLB_RIEMANN_HLLC = 2 };
/* numerics/time_integrator.h:23-39 */
enum { MLB_INTEGRATOR_FE = 0, MLB_INTEGRATOR_RK4 = 1, MLB_INTEGRATOR_SSPRK3 = 2 };
/* boundary/boundary.h:31-45 */
enum { MLB_BC_SYMMETRY = 0, MLB_BC_EXTRAPOLATION = 1, MLB_BC_WALL_ADIABATIC = 2, MLB_BC_UPT = 3, MLB_BC_P_OUT = 4,
       /* new (viscous runs; the reference is Euler only): no-slip wall moving with velocity u[2]; T > 0: isothermal at T, else adiabatic */
       MLB_BC_WALL_NOSLIP = 5 };
/* numerics/basis.h:24-37 */
enum { MLB_BASIS_MONOMIAL = 0, MLB_BASIS_LEGENDRE = 1 };
/* mesh/mesh.h:30-42 (generators kept on the host) */
enum { MLB_MESH_CARTESIAN = 0, MLB_MESH_CARTESIAN_TRI = 1, MLB_MESH_WEDGE = 2 };
/* new: cell renumbering for locality (no reference equivalent) */
enum { MLB_RENUMBER_NONE = 0, MLB_RENUMBER_RCM = 1 };
/* new: floating-point contraction mode of the device kernels.
 *   STRICT: no FMA contraction and reference summation order — reproduces the rounding of the reference's
 *           x86-64 Kokkos Serial build operation by operation (libm pow excepted);
 *   FAST:   FMA contraction allowed (<= 1e-12 relative per step against the reference). */
enum { MLB_FP_STRICT = 0, MLB_FP_FAST = 1 };

/* mesh/zone.h:57-142 — a named list of faces; the zone called "interior" holds the interior faces. */
typedef struct {
    const char *name;
    uint32_t n_faces;
    const uint32_t *faces;
} mlb_zone;

/* mesh/mesh.h:228-253 — the arrays the hot path consumes.  Geometry pointers may be NULL, in which case the
 * library computes them exactly as Mesh::compute_* does (mesh/mesh.cpp:167-261).  Cells are triangles or quadrilaterals, in any
 * mix; TENO on quadrilaterals is new (the reference throws, numerics/face_reconstruction.cpp:485-487): a quadrilateral is
 * integrated as the two triangles Mesh::compute_cell_volumes splits it into, its reference frame is spanned by the edges
 * node 0 -> node 1 and node 0 -> node 3, and it carries its own basis means (no oracle; tests/test_gpu_parity.py checks
 * k-exactness). */
typedef struct {
    uint32_t n_cells, n_faces, n_nodes;
    const double *node_coords;              /* [n_nodes][2] */
    const uint32_t *offsets_nodes_of_cell;  /* [n_cells+1] */
    const uint32_t *nodes_of_cell;
    const uint32_t *offsets_faces_of_cell;  /* [n_cells+1] */
    const uint32_t *faces_of_cell;
    const uint32_t *offsets_nodes_of_face;  /* [n_faces+1] */
    const uint32_t *nodes_of_face;
    const int32_t *cells_of_face;           /* [n_faces][2], -1 = boundary */
    const double *cell_coords;              /* [n_cells][2] or NULL */
    const double *cell_volume;              /* [n_cells]    or NULL */
    const double *face_area;                /* [n_faces]    or NULL */
    const double *face_normals;             /* [n_faces][2] or NULL (area-weighted, out of cell 0) */
    uint32_t n_zones;
    const mlb_zone *zones;
} mlb_mesh;

/* [numerics] + [numerics.face_reconstruction] of the TOML input (solver/solver.cpp:110-186,
 * numerics/face_reconstruction.cpp:109-116), already parsed. */
typedef struct {
    int32_t recon;                   /* MLB_RECON_* */
    int32_t riemann;                 /* MLB_RIEMANN_* */
    int32_t integrator;              /* MLB_INTEGRATOR_* */
    int32_t basis;                   /* MLB_BASIS_* (TENO) */
    int32_t basis_order;             /* TENO polynomial order p, 1..9 (what numerics/basis.h:81-176 tabulates).  Every value the
                                        reference accepts runs; p <= 4 with max_stencil_size_factor 2.0 on triangles (the examples'
                                        configuration) has the specialised streaming kernels, everything else - p = 5..9, other
                                        factors, quadrilaterals - a generic kernel (csrc/teno_generic.cuh) */
    double max_stencil_size_factor;  /* TENO, default 2.0: M = floor(factor * K) cells per stencil (face_reconstruction.cpp:170-180) */
    int32_t quadrature_order_cell;   /* Dunavant order 1..5, 0 = reference default p+1 (which does not exist for p >= 5: the reference
                                        throws there too, numerics/quadrature.cpp:269) */
    int32_t quadrature_order_face;   /* Gauss-Legendre order, 0 = reference default (p+1)/2 */
    int32_t fp_mode;                 /* MLB_FP_* */
    int32_t renumber;                /* MLB_RENUMBER_* */
    int32_t teno_fixed;              /* 0 = reference-faithful TENO (it turns non-finite within a step, as the
                                        reference does); 1 = normalised weights + mean-free basis (new, no oracle) */
    int32_t keep_stage_rhs;          /* 1 = keep every stage residual for mlb_get_array("rhsN") */
} mlb_numerics;

/* [physics] (physics/physics.cpp:27-53) */
typedef struct {
    double gamma, p_ref, T_ref, rho_ref, p_min, p_max;
    /* new - viscous terms (BASELINE configs[4]; the reference's physics is Euler only, physics/physics.h:23-29, its viscous
     * spectral radius a commented-out stub, solver/solver.cpp:638-651).  mu = 0 (a zero-initialised tail) is the reference's inviscid
     * model, bit for bit.  mu > 0 adds the Navier-Stokes fluxes: Newtonian stress with Stokes' hypothesis, Fourier heat flux with
     * conductivity mu cp / Pr, constant mu; face gradients of (u, v, T) = average of the two cells' least-squares gradients corrected
     * along the centroid line by the two-point difference; second-order accurate whatever the reconstruction of the inviscid part. */
    double mu, Pr /* 0 = 0.72 */;
} mlb_physics;

/* one [[boundaries]] entry (solver/solver.cpp:188-237); order of the array = order in the TOML file */
typedef struct {
    const char *zone_name;
    int32_t type;   /* MLB_BC_* */
    double u[2];    /* upt */
    double p;       /* upt, p_out */
    double T;       /* upt */
} mlb_bc;

/* Domain decomposition (new).  n_ranks == 1 → single GPU.  The exchange itself is driven by the caller's
 * communicator (see mlb_halo_*). */
typedef struct {
    int32_t rank, n_ranks;
    int32_t device;                  /* CUDA device ordinal for this context */
} mlb_parallel;

const char *mlb_version(void);
/* Revision of the structs and signatures in this header; bumped whenever one of them changes (4: mlb_physics gained mu / Pr,
 * MLB_BC_WALL_NOSLIP).  A host built against an older header would hand the library structs of the wrong size: it calls
 * mlb_check_abi(MLB_ABI_VERSION) once before anything else and gets a clean error instead (oracle/dropin_harness.cpp and the
 * ctypes mirror do). */
#define MLB_ABI_VERSION 4
int mlb_check_abi(int32_t header_abi_version);
/* OpenMP threads used by the host preprocessor (n <= 0: query only); returns the value in effect */
int mlb_set_host_threads(int32_t n);
/* message of the last failed call on `ctx` (or of the last failed mlb_create / stateless call when ctx == NULL) */
const char *mlb_last_error(const mlb_ctx *ctx);

/* ---- lifetime: replaces Solver::init's init_numerics/init_boundaries/allocate_memory/copy_host_to_device
 *      (solver/solver.cpp:39-80,297-320).  Runs the mesh preprocessor (renumbering, SoA layout, CSR/ELL flattening of
 *      the face and TENO-stencil connectivity, TENO tables) and uploads everything. */
int mlb_create(mlb_ctx **out, const mlb_mesh *mesh, const mlb_numerics *numerics, const mlb_physics *physics,
               const mlb_bc *bcs, int32_t n_bcs, const mlb_parallel *parallel /* NULL = 1 GPU, device 0 */);
void mlb_destroy(mlb_ctx *ctx);

/* ---- state: Solver::copy_host_to_device / copy_device_to_host (solver/solver.cpp:322-334) */
int mlb_set_state(mlb_ctx *ctx, const double *U /* [nc][4] */, const double *prim /* [nc][5] or NULL = recompute */);
int mlb_get_state(mlb_ctx *ctx, double *U /* or NULL */, double *prim /* or NULL */, double *cfl_local /* or NULL */);

/* ---- FaceReconstruction::calc_face_values (numerics/face_reconstruction.h:75; .cpp:95-99,1081-1106) on the current
 *      state.  F_out [nf][Q][2][4] in reference layout, entries the reference leaves undefined are written as 0. */
int mlb_calc_face_values(mlb_ctx *ctx, double *F_out);
int mlb_n_face_quadrature_points(const mlb_ctx *ctx);

/* ---- the rhs_func seam: Solver::calc_rhs (solver/solver_rhs.cpp:44-55; solver.h:268-270) */
int mlb_calc_rhs(mlb_ctx *ctx, double *rhs_out /* [nc][4] or NULL */);                   /* on the resident state */
int mlb_calc_rhs_host(mlb_ctx *ctx, const double *U_in /* [nc][4] */, double *rhs_out);   /* host buffers in/out   */

/* ---- Solver::calc_dt (solver/solver.cpp:580-590,592-742): spectral radius, max-reduction, dt = cfl/max,
 *      cfl_local *= dt.  Uses the primitives of the last update_primitives, as the reference does. */
int mlb_calc_dt(mlb_ctx *ctx, double cfl, double *dt_out);
int mlb_set_dt(mlb_ctx *ctx, double dt);

/* ---- Solver::take_step (solver/solver.cpp:521-531) = TimeIntegrator::take_step (numerics/time_integrator.cpp:57-163)
 *      + update_primitives.  Uses the dt of the last mlb_calc_dt / mlb_set_dt (dt < 0 → error, solver.cpp:587-589). */
int mlb_take_step(mlb_ctx *ctx);
/* same seam with host buffers: H2D of U, one step, D2H of U (the end-to-end path a host-resident driver pays) */
int mlb_take_step_host(mlb_ctx *ctx, double cfl /* <=0: keep dt */, double *U_inout, double *dt_out);
/* ---- Solver::run's loop (solver/solver.cpp:352-373) minus checks/output: n_steps x { calc_dt ; take_step },
 *      device-resident and asynchronous; returns after the last step completed.  cfl <= 0 → fixed dt.
 *      Launch-bound meshes: a step is captured once and replayed as a CUDA graph (n_steps >= 8); with MLB_SMALL_STEP=1 in the
 *      environment a first-order run of up to 262 144 cells executes ALL its steps in one cooperative kernel (same results;
 *      opt-in until measured on hardware, csrc/small_step.cuh). */
int mlb_run(mlb_ctx *ctx, uint32_t n_steps, double cfl, double *t_out, double *dt_last_out);
int mlb_get_time(mlb_ctx *ctx, double *t, uint64_t *step);

/* ---- diagnostics and output fed from the device-resident fields.
 * mlb_field_ranges: the scalar ranges Solver::do_checks prints (max_array / min_array, solver/solver.cpp:434-443,
 *   common/common_math.h:577-614) for RHO, RHOU_X, RHOU_Y, RHOE, U_X, U_Y, P, T, H, and the number of NaN entries
 *   Solver::check_fields looks for (solver/solver.cpp:470-498), as one device reduction - no copy of the state.
 * mlb_write_vtu: DataWriter::write_vtu (io/data_writer.cpp:93-244), byte-identical file "<prefix>_<step:06>.vtu";
 *   names[] as in the TOML `variables` list ("CFL", "RHO", ..., "H"); `mesh` = the arrays given to mlb_create. */
int mlb_field_ranges(mlb_ctx *ctx, double *min9, double *max9, uint64_t *n_nan);
int mlb_write_vtu(mlb_ctx *ctx, const mlb_mesh *mesh, const char *prefix, uint32_t step, int32_t n_vars,
                  const char *const *names);

/* ---- test hook mirroring the fake RHS of test/time_integrator_test.cpp:22-28 (NULL clears it) */
int mlb_set_rhs_override(mlb_ctx *ctx, const double *rhs /* [nc][4] */);

/* ---- introspection / parity hooks.  `name` is one of:
 *   "perm_cells" (u32[nc], library index -> reference cell), "perm_faces" (u32[n real faces]),
 *   "rhs0".."rhs3", "U_temp" (f64[nc][4]), "cfl_local" (f64[nc]),
 *   "teno:offsets_stencil_groups", "teno:offsets_stencils", "teno:stencils", "teno:offsets_reconstruction_matrices"
 *   (u32, reference CSR layout and numbering, numerics/face_reconstruction.h:201-260),
 *   "teno:reconstruction_matrices", "teno:transformed_areas", "teno:integral_psi_target",
 *   "teno:oscillation_indicator" (f64), "teno:poly_indices" (u8[K][2]),
 *   "stats" (f64[13]: kernel launches, preprocess seconds, device bytes, held cells, owned cells, faces, reconstructed cells, stages per step,
 *            host seconds of the stencil search, of the matrices, device seconds of the table build, steps replayed as a CUDA graph, steps run
 *            inside the cooperative small-mesh kernel).
 * out == NULL → only the byte size is returned in *nbytes. */
int mlb_get_array(mlb_ctx *ctx, const char *name, void *out, uint64_t *nbytes);

/* ---- measurement support: CUDA events on the library's compute stream */
int mlb_event_record(mlb_ctx *ctx, int32_t slot /* 0..15 */);
int mlb_event_elapsed_ms(mlb_ctx *ctx, int32_t slot_begin, int32_t slot_end, float *ms);   /* synchronises slot_end */
/* per-kernel device time accumulated while profiling is on: names[i] / ms[i] / launches[i]; returns count */
int mlb_profile_enable(mlb_ctx *ctx, int32_t on);
int mlb_profile_read(mlb_ctx *ctx, int32_t max_entries, const char **names, double *ms, uint64_t *launches);
uint64_t mlb_launch_count(const mlb_ctx *ctx);
int mlb_synchronize(mlb_ctx *ctx);
void *mlb_stream(mlb_ctx *ctx);   /* cudaStream_t of the compute stream */

/* ---- multi-GPU halo exchange (new).  The context of rank r owns the cells with part[c] == r plus ghost
 *      copies of every remote cell its stencils read.  Each stage the caller moves `send` to the peers and hands back
 *      `recv`; buffers are DEVICE pointers owned by the library so NCCL / peer copies can use them directly. */
int mlb_partition(const mlb_mesh *mesh, int32_t n_parts, int32_t *part_out /* [nc] */);
/* the same recursive coordinate bisection from the centroids alone (a rank that holds only its part of the mesh can still
 * compute the global partition: 16 bytes per cell) */
int mlb_partition_coords(uint64_t n_cells, const double *cell_xy /* [n_cells][2] */, int32_t n_parts, int32_t *part_out);
/* Graph partition: multilevel recursive bisection (heavy-edge matching, greedy graph growing, Fiduccia-Mattheyses refinement) of the
 * cell-face dual graph - no coordinates involved; part sizes are exactly those of mlb_partition (load balance is identical, what
 * differs is where the cuts run: shorter on domains that are not convex).  Deterministic and independent of the number of host
 * threads, so every rank can compute it for itself.  mlb_partition_graph takes the graph from mesh->cells_of_face (cut faces, -2, and
 * boundary faces, -1, carry no edge); _csr takes any symmetric adjacency structure. */
int mlb_partition_graph(const mlb_mesh *mesh, int32_t n_parts, int32_t *part_out /* [nc] */);
int mlb_partition_graph_csr(uint32_t n, const uint64_t *xadj /* [n+1] */, const uint32_t *adj, int32_t n_parts, int32_t *part_out);
int mlb_create_partitioned(mlb_ctx **out, const mlb_mesh *mesh, const int32_t *part, const mlb_numerics *numerics,
                           const mlb_physics *physics, const mlb_bc *bcs, int32_t n_bcs, const mlb_parallel *parallel);
/* Rank-local ingest: the context is created from THIS RANK'S PART of the mesh only - its own cells plus enough layers of ghost
 * cells for every stencil search to stay inside - so that no rank ever holds the global mesh.  Rules for `local_mesh`:
 *   - cells, faces (and zones) keep the relative order of their global ids (global_cell_ids strictly ascending; zone face lists
 *     in the global zone's order): stencil membership depends on that order, and with it preserved every table and
 *     every result is bit-identical to the context mlb_create_partitioned builds from the global mesh;
 *   - a face whose other cell is not part of the local mesh is marked cells_of_face[f] = {present cell, -2} ("cut"); -1 stays
 *     "boundary".  Creation fails ("needs more ghost layers") if a stencil search or an owned cell touches a cut face;
 *   - part_local[i] = owning rank of local cell i.
 * All arrays crossing the ABI afterwards (mlb_set_state, mlb_get_state ...) are in the LOCAL mesh's numbering; the halo id
 * lists (mlb_halo_recv_ids / mlb_halo_set_send_ids) are in GLOBAL ids. */
typedef struct {
    uint32_t n_global_cells;
    const uint32_t *global_cell_ids;   /* [local_mesh->n_cells], strictly ascending */
    double cell0_nodes[6];             /* x,y of the three nodes of GLOBAL cell 0 in nodes_of_cell order: the reference takes
                                          integral_psi_target from its cell 0 (numerics/face_reconstruction.cpp:598-602) */
} mlb_local_mesh;
int mlb_create_local(mlb_ctx **out, const mlb_mesh *local_mesh, const int32_t *part_local, const mlb_local_mesh *local,
                     const mlb_numerics *numerics, const mlb_physics *physics, const mlb_bc *bcs, int32_t n_bcs,
                     const mlb_parallel *parallel);
/* n_peers first (arrays NULL), then arrays of at least *n_peers entries (never more than n_ranks - 1) */
int mlb_halo_info(mlb_ctx *ctx, int32_t *n_peers, int32_t *peers /* [n_peers] */, uint64_t *send_counts,
                  uint64_t *recv_counts /* CELLS per peer (4 doubles each), per exchange */);
/* ghost cells this rank receives from peer `peer_index` (index into the peers array), as reference cell ids in the
 * order they occupy the receive buffer; the caller ships each list to its peer, which registers what it must send: */
int mlb_halo_recv_ids(mlb_ctx *ctx, int32_t peer_index, uint32_t *ref_ids_out /* [recv_counts[peer_index]] */);
int mlb_halo_set_send_ids(mlb_ctx *ctx, int32_t n_lists, const int32_t *peer_ranks, const uint64_t *counts,
                          const uint32_t *ref_ids /* concatenated */);
int mlb_halo_buffers(mlb_ctx *ctx, void **send_dev, void **recv_dev);   /* contiguous, peers in ascending order */
/* The exchange runs on the context's COMMUNICATION stream (mlb_comm_stream; cudaStream_t), concurrently with the interior
 * work of the stage on the compute stream:
 *   mlb_halo_pack   waits (on the device) for everything enqueued on the compute stream so far, then fills `send`;
 *   the caller enqueues its transfers send -> peers, peers -> recv ON mlb_comm_stream (ncclSend/ncclRecv, peer copies);
 *   mlb_halo_unpack scatters `recv` into the ghost cells; the compute stream's first reader of ghost data (the rim part
 *   of mlb_stage, the CFL kernel) waits for it on the device.  Nothing here blocks the host. */
int mlb_halo_pack(mlb_ctx *ctx, int32_t stage);
int mlb_halo_unpack(mlb_ctx *ctx, int32_t stage);
void *mlb_comm_stream(mlb_ctx *ctx);
/* split-phase stepping for callers that own the communicator: stage s of the current step.
 * mlb_stage_begin (optional) enqueues the reconstruction of the interior cells (TENO stencils made of owned cells only;
 * the preprocessor numbers them first) — call it right after the exchange has been enqueued; mlb_stage waits for the
 * unpack, reconstructs the remaining cells and runs the flux / residual / RK kernels.  Without mlb_stage_begin,
 * mlb_stage does both parts itself, in the same order. */
int mlb_n_stages(const mlb_ctx *ctx);
int mlb_stage_begin(mlb_ctx *ctx, int32_t stage);
int mlb_stage(mlb_ctx *ctx, int32_t stage);
int mlb_local_max_spectral_radius(mlb_ctx *ctx, double *max_out);   /* rank-local part of calc_dt; max_out == NULL: async */
int mlb_apply_dt(mlb_ctx *ctx, double cfl, double global_max);      /* dt = cfl/max ; cfl_local *= dt */
/* device-resident variant for communicators that reduce in place (ncclAllReduce(max) on mlb_scalars_device() + 2):
 * the scalar block holds doubles { dt, t, max spectral radius, cfl, ... }; no host round trip per step */
void *mlb_scalars_device(mlb_ctx *ctx);
int mlb_apply_dt_device(mlb_ctx *ctx, double cfl);
/* owned cells only, host buffers [n_owned][4] in the order of mlb_owned_cells (what a rank-local driver holds) */
int mlb_set_owned(mlb_ctx *ctx, const double *U_owned);
int mlb_get_owned(mlb_ctx *ctx, double *U_owned);
int mlb_finish_step(mlb_ctx *ctx);                                  /* update_primitives ; t += dt ; step++ */
int mlb_owned_cells(mlb_ctx *ctx, uint32_t *n_owned, uint32_t *cells_out /* reference ids, or NULL */);

/* ---- native multi-GPU driver: NCCL over NVLink / NVSwitch inside the library (libnccl.so.2 is resolved at run time; a process
 *      that already carries an NCCL, e.g. PyTorch's, keeps using that one).  One process per GPU:
 *        rank 0: mlb_comm_unique_id(id) ; the host ships the MLB_COMM_ID_BYTES bytes to every rank (MPI_Bcast, a file, ...)
 *        all:    mlb_comm_init(ctx, id)       communicators + the exchange plan (replaces mlb_halo_recv_ids / _set_send_ids)
 *                mlb_run_distributed(...)     Solver::run's loop: per stage pack -> grouped ncclSend/ncclRecv -> unpack on the
 *                                             communication stream under the reconstruction of the interior cells, per step one
 *                                             ncclAllReduce(max) of the spectral radius; the whole step (both streams, NCCL
 *                                             included) is captured once and replayed as a CUDA graph; no host code between stages
 *                mlb_take_step_distributed_host(...)  the take_step seam with HOST buffers of the rank's own cells
 *      Collective: every rank of the partition must make the same calls in the same order. */
#define MLB_COMM_ID_BYTES 128
int mlb_comm_unique_id(void *id_out /* MLB_COMM_ID_BYTES */);
int mlb_comm_init(mlb_ctx *ctx, const void *id /* MLB_COMM_ID_BYTES */);
int mlb_run_distributed(mlb_ctx *ctx, uint32_t n_steps, double cfl /* <= 0: fixed dt */, double *t_out, double *dt_last_out);
int mlb_take_step_distributed_host(mlb_ctx *ctx, double cfl, double *U_owned_inout /* [n_owned][4], mlb_owned_cells order */,
                                   double *dt_out);

/* ---- stateless device kernels for known-answer tests of the plug-in interfaces */
/* RiemannSolver::calc_flux (numerics/riemann_solver.h:85-90); L/R rows = rho,u,v,p,h */
int mlb_riemann_flux(int32_t device, int32_t riemann, int32_t fp_mode, uint64_t n, const double *n_unit /* [n][2] */,
                     const double *L /* [n][5] */, const double *R /* [n][5] */, double gamma, double *flux /* [n][4] */);
/* Physics::compute_primitives_from_conservatives (physics/physics.h:852-867) + set_R_cp_cv (physics.cpp:69-73) */
int mlb_compute_primitives(int32_t device, int32_t fp_mode, const mlb_physics *physics, uint64_t n, const double *U,
                           double *prim /* [n][5] */, double *R_cp_cv /* [3] or NULL */);

/* ---- the mesh preprocessor alone (host only, no device): what mlb_create uploads.  `part` may be NULL.  Arrays by name:
 *   "sizes" (u32[12]: N, N_owned, N_recon, NF, n_slots, Q, K, M, Npad, S, Mp, streaming tile size), "n_interior" (u32:
 *   owned cells [0, n_interior) have TENO stencils without ghosts), "perm_cells", "perm_faces",
 *   "slot_face", "slot_nbr" (i32), "rhs_order" (u8), "st_ids", "ghost_owner" (i32 per ghost cell),
 *   "fm_ids" (u32), "fm_mat", "fm_area0", "OIs" (f64; compact streaming tables, fp_mode FAST only),
 *   "halo_peers" (i32), "halo_recv_counts" (u64), "halo_recv_ids" (u32, concatenated; partitioned plans only),
 *   and the "teno:*" tables of mlb_get_array (reference CSR layout; unpartitioned plans only). */
int mlb_plan_create(mlb_plan **out, const mlb_mesh *mesh, const mlb_numerics *numerics, const mlb_bc *bcs, int32_t n_bcs,
                    const int32_t *part, const mlb_parallel *parallel);
/* the plan of a rank-local mesh (see mlb_create_local); cell ids in its arrays are in the LOCAL mesh's numbering */
int mlb_plan_create_local(mlb_plan **out, const mlb_mesh *local_mesh, const mlb_numerics *numerics, const mlb_bc *bcs,
                          int32_t n_bcs, const int32_t *part_local, const mlb_parallel *parallel, const mlb_local_mesh *local);
int mlb_plan_get(mlb_plan *plan, const char *name, void *out, uint64_t *nbytes);
void mlb_plan_destroy(mlb_plan *plan);

/* ---- host mesh generators: Mesh::init_cart / init_cart_tri / init_wedge (mesh/mesh.cpp:305-848) */
int mlb_host_mesh_generate(mlb_host_mesh **out, int32_t type, uint32_t nx, uint32_t ny, double Lx, double Ly);
/* connectivity + node coordinates supplied by the caller (any unstructured mesh of triangles / quads); geometry computed as
 * Mesh::compute_cell_centroids / compute_cell_volumes / compute_face_areas / compute_face_normals do (mesh/mesh.cpp:167-261) */
int mlb_host_mesh_from_arrays(mlb_host_mesh **out, const mlb_mesh *mesh);
/* Mesh file reader / writer: what `[mesh] type = "file"` needs (Mesh::init's MeshType::FILE branch is a stub that throws,
 * mesh/mesh.cpp:41-43).  Gmsh MSH 2.2 ASCII, 2-D: triangles and quadrilaterals become cells, tagged 2-node lines name the
 * boundary zones ($PhysicalNames), "interior" lists the two-cell faces; geometry as Mesh::compute_* computes it. */
int mlb_host_mesh_read_gmsh(mlb_host_mesh **out, const char *path);
/* the same construction from cell-node lists already in memory (any mix of triangles and quadrilaterals, counter-clockwise):
 * faces = unique cell edges, tagged boundary edges name the zones, the rest as for the reader */
int mlb_host_mesh_from_cells(mlb_host_mesh **out, uint32_t n_nodes, const double *node_coords /* [n_nodes][2] */, uint32_t n_cells,
                             const uint32_t *offsets_nodes_of_cell, const uint32_t *nodes_of_cell, uint32_t n_boundary_edges,
                             const uint32_t *edge_nodes /* [n][2] */, const int32_t *edge_tags, uint32_t n_names,
                             const int32_t *name_tags, const char *const *names);
int mlb_host_mesh_write_gmsh(const mlb_mesh *mesh, const char *path);
int mlb_host_mesh_view(const mlb_host_mesh *m, mlb_mesh *view);   /* pointers stay valid until mlb_host_mesh_free */
void mlb_host_mesh_free(mlb_host_mesh *m);

#ifdef __cplusplus
}
#endif
#endif /* MALLARD_B200_H */
