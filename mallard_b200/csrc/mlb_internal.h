// Internal declarations shared by the host preprocessor, the kernels and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mallard_b200.h"

namespace mlb {

typedef std::vector<double> dvec;
typedef std::vector<uint32_t> uvec;
typedef std::vector<int32_t> ivec;

// ---------------------------------------------------------------------------------------------------------------
// Host mesh in the reference's numbering (mesh/mesh.h:228-253)
// ---------------------------------------------------------------------------------------------------------------
struct HostZone { std::string name; uvec faces; };
struct HostMesh {
    uint32_t nc = 0, nf = 0, nn = 0;
    dvec node_xy, cell_xy, cell_vol, face_area, face_n;
    uvec noc, onc, foc, ofc, nof, onf;
    ivec cof;
    std::vector<HostZone> zones;
    std::vector<mlb_zone> zone_views;   // for mlb_host_mesh_view
    int nfc(uint32_t c) const { return (int)(ofc[c + 1] - ofc[c]); }
    int nnc(uint32_t c) const { return (int)(onc[c + 1] - onc[c]); }
    const HostZone * zone(const std::string & n) const { for (auto & z : zones) if (z.name == n) return &z; return nullptr; }
};
void host_mesh_generate(HostMesh & m, int type, uint32_t nx, uint32_t ny, double Lx, double Ly);
void host_mesh_from_view(HostMesh & m, const mlb_mesh & v);
void host_mesh_geometry(HostMesh & m);
void host_mesh_read_gmsh(HostMesh & m, const char * path);          // mesh_io.cpp
void host_mesh_write_gmsh(const mlb_mesh & v, const char * path);
void host_mesh_from_cells(HostMesh & m, uint32_t n_nodes, const double * node_xy, uint32_t n_cells, const uint32_t * onc, const uint32_t * noc,
                          uint32_t n_edges, const uint32_t * edge_nodes, const int32_t * edge_tags, uint32_t n_names, const int32_t * name_tags,
                          const char * const * names);

// partition_graph.cpp: multilevel recursive bisection of a graph in CSR form (unit weights), n_parts parts of the sizes recursive
// coordinate bisection cuts; deterministic, thread-count independent
void graph_partition(uint32_t n, const uint64_t * xadj, const uint32_t * adj, int32_t n_parts, int32_t * part_out);
void dual_graph(uint32_t n_cells, uint32_t n_faces, const int32_t * cells_of_face, std::vector<uint64_t> & xadj, std::vector<uint32_t> & adj);

// ---------------------------------------------------------------------------------------------------------------
// Gas constants (physics/physics.cpp:69-73)
// ---------------------------------------------------------------------------------------------------------------
struct GasParams { double gamma, p_min, p_max, R, cp, cv, mu, kappa; };   // mu = 0: inviscid (the reference); kappa = mu cp / Pr
GasParams make_gas(const mlb_physics & p);

// Boundary condition as the device sees it: type + data[6] (boundary_upt.cpp:38-73: rho,u,v,p,T,h ; p_out: p)
struct BcParams { int32_t type; double data[6]; };
constexpr int MAX_BCS = 16;
constexpr int MAX_SLOTS = 4;    // faces per cell (triangles 3, quads 4)
constexpr int MAX_Q = 7;        // Gauss-Legendre orders 1..7 (numerics/quadrature.cpp:23-151)
constexpr int MAX_K = 55;       // (p+1)(p+2)/2 for p <= 9

// ---------------------------------------------------------------------------------------------------------------
// Preprocessor output (all in LIBRARY numbering; perm arrays map back to the reference)
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t NO_FACE = 0xFFFFFFFFu;
constexpr int32_t CUT_FACE = -2;   // cells_of_face[f][1] of a rank-local mesh: the neighbour exists in the global mesh but not here
constexpr int TILE = 8;         // cells per interleaved TENO table tile (one warp = 8 cells x 4 variables)

// Streaming (FAST mode) table layout: tiles of FAST_CT = 8 cells; a tile belongs to ONE WARP, which streams it through its
// own shared-memory ring (teno_stream_warp.cuh).  S = 1 + 3 stencil slots (triangles).
constexpr int FAST_CT = 8;
constexpr int FAST_S = 4;
#ifndef MLB_FAST_RC3
#define MLB_FAST_RC3 3
#endif
constexpr int fast_rows_per_chunk(int order) { return order == 1 ? 2 : order == 2 ? 5 : order == 3 ? MLB_FAST_RC3 : 2; }
constexpr int fast_stages(int /*order*/) { return 4; }

struct TenoTables {
    int basis = 1, order = 0, K = 0, M = 0, Mp = 0, S = 0;   // Mp = M padded to even, S = stencil slots per cell (1 + max faces)
    int nq_cell = 0;
    std::vector<uint8_t> pidx;    // [K][2]
    dvec qc_xy, qc_w;             // Dunavant cell quadrature
    dvec psi_bar, OI;             // [K], [K][K]
    bool mixed = false;           // the mesh holds quadrilaterals (new; the reference's TENO is triangles only): psi_bar per cell
    dvec psi_bar_cell;            // [Npad][K] mean of every basis function over the cell itself, library numbering (mixed meshes only)
    // Tile-interleaved tables, n_tiles = ceil(N / TILE):
    //   st_ids  [tile][S][Mp][TILE]      u32   library cell ids; empty stencil -> all NO_FACE
    //   st_area [tile][S][Mp][TILE]      f64   transformed areas (0 padding)
    //   st_mat  [tile][S][K][Mp/2][TILE][2] f64  pseudo-inverse rows, m-pairs interleaved across the tile's cells
    uvec st_ids;
    dvec st_area, st_mat;
    // Compact streaming tables (FAST mode), n_ftiles = ceil(N_recon / FAST_CT):
    //   fm_ids   [ftile][S][M-1][FAST_CT]                 u32  columns m = 1..M-1; empty stencil / padding cell -> the cell itself
    //   fm_mat   [ftile][S][K-1][(M-1)/2 pairs | 1][FAST_CT]  f64  rows k = 1..K-1 of A+ with area_t folded into the columns;
    //                                                      per row: (M-1)/2 column pairs [pair][cell][2], then [cell] singles
    //   fm_area0 [ftile * FAST_CT]                         f64  area_t of the cell itself (central stencil, m = 0)
    //   OIs      [K-1][K-1]  upper-triangular fold of the oscillation matrix: OI[k][k] on the diagonal, OI[k][j] + OI[j][k] above
    bool fast = false;
    uvec fm_ids;
    dvec fm_mat, fm_area0, OIs;
    dvec tri_xy;                  // [Npad][6] node coordinates per held cell (library numbering) when the device builds fm_mat
    // Reference-layout CSR copies (reference numbering) for parity checks (optional, small meshes only)
    bool keep_ref = false;
    uvec ref_off_groups, ref_off_stencils, ref_stencils, ref_off_mats;
    dvec ref_mats, ref_areas;
};

struct Prep {
    uint32_t N = 0;        // cells owned+ghost held by this context
    uint32_t N_owned = 0;  // cells [0, N_owned) are updated here; [N_owned, N) are ghosts (multi-GPU)
    uint32_t N_recon = 0;  // cells [0, N_recon) are reconstructed (owned + first ghost ring)
    uint32_t N_interior = 0;  // owned cells [0, N_interior) have TENO stencils made of owned cells only: their reconstruction
                              // does not wait for the halo exchange (multi-GPU overlap); = N_owned when unpartitioned
    uint32_t Npad = 0;     // N rounded up to a multiple of 32
    uint32_t NF = 0;       // real faces held (phantom faces dropped)
    uint32_t NFpad = 0;
    int n_slots = 0;       // max faces per cell
    int Q = 1;
    uvec perm_cells;       // library cell -> reference cell
    uvec iperm_cells;      // reference cell -> library cell (NO_FACE if not held)
    uvec perm_faces;       // library face -> reference face
    // per cell, per face slot j (faces_of_cell order), SoA [n_slots][Npad]:
    uvec slot_face;        // library face id | side << 31 ; NO_FACE if the cell has fewer faces
    ivec slot_nbr;         // >= 0 neighbour library cell ; < 0 : -(bc index + 1) ; INT32_MIN = none / unassigned boundary
    std::vector<uint8_t> slot_nslot;   // [n_slots][Npad] neighbour's slot index of the shared face (TENO gather)
    std::vector<uint8_t> rhs_order;    // [Npad] 4 x 2 bit: slot visited i-th in the reference's accumulation order (Q16)
    std::vector<uint8_t> n_faces_of_cell;   // [Npad]
    dvec cell_vol;         // [Npad]
    dvec cell_xy;          // [2][Npad]
    dvec face_nx, face_ny, face_area;   // [NFpad] unit normal (common_math.h:101-106) and area
    dvec slot_fx;          // TENO: [n_slots][4][Npad] face end points in the cell's reference coordinates
    dvec slot_d;           // viscous: [n_slots][2][Npad] vector from the centroid of every reconstructed cell to its neighbour's across face j
                           // (boundary faces: to the centroid's mirror image): the least-squares gradient stencil
    dvec face_d;           // viscous: [4][NFpad] centroid-to-centroid vector d (boundary faces: to the mirror image) and the vector r
                           // from cell 0's centroid to the face's mid-point
    uvec face_cl;          // [NFpad] library cell on side 0 of the face
    ivec face_cr;          // [NFpad] library cell on side 1 ; < 0: -(bc index + 1) ; INT32_MIN: no flux through this face
    std::vector<uint8_t> face_slots;   // [NFpad] slot in cell 0 | slot in cell 1 << 4
    dvec qf_x, qf_w;       // face quadrature
    TenoTables teno;
    double seconds = 0.0, seconds_stencils = 0.0, seconds_matrices = 0.0;
};

struct PrepOptions {
    int renumber = MLB_RENUMBER_RCM;
    bool keep_ref_tables = false;
    bool fast_tables = false;         // build the compact streaming tables instead of the bit-faithful ones
    bool device_tables = false;       // ... and leave their matrices to the device (teno_tables.cu); the host emits ids + node coordinates
    const int32_t * part = nullptr;   // partition vector (reference numbering) or null
    int rank = 0, n_ranks = 1;
    bool viscous = false;             // hold the second ghost ring and the Green-Gauss geometry
    const double * psi_ref_tri = nullptr;   // rank-local meshes: node coordinates of the GLOBAL mesh's cell 0 (integral_psi_target)
};

void preprocess(const HostMesh & m, const mlb_numerics & num, const std::vector<std::string> & bc_zones,
                const PrepOptions & opt, Prep & out);

void gauss_legendre_rule(int order, dvec & x, dvec & w);
void dunavant_rule(int order, dvec & xy, dvec & w);

// RCM ordering of the cell dual graph (host). Returns library->reference permutation.
void rcm_order(const HostMesh & m, const std::vector<uint32_t> & cells, uvec & order);

}  // namespace mlb
