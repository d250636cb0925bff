// TEST INFRASTRUCTURE ONLY — the integration proof for INTEGRATION.md.
//
// The UNMODIFIED reference host (MatthewBonanni/Mallard: TOML input, mesh generation, exprtk initial condition, checks,
// VTU writer — compiled from /root/reference into oracle/_ref/lib/libmallard_ref.a by oracle/build_ref.sh) with the three
// hot-path seams of Solver::run (src/solver/solver.cpp:352-373) rerouted through the C ABI of libmallard_b200.so:
//
//     Solver::calc_dt()                     -> mlb_calc_dt / mlb_set_dt            (solver.cpp:580-590)
//     Solver::take_step()                   -> mlb_take_step                       (solver.cpp:521-531)
//     Solver::copy_device_to_host()         -> mlb_get_state                       (solver.cpp:322-334)
//
// Everything else (done(), do_checks(), check_fields(), write_data()) is the reference's own code operating on the
// reference's own host views, which this harness fills from the GPU.  Because all members live in headers, the private
// state is reached with `#define private public` instead of patching reference sources (same trick as ref_harness.cpp).
//
// --native-io also replaces the two callers either side of the path: do_checks' scalar ranges and
// check_fields' NaN test come from mlb_field_ranges (a device reduction), and the VTU files are written by
// mlb_write_vtu ("<prefix>_native_<step>.vtu") straight from the device-resident fields; the per-step
// copy_device_to_host() of the reference (solver.cpp:423) disappears.
//
// Usage: mallard_dropin -i input.toml [--fp strict|fast] [--quiet] [--native-io]
// Last line of stdout: JSON with steps, t, wall seconds of the loop and cell-updates/s per RK stage.
#include <sstream>
#define private public
#define protected public
#include "solver.h"
#include "face_reconstruction.h"
#include "mesh.h"
#include "zone.h"
#undef private
#undef protected

#include <Kokkos_Core.hpp>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../include/mallard_b200.h"

namespace {

int enum_of(const std::string & s, std::initializer_list<std::pair<const char *, int>> table, const char * what) {
    for (auto & kv : table) if (s == kv.first) return kv.second;
    throw std::runtime_error(std::string("Unknown ") + what + " type: " + s + ".");
}

#define MLB_OK(call)                                                                              \
    do {                                                                                          \
        if ((call) != 0) throw std::runtime_error(std::string(#call ": ") + mlb_last_error(ctx)); \
    } while (0)

}  // namespace

int main(int argc, char ** argv) {
    std::string input_file, fp = "strict";
    bool quiet = false, native_io = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-i") && i + 1 < argc) input_file = argv[++i];
        else if (!strcmp(argv[i], "--fp") && i + 1 < argc) fp = argv[++i];
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
        else if (!strcmp(argv[i], "--native-io")) native_io = true;
    }
    if (input_file.empty()) { fprintf(stderr, "usage: mallard_dropin -i input.toml [--fp strict|fast] [--quiet]\n"); return 2; }
    // this binary was compiled against one revision of the header; the library next to it may be newer (struct sizes!)
    if (mlb_check_abi(MLB_ABI_VERSION)) { fprintf(stderr, "mallard_dropin: %s\n", mlb_last_error(nullptr)); return 3; }
    Kokkos::initialize(argc, argv);
    int rc = 0;
    {
        std::streambuf * old = std::cout.rdbuf();
        std::ostringstream sink;
        if (quiet) std::cout.rdbuf(sink.rdbuf());
        mlb_ctx * ctx = nullptr;
        try {
            Solver s;
            if (s.init(input_file) != 0) throw std::runtime_error("Solver::init failed");
            Mesh & m = *s.mesh;

            // ---- the reference's host mesh arrays, as they are (mesh/mesh.h:242-253)
            std::vector<std::string> zone_names;                    // FaceZone::get_name() returns by value
            for (auto & z : *m.face_zones()) zone_names.push_back(z.get_name());
            std::vector<mlb_zone> zones;
            {
                size_t iz = 0;
                for (auto & z : *m.face_zones()) zones.push_back({zone_names[iz++].c_str(), (uint32_t)z.h_faces.extent(0), z.h_faces.data()});
            }
            mlb_mesh mm{};
            mm.n_cells = m.n_cells; mm.n_faces = m.n_faces; mm.n_nodes = m.n_nodes;
            mm.node_coords = m.h_node_coords.data();
            mm.offsets_nodes_of_cell = m.h_offsets_nodes_of_cell.data(); mm.nodes_of_cell = m.h_nodes_of_cell.data();
            mm.offsets_faces_of_cell = m.h_offsets_faces_of_cell.data(); mm.faces_of_cell = m.h_faces_of_cell.data();
            mm.offsets_nodes_of_face = m.h_offsets_nodes_of_face.data(); mm.nodes_of_face = m.h_nodes_of_face.data();
            mm.cells_of_face = m.h_cells_of_face.data();
            mm.cell_coords = m.h_cell_coords.data(); mm.cell_volume = m.h_cell_volume.data();
            mm.face_area = m.h_face_area.data(); mm.face_normals = m.h_face_normals.data();
            mm.n_zones = (uint32_t)zones.size(); mm.zones = zones.data();

            // ---- the same TOML keys the reference reads (solver.cpp:110-186, face_reconstruction.cpp:109-116, physics.cpp:27-53)
            const auto & in = s.input;
            mlb_numerics num{};
            const std::string recon = toml::find<std::string>(in, "numerics", "face_reconstruction", "type");
            num.recon = enum_of(recon, {{"FO", MLB_RECON_FO}, {"FirstOrder", MLB_RECON_FO}, {"TENO", MLB_RECON_TENO}}, "face reconstruction");
            num.riemann = enum_of(toml::find<std::string>(in, "numerics", "riemann_solver"),
                                  {{"Rusanov", MLB_RIEMANN_RUSANOV}, {"HLL", MLB_RIEMANN_HLL}, {"HLLC", MLB_RIEMANN_HLLC}}, "Riemann solver");
            num.integrator = enum_of(toml::find<std::string>(in, "numerics", "time_integrator"),
                                     {{"FE", MLB_INTEGRATOR_FE}, {"RK4", MLB_INTEGRATOR_RK4}, {"SSPRK3", MLB_INTEGRATOR_SSPRK3}}, "time integrator");
            num.basis = MLB_BASIS_LEGENDRE; num.basis_order = 3; num.max_stencil_size_factor = 2.0;
            if (num.recon == MLB_RECON_TENO) {
                const auto & fr = toml::find(in, "numerics", "face_reconstruction");
                num.basis = enum_of(toml::find_or<std::string>(fr, "basis_type", "monomial") /* the reference's default, face_reconstruction.cpp:110 */, {{"monomial", MLB_BASIS_MONOMIAL}, {"legendre", MLB_BASIS_LEGENDRE}}, "basis");
                num.basis_order = toml::find<int>(fr, "basis_order");   // required by the reference too (:111)
                num.max_stencil_size_factor = toml::find_or<double>(fr, "max_stencil_size_factor", 2.0);
                num.quadrature_order_cell = toml::find_or<int>(fr, "quadrature_order_cell", 0);
                num.quadrature_order_face = toml::find_or<int>(fr, "quadrature_order_face", 0);
            }
            num.fp_mode = fp == "fast" ? MLB_FP_FAST : MLB_FP_STRICT;
            num.renumber = MLB_RENUMBER_RCM;
            mlb_physics ph{};
            ph.gamma = toml::find_or<double>(in, "physics", "gamma", 1.4);
            ph.p_ref = toml::find_or<double>(in, "physics", "p_ref", 101325.0);
            ph.T_ref = toml::find_or<double>(in, "physics", "T_ref", 298.15);
            ph.rho_ref = toml::find_or<double>(in, "physics", "rho_ref", 1.225);
            ph.p_min = toml::find_or<double>(in, "physics", "p_min", -1e20);
            ph.p_max = toml::find_or<double>(in, "physics", "p_max", 1e20);
            std::vector<std::string> bc_names;
            std::vector<mlb_bc> bcs;
            for (const auto & b : toml::find<std::vector<toml::value>>(in, "boundaries")) {
                mlb_bc bc{};
                bc_names.push_back(toml::find<std::string>(b, "name"));
                bc.type = enum_of(toml::find<std::string>(b, "type"), {{"symmetry", MLB_BC_SYMMETRY}, {"extrapolation", MLB_BC_EXTRAPOLATION},
                                  {"wall_adiabatic", MLB_BC_WALL_ADIABATIC}, {"upt", MLB_BC_UPT}, {"p_out", MLB_BC_P_OUT}}, "boundary");
                if (b.contains("u")) { auto u = toml::find<std::vector<double>>(b, "u"); bc.u[0] = u[0]; bc.u[1] = u[1]; }
                bc.p = toml::find_or<double>(b, "p", 0.0);
                bc.T = toml::find_or<double>(b, "T", 0.0);
                bcs.push_back(bc);
            }
            for (size_t i = 0; i < bcs.size(); i++) bcs[i].zone_name = bc_names[i].c_str();

            if (mlb_create(&ctx, &mm, &num, &ph, bcs.data(), (int32_t)bcs.size(), nullptr) != 0)
                throw std::runtime_error(std::string("mlb_create: ") + mlb_last_error(nullptr));
            MLB_OK(mlb_set_state(ctx, s.conservatives.data(), s.primitives.data()));   // host backends: device view == host mirror

            // ---- Solver::run (solver.cpp:352-373) with the seams rerouted
            const int n_stages = mlb_n_stages(ctx);
            if (native_io) {
                // writers as the reference configured them (DataWriter::init): prefix, interval, variable names
                auto write_native = [&](bool force) {
                    for (auto & w : s.data_writers) {
                        if (!((s.step % w->interval == 0) || force)) continue;          // DataWriter::write, data_writer.cpp:77-80
                        std::vector<std::string> vn;
                        for (auto * d : w->data_ptrs) vn.push_back(d->name());
                        std::vector<const char *> vp;
                        for (auto & x : vn) vp.push_back(x.c_str());
                        MLB_OK(mlb_write_vtu(ctx, &mm, (w->prefix + "_native").c_str(), s.step, (int32_t)vp.size(), vp.data()));
                    }
                };
                static const char * const NAMES[9] = {"RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X", "U_Y", "P", "T", "H"};
                write_native(true);
                auto t0n = std::chrono::steady_clock::now();
                while (!s.done()) {
                    if (s.step % s.check_interval == 0) {                               // Solver::do_checks, solver.cpp:422-443
                        double mn[9], mx[9];
                        MLB_OK(mlb_field_ranges(ctx, mn, mx, nullptr));
                        s.print_step_info();
                        for (int i = 0; i < 9; i++)
                            std::cout << "> Scalar range: " << NAMES[i] << " = [" << mn[i] << ", " << mx[i] << "]" << std::endl;
                    }
                    if (s.use_cfl) MLB_OK(mlb_calc_dt(ctx, s.cfl, &s.dt));
                    else MLB_OK(mlb_set_dt(ctx, s.dt));
                    MLB_OK(mlb_take_step(ctx));
                    s.step++;
                    s.t += s.dt;
                    if (s.check_nan) {                                                  // Solver::check_fields, solver.cpp:470-498
                        uint64_t n_nan = 0;
                        MLB_OK(mlb_field_ranges(ctx, nullptr, nullptr, &n_nan));
                        if (n_nan) throw std::runtime_error("NaN found in fields.");
                    }
                    write_native(false);
                }
                auto t1n = std::chrono::steady_clock::now();
                write_native(true);
                std::cout.rdbuf(old);
                const double secn = std::chrono::duration<double>(t1n - t0n).count();
                printf("{\"n_cells\": %u, \"steps\": %u, \"t\": %.17g, \"dt_last\": %.17g, \"seconds\": %.6e, \"cell_updates_per_s_per_stage\": %.6e, "
                       "\"fp_mode\": \"%s\", \"launches\": %llu, \"native_io\": true}\n",
                       m.n_cells, (unsigned)s.step, (double)s.t, (double)s.dt, secn, (double)m.n_cells * n_stages * s.step / secn, fp.c_str(),
                       (unsigned long long)mlb_launch_count(ctx));
            } else {
            s.write_data(true);
            auto t0 = std::chrono::steady_clock::now();
            while (!s.done()) {
                MLB_OK(mlb_get_state(ctx, s.conservatives.data(), s.primitives.data(), s.cfl_local.data()));   // copy_device_to_host
                s.do_checks();
                if (s.use_cfl) MLB_OK(mlb_calc_dt(ctx, s.cfl, &s.dt));                                          // calc_dt
                else MLB_OK(mlb_set_dt(ctx, s.dt));
                MLB_OK(mlb_take_step(ctx));                                                                     // take_step
                s.step++;
                s.t += s.dt;
                if (s.check_nan || !s.data_writers.empty())
                    MLB_OK(mlb_get_state(ctx, s.conservatives.data(), s.primitives.data(), s.cfl_local.data()));
                s.check_fields();
                s.write_data();
            }
            auto t1 = std::chrono::steady_clock::now();
            MLB_OK(mlb_get_state(ctx, s.conservatives.data(), s.primitives.data(), s.cfl_local.data()));
            s.write_data(true);
            std::cout.rdbuf(old);
            const double sec = std::chrono::duration<double>(t1 - t0).count();
            printf("{\"n_cells\": %u, \"steps\": %u, \"t\": %.17g, \"dt_last\": %.17g, \"seconds\": %.6e, \"cell_updates_per_s_per_stage\": %.6e, "
                   "\"fp_mode\": \"%s\", \"launches\": %llu}\n",
                   m.n_cells, (unsigned)s.step, (double)s.t, (double)s.dt, sec, (double)m.n_cells * n_stages * s.step / sec, fp.c_str(),
                   (unsigned long long)mlb_launch_count(ctx));
            }
        } catch (const std::exception & e) {
            std::cout.rdbuf(old);
            fprintf(stderr, "mallard_dropin: %s\n", e.what());
            rc = 1;
        }
        if (ctx) mlb_destroy(ctx);
        std::cout.rdbuf(old);
    }
    Kokkos::finalize();
    return rc;
}
