#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/bin/ref_harness, built from /root/reference by oracle/build_ref.sh) on small cases.

Run here (the container that has /root/reference); the fixtures are committed so that the tests can
run on the GPU box where the reference sources do not exist:

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py sod teno   # name filter

Every fixture holds: `meta` (JSON: the case description a test needs to rebuild the same configuration
through the oracle and through the C-ABI library), the initial state, and the reference's stage-level
outputs (see oracle/ref_harness.cpp).  OMP_NUM_THREADS=1 makes the reference's atomic scatter
deterministic (src/numerics/flux_functor.h:156-161).
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mlbd  # noqa: E402

HARNESS = os.path.join(HERE, "_ref", "bin", "ref_harness")
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

SOD_IC = dict(type="analytical",
              rho="var l := x <  0.5; var r := x >= 0.5; 1.0 * l + 0.125 * r",
              u=["0.0", "0.0"],
              p="var l := x <  0.5; var r := x >= 0.5; 1.0 * l + 0.1 * r")
_Q = "var l := x <  0.8; var r := x >= 0.8; var b := y <  0.8; var t := y >= 0.8; "
RIEMANN2D_IC = dict(type="analytical",
                    rho=_Q + "1.5 * r * t + 0.532258064516129 * l * t + 0.137992831541219 * l * b + 0.532258064516129 * r * b",
                    u=[_Q + "0.0 * r * t + 1.206045378311055 * l * t + 1.206045378311055 * l * b + 0.0 * r * b",
                       _Q + "0.0 * r * t + 0.0 * l * t + 1.206045378311055 * l * b + 1.206045378311055 * r * b"],
                    p=_Q + "1.5 * r * t + 0.3 * l * t + 0.029032258064516 * l * b + 0.3 * r * b")
SMOOTH_IC = dict(type="analytical",
                 rho="1.0 + 0.2 * sin(2 * pi * x) * cos(2 * pi * y)",
                 u=["0.5 + 0.1 * cos(2 * pi * x)", "0.3 + 0.1 * sin(2 * pi * y)"],
                 p="1.0 + 0.1 * cos(2 * pi * (x + y))")
WEDGE_IC = dict(type="constant", u=[600.0, 0.0], p=101325.0, T=300.0)
SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]
WEDGE_BCS = [dict(name="left", type="upt", u=[600.0, 0.0], p=101325.0, T=300.0), dict(name="right", type="p_out", p=101325.0),
             dict(name="top", type="symmetry"), dict(name="bottom", type="symmetry")]
WEDGE_WALL_BCS = WEDGE_BCS[:3] + [dict(name="bottom", type="wall_adiabatic")]
EXTRAP4 = [dict(name=n, type="extrapolation") for n in ("left", "right", "top", "bottom")]
TENO3 = dict(type="TENO", basis_type="legendre", basis_order=3, max_stencil_size_factor=2.0)

CASES = {
    # BASELINE.json configs[0]: examples/sod verbatim (851 steps to t_stop=0.2); dumps step 0 and the last step.
    "sod": dict(mesh=dict(type="cartesian", Nx=1000, Ny=1, Lx=1.0, Ly=0.001), ic=SOD_IC, bcs=SYM4, cfl=1.0,
                riemann="HLLC", integrator="SSPRK3", recon=dict(type="FO"), n_steps=851, every=851, keep_mesh=False),
    "sod_hll_rk4": dict(mesh=dict(type="cartesian", Nx=200, Ny=1, Lx=1.0, Ly=0.005), ic=SOD_IC, bcs=SYM4, cfl=0.8,
                        riemann="HLL", integrator="RK4", recon=dict(type="FO"), n_steps=40, every=40, keep_mesh=True),
    "sod_rusanov_fe": dict(mesh=dict(type="cartesian", Nx=200, Ny=1, Lx=1.0, Ly=0.005), ic=SOD_IC, bcs=SYM4, cfl=0.5,
                           riemann="Rusanov", integrator="FE", recon=dict(type="FO"), n_steps=40, every=40, keep_mesh=False),
    # BASELINE.json configs[2]: examples/wedge downsized 5x per direction, plus the wall_adiabatic variant (K7).
    "wedge_30x10": dict(mesh=dict(type="wedge", Nx=30, Ny=10, Lx=4.0, Ly=1.5), ic=WEDGE_IC, bcs=WEDGE_BCS, cfl=1.0,
                        riemann="HLLC", integrator="SSPRK3", recon=dict(type="FO"), n_steps=200, every=100, keep_mesh=True),
    "wedge_wall_30x10": dict(mesh=dict(type="wedge", Nx=30, Ny=10, Lx=4.0, Ly=1.5), ic=WEDGE_IC, bcs=WEDGE_WALL_BCS, cfl=1.0,
                             riemann="HLLC", integrator="SSPRK3", recon=dict(type="FO"), n_steps=200, every=100, keep_mesh=False),
    "fo_tri_extrap": dict(mesh=dict(type="cartesian_tri", Nx=12, Ny=9, Lx=1.0, Ly=0.75), ic=RIEMANN2D_IC, bcs=EXTRAP4, cfl=0.5,
                          riemann="HLLC", integrator="SSPRK3", recon=dict(type="FO"), n_steps=30, every=30, keep_mesh=True),
    # BASELINE.json configs[1]: examples/riemann_2d numerics on a small cartesian_tri mesh.  Smooth IC: every cell
    # stays on the central branch in stage 1; riemann IC: the reference produces Inf/NaN.
    "teno_smooth_6x6": dict(mesh=dict(type="cartesian_tri", Nx=6, Ny=6, Lx=1.0, Ly=1.0), ic=SMOOTH_IC, bcs=SYM4, cfl=0.1,
                            riemann="HLLC", integrator="SSPRK3", recon=TENO3, n_steps=1, every=1, keep_mesh=True, keep_teno=True),
    "teno_riemann_8x8": dict(mesh=dict(type="cartesian_tri", Nx=8, Ny=8, Lx=1.0, Ly=1.0), ic=RIEMANN2D_IC, bcs=SYM4, cfl=0.1,
                             riemann="HLLC", integrator="SSPRK3", recon=TENO3, n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    "teno_smooth_10x7_p2": dict(mesh=dict(type="cartesian_tri", Nx=10, Ny=7, Lx=1.0, Ly=0.7), ic=SMOOTH_IC, bcs=EXTRAP4, cfl=0.1,
                                riemann="Rusanov", integrator="SSPRK3",
                                recon=dict(type="TENO", basis_type="legendre", basis_order=2, max_stencil_size_factor=2.0),
                                n_steps=1, every=1, keep_mesh=False, keep_teno=True),
    # the lowest and the highest order the device kernels are instantiated for (K = 3, M = 6 and K = 15, M = 30)
    "teno_legendre_6x5_p1": dict(mesh=dict(type="cartesian_tri", Nx=6, Ny=5, Lx=1.0, Ly=0.8), ic=SMOOTH_IC, bcs=SYM4, cfl=0.1,
                                 riemann="Rusanov", integrator="FE",
                                 recon=dict(type="TENO", basis_type="legendre", basis_order=1, max_stencil_size_factor=2.0),
                                 n_steps=1, every=1, keep_mesh=False, keep_teno=True),
    "teno_legendre_9x8_p4": dict(mesh=dict(type="cartesian_tri", Nx=9, Ny=8, Lx=1.0, Ly=0.9), ic=SMOOTH_IC, bcs=EXTRAP4, cfl=0.1,
                                 riemann="HLLC", integrator="SSPRK3",
                                 recon=dict(type="TENO", basis_type="legendre", basis_order=4, max_stencil_size_factor=2.0),
                                 n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    # the reference's DEFAULT basis (face_reconstruction.cpp:110): monomials, incl. the derivative quirk in the oscillation
    # indicator (basis.h:72-78)
    "teno_monomial_7x6_p3": dict(mesh=dict(type="cartesian_tri", Nx=7, Ny=6, Lx=1.0, Ly=1.0), ic=SMOOTH_IC, bcs=SYM4, cfl=0.1,
                                 riemann="HLLC", integrator="SSPRK3",
                                 recon=dict(type="TENO", basis_type="monomial", basis_order=3, max_stencil_size_factor=2.0),
                                 n_steps=1, every=1, keep_mesh=False, keep_teno=True),
    # configurations beyond the specialised device kernels (generic kernel): order 5 (K = 21, M = 42; the reference has Dunavant
    # rules up to order 5 only, so quadrature_order_cell must be given) and a stencil-size factor other than 2
    "teno_legendre_12x10_p5": dict(mesh=dict(type="cartesian_tri", Nx=12, Ny=10, Lx=1.2, Ly=1.0), ic=SMOOTH_IC, bcs=SYM4, cfl=0.1,
                                   riemann="HLLC", integrator="SSPRK3",
                                   recon=dict(type="TENO", basis_type="legendre", basis_order=5, max_stencil_size_factor=2.0,
                                              quadrature_order_cell=5),
                                   n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    "teno_legendre_8x7_p2_f15": dict(mesh=dict(type="cartesian_tri", Nx=8, Ny=7, Lx=1.0, Ly=0.9), ic=SMOOTH_IC, bcs=EXTRAP4, cfl=0.1,
                                     riemann="HLLC", integrator="SSPRK3",
                                     recon=dict(type="TENO", basis_type="legendre", basis_order=2, max_stencil_size_factor=1.5),
                                     n_steps=1, every=1, keep_mesh=False, keep_teno=True),
    # the upper end of what the reference accepts (basis_order up to 9: K = 55, M = 110; basis.h holds Legendre polynomials up to degree 9)
    # and an odd order in between with the reference's default basis - pins the oracle where only the generic device kernel runs
    "teno_monomial_14x12_p7": dict(mesh=dict(type="cartesian_tri", Nx=14, Ny=12, Lx=1.2, Ly=1.0), ic=SMOOTH_IC, bcs=EXTRAP4, cfl=0.1,
                                   riemann="HLLC", integrator="SSPRK3",
                                   recon=dict(type="TENO", basis_type="monomial", basis_order=7, max_stencil_size_factor=2.0,
                                              quadrature_order_cell=5),
                                   n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    "teno_legendre_16x14_p9": dict(mesh=dict(type="cartesian_tri", Nx=16, Ny=14, Lx=1.2, Ly=1.0), ic=SMOOTH_IC, bcs=SYM4, cfl=0.1,
                                   riemann="Rusanov", integrator="SSPRK3",
                                   recon=dict(type="TENO", basis_type="legendre", basis_order=9, max_stencil_size_factor=2.0,
                                              quadrature_order_cell=5),
                                   n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    # TENO face states through every boundary functor (upt, p_out, wall_adiabatic, symmetry: boundary/*.cpp) and through RK4's four
    # stages; TENO + HLL on the four-quadrant data (non-finite pattern of the HLL flux)
    "teno_bcs_rk4_10x8": dict(mesh=dict(type="cartesian_tri", Nx=10, Ny=8, Lx=1.0, Ly=0.8), ic=SMOOTH_IC,
                              bcs=[dict(name="left", type="upt", u=[0.5, 0.3], p=1.0, T=0.0036), dict(name="right", type="p_out", p=0.95),
                                   dict(name="top", type="symmetry"), dict(name="bottom", type="wall_adiabatic")],
                              cfl=0.1, riemann="HLLC", integrator="RK4", recon=TENO3, n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    "teno_hll_riemann_9x7": dict(mesh=dict(type="cartesian_tri", Nx=9, Ny=7, Lx=1.0, Ly=0.8), ic=RIEMANN2D_IC, bcs=EXTRAP4, cfl=0.1,
                                 riemann="HLL", integrator="SSPRK3", recon=TENO3, n_steps=1, every=1, keep_mesh=False, keep_teno=False),
    "teno_monomial_9x8_p2": dict(mesh=dict(type="cartesian_tri", Nx=9, Ny=8, Lx=1.2, Ly=1.0), ic=SMOOTH_IC, bcs=EXTRAP4, cfl=0.1,
                                 riemann="HLL", integrator="RK4",
                                 recon=dict(type="TENO", basis_type="monomial", basis_order=2, max_stencil_size_factor=2.0),
                                 n_steps=1, every=1, keep_mesh=False, keep_teno=False),
}


def toml_value(v):
    if isinstance(v, str):
        return '"' + v + '"'
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (list, tuple)):
        return "[" + ", ".join(toml_value(x) for x in v) + "]"
    return repr(v)


def write_toml(case, path, n_steps=None):
    L = ["[run]", "n_steps = %d" % (n_steps or case["n_steps"]), "cfl = %r" % case["cfl"], "", "[mesh]"]
    L += ["%s = %s" % (k, toml_value(v)) for k, v in case["mesh"].items()]
    L += ["", "[initialize]"] + ["%s = %s" % (k, toml_value(v)) for k, v in case["ic"].items()]
    for bc in case["bcs"]:
        L += ["", "[[boundaries]]"] + ["%s = %s" % (k, toml_value(v)) for k, v in bc.items()]
    L += ["", "[numerics]", 'riemann_solver = "%s"' % case["riemann"], 'time_integrator = "%s"' % case["integrator"],
          "check_nan = false", "", "[numerics.face_reconstruction]"]
    L += ["%s = %s" % (k, toml_value(v)) for k, v in case["recon"].items()]
    ph = dict(gamma=1.4, p_ref=101325.0, T_ref=298.15, rho_ref=1.225)
    ph.update(case.get("physics", {}))          # (oracle/pin_sweep.py varies the gas: gamma, reference state, pressure clamp)
    L += ["", "[physics]", 'type = "euler"'] + ["%s = %r" % (k, float(v)) for k, v in ph.items()] + ["", "[output]", "check_interval = 1000000", ""]
    with open(path, "w") as f:
        f.write("\n".join(L))


MESH_KEYS = ["node_coords", "cell_coords", "cell_volume", "face_area", "face_normals", "nodes_of_cell",
             "offsets_nodes_of_cell", "faces_of_cell", "offsets_faces_of_cell", "nodes_of_face", "offsets_nodes_of_face",
             "cells_of_face"]


def run_case(name, case):
    with tempfile.TemporaryDirectory() as td:
        toml = os.path.join(td, "input.toml")
        out = os.path.join(td, "out.mlbd")
        write_toml(case, toml)
        env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
        subprocess.check_call([HARNESS, "dump", toml, out, str(case["n_steps"]), str(case["every"])], env=env,
                              stdout=subprocess.DEVNULL)
        d = mlbd.read(out)
    keep = {}
    for k, v in d.items():
        if k in MESH_KEYS or k.startswith("zone:"):
            if case.get("keep_mesh"):
                keep[k] = v
        elif k.startswith("teno:"):
            if case.get("keep_teno") or k in ("teno:meta", "teno:integral_psi_target", "teno:oscillation_indicator",
                                              "teno:poly_indices", "teno:quad_face_points"):
                keep[k] = v
            elif k == "teno:stencils" or k.startswith("teno:offsets"):
                keep[k] = v   # integer connectivity is always kept (bit-exact check)
        else:
            keep[k] = v
    meta = {kk: vv for kk, vv in case.items() if kk not in ("keep_mesh", "keep_teno")}
    keep["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **keep)
    print("%-24s %8.1f kB  %d arrays" % (name, os.path.getsize(path) / 1024, len(keep)))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if not os.path.exists(HARNESS):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    sel = sys.argv[1:]
    for name, case in CASES.items():
        if sel and not any(s in name for s in sel):
            continue
        run_case(name, case)


if __name__ == "__main__":
    main()
