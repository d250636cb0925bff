"""Pins the CPU restatement (oracle/mallard_oracle.cpp) against the real reference:
 (1) the golden vectors of the reference's own gtest suite,
 (2) stage-level dumps of the unmodified reference (tests/golden/*.npz; oracle/make_golden.py).
The oracle is compiled without FMA like the reference's x86-64 build, so (2) is asserted BIT-EXACT.
"""
import os

import numpy as np
import pytest

import golden_util as gu

SOD_L = [1.0, 0.0, 0.0, 1.0]      # rho, u, v, p
SOD_R = [0.125, 0.0, 0.0, 0.1]
GAMMA = 1.4


def _row(rho, u, v, p):
    # h = e + p/rho, e = p/((gamma-1) rho)   (test/riemann_solver_test.cpp:30-40)
    e = p / ((GAMMA - 1.0) * rho)
    return [rho, u, v, p, e + p / rho]


S2 = 1.0 / np.sqrt(2.0)
# (solver, n_unit, expected flux) — test/riemann_solver_test.cpp:43-46,74-77,106-109,168-171,199-202,231-234
RIEMANN_KAT = [
    ("Rusanov", (1.0, 0.0), (0.51765698, 0.55, 0.0, 1.33111795)),
    ("Rusanov", (0.0, 1.0), (0.51765698, 0.0, 0.55, 1.33111795)),
    ("Rusanov", (S2, S2), (0.51765698, 0.38890873, 0.38890873, 1.33111795)),
    ("HLLC", (1.0, 0.0), (0.415322226496596, 0.508584114470313, 0.0, 1.139144729421316)),
    ("HLLC", (0.0, 1.0), (0.415322226496596, 0.0, 0.508584114470313, 1.139144729421316)),
    ("HLLC", (S2, S2), (0.4153222265, 0.3596232761, 0.3596232761, 1.1391447294)),
]


@pytest.mark.parametrize("kind,n,expect", RIEMANN_KAT)
def test_riemann_known_answers(oracle_mod, kind, n, expect):
    f = oracle_mod.riemann_flux(kind, [n], [_row(*SOD_L)], [_row(*SOD_R)], GAMMA)[0]
    np.testing.assert_allclose(f, expect, atol=1e-6, rtol=0)


@pytest.mark.parametrize("kind", ["Rusanov", "HLL", "HLLC"])
def test_riemann_uniform_state(oracle_mod, kind):
    # test/riemann_solver_test.cpp:137-140,262-265: identical states at rest -> pure pressure flux (0, p nx, p ny, 0)
    st = _row(1.0, 0.0, 0.0, 1.0)
    f = oracle_mod.riemann_flux(kind, [(1.0, 0.0)], [st], [st], GAMMA)[0]
    np.testing.assert_allclose(f, (0.0, 1.0, 0.0, 0.0), atol=1e-6, rtol=0)


def test_physics_constants_and_round_trip(oracle_mod):
    # test/physics_test.cpp:30-32 and :60-112
    rho, u, v, p = 1.225, 10.0, 5.0, 101325.0
    _, rc = oracle_mod.prims([[1, 0, 0, 1]])
    np.testing.assert_allclose(rc, (277.42507366857529, 970.98775784001373, 693.56268417143838), rtol=1e-12)
    R, cp, cv = rc
    T = p / (rho * R)
    e = cv * T
    U = [rho, rho * u, rho * v, rho * (e + 0.5 * (u * u + v * v))]
    P, _ = oracle_mod.prims([U])
    np.testing.assert_allclose(P[0], (u, v, p, T, e + p / rho), rtol=1e-6)


@pytest.mark.parametrize("integ", ["FE", "RK4", "SSPRK3"])
def test_integrators_constant_rhs(oracle_mod, integ):
    # test/time_integrator_test.cpp:22-28,75-76,126-127,177-178: rhs[c][i] = 4c+i = U0, dt = 0.1 -> U = 1.1 U0
    m = oracle_mod.Mesh.generate("cartesian", 2, 1, 1.0, 1.0)
    s = oracle_mod.Solver(m, "FO", "HLLC", integ)
    U0 = np.arange(8, dtype=np.float64).reshape(2, 4)
    s.set_state(np.where(U0 == 0, 0.0, U0), P=np.ones((2, 5)))
    s.set_rhs_override(U0)
    s.take_step(0.1)
    np.testing.assert_allclose(s.get("U"), 1.1 * U0, rtol=1e-6)


@pytest.mark.parametrize("name", gu.names())
def test_bit_exact_against_reference_dump(oracle_mod, name):
    meta, g = gu.load(name)
    mesh = gu.oracle_mesh(oracle_mod, meta)
    if "cells_of_face" in g:   # mesh generator parity (connectivity AND geometry bit-exact)
        for k in gu.MESH_KEYS:
            ref = g[k]
            got = mesh.get(k).reshape(ref.shape)
            if k == "face_normals":   # phantom faces: 0/0 = NaN in both
                assert np.array_equal(np.isnan(ref), np.isnan(got))
                ref, got = np.nan_to_num(ref), np.nan_to_num(got)
            assert np.array_equal(ref, got), k
        for i, z in enumerate(gu.ZONES):
            assert np.array_equal(g["zone:%d:%s" % (i, z)], mesh.zone(z)), z
    s = gu.oracle_solver(oracle_mod, meta, mesh)
    if meta["recon"]["type"] == "TENO":
        for k in g:
            if k.startswith("teno:") and k not in ("teno:meta", "teno:quad_cell_points", "teno:quad_cell_weights"):
                got = s.get(k).reshape(g[k].shape)
                assert np.array_equal(got, g[k]), k
    s.set_state(g["U0"], g["P0"])
    F = s.calc_face_values().copy()
    real = gu.real_faces(mesh.get("cells_of_face"), mesh.get("nodes_of_face"))
    cof = mesh.get("cells_of_face")
    interior = real & (cof[:, 1] >= 0)
    assert np.array_equal(F[real][:, :, 0], g["F_stage1"][real][:, :, 0], equal_nan=True)
    assert np.array_equal(F[interior][:, :, 1], g["F_stage1"][interior][:, :, 1], equal_nan=True)
    assert np.array_equal(s.calc_rhs(), g["rhs_stage1"], equal_nan=True)
    n_rhs = {"FE": 1, "RK4": 4, "SSPRK3": 3}[meta["integrator"]]
    for i in range(meta["n_steps"]):
        dt = s.calc_dt(meta["cfl"])
        key = "step%d:" % i
        if key + "dt" in g:
            assert dt == g[key + "dt"][0] or (np.isnan(dt) and np.isnan(g[key + "dt"][0]))
            assert np.array_equal(s.get("cfl_local"), g[key + "cfl_local"], equal_nan=True)
        s.take_step(dt)
        if key + "U" in g:
            for r in range(n_rhs):
                assert np.array_equal(s.get("rhs%d" % r), g[key + "rhs%d" % r], equal_nan=True), (i, r)
            assert np.array_equal(s.get("U"), g[key + "U"], equal_nan=True), i
            assert np.array_equal(s.get("P"), g[key + "P"], equal_nan=True), i


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "ref_harness")),
                    reason="the unmodified reference is not built here (oracle/build_ref.sh needs /root/reference)")
@pytest.mark.parametrize("case_name", ["p6 monomial", "p3 legendre quadrature_order_face 4", "p3 legendre factor 2.5", "first order wedge Rusanov RK4"])
def test_pin_sweep_cases_run_live_against_the_unmodified_reference(oracle_mod, case_name):
    """oracle/pin_sweep.py (profiles/r02h_pin_sweep.txt: 44 configurations, every array bit for bit) - a few of its cases run here, live:
    the reference dumps a step on the spot, the oracle and the preprocessor's reference-layout tables must reproduce every array."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pin_sweep
    case = dict(pin_sweep.cases())[case_name]
    res = pin_sweep.compare(case)
    assert len(res) >= 7
    assert all(v == "ok" for _, v in res), [(w, v) for w, v in res if v != "ok"]


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "ref_harness")),
                    reason="the unmodified reference is not built here (oracle/build_ref.sh needs /root/reference)")
def test_unstructured_meshes_injected_into_the_unmodified_reference(oracle_mod, capsys):
    """The jittered, id-shuffled triangulations of the strong-scaling records (BASELINE configs[3] family) have no generator in the reference:
    the harness injects their connectivity (`ref_harness mesh`), the reference computes its own geometry, stencils and matrices on it, and the
    oracle, the preprocessor and the emulated kernels (STRICT: every bit) must agree with it - oracle/pin_sweep.py --unstructured, live."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pin_sweep
    assert pin_sweep.compare_unstructured() == 0
    out = capsys.readouterr().out
    assert out.count("all bit-exact") == 9 and "DIFFER" not in out
    strict = [l.split("kernels: strict ")[1].split("   fast")[0] for l in out.splitlines() if "kernels: strict" in l]
    assert sum(s == "F 0.0e+00 rhs 0.0e+00 step 0.0e+00" for s in strict) >= 8 and all(s.endswith("step 0.0e+00") for s in strict)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "ref_harness")),
                    reason="the unmodified reference is not built here (oracle/build_ref.sh needs /root/reference)")
def test_riemann_fluxes_on_random_states_against_the_unmodified_reference(oracle_mod, capsys):
    """oracle/pin_sweep.py --riemann, live on 30 000 states per gamma: densities and pressures over six decades, Mach numbers up to 6, weak jumps
    and identical states - the reference's own flux functions (`ref_harness riemann`) against the oracle (every bit, non-finite results
    included) and against the kernel source's flux functions on the host: STRICT every bit, FAST (the lean re-formulation) within 1e-10 of
    the flux scale everywhere and within 1e-12 on all but a handful of near-vacuum contact states."""
    import re
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pin_sweep
    assert pin_sweep.compare_riemann(n=30000, seed=7) == 0
    out = capsys.readouterr().out
    assert out.count("oracle: every bit of 30000 fluxes") == 6
    assert out.count("kernel STRICT: 0 states differ (max 0.0e+00 of the flux scale), non-finite pattern equal") == 6
    fast = re.findall(r"FAST \(lean\): max (\S+) of the flux scale \((\d+) of (\d+) states above 1e-12\), (\d+) states finite in one", out)
    assert len(fast) == 6
    for mx, above, total, mismatch in fast:
        assert float(mx) <= 1e-10 and int(above) <= 1e-3 * int(total) and int(mismatch) == 0


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "ref_harness")),
                    reason="the unmodified reference is not built here (oracle/build_ref.sh needs /root/reference)")
def test_primitives_of_random_states_against_the_unmodified_reference(oracle_mod, capsys):
    """oracle/pin_sweep.py --prims, live on 20 000 states per gas: Euler::compute_primitives_from_conservatives called directly (`ref_harness prims`),
    incl. negative internal energies, an active pressure clamp and NaN energies - oracle and STRICT kernel source: every bit."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pin_sweep
    assert pin_sweep.compare_prims(n=20000, seed=3) == 0
    out = capsys.readouterr().out
    assert out.count("oracle every bit;  kernel STRICT every bit") == 4 and out.count("non-finite pattern equal") == 4
