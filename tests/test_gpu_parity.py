"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI (mallard_b200.Solver is a thin ctypes
wrapper).  Three layers:
  1. the reference's own known-answer tests (Riemann fluxes, EOS, integrators) re-run against the device code;
  2. stage-level comparison with dumps of the unmodified reference (tests/golden, see oracle/make_golden.py);
  3. comparison with the oracle (CPU restatement, itself pinned bit-exact to the reference) on seeded inputs, and
     size-independent properties at larger sizes.
Tolerance: north_star asks for <= 1e-12 relative per step.  In STRICT mode (no FMA contraction, reference summation
order) the first-order path is additionally asserted to be BIT-EXACT wherever libm pow is not involved.
"""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import mallard_b200 as mb
from conftest import UNVERIFIED_ON_HARDWARE

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")

TOL = 1e-12
SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]
GAMMA = 1.4


def _row(rho, u, v, p):
    e = p / ((GAMMA - 1.0) * rho)
    return [rho, u, v, p, e + p / rho]


S2 = 1.0 / np.sqrt(2.0)
RIEMANN_KAT = [   # test/riemann_solver_test.cpp:43-46,74-77,106-109,168-171,199-202,231-234
    ("Rusanov", (1.0, 0.0), (0.51765698, 0.55, 0.0, 1.33111795)),
    ("Rusanov", (0.0, 1.0), (0.51765698, 0.0, 0.55, 1.33111795)),
    ("Rusanov", (S2, S2), (0.51765698, 0.38890873, 0.38890873, 1.33111795)),
    ("HLLC", (1.0, 0.0), (0.415322226496596, 0.508584114470313, 0.0, 1.139144729421316)),
    ("HLLC", (0.0, 1.0), (0.415322226496596, 0.0, 0.508584114470313, 1.139144729421316)),
    ("HLLC", (S2, S2), (0.4153222265, 0.3596232761, 0.3596232761, 1.1391447294)),
]


@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("kind,n,expect", RIEMANN_KAT)
def test_riemann_known_answers(kind, n, expect, fp):
    f = mb.riemann_flux(kind, [n], [_row(1.0, 0, 0, 1.0)], [_row(0.125, 0, 0, 0.1)], GAMMA, fp_mode=fp)[0]
    np.testing.assert_allclose(f, expect, atol=1e-6, rtol=0)


@pytest.mark.parametrize("kind", ["Rusanov", "HLL", "HLLC"])
def test_riemann_uniform_state(kind):   # test/riemann_solver_test.cpp:137-140,262-265
    st = _row(1.0, 0.0, 0.0, 1.0)
    f = mb.riemann_flux(kind, [(1.0, 0.0)], [st], [st], GAMMA)[0]
    np.testing.assert_allclose(f, (0.0, 1.0, 0.0, 0.0), atol=1e-6, rtol=0)


@pytest.mark.parametrize("kind", ["Rusanov", "HLL", "HLLC"])
def test_riemann_random_states_vs_oracle(oracle_mod, kind):
    rng = np.random.default_rng(7)
    n = 20000
    def states():
        rho = rng.uniform(0.05, 3.0, n); p = rng.uniform(0.02, 4.0, n) * 10.0 ** rng.integers(-1, 2, n)
        u = rng.normal(0, 1.5, n); v = rng.normal(0, 1.5, n)
        e = p / ((GAMMA - 1) * rho)
        return np.stack([rho, u, v, p, e + p / rho], 1)
    th = rng.uniform(0, 2 * np.pi, n)
    nu = np.stack([np.cos(th), np.sin(th)], 1)
    L, R = states(), states()
    ref = oracle_mod.riemann_flux(kind, nu, L, R, GAMMA)
    strict = mb.riemann_flux(kind, nu, L, R, GAMMA, fp_mode="strict")
    fast = mb.riemann_flux(kind, nu, L, R, GAMMA, fp_mode="fast")
    ok = np.isfinite(ref).all(axis=1)                        # extreme random pairs can leave pow() of a negative base: NaN in both
    assert ok.mean() > 0.9
    assert np.array_equal(np.isfinite(strict).all(axis=1), ok) and np.array_equal(np.isfinite(fast).all(axis=1), ok)
    scale = np.abs(ref[ok]).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(strict[ok] - ref[ok]) / scale) < 1e-13   # libm pow (TRRS branch) is the only non-identical operation
    assert np.max(np.abs(fast[ok] - ref[ok]) / scale) < 1e-12
    if kind == "Rusanov":
        assert np.array_equal(strict, ref)                   # no pow on this path: bit-exact


def test_physics_constants_and_round_trip():   # test/physics_test.cpp:30-32,60-112
    rho, u, v, p = 1.225, 10.0, 5.0, 101325.0
    _, rc = mb.compute_primitives(np.zeros((0, 4)))
    np.testing.assert_allclose(rc, (277.42507366857529, 970.98775784001373, 693.56268417143838), rtol=1e-12)
    R, cp, cv = rc
    T = p / (rho * R); e = cv * T
    U = [rho, rho * u, rho * v, rho * (e + 0.5 * (u * u + v * v))]
    P, _ = mb.compute_primitives([U])
    np.testing.assert_allclose(P[0], (u, v, p, T, e + p / rho), rtol=1e-6)


@pytest.mark.parametrize("integ", ["FE", "RK4", "SSPRK3"])
def test_integrators_constant_rhs(integ):   # test/time_integrator_test.cpp:22-28,75-76,126-127,177-178
    m = mb.Mesh.generate("cartesian", 2, 1, 1.0, 1.0)
    s = mb.Solver(m, "FO", "HLLC", integ, bcs=SYM4)
    U0 = np.arange(8, dtype=np.float64).reshape(2, 4)
    s.set_state(U0, P=np.ones((2, 5)))
    s.set_rhs_override(U0)
    s.take_step(0.1)
    np.testing.assert_allclose(s.get_state(), 1.1 * U0, rtol=1e-6)
    t, step = s.time()
    assert step == 1 and abs(t - 0.1) < 1e-15


def _solver(meta, mesh, fp, **kw):
    return mb.Solver(mesh, fp_mode=fp, **gu.solver_kwargs(meta), **kw)


BIT_EXACT_STRICT = {"sod_rusanov_fe", "wedge_30x10", "wedge_wall_30x10"}   # no libm pow anywhere on these paths


GENERIC_KERNEL_FIXTURES = {"teno_legendre_12x10_p5", "teno_legendre_8x7_p2_f15", "teno_monomial_14x12_p7", "teno_legendre_16x14_p9"}     # served by csrc/teno_generic.cuh
# reference dumps added after the round-2 GPU budget ran out (TENO through upt / p_out / wall boundaries + RK4; TENO + HLL on the four-quadrant
# data): the kernels they exercise HAVE run on a B200, these comparisons have not - gated like the rest (tests/test_zz_gpu_hardware_trial.py)
LATE_FIXTURES = {"teno_bcs_rk4_10x8", "teno_hll_riemann_9x7"}


@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("name", gu.names())
def test_against_reference_dumps(name, fp):
    if name in GENERIC_KERNEL_FIXTURES and os.environ.get("MLB_RUN_UNVERIFIED") != "1":
        pytest.skip("generic TENO kernel: written after the round-2 GPU budget ran out, not yet run on a B200 (set MLB_RUN_UNVERIFIED=1)")
    if name in LATE_FIXTURES and os.environ.get("MLB_RUN_UNVERIFIED") != "1":
        pytest.skip("fixture added after the round-2 GPU budget ran out: comparison not yet run on a B200 (set MLB_RUN_UNVERIFIED=1)")
    meta, g = gu.load(name)
    mm = meta["mesh"]
    mesh = mb.Mesh.generate(mm["type"], mm["Nx"], mm["Ny"], mm["Lx"], mm["Ly"])
    s = _solver(meta, mesh, fp)
    teno = meta["recon"]["type"] == "TENO"
    s.set_state(g["U0"], g["P0"])
    F = s.calc_face_values()
    cof = mesh.arrays["cells_of_face"]
    real = gu.real_faces(cof, mesh.arrays["nodes_of_face"])
    interior = real & (cof[:, 1] >= 0)
    exact = fp == "strict" and name in BIT_EXACT_STRICT
    worst = 0.0

    def check(a, b, what):
        nonlocal worst
        if exact:
            assert np.array_equal(a, b, equal_nan=True), what
        e = gu.rel_err(a, b) if fp == "strict" else gu.field_err(a, b)   # strict: element-wise; fast: relative to field scale
        if fp == "fast" and what.startswith("rhs"):
            # reference-faithful TENO on discontinuous data: a later stage's residual can be non-finite everywhere except for a few entries of
            # pure rounding noise (teno_hll_riemann_9x7: rhs1 is NaN but for entries of 1e-15) - noise has no field scale of its own, it is
            # measured against the scale of the first stage's residual
            fin_b, fin_1 = np.isfinite(b), np.isfinite(g["rhs_stage1"])
            scale_1 = np.abs(g["rhs_stage1"][fin_1]).max() if fin_1.any() else 0.0
            if fin_b.any() and np.abs(b[fin_b]).max() < 1e-9 * scale_1 and np.array_equal(np.isfinite(a), fin_b):
                e = float(np.abs(a[fin_b] - b[fin_b]).max() / scale_1)
        worst = max(worst, e)
        # residuals are differences of face fluxes: with FMA contraction their cancellation amplifies rounding, so the
        # 1e-12 bar applies to the conserved/primitive fields and dt; residual arrays get 1e-10 of the field scale
        assert e <= (1e-10 if (fp == "fast" and what.startswith("rhs")) else TOL), (what, e)

    if not teno:   # first-order face values are copies: always bit-exact
        assert np.array_equal(F[real][:, :, 0], g["F_stage1"][real][:, :, 0])
        assert np.array_equal(F[interior][:, :, 1], g["F_stage1"][interior][:, :, 1])
    else:
        check(F[real][:, :, 0], g["F_stage1"][real][:, :, 0], "F side 0")
        check(F[interior][:, :, 1], g["F_stage1"][interior][:, :, 1], "F side 1")
    check(s.calc_rhs(), g["rhs_stage1"], "rhs_stage1")
    n_rhs = {"FE": 1, "RK4": 4, "SSPRK3": 3}[meta["integrator"]]
    for i in range(meta["n_steps"]):
        key = "step%d:" % i
        try:
            dt = s.calc_dt(meta["cfl"])
        except mb.MallardError:
            assert not np.isfinite(g.get(key + "dt", [np.nan])[0]) or g[key + "dt"][0] < 0
            break
        if key + "dt" in g:
            check(np.array([dt]), g[key + "dt"], "dt")
            U, P, cl = s.get_state(prim=True, cfl_local=True)
            check(cl, g[key + "cfl_local"], "cfl_local")
        s.take_step()
        if key + "U" in g:
            if i == 0:   # drift is reported for later steps, per-step parity is asserted on the first
                for r in range(n_rhs):
                    check(s.get("rhs%d" % r), g[key + "rhs%d" % r], "rhs%d" % r)
                U, P = s.get_state(prim=True)
                check(U, g[key + "U"], "U")
                check(P, g[key + "P"], "P")
                if n_rhs > 1:
                    check(s.get("U_temp"), g[key + "U_temp"], "U_temp")
            else:
                U = s.get_state()
                drift = gu.field_err(U, g[key + "U"])
                print("%s[%s] drift after %d steps: %.3e" % (name, fp, i + 1, drift))
                assert drift <= (0.0 if exact else 1e-9), drift
    print("%s[%s] worst per-step relative error %.3e" % (name, fp, worst))


def _random_smooth_state(xy, rng):
    k = rng.uniform(0.5, 2.0, 4)
    rho = 1.0 + 0.3 * np.sin(2 * np.pi * k[0] * xy[:, 0]) * np.cos(2 * np.pi * k[1] * xy[:, 1])
    u = 0.4 + 0.2 * np.cos(2 * np.pi * k[2] * xy[:, 1]); v = -0.3 + 0.2 * np.sin(2 * np.pi * k[3] * xy[:, 0])
    p = 1.0 + 0.2 * np.cos(2 * np.pi * (xy[:, 0] - xy[:, 1]))
    e = p / ((GAMMA - 1) * rho)
    return np.stack([rho, rho * u, rho * v, rho * (e + 0.5 * (u * u + v * v))], 1)


@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("mtype,nx,ny,recon,riemann,integ,fixed", [
    ("cartesian", 96, 64, "FO", "HLLC", "SSPRK3", False), ("cartesian", 50, 70, "FO", "HLL", "RK4", False),
    ("cartesian_tri", 40, 30, "FO", "Rusanov", "FE", False), ("wedge", 60, 20, "FO", "HLLC", "SSPRK3", False),
    ("cartesian_tri", 24, 20, "TENO", "HLLC", "SSPRK3", False), ("cartesian_tri", 16, 18, "TENO", "Rusanov", "RK4", False),
    ("cartesian_tri", 24, 20, "TENO", "HLLC", "SSPRK3", True), ("cartesian_tri", 18, 16, "TENO", "HLL", "FE", True)])
def test_against_oracle_seeded(oracle_mod, mtype, nx, ny, recon, riemann, integ, fixed, fp):
    """Reference-faithful TENO (fixed=False) turns non-finite during the first step exactly like the reference does
    (SURVEY §0.2): there the NaN/Inf pattern must coincide and the finite entries agree.  fixed=True is the normalised
    weight variant (N2, no reference oracle): compared against the oracle's implementation of the same definition."""
    om = oracle_mod.Mesh.generate(mtype, nx, ny, 2.0, 1.0)
    mesh = mb.Mesh.generate(mtype, nx, ny, 2.0, 1.0)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="p_out", p=0.9), dict(name="top", type="symmetry"),
           dict(name="bottom", type="wall_adiabatic")]
    kw = dict(recon=recon, riemann=riemann, integrator=integ, bcs=bcs, order=3, teno_fixed=fixed)
    so = oracle_mod.Solver(om, **kw)
    sg = mb.Solver(mesh, fp_mode=fp, **kw)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(11))
    so.set_state(U0); sg.set_state(U0)
    err = gu.rel_err if fp == "strict" else gu.field_err
    assert err(sg.calc_rhs(), so.calc_rhs()) <= TOL
    n_steps = 1 if (recon == "TENO" and not fixed) else 3
    for step in range(n_steps):
        dto, dtg = so.calc_dt(0.4), sg.calc_dt(0.4)
        assert abs(dtg - dto) <= TOL * dto
        so.take_step(dto); sg.take_step()
        Ug, Pg = sg.get_state(prim=True)
        assert err(Ug, so.get("U")) <= TOL * (step + 1), step
        assert err(Pg, so.get("P")) <= 10 * TOL * (step + 1), step
        if fp == "strict" and recon == "FO" and riemann == "Rusanov":
            assert np.array_equal(Ug, so.get("U"))


def test_renumbering_does_not_change_a_single_bit():
    """STRICT mode sums every cell's faces in the reference order whatever the storage order: RCM vs identity must agree
    bit for bit over several steps (TENO and FO)."""
    for recon, mtype, nx, ny in (("FO", "wedge", 90, 30), ("TENO", "cartesian_tri", 30, 22)):
        mesh = mb.Mesh.generate(mtype, nx, ny, 4.0, 1.5)
        U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(3))
        res = []
        for ren in ("rcm", "none"):
            s = mb.Solver(mesh, recon, "HLLC", "SSPRK3", bcs=SYM4, renumber=ren, fp_mode="strict", teno_fixed=True)
            s.set_state(U0)
            s.run(4, cfl=0.3)
            res.append(s.get_state())
        assert np.isfinite(res[0]).all()
        assert np.array_equal(res[0], res[1]), recon


def test_host_buffer_seams_match_resident_path():
    mesh = mb.Mesh.generate("cartesian_tri", 20, 20, 1.0, 1.0)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(5))
    a = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", bcs=SYM4, teno_fixed=True)
    b = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", bcs=SYM4, teno_fixed=True)
    a.set_state(U0)
    rhs_res = a.calc_rhs()
    assert np.array_equal(b.calc_rhs(U0), rhs_res)          # rhs_func seam with host buffers
    dt = a.calc_dt(0.2); a.take_step()
    U1, dt_b = b.take_step_host(U0.copy(), cfl=0.2)         # take_step seam with host buffers
    assert dt_b == dt and np.array_equal(U1, a.get_state())


@pytest.mark.parametrize("recon,n", [("FO", 700), ("TENO", 256), ("TENO", 1024)])   # 1024: BASELINE configs[1]'s full size
def test_large_mesh_properties(recon, n):
    """Size-independent properties at sizes the oracle cannot reach quickly: free-stream preservation and discrete
    conservation (sum of V*rhs vanishes in the interior; with symmetry walls the mass residual sums to zero)."""
    mesh = mb.Mesh.generate("cartesian_tri", n, n, 1.0, 1.0)
    s = mb.Solver(mesh, recon, "HLLC", "SSPRK3", bcs=SYM4, fp_mode="fast", keep_stage_rhs=False, teno_fixed=True)
    nc = mesh.n_cells
    e = 1.0 / (0.4 * 1.2)
    Uc = np.tile([1.2, 1.2 * 0.3, 1.2 * -0.2, 1.2 * (e + 0.5 * 0.13)], (nc, 1))
    s.set_state(Uc)
    rhs = s.calc_rhs()
    interior = np.ones(nc, bool)
    cof = mesh.arrays["cells_of_face"]
    bfaces = np.nonzero(gu.real_faces(cof, mesh.arrays["nodes_of_face"]) & (cof[:, 1] < 0))[0]
    interior[cof[bfaces, 0]] = False
    assert np.abs(rhs[interior]).max() < 1e-9                # uniform flow: zero residual away from the walls
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(1))
    s.set_state(U0)
    rhs = s.calc_rhs()
    V = mesh.arrays["cell_volume"]
    assert abs(np.sum(V * rhs[:, 0])) < 1e-10 * np.sum(V * np.abs(rhs[:, 0]))
    assert abs(np.sum(V * rhs[:, 3])) < 1e-10 * np.sum(V * np.abs(rhs[:, 3]))
    t, dt = s.run(3, cfl=0.2)
    U = s.get_state()
    assert np.isfinite(U).all() and t > 0
    mass0, mass1 = np.sum(V * U0[:, 0]), np.sum(V * U[:, 0])
    assert abs(mass1 - mass0) < 1e-12 * mass0


@pytest.mark.parametrize("fixed", [True, False], ids=["normalised-weights", "reference-faithful(benchmarked)"])
def test_full_size_riemann2d_fast_mode_matches_the_bit_faithful_mode(fixed, capsys):
    """BASELINE configs[1] at its full size (cartesian_tri 1024^2, four-quadrant IC, TENO p=3 + HLLC + SSPRK3): the oracle
    needs ~8 minutes of serial preprocessing there, so parity is carried by transitivity - STRICT mode is pinned against the
    reference on the fixtures and on the 128^2 case above, and here the FAST path (compact device-built tables, warp-private
    streaming kernel, FMA) must agree with STRICT on stage-1 face values, the residual and one full step.
    fixed=False is the configuration bench.py times (reference-faithful weights): the reference itself produces
    Inf / NaN face values next to the initial jumps there, so the non-finite PATTERN is compared as well as the finite entries."""
    import bench
    mesh = mb.Mesh.generate("cartesian_tri", 1024, 1024, 1.0, 1.0)
    U0, P0 = bench.riemann2d_state(mesh.arrays["cell_coords"])
    out = {}
    for fp in ("strict", "fast"):
        s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=SYM4, fp_mode=fp, teno_fixed=fixed, keep_stage_rhs=False)
        s.set_state(U0, P0)
        real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
        F = s.calc_face_values()[real][:, :, 0]
        rhs = s.calc_rhs()
        dt = s.calc_dt(0.1)
        s.take_step()
        out[fp] = (F, rhs, dt, s.get_state())
        s.close()
    Ff, Fs = out["fast"][0], out["strict"][0]
    rf, rs = out["fast"][1], out["strict"][1]
    Uf, Us = out["fast"][3], out["strict"][3]
    lines = []
    if fixed:
        assert gu.field_err(Ff, Fs) <= TOL
        assert gu.field_err(rf, rs) <= 1e-10     # residual: flux differences divided by cell volumes ~ 5e-7
        lines.append("F  field-scale %.2e element-wise %.2e" % (gu.field_err(Ff, Fs), gu.elem_err(Ff, Fs)))
        lines.append("rhs field-scale %.2e" % gu.field_err(rf, rs))
    else:
        # face values: reference-faithful weights give entries up to ~1e24 and Inf / NaN where 1/(SI+eps)^6 overflows; FAST and
        # STRICT must put them in the same places (a handful of entries may sit within rounding of the overflow threshold)
        mm = gu.pattern_mismatch(Ff, Fs)
        fin = np.isfinite(Ff).all(axis=(1, 2)) & np.isfinite(Fs).all(axis=(1, 2))
        lines.append("F  non-finite faces strict %d fast %d, class mismatch fraction %.2e" % ((~np.isfinite(Fs).all(axis=(1, 2))).sum(),
                                                                                            (~np.isfinite(Ff).all(axis=(1, 2))).sum(), mm))
        assert mm <= 1e-5
        # finite entries, each against its own magnitude: the huge (1e10 .. 1e24) discontinuity-branch values as well as the O(1) ones
        ee = gu.elem_err(Ff[fin], Fs[fin], floor=1e-30)
        lines.append("F  finite entries element-wise %.2e (max |F| %.2e)" % (ee, np.abs(Fs[fin]).max()))
        assert ee <= 1e-9
        smooth = np.abs(Fs[fin]).max(axis=(1, 2)) < 1e3       # faces whose cell stayed on the central branch
        assert smooth.mean() > 0.9 and gu.field_err(Ff[fin][smooth], Fs[fin][smooth]) <= TOL
        mr = gu.pattern_mismatch(rf, rs)
        finr = np.isfinite(rf).all(axis=1) & np.isfinite(rs).all(axis=1)
        lines.append("rhs non-finite cells strict %d fast %d, class mismatch fraction %.2e" % ((~np.isfinite(rs).all(axis=1)).sum(),
                                                                                             (~np.isfinite(rf).all(axis=1)).sum(), mr))
        assert mr <= 1e-4
        quiet = finr & (np.abs(rs).max(axis=1) < 1e6)
        assert quiet.mean() > 0.9 and gu.field_err(rf[quiet], rs[quiet]) <= 1e-10
    assert abs(out["fast"][2] - out["strict"][2]) <= TOL * out["strict"][2]
    # One step later: the four-quadrant jumps drive a band of cells non-finite within the step in BOTH modes (neither weight
    # variant is positivity preserving across a 1:10 pressure jump at cfl 0.1 on this mesh); where both are finite the
    # states agree to the tolerance, and the non-finite sets differ by at most a handful of cells on their rim.
    both = np.isfinite(Uf).all(axis=1) & np.isfinite(Us).all(axis=1)
    assert both.mean() > 0.9
    diff = (np.isfinite(Uf).all(axis=1) != np.isfinite(Us).all(axis=1)).mean()
    lines.append("U after one step: finite in both %.4f, finite-set difference %.2e" % (both.mean(), diff))
    assert diff <= 1e-4
    if fixed:
        assert gu.field_err(Uf[both], Us[both]) <= TOL
    else:
        calm = both & (np.abs(Us).max(axis=1) < 1e3)
        assert calm.mean() > 0.9 and gu.field_err(Uf[calm], Us[calm]) <= TOL
    with capsys.disabled():
        print("\nfull size 1024^2 FAST vs STRICT [%s]:\n  " % ("teno_fixed" if fixed else "reference-faithful") + "\n  ".join(lines))


# ------------------------------------------------------------------------------------------------------------------
# The persistent, MULTI-TILE path of the streaming kernels against the reference / the oracle directly.  Both streaming
# kernels run 148 SMs x 2 CTAs x 4 warps = 1184 warps of 8-cell tiles: a warp only enters its second tile (ring wrap across
# tiles, the fxbuf parity flip, prefetch_tile(next)) above 9472 cells.  The fixtures under tests/golden are <= 140 cells, so
# these cases run the unmodified reference (oracle/_ref/bin/ref_harness travels to the GPU box prebuilt) / the oracle on
# >= 20 k cells: three and more tiles per warp.
# ------------------------------------------------------------------------------------------------------------------
def _check_against(label, fp, got, ref, tol, report):
    """STRICT: element-wise; FAST: relative to the field scale (asserted) AND element-wise (printed, bounded)."""
    if fp == "strict":
        e = gu.rel_err(got, ref)
        report.append("%-10s strict element-wise %.2e" % (label, e))
        assert e <= tol, (label, e)
    else:
        e, ee = gu.field_err(got, ref), gu.elem_err(got, ref)
        report.append("%-10s fast field-scale %.2e element-wise %.2e" % (label, e, ee))
        assert e <= tol, (label, e)
        assert ee <= 1e4 * tol, (label, ee)      # entries down to 1e-9 of the field scale, each against its own magnitude


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref/bin/ref_harness not built (needs /root/reference)")
def test_multi_tile_streaming_path_vs_unmodified_reference_128x128(tmp_path, capsys):
    """cartesian_tri 128^2 = 32 768 cells (>= 3 tiles per warp), examples/riemann_2d numerics (TENO legendre p=3 + HLLC +
    SSPRK3, cfl 0.1), smooth initial condition: stage-1 face values, the bare residual, dt, every stage residual, U_temp
    and the state after the step, STRICT and FAST, against dumps of the UNMODIFIED reference made on the spot."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden as mg
    import mlbd
    case = dict(mesh=dict(type="cartesian_tri", Nx=128, Ny=128, Lx=1.0, Ly=1.0), ic=mg.SMOOTH_IC, bcs=mg.SYM4, cfl=0.1,
                riemann="HLLC", integrator="SSPRK3", recon=mg.TENO3, n_steps=1)
    toml, out = str(tmp_path / "input.toml"), str(tmp_path / "out.mlbd")
    mg.write_toml(case, toml)
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
    subprocess.check_call([REF_HARNESS, "dump", toml, out, "1", "1"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    g = mlbd.read(out)
    os.remove(out)
    mesh = mb.Mesh.generate("cartesian_tri", 128, 128, 1.0, 1.0)
    assert mesh.n_cells == 32768 and np.array_equal(mesh.arrays["cells_of_face"], g["cells_of_face"])
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    interior = real & (mesh.arrays["cells_of_face"][:, 1] >= 0)
    report = []
    for fp in ("strict", "fast"):
        s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=SYM4, fp_mode=fp)
        s.set_state(g["U0"], g["P0"])
        F = s.calc_face_values()
        _check_against("F side 0", fp, F[real][:, :, 0], g["F_stage1"][real][:, :, 0], TOL, report)
        _check_against("F side 1", fp, F[interior][:, :, 1], g["F_stage1"][interior][:, :, 1], TOL, report)
        _check_against("rhs", fp, s.calc_rhs(), g["rhs_stage1"], TOL if fp == "strict" else 1e-10, report)
        dt = s.calc_dt(0.1)
        assert abs(dt - g["step0:dt"][0]) <= TOL * dt
        s.take_step()
        for r in range(3):
            _check_against("rhs%d" % r, fp, s.get("rhs%d" % r), g["step0:rhs%d" % r], TOL if fp == "strict" else 1e-10, report)
        _check_against("U_temp", fp, s.get("U_temp"), g["step0:U_temp"], TOL, report)
        U, P = s.get_state(prim=True)
        _check_against("U", fp, U, g["step0:U"], TOL, report)
        _check_against("P", fp, P, g["step0:P"], TOL, report)
        s.close()
    with capsys.disabled():
        print("\n128x128 (32768 cells) against the unmodified reference:\n  " + "\n  ".join(report))


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref/bin/ref_harness not built (needs /root/reference)")
@UNVERIFIED_ON_HARDWARE
def test_multi_tile_streaming_path_vs_unmodified_reference_on_an_injected_jittered_mesh(tmp_path, capsys):
    """The unstructured counterpart of the 128^2 test: a jittered, id-shuffled 110 x 100 triangulation (22 000 cells: several tiles per
    warp of the streaming kernels, BASELINE configs[3] family) is INJECTED into the unmodified reference (`ref_harness mesh`; the reference
    has no generator for it and computes its own geometry, stencils and matrices on the injected connectivity); stage-1 face values, the
    residual, dt and the state after the step, STRICT and FAST, against its dump.  (oracle/pin_sweep.py --unstructured is the same
    comparison for the oracle, the preprocessor and the emulated kernels on smaller meshes.)"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden as mg
    import mlbd
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(110, 100, 1.0, 0.9, seed=12345)
    a = mesh.arrays
    case = dict(mesh=dict(type="cartesian_tri", Nx=110, Ny=100, Lx=1.0, Ly=0.9), ic=mg.SMOOTH_IC, bcs=mg.EXTRAP4, cfl=0.1,
                riemann="HLLC", integrator="SSPRK3", recon=mg.TENO3, n_steps=1)
    toml, inj, out = str(tmp_path / "input.toml"), str(tmp_path / "mesh.mlbd"), str(tmp_path / "out.mlbd")
    mg.write_toml(case, toml)
    rec = {k: (a[k].reshape(-1, 2) if k in ("node_coords", "cells_of_face") else a[k]) for k in
           ("node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell", "offsets_nodes_of_face", "nodes_of_face", "cells_of_face")}
    for i, (zn, zf) in enumerate(mesh.zones):
        rec["zone:%d:%s" % (i, zn)] = np.ascontiguousarray(zf, dtype=np.uint32)
    mlbd.write(inj, rec)
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
    subprocess.check_call([REF_HARNESS, "mesh", inj, toml, out, "1", "1"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    g = mlbd.read(out)
    os.remove(out)
    assert mesh.n_cells == 22000 and np.array_equal(a["cell_volume"].reshape(-1), g["cell_volume"].reshape(-1))      # the reference's own geometry
    interior = a["cells_of_face"].reshape(-1, 2)[:, 1] >= 0
    report = []
    for fp in ("strict", "fast"):
        s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode=fp)
        s.set_state(g["U0"], g["P0"])
        F = s.calc_face_values()
        _check_against("F side 0", fp, F[:, :, 0], g["F_stage1"][:, :, 0], TOL, report)
        _check_against("F side 1", fp, F[interior][:, :, 1], g["F_stage1"][interior][:, :, 1], TOL, report)
        _check_against("rhs", fp, s.calc_rhs(), g["rhs_stage1"], TOL if fp == "strict" else 1e-10, report)
        dt = s.calc_dt(0.1)
        assert abs(dt - g["step0:dt"][0]) <= TOL * dt
        s.take_step()
        U, P = s.get_state(prim=True)
        _check_against("U", fp, U, g["step0:U"], TOL, report)
        _check_against("P", fp, P, g["step0:P"], TOL, report)
        s.close()
    with capsys.disabled():
        print("\njittered 110x100 (22000 cells) injected into the unmodified reference:\n  " + "\n  ".join(report))


@pytest.mark.parametrize("fp", ["strict", "fast"])
def test_multi_tile_streaming_path_vs_oracle_jittered_22k(oracle_mod, fp, capsys):
    """The same on an unstructured numbering: jittered, id-shuffled 110 x 100 triangulation (22 000 cells), vortex, against the
    oracle (itself pinned bit-exact to the reference)."""
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(110, 100, 10.0, 10.0, seed=12345)
    om = _oracle_mesh_of(oracle_mod, mesh)
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", bcs=syn.EXTRAP4, order=3)
    so = oracle_mod.Solver(om, **kw)
    sg = mb.Solver(mesh, fp_mode=fp, **kw)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    so.set_state(U0); sg.set_state(U0)
    report = []
    Fg, Fo = sg.calc_face_values(), so.calc_face_values()
    cof = mesh.arrays["cells_of_face"]
    _check_against("F side 0", fp, Fg[:, :, 0], Fo[:, :, 0], TOL, report)
    _check_against("F side 1", fp, Fg[cof[:, 1] >= 0][:, :, 1], Fo[cof[:, 1] >= 0][:, :, 1], TOL, report)
    # the reference-faithful weights leave face states far above the solution scale in the discontinuity branch;
    # the residual's flux cancellation amplifies last-ulp differences element-wise, so it is measured against the field scale
    e = gu.field_err(sg.calc_rhs(), so.calc_rhs())
    report.append("rhs        field-scale %.2e" % e)
    assert e <= (TOL if fp == "strict" else 1e-10)
    with capsys.disabled():
        print("\njittered 110x100 (22000 cells) against the oracle [%s]:\n  " % fp + "\n  ".join(report))


# ------------------------------------------------------------------------------------------------------------------
# Unstructured meshes (BASELINE configs[3] family): jittered, id-shuffled triangulations handed over as plain arrays
# ------------------------------------------------------------------------------------------------------------------
CONN_KEYS = ["node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell",
             "offsets_nodes_of_face", "nodes_of_face", "cells_of_face"]


def _oracle_mesh_of(oracle_mod, mesh):
    return oracle_mod.Mesh.from_arrays({k: mesh.arrays[k] for k in CONN_KEYS}, mesh.zones)


@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("recon,riemann,integ,fixed", [("FO", "HLLC", "SSPRK3", False), ("TENO", "HLLC", "SSPRK3", True),
                                                       ("TENO", "HLLC", "SSPRK3", False), ("TENO", "HLL", "RK4", True)])
def test_jittered_unstructured_mesh_vs_oracle(oracle_mod, recon, riemann, integ, fixed, fp):
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(22, 18, 10.0, 10.0, seed=12345)
    om = _oracle_mesh_of(oracle_mod, mesh)
    kw = dict(recon=recon, riemann=riemann, integrator=integ, bcs=syn.EXTRAP4, order=3, teno_fixed=fixed)
    so = oracle_mod.Solver(om, **kw)
    sg = mb.Solver(mesh, fp_mode=fp, **kw)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    so.set_state(U0); sg.set_state(U0)
    err = gu.rel_err if fp == "strict" else gu.field_err
    if recon == "TENO":
        Fg, Fo = sg.calc_face_values(), so.calc_face_values()
        cof = mesh.arrays["cells_of_face"]
        assert err(Fg[:, :, 0], Fo[:, :, 0]) <= TOL
        assert err(Fg[cof[:, 1] >= 0][:, :, 1], Fo[cof[:, 1] >= 0][:, :, 1]) <= TOL
    # Reference-faithful TENO weights leave face states many orders of magnitude above the solution in the
    # discontinuity branch; there HLLC's star-state estimate runs through TRRS' pow (libm on the host, CUDA's pow on the
    # device: last-ulp differences), which the residual's flux cancellation amplifies ELEMENT-wise.  Bit-faithfulness is
    # asserted on the face values above; the residual of that variant is measured against the field scale.
    if recon == "TENO" and not fixed and fp == "strict":
        assert gu.field_err(sg.calc_rhs(), so.calc_rhs()) <= TOL
    else:
        assert err(sg.calc_rhs(), so.calc_rhs()) <= (TOL if fp == "strict" else 1e-10)
    n_steps = 1 if (recon == "TENO" and not fixed) else 3
    for step in range(n_steps):
        dto, dtg = so.calc_dt(0.4), sg.calc_dt(0.4)
        assert abs(dtg - dto) <= TOL * dto
        so.take_step(dto); sg.take_step()
        Ug = sg.get_state()
        assert err(Ug, so.get("U")) <= TOL * (step + 1), step


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["cartesian_tri", "jittered"])
def test_device_built_teno_tables_are_bit_identical_to_host_tables(order, kind):
    """the compact reconstruction matrices are computed on the GPU (teno_tables.cu); the host preprocessor
    (whose reference-layout output is pinned bit-exact against the real reference in test_oracle_vs_reference.py) is the
    checker.  Every double must be identical."""
    from mallard_b200 import synthetic as syn
    mesh = mb.Mesh.generate("cartesian_tri", 26, 22, 2.0, 1.0) if kind == "cartesian_tri" else syn.jittered_tri(26, 22, 2.0, 1.0, seed=7)
    bcs = SYM4
    s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=order, bcs=bcs, fp_mode="fast")
    assert s.get("stats")[10] > 0.0, "tables were not built on the device"
    plan = mb.Plan(mesh, "TENO", order=order, bcs=bcs, fp_mode="fast")
    assert np.array_equal(s.get("dev:fm_ids"), plan.get("fm_ids"))
    host_mat, dev_mat = plan.get("fm_mat"), s.get("dev:fm_mat")
    assert host_mat.shape == dev_mat.shape and np.abs(host_mat).max() > 0
    assert np.array_equal(host_mat.view(np.uint64), dev_mat.view(np.uint64))
    assert np.array_equal(plan.get("fm_area0").view(np.uint64), s.get("dev:fm_area0").view(np.uint64))


def test_vortex_on_unstructured_mesh_properties():
    """Config-4 style run at a size the oracle would need minutes for: fixed-weight TENO on a jittered 160x160 mesh keeps
    the vortex finite, conserves mass to round-off (extrapolation boundaries far from the vortex see uniform flow) and is
    independent of the storage order (RCM vs none, STRICT mode: bit-identical)."""
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(160, 160, 10.0, 10.0, seed=3)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    res = []
    for ren in ("rcm", "none"):
        s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode="strict", teno_fixed=True, renumber=ren,
                      keep_stage_rhs=False)
        s.set_state(U0)
        s.run(3, cfl=0.3)
        res.append(s.get_state())
    assert np.isfinite(res[0]).all()
    assert np.array_equal(res[0], res[1])
    sf = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode="fast", teno_fixed=True, keep_stage_rhs=False)
    sf.set_state(U0)
    sf.run(3, cfl=0.3)
    assert gu.field_err(sf.get_state(), res[0]) <= 3 * TOL


@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("order,fixed", [(1, True), (2, True), (3, False), (3, True), (4, True)])
def test_monomial_basis_vs_oracle(oracle_mod, order, fixed, fp):
    """basis_type = "monomial" is the reference's default (face_reconstruction.cpp:110); basis values go through pow
    (basis.h:66-70), so STRICT is within rounding of libm rather than bit-exact."""
    om = oracle_mod.Mesh.generate("cartesian_tri", 20, 16, 2.0, 1.0)
    mesh = mb.Mesh.generate("cartesian_tri", 20, 16, 2.0, 1.0)
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", bcs=SYM4, basis="monomial", order=order, teno_fixed=fixed)
    so = oracle_mod.Solver(om, **kw)
    sg = mb.Solver(mesh, fp_mode=fp, **kw)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(23))
    so.set_state(U0); sg.set_state(U0)
    err = gu.rel_err if fp == "strict" else gu.field_err
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    assert err(sg.calc_face_values()[real][:, :, 0], so.calc_face_values()[real][:, :, 0]) <= TOL
    dto, dtg = so.calc_dt(0.3), sg.calc_dt(0.3)
    so.take_step(dto); sg.take_step()
    assert err(sg.get_state(), so.get("U")) <= TOL


@UNVERIFIED_ON_HARDWARE
def test_generic_teno_kernel_is_bit_identical_to_the_specialised_one(monkeypatch):
    """teno_generic.cuh keeps the specialised kernels' operation and summation order with run-time (K, M, S): forced onto a
    configuration that has a specialised kernel (MLB_TENO_GENERIC=1, read per launch) it must reproduce it bit for bit."""
    mesh = mb.Mesh.generate("cartesian_tri", 22, 18, 2.0, 1.0)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(31))
    for order, basis in ((3, "legendre"), (2, "monomial"), (4, "legendre")):
        s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", basis=basis, order=order, bcs=SYM4, fp_mode="strict")
        s.set_state(U0)
        monkeypatch.delenv("MLB_TENO_GENERIC", raising=False)
        F_spec, rhs_spec = s.calc_face_values(), s.calc_rhs()
        monkeypatch.setenv("MLB_TENO_GENERIC", "1")
        F_gen, rhs_gen = s.calc_face_values(), s.calc_rhs()
        monkeypatch.delenv("MLB_TENO_GENERIC", raising=False)
        assert np.abs(F_spec).max() > 0
        assert np.array_equal(F_spec, F_gen, equal_nan=True) and np.array_equal(rhs_spec, rhs_gen, equal_nan=True), (order, basis)
        s.close()


_ORACLE_CACHE = {}


@UNVERIFIED_ON_HARDWARE
@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("order,factor,qc,basis,fixed", [(5, 2.0, 5, "legendre", True), (6, 2.0, 5, "legendre", True), (7, 2.0, 5, "monomial", True),
                                                         (9, 2.0, 5, "legendre", True), (2, 1.5, 0, "legendre", False), (3, 3.0, 0, "monomial", True),
                                                         (4, 2.5, 0, "legendre", True)])
def test_teno_orders_5_to_9_and_other_stencil_factors_vs_oracle(oracle_mod, order, factor, qc, basis, fixed, fp):
    """Everything the reference's TOML accepts (face_reconstruction.cpp:110-116,170-180): basis_order up to 9 (Dunavant rules
    stop at order 5, so quadrature_order_cell is given there, as it must be for the reference) and max_stencil_size_factor
    other than 2, through the generic kernel, against the oracle (pinned to the reference for p = 5 and factor 1.5 by the
    fixtures teno_legendre_12x10_p5 / teno_legendre_8x7_p2_f15)."""
    nx, ny = (12, 10) if order >= 7 else (20, 16)          # 240 cells hold the 110-cell stencils of p = 9; the oracle's set-up dominates the run time
    mesh = mb.Mesh.generate("cartesian_tri", nx, ny, 2.0, 1.0)
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", bcs=SYM4, basis=basis, order=order, factor=factor, quad_cell_order=qc,
              teno_fixed=fixed)
    key = (order, factor, qc, basis, fixed)
    if key not in _ORACLE_CACHE:                            # one oracle set-up (33 s at p = 9) serves both floating-point modes
        _ORACLE_CACHE[key] = oracle_mod.Solver(oracle_mod.Mesh.generate("cartesian_tri", nx, ny, 2.0, 1.0), **kw)
    so = _ORACLE_CACHE[key]
    sg = mb.Solver(mesh, fp_mode=fp, **kw)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(23))
    so.set_state(U0); sg.set_state(U0)
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    Fg, Fo = sg.calc_face_values()[real][:, :, 0], so.calc_face_values()[real][:, :, 0]
    # the higher the order, the worse conditioned the pseudo-inverse rows (entries ~1e6 at p = 9): differences are measured
    # against the field scale in both modes; STRICT is additionally bit-exact wherever pow() is not involved (legendre)
    # (reference-faithful weights, fixed = False: 1 / (SI + eps)^6 is three multiplications on the device and libm pow in the oracle / reference -
    #  the last ulp of a weight that only the normalised variant divides away; measured on the host emulation: 1e-17 of the field scale)
    if fp == "strict" and basis == "legendre" and fixed:
        assert np.array_equal(Fg, Fo, equal_nan=True)
    else:
        assert gu.field_err(Fg, Fo) <= (TOL if order <= 5 else 1e-9)
    dto, dtg = so.calc_dt(0.3), sg.calc_dt(0.3)
    so.take_step(dto); sg.take_step()
    assert gu.field_err(sg.get_state(), so.get("U")) <= (TOL if order <= 5 else 1e-9)


# ------------------------------------------------------------------------------------------------------------------
# Quadrilaterals and mixed meshes under TENO (BASELINE configs[3] as worded).  The reference throws on anything but
# triangles (face_reconstruction.cpp:485-487): no oracle exists, the checks are analytic.
# ------------------------------------------------------------------------------------------------------------------
_DUNAVANT5 = (np.array([[1 / 3, 1 / 3], [0.059715871789770, 0.470142064105115], [0.470142064105115, 0.059715871789770], [0.470142064105115, 0.470142064105115],
                        [0.797426985353087, 0.101286507323456], [0.101286507323456, 0.797426985353087], [0.101286507323456, 0.101286507323456]]),
              np.array([0.225, 0.132394152788506, 0.132394152788506, 0.132394152788506, 0.125939180544827, 0.125939180544827, 0.125939180544827]))


def _cell_averages(mesh, f):
    """Exact (degree <= 5) cell averages of f(x, y) -> [n][4] over triangles and quadrilaterals (fan of triangles from node 0)."""
    A = mesh.arrays
    X, onc, noc = A["node_coords"], A["offsets_nodes_of_cell"].astype(np.int64), A["nodes_of_cell"].astype(np.int64)
    out = np.zeros((mesh.n_cells, 4))
    area = np.zeros(mesh.n_cells)
    xy, w = _DUNAVANT5
    for c in range(mesh.n_cells):
        n = noc[onc[c]:onc[c + 1]]
        for t in range(len(n) - 2):
            v0, v1, v2 = X[n[0]], X[n[t + 1]], X[n[t + 2]]
            a = 0.5 * abs((v1[0] - v0[0]) * (v2[1] - v0[1]) - (v2[0] - v0[0]) * (v1[1] - v0[1]))
            p = v0[None, :] + xy[:, :1] * (v1 - v0)[None, :] + xy[:, 1:] * (v2 - v0)[None, :]
            out[c] += a * (w[:, None] * f(p[:, 0], p[:, 1])).sum(axis=0)
            area[c] += a
    return out / area[:, None]


@UNVERIFIED_ON_HARDWARE
@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("order,tri_fraction", [(1, 0.5), (2, 0.5), (3, 0.5), (3, 0.0), (2, 1.0), (4, 0.6)])
def test_teno_on_quadrilateral_and_mixed_meshes_is_k_exact(order, tri_fraction, fp):
    """A reconstruction of order p must reproduce every polynomial of degree <= p exactly from its cell averages, at every
    face quadrature point, from both sides of every interior face, on quadrilaterals, triangles and any mix of them, jittered
    (normalised weights / mean-free basis: the reference-faithful variant adds twice the basis mean, SURVEY Q3, and is not
    k-exact even on triangles).  This pins the quadrilateral extension of the stencil matrices (fan quadrature, per-cell basis
    means, frame from nodes 0, 1, last) without an oracle."""
    from mallard_b200 import synthetic as syn
    mesh = syn.mixed_tri_quad(14, 12, 3.0, 2.0, seed=3, tri_fraction=tri_fraction)
    nn = np.diff(mesh.arrays["offsets_nodes_of_cell"])
    assert (tri_fraction == 1.0 or (nn == 4).any()) and (tri_fraction == 0.0 or (nn == 3).any())
    rng = np.random.default_rng(17)
    coef = rng.uniform(-1.0, 1.0, size=(4, order + 1, order + 1))

    def f(x, y):
        out = np.zeros(x.shape + (4,))
        for v in range(4):
            for i in range(order + 1):
                for j in range(order + 1 - i):
                    out[..., v] += coef[v, i, j] * x ** i * y ** j
        out[..., 0] += 5.0
        return out
    U0 = _cell_averages(mesh, f)
    bcs = [dict(name=n, type="extrapolation") for n in ("left", "right", "top", "bottom")]
    s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=order, bcs=bcs, fp_mode=fp, teno_fixed=True)
    s.set_state(U0)
    F = s.calc_face_values()                                        # [nf][Q][2][4]
    A = mesh.arrays
    nof, cof = A["nodes_of_face"].reshape(-1, 2).astype(np.int64), A["cells_of_face"]
    xi = {1: [0.0], 2: [-0.5773502691896257, 0.5773502691896257], 3: [-0.7745966692414834, 0.0, 0.7745966692414834]}[s.n_quad]
    x0, x1 = A["node_coords"][nof[:, 0]], A["node_coords"][nof[:, 1]]
    worst = 0.0
    for q, z in enumerate(xi):
        pq = (z + 1.0) * 0.5 * (x1 - x0) + x0
        exact = f(pq[:, 0], pq[:, 1])
        worst = max(worst, np.abs(F[:, q, 0] - exact).max(), np.abs(F[cof[:, 1] >= 0][:, q, 1] - exact[cof[:, 1] >= 0]).max())
    assert worst <= 2e-9 * (10.0 ** max(0, order - 3)), worst       # conditioning of the pseudo-inverse grows with the order
    # and the solver runs on it: free-stream preservation, conservation, a few finite steps
    Uc = np.tile([1.2, 0.36, -0.24, 2.6], (mesh.n_cells, 1))
    s.set_state(Uc)
    rhs = s.calc_rhs()
    interior = np.ones(mesh.n_cells, bool)
    interior[cof[cof[:, 1] < 0, 0]] = False
    assert np.abs(rhs[interior]).max() < 1e-9
    xy = A["cell_coords"]
    U1 = syn.isentropic_vortex(xy * [20.0 / 3.0, 10.0], centre=(10.0, 10.0))
    s.set_state(U1)
    rhs = s.calc_rhs()
    V = A["cell_volume"]
    # interior faces cancel pairwise: the volume-weighted residual sums to the boundary flux only, and the flow at the boundary is uniform
    # (the vortex's perturbation has decayed to exp(-49) there; with the boundary five radii away it is still 6e-6 and the sum 2e-5 of its
    # absolute sum - measured on the host emulation): mass residual ~ 0 relative to its absolute sum, down to what the one-sided stencils of
    # the boundary cells - which reach three to four cells into this small mesh, where the vortex is not yet negligible - leave of the
    # uniform far field (7e-9 ... 2.5e-6 measured for p = 1 ... 4).  Conservation itself is by construction of the gather (every interior
    # face product enters one cell with + and the other with -), and bit-exact against the oracle on this mesh type in the first-order test below
    assert abs(np.sum(V * rhs[:, 0])) < 1e-4 * np.sum(V * np.abs(rhs[:, 0]))
    s.run(5, cfl=0.2)
    assert np.isfinite(s.get_state()).all()
    s.close()


@UNVERIFIED_ON_HARDWARE
def test_first_order_on_a_mixed_mesh_matches_oracle(oracle_mod):
    """The first-order path has a reference (and an oracle) on any cell type: bit-exact on the mixed mesh in STRICT mode."""
    from mallard_b200 import synthetic as syn
    mesh = syn.mixed_tri_quad(16, 12, 3.0, 2.0, seed=9, tri_fraction=0.4)
    om = _oracle_mesh_of(oracle_mod, mesh)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="p_out", p=0.9), dict(name="top", type="symmetry"), dict(name="bottom", type="wall_adiabatic")]
    so = oracle_mod.Solver(om, "FO", "HLLC", "SSPRK3", bcs=bcs)
    sg = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", bcs=bcs, fp_mode="strict")
    U0 = _random_smooth_state(mesh.arrays["cell_coords"] / 3.0, np.random.default_rng(2))
    so.set_state(U0); sg.set_state(U0)
    assert gu.rel_err(sg.calc_rhs(), so.calc_rhs()) <= 1e-13
    for _ in range(3):
        dto, dtg = so.calc_dt(0.4), sg.calc_dt(0.4)
        assert dtg == dto
        so.take_step(dto); sg.take_step()
    assert gu.rel_err(sg.get_state(), so.get("U")) <= 1e-13


# ------------------------------------------------------------------------------------------------------------------
# Viscous terms (BASELINE configs[4]; the reference is Euler only: analytic validation, no oracle)
# ------------------------------------------------------------------------------------------------------------------
def _gas(mu, Pr=0.72):
    return dict(gamma=1.4, p_ref=101325.0, T_ref=298.15, rho_ref=1.225, p_min=-1e20, p_max=1e20, mu=mu, Pr=Pr)


R_GAS = 101325.0 / (298.15 * 1.225)


def _state_from_prim(rho, u, v, T):
    cv = R_GAS / 0.4
    return np.stack([rho, rho * u, rho * v, rho * (cv * T + 0.5 * (u * u + v * v))], 1)


@UNVERIFIED_ON_HARDWARE
@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("mtype", ["cartesian", "cartesian_tri", "mixed"])
def test_viscous_residual_of_couette_flow(mtype, fp):
    """Plane Couette flow between a fixed and a moving no-slip wall: the viscous part of the residual, rhs(mu) - rhs(0), is exactly
    (0, 0, 0, mu (U/H)^2) on quadrilaterals, triangles and jittered mixed meshes; on the wall-aligned quadrilateral mesh the whole
    residual is (tests/test_kernel_emulation.py runs the same check on the host emulation of the kernels)."""
    from mallard_b200 import synthetic as syn
    mu, Uw, H, L = 0.05, 3.0, 1.0, 2.0
    mesh = syn.mixed_tri_quad(24, 20, L, H, seed=5, tri_fraction=0.5) if mtype == "mixed" else mb.Mesh.generate(mtype, 24, 20, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0]),
           dict(name="top", type="wall_noslip", u=[Uw, 0.0])]
    xy = mesh.arrays["cell_coords"]
    n = mesh.n_cells
    U0 = _state_from_prim(np.full(n, 1.2), Uw * xy[:, 1] / H, np.zeros(n), np.full(n, 300.0))
    res = []
    for m in (mu, 0.0):
        s = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(m), bcs=bcs, fp_mode=fp)
        s.set_state(U0)
        res.append(s.calc_rhs())
        s.close()
    heat, dv = mu * (Uw / H) ** 2, res[0] - res[1]
    scale = np.abs(res[1]).max() + 1.0
    assert np.abs(dv[:, :3]).max() < 1e-11 * scale + 1e-9 and np.abs(dv[:, 3] - heat).max() < 1e-11 * scale + 1e-8 * heat
    if mtype == "cartesian":
        # STRICT: the inviscid fluxes through opposite faces cancel exactly; FAST re-rounds them (FMA, reciprocals): what is left is a few ulp
        # of the energy flux per cell height, (rho E + p) U / dy ~ 1e7 here
        flux = (np.abs(U0[:, 3]).max() + 1.2 * R_GAS * 300.0) * Uw / (H / 20)
        slack = 0.0 if fp == "strict" else 1e-13 * flux
        assert np.abs(res[0][:, :3]).max() < 1e-7 + slack and np.abs(res[0][:, 3] - heat).max() < 1e-8 * heat + 1e-9 + slack


@UNVERIFIED_ON_HARDWARE
@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("mtype", ["cartesian", "mixed"])
def test_steady_heat_conduction_between_isothermal_walls(mtype, fp):
    """MLB_BC_WALL_NOSLIP with T > 0 (isothermal walls): gas at rest, uniform pressure, T linear between a cold and a hot wall - a uniform
    heat flux, so the viscous part of the residual vanishes in every cell, wall cells included (tests/test_kernel_emulation.py runs the same
    check, and the exactness of the temperature gradient, on the host emulation)."""
    from mallard_b200 import synthetic as syn
    mu, H, L, T0, T1 = 0.05, 1.0, 2.0, 280.0, 340.0
    mesh = syn.mixed_tri_quad(24, 20, L, H, seed=7, tri_fraction=0.5) if mtype == "mixed" else mb.Mesh.generate(mtype, 24, 20, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0], T=T0),
           dict(name="top", type="wall_noslip", u=[0.0, 0.0], T=T1)]
    y = mesh.arrays["cell_coords"][:, 1]
    T = T0 + (T1 - T0) * y / H
    U0 = _state_from_prim(1.0e5 / (R_GAS * T), np.zeros_like(y), np.zeros_like(y), T)
    res = []
    for m in (mu, 0.0):
        s = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(m), bcs=bcs, fp_mode=fp)
        s.set_state(U0)
        res.append(s.calc_rhs())
        s.close()
    dv = res[0] - res[1]
    flux = mu * (R_GAS * 1.4 / 0.4) / 0.72 * (T1 - T0) / H
    h = mesh.arrays["cell_volume"].min() ** 0.5
    assert np.abs(dv[:, :3]).max() == 0.0 and np.abs(dv[:, 3]).max() < 1e-9 * flux / h


@UNVERIFIED_ON_HARDWARE
@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("order,mtype", [(2, "cartesian_tri"), (3, "cartesian_tri"), (3, "cartesian"), (5, "cartesian_tri")])   # p = 5: three quadrature points per face (run-time loop), generic reconstruction kernel
def test_couette_flow_under_teno_is_a_steady_state_up_to_viscous_heating(order, mtype, fp):
    """The viscous terms behind the TENO face kernel (face_flux_kernel<RS, true, Q, true>; on the device the quadrature lanes of a face are
    combined with shuffles before lane 0 subtracts the viscous flux): a k-exact reconstruction reproduces plane Couette flow, the inviscid
    fluxes of this steady Euler solution cancel, what is left is the viscous heating mu (U / H)^2 in the energy equation of every cell that
    does not touch a wall.  Regular meshes: see tests/test_kernel_emulation.py (same check on the host emulation) for why."""
    mu, Uw, H, L = 0.05, 3.0, 1.0, 2.0
    mesh = mb.Mesh.generate(mtype, 14, 12, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0]),
           dict(name="top", type="wall_noslip", u=[Uw, 0.0])]
    U0 = _cell_averages(mesh, lambda x, y: np.stack([np.full_like(x, 1.2), 1.2 * Uw * y / H, 0.0 * x, 1.0 / 0.4 + 0.6 * (Uw * y / H) ** 2], -1))
    s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=order, quad_cell_order=5 if order >= 5 else 0, gas=_gas(mu), bcs=bcs, teno_fixed=True, fp_mode=fp)
    s.set_state(U0)
    rhs = s.calc_rhs()
    s.close()
    cof = mesh.arrays["cells_of_face"]
    inner = np.ones(mesh.n_cells, bool)
    inner[cof[cof[:, 1] < 0, 0]] = False
    assert np.abs(rhs[inner, :3]).max() < 1e-9 and np.abs(rhs[inner, 3] - mu * (Uw / H) ** 2).max() < 1e-9


@UNVERIFIED_ON_HARDWARE
def test_decaying_shear_layer_follows_the_diffusion_equation():
    """Time-dependent check (a Taylor-Green vortex needs periodic boundaries, which the reference does not have): a low-Mach shear
    layer u(y, 0) = U erf(y / delta0) between slip walls far away diffuses as u = U erf(y / sqrt(delta0^2 + 4 nu t)).  First-order
    inviscid part (faces aligned with the flow: no numerical diffusion of u), viscous dt limit from the spectral radius; the
    profile error falls with the mesh (second-order viscous operator)."""
    from math import erf
    nu, Uw, d0, T0, rho0 = 0.02, 1.0, 0.08, 300.0, 1.0
    errs = []
    for ny in (40, 80):
        mesh = mb.Mesh.generate("cartesian", 4, ny, 0.1, 2.0)
        y = mesh.arrays["cell_coords"][:, 1] - 1.0
        u0 = Uw * np.vectorize(erf)(y / d0)
        bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="symmetry"),
               dict(name="top", type="symmetry")]
        s = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(nu * rho0), bcs=bcs, fp_mode="fast", keep_stage_rhs=False)
        s.set_state(_state_from_prim(np.full(mesh.n_cells, rho0), u0, np.zeros(mesh.n_cells), np.full(mesh.n_cells, T0)))
        t = 0.0
        while t < 0.2:
            t, _ = s.run(200, cfl=0.5)
        U = s.get_state()
        exact = Uw * np.vectorize(erf)(y / np.sqrt(d0 * d0 + 4.0 * nu * t))
        errs.append(np.abs(U[:, 1] / U[:, 0] - exact).max())
        s.close()
    assert errs[0] < 0.02 and errs[1] < errs[0] / 2.5, errs


@UNVERIFIED_ON_HARDWARE
def test_viscous_run_on_partitioned_ranks_reproduces_the_single_context_run():
    """Viscous contexts hold a second ghost ring (the least-squares gradients of the first ring need it): bit-identical to the
    single-context run in STRICT mode."""
    from test_gpu_multi import Loopback
    mesh = mb.Mesh.generate("cartesian_tri", 30, 20, 3.0, 2.0)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"] / 3.0, np.random.default_rng(6))
    for recon in ("FO", "TENO"):
        kw = dict(recon=recon, riemann="HLLC", integrator="SSPRK3", order=2, bcs=SYM4, fp_mode="strict", teno_fixed=True, gas=_gas(0.01))
        one = mb.Solver(mesh, **kw)
        one.set_state(U0)
        many = Loopback(mesh, mb.partition(mesh, 3), 3, **kw)
        many.set_state(U0)
        for _ in range(3):
            one.calc_dt(0.3)
            one.take_step()
            many.step(0.3)
        assert np.isfinite(one.get_state()).all() and np.array_equal(one.get_state(), many.state()), recon


@UNVERIFIED_ON_HARDWARE
@pytest.mark.parametrize("fp", ["strict", "fast"])
@pytest.mark.parametrize("case", ["sod", "wedge", "tri_hll_rk4"])
def test_cooperative_small_mesh_kernel_equals_the_multi_kernel_path(oracle_mod, monkeypatch, case, fp):
    """MLB_SMALL_STEP=1: mlb_run executes all steps of a small first-order mesh in ONE cooperative kernel (csrc/small_step.cuh: the
    bodies of the CFL / face / gather kernels between grid barriers).  Same bits as the replayed multi-kernel step in STRICT mode
    (and as the oracle), within rounding of FMA contraction in FAST mode; t, dt and the step counter agree; stats[12] proves which
    path ran.  (tests/test_kernel_emulation.py runs the same phases on the host against the oracle and the reference's dumps.)"""
    if case == "sod":
        args, bcs, riemann, integ, cfl, n = ("cartesian", 1000, 1, 1.0, 1.0e-3), SYM4, "HLLC", "SSPRK3", 1.0, 120
    elif case == "wedge":
        args, riemann, integ, cfl, n = ("wedge", 150, 50, 4.0, 1.5), "HLLC", "SSPRK3", 1.0, 80
        bcs = [dict(name="left", type="upt", u=[600.0, 0.0], p=101325.0, T=300.0), dict(name="right", type="p_out", p=101325.0),
               dict(name="top", type="symmetry"), dict(name="bottom", type="wall_adiabatic")]
    else:
        args, bcs, riemann, integ, cfl, n = ("cartesian_tri", 40, 30, 1.0, 0.8), [dict(name=z, type="extrapolation") for z in ("left", "right", "top", "bottom")], "HLL", "RK4", 0.6, 30
    mesh = mb.Mesh.generate(*args)
    xy = mesh.arrays["cell_coords"]
    if case == "sod":
        rho, p = np.where(xy[:, 0] < 0.5, 1.0, 0.125), np.where(xy[:, 0] < 0.5, 1.0, 0.1)
        U0 = np.stack([rho, 0 * rho, 0 * rho, p / 0.4], 1)
    elif case == "wedge":
        R = 101325.0 / (298.15 * 1.225)
        rho = np.full(mesh.n_cells, 101325.0 / (R * 300.0))
        U0 = np.stack([rho, rho * 600.0, 0 * rho, 101325.0 / 0.4 + 0.5 * rho * 600.0 ** 2], 1)
    else:
        U0 = _random_smooth_state(xy, np.random.default_rng(3))
    kw = dict(recon="FO", riemann=riemann, integrator=integ, bcs=bcs)
    a, b = mb.Solver(mesh, fp_mode=fp, **kw), mb.Solver(mesh, fp_mode=fp, **kw)
    a.set_state(U0); b.set_state(U0)
    monkeypatch.delenv("MLB_SMALL_STEP", raising=False)
    ta = a.run(n, cfl=cfl)
    monkeypatch.setenv("MLB_SMALL_STEP", "1")
    tb = b.run(n, cfl=cfl)
    for blocks in ("1", "5"):                                       # grid-stride loops over fewer blocks than there is work for
        monkeypatch.setenv("MLB_SMALL_STEP_BLOCKS", blocks)
        c = mb.Solver(mesh, fp_mode=fp, **kw)
        c.set_state(U0)
        tc = c.run(n, cfl=cfl)
        assert np.array_equal(c.get_state(), b.get_state()) and tc == tb, blocks
        c.close()
    monkeypatch.delenv("MLB_SMALL_STEP_BLOCKS", raising=False)
    monkeypatch.delenv("MLB_SMALL_STEP", raising=False)
    assert int(a.get("stats")[12]) == 0 and int(b.get("stats")[12]) == n
    Ua, Ub = a.get_state(), b.get_state()
    assert np.isfinite(Ua).all() and np.abs(Ua - U0).max() > 0
    assert a.time()[1] == b.time()[1] == n
    if fp == "strict":
        assert np.array_equal(Ua, Ub) and ta == tb
        so = oracle_mod.Solver(oracle_mod.Mesh.generate(*args), **kw)
        so.set_state(U0)
        for _ in range(n):
            so.take_step(so.calc_dt(cfl))
        assert gu.field_err(Ub, so.get("U")) <= 1e-12                # (HLLC's TRRS branch: pow() of libm vs CUDA)
    else:
        assert gu.field_err(Ub, Ua) <= 1e-12 and abs(ta[0] - tb[0]) <= 1e-12 * ta[0]
    # continuing with the other path from the cooperative kernel's state: the buffers are left as the multi-kernel step leaves them
    a.run(9, cfl=cfl); b.run(9, cfl=cfl)
    assert np.array_equal(a.get_state(), b.get_state()) if fp == "strict" else gu.field_err(b.get_state(), a.get_state()) <= 1e-12
    a.close(); b.close()


def test_device_side_field_ranges_and_nan_count():
    """mlb_field_ranges = max_array / min_array of Solver::do_checks (solver.cpp:434-437; `a > max` / `a < min`, so NaN never
    wins) + the NaN test of check_fields (solver.cpp:470-498), against numpy on the exported state."""
    mesh = mb.Mesh.generate("wedge", 60, 40, 4.0, 1.5)
    s = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", bcs=SYM4, fp_mode="strict")
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(5))
    s.set_state(U0)
    s.run(2, cfl=0.5)
    U, P = s.get_state(prim=True)
    rng, n_nan = s.field_ranges()
    assert n_nan == 0
    for i, n in enumerate(s.FIELD_NAMES):
        col = U[:, i] if i < 4 else P[:, i - 4]
        assert rng[n] == (col.min(), col.max()), n
    U0[17, 2] = np.nan
    s.set_state(U0)
    rng, n_nan = s.field_ranges()
    U, P = s.get_state(prim=True)
    assert n_nan == int(np.isnan(U).sum() + np.isnan(P).sum()) and n_nan >= 2
    assert rng["RHOU_Y"] == (np.nanmin(U[:, 2]), np.nanmax(U[:, 2]))


def test_fast_mode_drift_over_a_run(capsys):
    """north_star: "<= 1e-12 per step, with drift reported over the run".  Vortex advection on a jittered unstructured mesh,
    normalised TENO weights (the reference-faithful ones go non-finite within a step), 300 steps, sampled every 25 steps:
      |F - S|   FAST (FMA, compact device-built tables, re-associated sums) against the bit-faithful STRICT mode,
      |S' - S|  STRICT against ITSELF started from an initial state perturbed by +-1 ulp per entry.
    The scheme amplifies rounding-level perturbations exponentially (x100 per 100 steps at p = 3; a boundary instability of the
    stencil construction, profiles/r02_drift_study.txt has the full study to the point where the solution leaves the finite
    range in both modes), so the bar for FAST is: <= 1e-12 per step at the start, and never above the curve a 1-ulp
    perturbation of the bit-faithful mode follows."""
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(96, 96, 10.0, 10.0, seed=5)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    rng = np.random.default_rng(2026)
    U0p = U0 * (1.0 + np.where(rng.random(U0.shape) < 0.5, -1.0, 1.0) * 2.0 ** -52)
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", order=3, bcs=syn.EXTRAP4, teno_fixed=True, keep_stage_rhs=False)
    ss, sp, sf = mb.Solver(mesh, fp_mode="strict", **kw), mb.Solver(mesh, fp_mode="strict", **kw), mb.Solver(mesh, fp_mode="fast", **kw)
    ss.set_state(U0); sp.set_state(U0p); sf.set_state(U0)
    rows = []
    for k in range(12):
        ss.run(25, cfl=0.1); sp.run(25, cfl=0.1); sf.run(25, cfl=0.1)
        Us, Up, Uf = ss.get_state(), sp.get_state(), sf.get_state()
        assert np.isfinite(Us).all()
        rows.append((25 * (k + 1), gu.field_err(Uf, Us), gu.field_err(Up, Us), abs(sf.time()[0] - ss.time()[0]) / ss.time()[0]))
    with capsys.disabled():
        print("\ndrift (max field-relative difference of U): FAST vs STRICT | STRICT from a 1-ulp perturbed start vs STRICT | relative difference of t")
        for n, e, ep, te in rows:
            print("  step %4d   |F-S| %.2e   |S'-S| %.2e   dt %.2e" % (n, e, ep, te))
    assert rows[0][1] <= 25 * TOL                                     # <= 1e-12 per step
    for n, e, ep, te in rows:
        assert e <= 8.0 * max(ep, n * 1e-15), (n, e, ep)              # FAST never leaves the band of the scheme's own sensitivity
        assert te <= max(n * 1e-15, 10.0 * e)                        # the time axis (dt = cfl / max spectral radius) follows the state
    g_f = (rows[-1][1] / rows[3][1]) ** (1.0 / 8)                     # growth per 25 steps, steps 100 .. 300
    g_p = (rows[-1][2] / rows[3][2]) ** (1.0 / 8)
    assert g_f <= 2.0 * g_p + 1.0, (g_f, g_p)                         # same growth rate as the 1-ulp perturbation: no extra error source


@pytest.mark.parametrize("recon,integ", [("FO", "SSPRK3"), ("TENO", "RK4")])
def test_graph_replayed_run_equals_step_by_step(recon, integ):
    """mlb_run captures one step (CFL kernel + stages) into a CUDA graph and replays it (n_steps >= 8); the result must be
    the one of calc_dt + take_step issued step by step, bit for bit."""
    mesh = mb.Mesh.generate("cartesian_tri", 24, 18, 1.0, 0.75)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(9))
    kw = dict(recon=recon, riemann="HLLC", integrator=integ, order=2, bcs=SYM4, fp_mode="fast", teno_fixed=True, keep_stage_rhs=False)
    a, b = mb.Solver(mesh, **kw), mb.Solver(mesh, **kw)
    a.set_state(U0); b.set_state(U0)
    t, dt = a.run(20, cfl=0.3)
    for _ in range(20):
        b.calc_dt(0.3)
        b.take_step()
    assert int(a.get("stats")[11]) == 19 and int(b.get("stats")[11]) == 0
    assert np.array_equal(a.get_state(), b.get_state())
    assert a.time() == b.time() and a.time()[1] == 20


@pytest.mark.parametrize("mtype,nx,ny", [("cartesian", 1, 1), ("cartesian_tri", 1, 1), ("cartesian", 3, 1)])
def test_tiny_meshes_vs_oracle(oracle_mod, mtype, nx, ny):
    """Edge case: one or two cells, every (or almost every) face on a boundary, empty or one-face interior zone."""
    bcs = [dict(name="left", type="symmetry"), dict(name="right", type="extrapolation"), dict(name="top", type="wall_adiabatic"),
           dict(name="bottom", type="symmetry")]
    mesh, om = mb.Mesh.generate(mtype, nx, ny, 1.0, 0.5), oracle_mod.Mesh.generate(mtype, nx, ny, 1.0, 0.5)
    U0 = _random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(3))
    sg = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", bcs=bcs, fp_mode="strict")
    so = oracle_mod.Solver(om, "FO", "HLLC", "SSPRK3", bcs=bcs)
    sg.set_state(U0); so.set_state(U0)
    assert np.array_equal(sg.calc_rhs(), so.calc_rhs())
    dto, dtg = so.calc_dt(0.5), sg.calc_dt(0.5)
    assert dtg == dto
    so.take_step(dto); sg.take_step()
    assert np.array_equal(sg.get_state(), so.get("U"))


def test_mesh_read_from_a_gmsh_file_runs_like_the_same_mesh_from_arrays(tmp_path):
    """the vortex on a jittered triangulation handed over as arrays, and on the same mesh written to a Gmsh
    file and read back (faces renumbered, hence a different - equally valid - accumulation order in the residual)."""
    from mallard_b200 import meshio, synthetic as syn
    mesh = syn.jittered_tri(40, 32, 10.0, 8.0, seed=21)
    path = tmp_path / "vortex.msh"
    meshio.write_gmsh(mesh, str(path))
    back = meshio.read_gmsh(str(path))
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"], centre=(5.0, 4.0))
    res = []
    for m in (mesh, back):
        bcs = [dict(name=n, type="extrapolation") for n, _ in m.zones if n != "interior"]
        s = mb.Solver(m, "TENO", "HLLC", "SSPRK3", order=3, bcs=bcs, fp_mode="strict", teno_fixed=True, keep_stage_rhs=False)
        s.set_state(U0)
        s.run(3, cfl=0.3)
        res.append(s.get_state())
    assert np.isfinite(res[0]).all()
    assert gu.field_err(res[1], res[0]) <= 3 * TOL


def test_python_write_vtu_holds_the_exported_fields(tmp_path):
    """Solver.write_vtu (mlb_write_vtu) from Python: the appended arrays are the state get_state returns, in file order."""
    from test_gpu_dropin import read_vtu
    mesh = mb.Mesh.generate("wedge", 20, 8, 4.0, 1.5)
    s = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", bcs=SYM4, fp_mode="strict")
    s.set_state(_random_smooth_state(mesh.arrays["cell_coords"], np.random.default_rng(8)))
    s.run(4, cfl=0.5)
    name = s.write_vtu(str(tmp_path / "w"), 4, ["RHO", "RHOE", "P", "H", "CFL"])
    v = read_vtu(name)
    U, P, cfl = s.get_state(prim=True, cfl_local=True)
    assert np.array_equal(v["RHO"], U[:, 0]) and np.array_equal(v["RHOE"], U[:, 3])
    assert np.array_equal(v["P"], P[:, 2]) and np.array_equal(v["H"], P[:, 4]) and np.array_equal(v["CFL"], cfl)
    assert np.array_equal(v["connectivity"], mesh.arrays["nodes_of_cell"]) and (v["types"] == 7).all()
    with pytest.raises(mb.MallardError, match="Unknown variable"):
        s.write_vtu(str(tmp_path / "w"), 5, ["RHO", "VORTICITY"])
