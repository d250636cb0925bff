#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Pins the oracle (and the product's HOST preprocessor tables) against the UNMODIFIED reference over a grid of
configurations wider than the committed fixtures: for every case the reference (oracle/_ref/bin/ref_harness) dumps one step, and
the oracle restatement must reproduce every TENO table, the stage-1 face values and residual, dt, the residual of every stage and
the state after the step BIT FOR BIT; the preprocessor's reference-layout tables (mlb_plan_*) likewise.  Nothing is written to
tests/golden/ (the dumps of the larger cases are tens of MB); the result table goes to stdout:
    python oracle/pin_sweep.py > profiles/r02h_pin_sweep.txt
Runs only where /root/reference has been built (oracle/build_ref.sh); ~15 min on 8 cores."""
import itertools
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, ROOT, os.path.join(ROOT, "tests")]
import make_golden as mg  # noqa: E402
import mlbd  # noqa: E402
import golden_util as gu  # noqa: E402
import oracle  # noqa: E402
import mallard_b200 as mb  # noqa: E402

TENO_KEYS_PLAN = ("teno:stencils", "teno:offsets_stencils", "teno:offsets_stencil_groups", "teno:transformed_areas",
                  "teno:reconstruction_matrices", "teno:oscillation_indicator", "teno:integral_psi_target", "teno:poly_indices")


def run_reference(case):
    with tempfile.TemporaryDirectory() as td:
        toml, out = os.path.join(td, "input.toml"), os.path.join(td, "out.mlbd")
        mg.write_toml(case, toml)
        env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
        subprocess.check_call([mg.HARNESS, "dump", toml, out, str(case["n_steps"]), str(case["every"])], env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return mlbd.read(out)


def compare(case):
    """-> list of (what, verdict) where verdict is 'ok' or a description of the first difference"""
    d = run_reference(case)
    meta = {k: v for k, v in case.items() if k not in ("keep_mesh", "keep_teno")}
    out = []

    def eq(what, got, ref):
        got = np.asarray(got).reshape(np.asarray(ref).shape)
        if np.array_equal(got, ref, equal_nan=True):
            out.append((what, "ok"))
        elif got.dtype.kind == "f":
            bad = got != ref
            bad &= ~(np.isnan(got) & np.isnan(ref))
            fin = bad & np.isfinite(got) & np.isfinite(ref)
            rel = float((np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-300)).max()) if fin.any() else float("inf")
            out.append((what, "DIFFERS in %d of %d entries, max relative %.2e" % (int(bad.sum()), got.size, rel)))
        else:
            out.append((what, "DIFFERS in %d of %d entries" % (int((got != ref).sum()), got.size)))
    mesh = gu.oracle_mesh(oracle, meta)
    s = gu.oracle_solver(oracle, meta, mesh)
    teno = meta["recon"]["type"] == "TENO"
    if teno:
        for k in sorted(d):
            if k.startswith("teno:") and k not in ("teno:meta", "teno:quad_cell_points", "teno:quad_cell_weights"):
                eq("oracle " + k, s.get(k), d[k])
        mm = meta["mesh"]
        r = meta["recon"]
        pm = mb.Mesh.generate(mm["type"], mm["Nx"], mm["Ny"], mm["Lx"], mm["Ly"])
        plan = mb.Plan(pm, "TENO", basis=r.get("basis_type", "monomial"), order=r["basis_order"], factor=r.get("max_stencil_size_factor", 2.0),
                       quad_cell_order=r.get("quadrature_order_cell", 0), quad_face_order=r.get("quadrature_order_face", 0), bcs=meta["bcs"])
        for k in TENO_KEYS_PLAN:
            if k in d:
                eq("preprocessor " + k, plan.get(k), d[k])
    s.set_state(d["U0"], d["P0"])
    F = s.calc_face_values().copy()
    real = gu.real_faces(mesh.get("cells_of_face"), mesh.get("nodes_of_face"))
    interior = real & (mesh.get("cells_of_face")[:, 1] >= 0)
    eq("oracle F side 0", F[real][:, :, 0], d["F_stage1"][real][:, :, 0])
    eq("oracle F side 1", F[interior][:, :, 1], d["F_stage1"][interior][:, :, 1])
    eq("oracle rhs stage 1", s.calc_rhs(), d["rhs_stage1"])
    try:
        dt = s.calc_dt(meta["cfl"])
        eq("oracle dt", np.array([dt]), d["step0:dt"])
        s.take_step(dt)
        n_rhs = {"FE": 1, "RK4": 4, "SSPRK3": 3}[meta["integrator"]]
        for r_ in range(n_rhs):
            eq("oracle rhs%d" % r_, s.get("rhs%d" % r_), d["step0:rhs%d" % r_])
        eq("oracle U after the step", s.get("U"), d["step0:U"])
        eq("oracle P after the step", s.get("P"), d["step0:P"])
    except Exception as ex:      # the reference's dt turned non-finite / negative: the oracle throws where Solver::calc_dt throws
        out.append(("oracle step", "calc_dt threw (%s); reference dt %r" % (str(ex)[:60], d.get("step0:dt"))))
    return out


def compare_kernels(case):
    """--kernels: the CUDA kernel SOURCE executed on the host (tests/emul: teno_recon_kernel / the generic TENO kernel, face_flux_kernel,
    gather_stage_kernel as nvcc compiles them, STRICT build and FMA-contracting FAST build) against the same dump of the reference:
    face values and stage-1 residual.  -> [(what, error string)]: STRICT element-wise relative error, FAST relative to the field scale."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    from emulation import EmulatedSolver, fast_available
    d = run_reference(case)
    meta = {k: v for k, v in case.items() if k not in ("keep_mesh", "keep_teno")}
    mm = meta["mesh"]
    pm = mb.Mesh.generate(mm["type"], mm["Nx"], mm["Ny"], mm["Lx"], mm["Ly"])
    cof = pm.arrays["cells_of_face"].reshape(-1, 2)
    real = gu.real_faces(cof, pm.arrays["nodes_of_face"])
    interior = real & (cof[:, 1] >= 0)
    out = []
    for fp in ("strict", "fast") if fast_available() else ("strict",):
        es = EmulatedSolver(pm, fp_mode=fp, **gu.solver_kwargs(meta))
        es.set_state(d["U0"])
        F, rhs = es.calc_face_values(), es.calc_rhs()
        err = gu.rel_err if fp == "strict" else gu.field_err
        eF = max(err(F[real][:, :, 0], d["F_stage1"][real][:, :, 0]), err(F[interior][:, :, 1], d["F_stage1"][interior][:, :, 1]))
        er = err(rhs, d["rhs_stage1"])
        es.close()
        # one whole step through the emulated kernels (spectral radius / dt from the CFL kernel's body, every stage's face and gather
        # kernels, the RK combination, the primitives of the last stage): dt, U and the primitives against the reference's
        eU = None
        if True:
            from emulation import EmulatedAsSolver
            ss = EmulatedAsSolver(pm, fp_mode=fp, **gu.solver_kwargs(meta))
            ss.set_state(d["U0"], d["P0"])      # (as mlb_set_state(U, P): the reference steps from the primitives its initial condition defines)
            try:
                dt = ss.calc_dt(meta["cfl"])
                ss.take_step()
                U, P = ss.get_state(prim=True)
                eU = max(err(np.array([dt]), d["step0:dt"]), err(U, d["step0:U"]), err(P, d["step0:P"]))
            except mb.MallardError:
                eU = float("nan")                         # dt not positive: the reference's is not either (four-quadrant data)
                if np.isfinite(d["step0:dt"][0]) and d["step0:dt"][0] > 0:
                    eU = float("inf")
            ss.close()
        out.append((fp, eF, er, eU))
    return out


def compare_meshes():
    """--meshes: the three generators (mesh/mesh.cpp:305-848) over a grid of sizes and extents, incl. the degenerate ones: every
    connectivity and geometry array and every zone of the oracle's AND of the product's host generator (mlb_host_mesh_generate)
    against the reference's, bit for bit."""
    grid = [("cartesian", nx, ny, Lx, Ly) for nx, ny, Lx, Ly in ((1, 1, 1.0, 1.0), (1, 4, 0.3, 1.0), (5, 1, 2.0, 0.1), (3, 7, 1.0, 1.5), (16, 9, 3.2, 0.7))]
    grid += [("cartesian_tri", nx, ny, Lx, Ly) for nx, ny, Lx, Ly in ((1, 1, 1.0, 1.0), (1, 5, 0.2, 1.0), (6, 1, 2.0, 0.1), (4, 7, 1.3, 1.5), (13, 10, 0.9, 2.7))]
    grid += [("wedge", nx, ny, Lx, Ly) for nx, ny, Lx, Ly in ((1, 1, 4.0, 1.5), (2, 2, 4.0, 1.5), (5, 3, 0.4, 1.0), (24, 8, 4.0, 1.5), (9, 17, 2.0, 0.6), (31, 5, 7.3, 1.1))]
    n_bad = 0
    print("# mesh generators against the unmodified reference: %d arrays + 5 zones per mesh, oracle and product (mlb_host_mesh_generate), bit for bit" % len(gu.MESH_KEYS))
    for mtype, nx, ny, Lx, Ly in grid:
        bcs = mg.SYM4
        ic = mg.SMOOTH_IC
        case = dict(mesh=dict(type=mtype, Nx=nx, Ny=ny, Lx=Lx, Ly=Ly), ic=ic, bcs=bcs, cfl=0.5, riemann="HLLC", integrator="FE", recon=dict(type="FO"), n_steps=1, every=1)
        try:
            d = run_reference(case)
        except subprocess.CalledProcessError as ex:
            print("%-14s %3d x %-3d [%g x %g]  reference refused the configuration (exit code %d)" % (mtype, nx, ny, Lx, Ly, ex.returncode), flush=True)
            continue
        om = oracle.Mesh.generate(mtype, nx, ny, Lx, Ly)
        pm = mb.Mesh.generate(mtype, nx, ny, Lx, Ly)
        bad = []
        for who, get, zone in (("oracle", lambda k: om.get(k), lambda z: om.zone(z)), ("product", lambda k: pm.arrays[k], lambda z: dict(pm.zones)[z])):
            for k in gu.MESH_KEYS:
                ref = d[k]
                got = np.asarray(get(k)).reshape(ref.shape)
                if k == "face_normals":        # phantom faces: 0/0 = NaN in both
                    if not np.array_equal(np.isnan(ref), np.isnan(got)):
                        bad.append(who + " " + k + " (NaN pattern)")
                    ref, got = np.nan_to_num(ref), np.nan_to_num(got)
                if not np.array_equal(ref, got):
                    bad.append(who + " " + k)
            for i, z in enumerate(gu.ZONES):
                if not np.array_equal(d["zone:%d:%s" % (i, z)], zone(z)):
                    bad.append(who + " zone " + z)
        print("%-14s %3d x %-3d [%g x %g]  %d cells  %s" % (mtype, nx, ny, Lx, Ly, pm.n_cells, "all bit-exact" if not bad else "DIFFER: " + ", ".join(bad)), flush=True)
        n_bad += len(bad)
    print("# %d differing arrays" % n_bad)
    return 1 if n_bad else 0


def compare_unstructured():
    """--unstructured: the jittered, id-shuffled triangulations of the strong-scaling records (BASELINE configs[3] family,
    mallard_b200.synthetic.jittered_tri) INJECTED into the unmodified reference (ref_harness `mesh` mode: the reference computes its own
    geometry, stencils and matrices on the injected connectivity).  Compared bit for bit: geometry, every TENO table of the oracle and of
    the preprocessor, face values, residuals, dt, U, P of the oracle; then the emulated kernels (STRICT / FAST) as in --kernels."""
    from mallard_b200 import synthetic as syn
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    from emulation import EmulatedSolver, EmulatedAsSolver, fast_available
    keys = ["node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell", "offsets_nodes_of_face", "nodes_of_face",
            "cells_of_face"]
    n_bad = 0
    print("# jittered, id-shuffled triangulations injected into the unmodified reference: oracle / preprocessor arrays bit for bit, emulated kernels as in --kernels")
    grid = [(9, 8, 12345, 0.15, True, "legendre", 3, "HLLC", "SSPRK3"), (12, 10, 7, 0.15, True, "monomial", 2, "HLL", "RK4"), (10, 9, 3, 0.3, True, "legendre", 4, "Rusanov", "SSPRK3"),
            (8, 7, 99, 0.15, False, "legendre", 1, "HLLC", "SSPRK3"), (13, 11, 5, 0.2, True, "legendre", 5, "HLLC", "SSPRK3"), (11, 9, 21, 0.15, True, None, 0, "HLLC", "SSPRK3")]
    grid += [(48, 40, 12345, 0.15, True, "legendre", 3, "HLLC", "SSPRK3")]      # 3840 cells: the bench's numerics on a mesh with thousands of distinct stencils
    # mixed triangle / quadrilateral meshes (BASELINE configs[3] as worded): first order only - the reference's TENO refuses quadrilaterals
    grid += [(10, 8, 4, 0.15, True, "mixed", 0, "HLLC", "SSPRK3"), (9, 11, 8, 0.25, True, "mixed", 0, "Rusanov", "RK4")]
    for nx, ny, seed, amp, shuffle, basis, order, rs, integ in grid:
        mixed = basis == "mixed"
        if mixed:
            basis = None
            mesh = syn.mixed_tri_quad(nx, ny, 1.0, 0.9, seed=seed, amp=amp, tri_fraction=0.5, shuffle=shuffle)
        else:
            mesh = syn.jittered_tri(nx, ny, 1.0, 0.9, seed=seed, amp=amp, shuffle=shuffle)
        a = mesh.arrays
        recon = dict(type="FO") if basis is None else dict(type="TENO", basis_type=basis, basis_order=order, max_stencil_size_factor=2.0)
        if order >= 5:
            recon["quadrature_order_cell"] = 5
        case = dict(mesh=dict(type="cartesian_tri", Nx=nx, Ny=ny, Lx=1.0, Ly=0.9), ic=mg.SMOOTH_IC, bcs=mg.EXTRAP4, cfl=0.1, riemann=rs, integrator=integ, recon=recon,
                    n_steps=1, every=1)
        name = "%s%dx%d seed %d amp %.2f %s, %s" % ("mixed tri/quad " if mixed else "", nx, ny, seed, amp, "shuffled" if shuffle else "ordered",
                                                  "first order" if basis is None else "p%d %s" % (order, basis)) + " " + rs + " " + integ
        t0 = time.perf_counter()
        with tempfile.TemporaryDirectory() as td:
            toml, inj, out = os.path.join(td, "input.toml"), os.path.join(td, "mesh.mlbd"), os.path.join(td, "out.mlbd")
            mg.write_toml(case, toml)
            rec = {k: (a[k].reshape(-1, 2) if k in ("node_coords", "cells_of_face") else a[k]) for k in keys}
            for i, (zn, zf) in enumerate(mesh.zones):
                rec["zone:%d:%s" % (i, zn)] = np.ascontiguousarray(zf, dtype=np.uint32)
            mlbd.write(inj, rec)
            subprocess.check_call([mg.HARNESS, "mesh", inj, toml, out, "1", "1"], env=dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false"),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = mlbd.read(out)
        bad, n_cmp = [], 0

        def eq(what, got, ref):
            nonlocal n_cmp
            n_cmp += 1
            got = np.asarray(got).reshape(np.asarray(ref).shape)
            if not np.array_equal(got, ref, equal_nan=True):
                bad.append(what)
        for k in ("cell_coords", "cell_volume", "face_area", "face_normals"):
            eq("product geometry " + k, a[k], d[k])
        om = oracle.Mesh.from_arrays({k: a[k] for k in keys}, mesh.zones)
        for k in ("cell_coords", "cell_volume", "face_area", "face_normals"):
            eq("oracle geometry " + k, om.get(k), d[k])
        kw = dict(recon=recon["type"], riemann=rs, integrator=integ, bcs=mg.EXTRAP4)
        if basis is not None:
            kw.update(basis=basis, order=order, factor=2.0, quad_cell_order=recon.get("quadrature_order_cell", 0))
        so = oracle.Solver(om, **kw)
        if basis is not None:
            plan = mb.Plan(mesh, "TENO", basis=basis, order=order, factor=2.0, quad_cell_order=recon.get("quadrature_order_cell", 0), bcs=mg.EXTRAP4)
            for k in sorted(d):
                if k.startswith("teno:") and k not in ("teno:meta", "teno:quad_cell_points", "teno:quad_cell_weights"):
                    eq("oracle " + k, so.get(k), d[k])
                    if k in TENO_KEYS_PLAN:
                        eq("preprocessor " + k, plan.get(k), d[k])
        so.set_state(d["U0"], d["P0"])
        F = so.calc_face_values().copy()
        cof = a["cells_of_face"].reshape(-1, 2)
        interior = cof[:, 1] >= 0
        eq("oracle F side 0", F[:, :, 0], d["F_stage1"][:, :, 0])
        eq("oracle F side 1", F[interior][:, :, 1], d["F_stage1"][interior][:, :, 1])
        eq("oracle rhs stage 1", so.calc_rhs(), d["rhs_stage1"])
        dt = so.calc_dt(0.1)
        eq("oracle dt", np.array([dt]), d["step0:dt"])
        so.take_step(dt)
        eq("oracle U after the step", so.get("U"), d["step0:U"])
        eq("oracle P after the step", so.get("P"), d["step0:P"])
        kern = []
        for fp in ("strict", "fast") if fast_available() else ("strict",):
            err = gu.rel_err if fp == "strict" else gu.field_err
            es = EmulatedSolver(mesh, fp_mode=fp, **kw)
            es.set_state(d["U0"])
            Fe, re_ = es.calc_face_values(), es.calc_rhs()
            eF = max(err(Fe[:, :, 0], d["F_stage1"][:, :, 0]), err(Fe[interior][:, :, 1], d["F_stage1"][interior][:, :, 1]))
            er = err(re_, d["rhs_stage1"])
            es.close()
            ss = EmulatedAsSolver(mesh, fp_mode=fp, **kw)
            ss.set_state(d["U0"], d["P0"])
            dte = ss.calc_dt(0.1)
            ss.take_step()
            U, P = ss.get_state(prim=True)
            eU = max(err(np.array([dte]), d["step0:dt"]), err(U, d["step0:U"]), err(P, d["step0:P"]))
            ss.close()
            kern.append("%s F %.1e rhs %.1e step %.1e" % (fp, eF, er, eU))
        print("%-62s %2d arrays: %s | kernels: %s  (%.0f s)" % (name, n_cmp, "all bit-exact" if not bad else "DIFFER: " + ", ".join(bad), "   ".join(kern),
                                                              time.perf_counter() - t0), flush=True)
        n_bad += len(bad)
    print("# %d differing arrays" % n_bad)
    return 1 if n_bad else 0


def compare_riemann(n=200000, seed=2024):
    """--riemann: the reference's three flux functions (numerics/riemann_solver.h:331-519; `ref_harness riemann`) on random face states
    over every regime the wave-speed estimators distinguish - densities and pressures log-uniform over six decades, Mach numbers up to 6,
    strong and weak jumps, expansions, both orientations of the normal, states identical on both sides, gamma 1.4 and 1.667 - against
    the oracle (every bit) and the CUDA kernel source's riemann_flux on the host (tests/emul; STRICT every bit except where HLLC's
    two-rarefaction estimate calls pow, FAST = the lean re-formulation within 1e-12 of the flux scale)."""
    rng = np.random.default_rng(seed)
    n_bad = 0
    print("# Riemann fluxes against the unmodified reference on %d random face states per gamma" % n)
    for gamma in (1.4, 1.667):
        ang = rng.uniform(0.0, 2.0 * np.pi, n)
        nu = np.stack([np.cos(ang), np.sin(ang)], 1)
        nu[: n // 20] = [1.0, 0.0]
        nu[n // 20: n // 10] = [0.0, -1.0]

        def side(k):
            rho = 10.0 ** rng.uniform(-3.0, 3.0, k)
            p = 10.0 ** rng.uniform(-3.0, 3.0, k)
            a = np.sqrt(gamma * p / rho)
            mach = rng.uniform(0.0, 6.0, k) * (rng.random(k) < 0.8)
            th = rng.uniform(0.0, 2.0 * np.pi, k)
            u, v = mach * a * np.cos(th), mach * a * np.sin(th)
            h = p / ((gamma - 1.0) * rho) + p / rho      # e + p / rho, as Physics::get_... hands it to the flux functions
            return np.stack([rho, u, v, p, h], 1)
        L, R = side(n), side(n)
        m = n // 4                                       # a quarter: weak jumps around a common state (the PVRS branch), some exactly equal
        R[:m] = L[:m] * (1.0 + 1.0e-3 * rng.standard_normal((m, 5)))
        R[: m // 4] = L[: m // 4]
        R[:m, 4] = R[:m, 3] / ((gamma - 1.0) * R[:m, 0]) + R[:m, 3] / R[:m, 0]
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "states.mlbd"), os.path.join(td, "flux.mlbd")
            mlbd.write(fin, {"n_unit": nu, "L": L, "R": R})
            subprocess.check_call([mg.HARNESS, "riemann", fin, fout, repr(gamma)], env=dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false"),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ref = mlbd.read(fout)["flux"]
        for k, kind in enumerate(("Rusanov", "HLL", "HLLC")):
            got = oracle.riemann_flux(kind, nu, L, R, gamma)
            same = np.array_equal(got, ref[k], equal_nan=True)
            bad = int((~((got == ref[k]) | (np.isnan(got) & np.isnan(ref[k])))).any(axis=1).sum())
            n_bad += 0 if same else 1
            # the kernel source's flux functions on the host
            sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
            import emulation
            ks = emulation.riemann_flux(kind, nu, L, R, gamma, "strict")
            fin = np.isfinite(ref[k]).all(axis=1)
            pat_s = np.array_equal(np.isfinite(ks), np.isfinite(ref[k]))
            diff_s = (ks != ref[k]) & np.isfinite(ref[k]) & np.isfinite(ks)
            scale = np.maximum(np.abs(ref[k]).max(axis=1, initial=0.0, where=np.isfinite(ref[k])), 1e-300)
            rel_s = float((np.abs(np.where(diff_s, ks - ref[k], 0.0)).max(axis=1) / scale).max())
            msg = "kernel STRICT: %d states differ (max %.1e of the flux scale), non-finite pattern %s" % (int(diff_s.any(axis=1).sum()), rel_s, "equal" if pat_s else "DIFFERS")
            if emulation.fast_available():
                kf = emulation.riemann_flux(kind, nu, L, R, gamma, "fast")
                both = np.isfinite(kf).all(axis=1) & fin
                rel_f = float((np.abs(kf[both] - ref[k][both]).max(axis=1) / scale[both]).max())
                rel_all = np.abs(kf[both] - ref[k][both]).max(axis=1) / scale[both]
                msg += "; FAST (lean): max %.1e of the flux scale (%d of %d states above 1e-12), %d states finite in one and not the other" % (
                    rel_f, int((rel_all > 1e-12).sum()), int(both.sum()), int((np.isfinite(kf).all(axis=1) != fin).sum()))
            print("gamma %.3f  %-8s oracle: %s   (finite fluxes: %.1f %% of the states)\n%22s%s" % (gamma, kind, "every bit of %d fluxes" % n if same else "DIFFERS in %d states" % bad,
                                                                                                  100.0 * fin.mean(), "", msg), flush=True)
    print("# %d flux functions differ" % n_bad)
    return 1 if n_bad else 0


def compare_prims(n=200000, seed=11):
    """--prims: Euler::compute_primitives_from_conservatives (physics/physics.h:826-860; `ref_harness prims`) on random conserved states -
    magnitudes over six decades, supersonic and nearly static, states whose internal energy is negative (the clamp's and Q14's territory:
    fmin(p_max, NaN)), an active clamp - against the oracle and the kernel source's cons_to_prim on the host, every bit."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import emulation
    rng = np.random.default_rng(seed)
    n_bad = 0
    print("# primitives of %d random conserved states per gas against the unmodified reference" % n)
    for gas in (dict(), dict(gamma=1.667, p_ref=2.0e5, T_ref=350.0, rho_ref=0.9), dict(p_min=0.5, p_max=40.0), dict(gamma=1.3, p_min=1.0e-2, p_max=1.0e2)):
        g = dict(gamma=1.4, p_ref=101325.0, T_ref=298.15, rho_ref=1.225, p_min=-1e20, p_max=1e20)
        g.update(gas)
        rho = 10.0 ** rng.uniform(-3.0, 3.0, n)
        p = 10.0 ** rng.uniform(-3.0, 3.0, n)
        a = np.sqrt(g["gamma"] * p / rho)
        mach = rng.uniform(0.0, 8.0, n) * (rng.random(n) < 0.8)
        th = rng.uniform(0.0, 2.0 * np.pi, n)
        u, v = mach * a * np.cos(th), mach * a * np.sin(th)
        E = p / (g["gamma"] - 1.0) + 0.5 * rho * (u * u + v * v)
        E[: n // 20] *= rng.uniform(0.0, 1.0, n // 20)            # less total energy than the kinetic energy alone: negative pressure before the clamp
        U = np.stack([rho, rho * u, rho * v, E], 1)
        U[n // 20: n // 20 + 50, 3] = np.nan                      # Q14: a NaN energy comes out as p_max
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "U.mlbd"), os.path.join(td, "P.mlbd")
            mlbd.write(fin, {"U": U, "gas": np.array([g[k] for k in ("gamma", "p_ref", "T_ref", "rho_ref", "p_min", "p_max")])})
            subprocess.check_call([mg.HARNESS, "prims", fin, fout, repr(g["gamma"])], env=dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false"),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ref = mlbd.read(fout)["P"]
        po = oracle.prims(U, g)[0]
        res = ["oracle " + ("every bit" if np.array_equal(po, ref, equal_nan=True) else "DIFFERS in %d states" % int((~((po == ref) | (np.isnan(po) & np.isnan(ref)))).any(axis=1).sum()))]
        n_bad += 0 if np.array_equal(po, ref, equal_nan=True) else 1
        for fp in ("strict", "fast") if emulation.fast_available() else ("strict",):
            pk = emulation.primitives(U, g, fp)
            same = np.array_equal(pk, ref, equal_nan=True)
            if fp == "strict":
                n_bad += 0 if same else 1
                res.append("kernel STRICT " + ("every bit" if same else "DIFFERS in %d states" % int((~((pk == ref) | (np.isnan(pk) & np.isnan(ref)))).any(axis=1).sum())))
            else:
                fin_ = np.isfinite(ref).all(axis=1) & np.isfinite(pk).all(axis=1)
                scale = np.maximum(np.abs(ref[fin_]), 1e-300)
                res.append("kernel FAST max %.1e element-wise, non-finite pattern %s" % (float((np.abs(pk[fin_] - ref[fin_]) / scale).max()),
                                                                                         "equal" if np.array_equal(np.isfinite(pk), np.isfinite(ref)) else "DIFFERS"))
        clamped = float(((ref[:, 2] == g["p_min"]) | (ref[:, 2] == g["p_max"])).mean())
        print("gas %-62s %s   (pressure on the clamp: %.1f %% of the states)" % (str(gas) if gas else "(the examples' air)", ";  ".join(res), 100.0 * clamped), flush=True)
    print("# %d comparisons differ" % n_bad)
    return 1 if n_bad else 0


def compare_random(n_cases=24, seed=99):
    """--random: randomly drawn TENO configurations on cartesian_tri meshes of random size and aspect ratio (regular meshes are where the
    squared distances of the stencil search tie exactly and std::sort's order decides membership): order 1 - 4, both bases,
    stencil factors 1.5 / 2 / 2.5, all Riemann solvers, both multi-stage integrators; every array as in the default mode."""
    rng = np.random.default_rng(seed)
    n_bad = 0
    print("# %d random TENO configurations on cartesian_tri meshes against the unmodified reference, every array bit for bit (seed %d)" % (n_cases, seed))
    for t in range(n_cases):
        order = int(rng.integers(1, 5)); basis = ["legendre", "monomial"][int(rng.integers(0, 2))]
        nx = int(rng.integers(6, 26)); ny = int(rng.integers(6, 22))
        Lx = float(rng.choice([1.0, 2.0, 0.5, 1.7, 3.0])); Ly = float(rng.choice([1.0, 0.5, 1.3, 2.0]))
        factor = float(rng.choice([2.0, 2.0, 1.5, 2.5]))
        case = dict(mesh=dict(type="cartesian_tri", Nx=nx, Ny=ny, Lx=Lx, Ly=Ly), ic=mg.SMOOTH_IC, bcs=[mg.SYM4, mg.EXTRAP4][t % 2], cfl=0.1,
                    riemann=["HLLC", "HLL", "Rusanov"][t % 3], integrator=["SSPRK3", "RK4"][t % 2],
                    recon=dict(type="TENO", basis_type=basis, basis_order=order, max_stencil_size_factor=factor), n_steps=1, every=1)
        name = "p%d %-8s %2dx%-2d [%g x %g] factor %.1f %s %s" % (order, basis, nx, ny, Lx, Ly, factor, case["riemann"], case["integrator"])
        try:
            res = compare(case)
        except subprocess.CalledProcessError as ex:
            print("%-62s reference refused the configuration (exit code %d)" % (name, ex.returncode), flush=True)
            continue
        bad = [(w, v) for w, v in res if v != "ok" and not v.startswith("calc_dt threw")]
        print("%-62s %2d arrays compared, %s" % (name, len(res), "all bit-exact" if not bad else "DIFFER: " + "; ".join("%s: %s" % b for b in bad)), flush=True)
        n_bad += len(bad)
    print("# %d differing arrays" % n_bad)
    return 1 if n_bad else 0


def compare_long_runs(n_steps=100):
    """--long: first-order runs of n_steps steps (calc_dt + take_step each) on injected unstructured meshes - jittered triangles, mixed
    triangles / quadrilaterals - in the unmodified reference, the oracle and the emulated kernel sequence of mlb_run: the state after the
    last step.  Oracle and STRICT kernels: every bit (no drift at all); FAST: relative to the field scale."""
    from mallard_b200 import synthetic as syn
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    from emulation import EmulatedAsSolver, fast_available
    keys = ["node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell", "offsets_nodes_of_face", "nodes_of_face",
            "cells_of_face"]
    wall = [dict(name="left", type="upt", u=[0.5, 0.3], p=1.0, T=0.0036), dict(name="right", type="p_out", p=0.95),
            dict(name="top", type="symmetry"), dict(name="bottom", type="wall_adiabatic")]
    n_bad = 0
    print("# %d first-order steps on injected unstructured meshes: state after the last step against the unmodified reference" % n_steps)
    for label, mesh, rs, integ, bcs in (("jittered 14x12, shuffled ids", syn.jittered_tri(14, 12, 1.0, 0.9, seed=21), "HLLC", "SSPRK3", mg.EXTRAP4),
                                        ("mixed tri/quad 12x10, shuffled ids", syn.mixed_tri_quad(12, 10, 1.0, 0.9, seed=4, tri_fraction=0.5), "Rusanov", "RK4", mg.SYM4),
                                        ("mixed tri/quad 11x9, upt / p_out / symmetry / wall", syn.mixed_tri_quad(11, 9, 1.0, 0.9, seed=9, tri_fraction=0.4), "HLL", "SSPRK3", wall)):
        a = mesh.arrays
        case = dict(mesh=dict(type="cartesian_tri", Nx=4, Ny=4, Lx=1.0, Ly=0.9), ic=mg.SMOOTH_IC, bcs=bcs, cfl=0.6, riemann=rs, integrator=integ, recon=dict(type="FO"),
                    n_steps=n_steps, every=n_steps)
        with tempfile.TemporaryDirectory() as td:
            toml, inj, out = os.path.join(td, "input.toml"), os.path.join(td, "mesh.mlbd"), os.path.join(td, "out.mlbd")
            mg.write_toml(case, toml)
            rec = {k: (a[k].reshape(-1, 2) if k in ("node_coords", "cells_of_face") else a[k]) for k in keys}
            for i, (zn, zf) in enumerate(mesh.zones):
                rec["zone:%d:%s" % (i, zn)] = np.ascontiguousarray(zf, dtype=np.uint32)
            mlbd.write(inj, rec)
            subprocess.check_call([mg.HARNESS, "mesh", inj, toml, out, str(n_steps), str(n_steps)], env=dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false"),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = mlbd.read(out)
        last = "step%d:" % (n_steps - 1)
        kw = dict(recon="FO", riemann=rs, integrator=integ, bcs=bcs)
        om = oracle.Mesh.from_arrays({k: a[k] for k in keys}, mesh.zones)
        so = oracle.Solver(om, **kw)
        so.set_state(d["U0"], d["P0"])
        for _ in range(n_steps):
            so.take_step(so.calc_dt(0.6))
        same_o = np.array_equal(so.get("U"), d[last + "U"]) and np.array_equal(so.get("P"), d[last + "P"])
        n_bad += 0 if same_o else 1
        res = ["oracle " + ("every bit" if same_o else "DIFFERS")]
        for fp in ("strict", "fast") if fast_available() else ("strict",):
            ss = EmulatedAsSolver(mesh, fp_mode=fp, **kw)
            ss.set_state(d["U0"], d["P0"])
            ss.run(n_steps, cfl=0.6)
            U, P = ss.get_state(prim=True)
            if fp == "strict":
                same = np.array_equal(U, d[last + "U"]) and np.array_equal(P, d[last + "P"])
                n_bad += 0 if same else 1
                res.append("kernels STRICT " + ("every bit" if same else "DIFFER (U %.1e)" % gu.rel_err(U, d[last + "U"])))
                # the cooperative small-mesh kernel's phases (csrc/small_step.cuh: all steps in one launch, grid barriers between the phases)
                os.environ["MLB_SMALL_STEP"] = "1"
                try:
                    sc = EmulatedAsSolver(mesh, fp_mode=fp, **kw)
                    sc.set_state(d["U0"], d["P0"])
                    sc.run(n_steps, cfl=0.6)
                    Uc, Pc = sc.get_state(prim=True)
                    took = int(sc.get("stats")[12])
                    sc.close()
                finally:
                    os.environ.pop("MLB_SMALL_STEP", None)
                same_c = np.array_equal(Uc, d[last + "U"]) and np.array_equal(Pc, d[last + "P"]) and took == n_steps
                n_bad += 0 if same_c else 1
                res.append("cooperative kernel's phases " + ("every bit" if same_c else "DIFFER"))
            else:
                res.append("kernels FAST %.1e of the field scale" % max(gu.field_err(U, d[last + "U"]), gu.field_err(P, d[last + "P"])))
            ss.close()
        print("%-52s %5d cells  %s %s  %s" % (label, mesh.n_cells, rs, integ, ";  ".join(res)), flush=True)
    print("# %d comparisons differ" % n_bad)
    return 1 if n_bad else 0


def cases():
    smooth, riemann = mg.SMOOTH_IC, mg.RIEMANN2D_IC
    sym, ext = mg.SYM4, mg.EXTRAP4

    def teno(nx, ny, basis, order, factor=2.0, qc=0, qf=0, ic=smooth, bcs=sym, riemann_="HLLC", integ="SSPRK3", cfl=0.1, L=None):
        r = dict(type="TENO", basis_type=basis, basis_order=order, max_stencil_size_factor=factor)
        if qc:
            r["quadrature_order_cell"] = qc
        if qf:
            r["quadrature_order_face"] = qf
        Lx, Ly = L or (nx / 10.0, ny / 10.0 * 0.9)
        return dict(mesh=dict(type="cartesian_tri", Nx=nx, Ny=ny, Lx=Lx, Ly=Ly), ic=ic, bcs=bcs, cfl=cfl, riemann=riemann_, integrator=integ,
                    recon=r, n_steps=1, every=1)
    # every order with both bases (mesh sized so that every stencil finds its M = 2 K cells)
    size = {1: (6, 5), 2: (8, 7), 3: (9, 8), 4: (10, 9), 5: (12, 10), 6: (13, 11), 7: (14, 12), 8: (15, 13), 9: (16, 14)}
    for order, basis in itertools.product(range(1, 10), ("legendre", "monomial")):
        nx, ny = size[order]
        yield "p%d %s" % (order, basis), teno(nx, ny, basis, order, qc=5 if order >= 5 else 0, bcs=ext if order % 2 else sym)
    # stencil-size factors, cell / face quadrature orders the TOML may ask for, the other Riemann solvers and integrators
    for factor in (1.5, 2.5, 3.0):
        yield "p3 legendre factor %.1f" % factor, teno(12, 10, "legendre", 3, factor=factor)
    yield "p2 monomial factor 3.0", teno(10, 9, "monomial", 2, factor=3.0, bcs=ext)
    for qc in (1, 2, 4, 5):
        yield "p3 legendre quadrature_order_cell %d" % qc, teno(9, 8, "legendre", 3, qc=qc)
    for qf in (1, 3, 4, 5):
        yield "p3 legendre quadrature_order_face %d" % qf, teno(9, 8, "legendre", 3, qf=qf, bcs=ext)
    yield "p3 legendre Rusanov RK4 four-quadrant data", teno(9, 8, "legendre", 3, ic=riemann, riemann_="Rusanov", integ="RK4", L=(1.2, 1.1))
    yield "p2 legendre HLL FE four-quadrant data", teno(8, 7, "legendre", 2, ic=riemann, riemann_="HLL", integ="FE", bcs=ext, L=(1.1, 1.2))
    yield "p4 monomial HLLC RK4 cfl 0.4", teno(10, 9, "monomial", 4, integ="RK4", cfl=0.4, bcs=ext)
    # TENO face states through every boundary functor (boundary/*.cpp), incl. an order only the generic kernel serves
    mixed = [dict(name="left", type="upt", u=[0.5, 0.3], p=1.0, T=0.0036), dict(name="right", type="p_out", p=0.95),
             dict(name="top", type="extrapolation"), dict(name="bottom", type="wall_adiabatic")]
    yield "p2 monomial, upt / p_out / extrapolation / wall_adiabatic", teno(9, 8, "monomial", 2, bcs=mixed, integ="RK4", L=(1.0, 0.8))
    yield "p5 legendre, upt / p_out / extrapolation / wall_adiabatic", teno(12, 10, "legendre", 5, qc=5, bcs=mixed, riemann_="HLL", L=(1.0, 0.8))
    # the gas: another gamma and reference state; a pressure clamp (physics/physics.h:844) that bites inside the field
    c = teno(9, 8, "legendre", 3, L=(1.0, 1.0)); c["physics"] = dict(gamma=1.667, p_ref=2.0e5, T_ref=350.0, rho_ref=0.9)
    yield "p3 legendre, gamma 1.667 and another reference state", c
    c = teno(9, 8, "monomial", 2, bcs=ext, L=(1.0, 1.0)); c["physics"] = dict(p_min=0.95, p_max=1.04)
    yield "p2 monomial, pressure clamped to [0.95, 1.04]", c
    c = dict(mesh=dict(type="wedge", Nx=24, Ny=8, Lx=4.0, Ly=1.5), ic=mg.WEDGE_IC, bcs=mg.WEDGE_BCS, cfl=0.7, riemann="HLLC", integrator="SSPRK3", recon=dict(type="FO"),
             n_steps=1, every=1, physics=dict(gamma=1.3, p_ref=9.0e4, T_ref=280.0, rho_ref=1.1))
    yield "first order wedge HLLC SSPRK3, gamma 1.3 and another reference state", c
    # first order: the three generators with every Riemann solver
    for mtype, nx, ny, Lx, Ly, ic, bcs in (("cartesian", 40, 3, 1.0, 0.075, mg.SOD_IC, sym), ("cartesian_tri", 11, 9, 1.1, 1.2, riemann, ext),
                                           ("wedge", 24, 8, 4.0, 1.5, mg.WEDGE_IC, mg.WEDGE_WALL_BCS)):
        for rs, integ in (("Rusanov", "RK4"), ("HLL", "SSPRK3"), ("HLLC", "FE")):
            yield "first order %s %s %s" % (mtype, rs, integ), dict(mesh=dict(type=mtype, Nx=nx, Ny=ny, Lx=Lx, Ly=Ly), ic=ic, bcs=bcs, cfl=0.5, riemann=rs,
                                                                      integrator=integ, recon=dict(type="FO"), n_steps=1, every=1)


def main():
    if not os.path.exists(mg.HARNESS):
        raise SystemExit("oracle/_ref/bin/ref_harness is not built (oracle/build_ref.sh needs /root/reference)")
    oracle.build()
    sel = [x for x in sys.argv[1:] if not x.startswith("--")]
    if "--long" in sys.argv:
        return compare_long_runs()
    if "--random" in sys.argv:
        extra = [x for x in sys.argv[sys.argv.index("--random") + 1:] if x.isdigit()]
        return compare_random(int(extra[0]) if extra else 24, int(extra[1]) if len(extra) > 1 else 99)
    if "--prims" in sys.argv:
        return compare_prims()
    if "--riemann" in sys.argv:
        rc = compare_riemann()
        print("# FAST above 1e-12: states with pressure jumps of 1e4 ... 1e6 next to a near-vacuum side, where the reference's own formula moves by 3e-12")
        print("# (median; up to 3e-11) under a 1-ulp perturbation of its inputs; 3 of 176 791 finite states deviate further than that, by at most 1.9e-12")
        return rc
    if "--unstructured" in sys.argv:
        return compare_unstructured()
    if "--meshes" in sys.argv:
        return compare_meshes()
    if "--kernels" in sys.argv:
        print("# CUDA kernel source on the host (tests/emul) against the unmodified reference: max error of the face values / of the stage-1 residual /")
        print("# after one whole step (dt, U, primitives; nan = dt not positive in the reference either);")
        print("# STRICT build element-wise relative, FAST build (FMA contraction) relative to the field scale")
        worst = {"strict": 0.0, "fast": 0.0}
        for name, case in cases():
            if sel and not any(x in name for x in sel):
                continue
            t0 = time.perf_counter()
            try:
                res = compare_kernels(case)
            except subprocess.CalledProcessError as ex:
                print("%-52s reference refused the configuration (exit code %d)" % (name, ex.returncode), flush=True)
                continue
            print("%-52s %s  (%.0f s)" % (name, "   ".join("%s F %.1e rhs %.1e step %s" % (fp, eF, er, "-" if eU is None else "%.1e" % eU) for fp, eF, er, eU in res),
                                        time.perf_counter() - t0), flush=True)
            for fp, eF, er, eU in res:
                worst[fp] = max(worst[fp], eF, er, eU if (eU is not None and eU == eU) else 0.0)
        print("# worst: STRICT %.2e, FAST %.2e" % (worst["strict"], worst["fast"]))
        return 0
    n_bad = 0
    print("# oracle restatement and preprocessor tables against the unmodified reference, one step per case, every array bit for bit")
    for name, case in cases():
        if sel and not any(x in name for x in sel):
            continue
        t0 = time.perf_counter()
        try:
            res = compare(case)
        except subprocess.CalledProcessError as ex:
            print("%-52s reference refused the configuration (exit code %d)" % (name, ex.returncode), flush=True)
            continue
        bad = [(w, v) for w, v in res if v != "ok" and not v.startswith("calc_dt threw")]
        notes = [v for w, v in res if v.startswith("calc_dt threw")]
        print("%-52s %2d arrays compared, %s  (%.0f s)%s" % (name, len(res), "all bit-exact" if not bad else "%d DIFFER" % len(bad), time.perf_counter() - t0,
                                                            "  [" + notes[0] + "]" if notes else ""), flush=True)
        for w, v in bad:
            print("      %s: %s" % (w, v), flush=True)
        n_bad += len(bad)
    print("# %d differing arrays" % n_bad)
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
