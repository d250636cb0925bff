#!/usr/bin/env python
"""bench.py — cell-updates/s per RK stage of the explicit residual hot path (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] = examples/riemann_2d: cartesian_tri 1024x1024 (2 097 152 triangles), four-quadrant
Riemann initial condition, TENO (legendre, p=3) + HLLC + SSPRK3, cfl = 0.1.  A "step" is one full time step of
Solver::run's loop: calc_dt + the three RK stages (each: TENO reconstruction kernel, residual/flux/RK-update kernel).
    value = n_cells * n_stages * K / t        [cell-updates/s per RK stage], state and tables resident in HBM.
    e2e   = same through the take_step seam with HOST buffers: per step H2D of U from pinned memory, the step, D2H of U.
For N>1 (torchrun, one rank per GPU) every rank owns one 1024x1024-quad block of a (1024*N)x1024 mesh (weak scaling); the
per-stage halo exchange of ghost-cell states goes over NCCL.

`--workload vortex --nx 2828 --ny 2828` is BASELINE configs[3] (isentropic vortex on a 16 M-cell jittered, id-shuffled
triangulation; with torchrun the ONE mesh is split by recursive coordinate bisection: strong scaling); `--recon FO` the
first-order numerics of examples/sod and examples/wedge; `--fp strict` the bit-faithful mode.

Every default line also carries (each measured AFTER the main line, in child processes that cannot take it with them):
    strong       the 16 M-cell (BASELINE configs[3]) and 64 M-cell jittered, id-shuffled vortex meshes split over the N GPUs (rank-local ingest),
                 efficiency against the base point an earlier run of the same series left on the box (bench_multi.py)
    experiments  N = 1: configs[4] numerics (viscous) and configs[3] as worded (mixed triangles / quadrilaterals) on one GPU, the cooperative
                 small-mesh kernel on examples/sod and examples/wedge, and two measured paths so far only in the builder's own runs (STRICT
                 mode of the main configuration; the first-order numerics on 33.5 M cells); N > 1: configs[4] itself (viscous, the 64 M-cell mesh where it fits), the 16 M-cell
                 mesh cut by the graph partitioner (mlb_partition_graph_csr) next to its coordinate-bisection `strong` record, and the library's own
                 NCCL driver with its phase trace

`--impl reference` times the UNMODIFIED reference (oracle/_ref, Kokkos OpenMP, FP64) on the host cores on a bounded sample
of the same workload (same numerics, smaller mesh) — or the oracle port if the reference binary is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

N_STAGES = 3
REF_SAMPLE = (160, 160)     # ONE bounded CPU sample for both `cpu_baseline` and `--impl reference`: cartesian_tri 160x160 = 51 200 cells
if os.environ.get("MLB_REF_SAMPLE"):      # (the CPU suite checks the line's format on a smaller sample)
    REF_SAMPLE = tuple(int(x) for x in os.environ["MLB_REF_SAMPLE"].split("x"))
ALG_BYTES_STAGE = 7592.0    # TENO p=3 tri, whole stage (reference layout)
ALG_BYTES_RECON = 7472.0    # of which the reconstruction kernel: A+ 6400, areas 640, ids 320, offsets 20, state 32, geometry 60
SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]


def riemann2d_state(xy, gas_R=None):
    """examples/riemann_2d/input.toml:17-48 evaluated at the cell centroids (Solver::init_solution_analytical)."""
    x, y = xy[:, 0], xy[:, 1]
    l, r, b, t = (x < 0.8).astype(float), (x >= 0.8).astype(float), (y < 0.8).astype(float), (y >= 0.8).astype(float)
    rho = 1.5 * r * t + 0.532258064516129 * l * t + 0.137992831541219 * l * b + 0.532258064516129 * r * b
    u = 0.0 * r * t + 1.206045378311055 * l * t + 1.206045378311055 * l * b + 0.0 * r * b
    v = 0.0 * r * t + 0.0 * l * t + 1.206045378311055 * l * b + 1.206045378311055 * r * b
    p = 1.5 * r * t + 0.3 * l * t + 0.029032258064516 * l * b + 0.3 * r * b
    gamma = 1.4
    R = 101325.0 / (298.15 * 1.225)
    cp = R * gamma / (gamma - 1.0)
    cv = cp / gamma
    T = p / (rho * R)
    e = cv * T
    E = e + 0.5 * (u * u + v * v)
    U = np.stack([rho, rho * u, rho * v, rho * E], 1)
    P = np.stack([u, v, p, T, e + p / rho], 1)
    return U, P


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (recipe of B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    """(GB/s, where it comes from): the driver-written measurement, else the profiling recipe's fallback."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def write_ref_toml(path, nx, ny, n_steps):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden
    case = dict(mesh=dict(type="cartesian_tri", Nx=nx, Ny=ny, Lx=1.0, Ly=1.0), ic=make_golden.RIEMANN2D_IC, bcs=SYM4, cfl=0.1,
                riemann="HLLC", integrator="SSPRK3", recon=make_golden.TENO3, n_steps=n_steps)
    make_golden.write_toml(case, path, n_steps)


def reference_cpu(n_steps, n_warmup, nx=96, ny=96):
    """Times the reference's own CPU implementation on a bounded sample.  Returns (value, dict)."""
    harness = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")
    cores = host_cores()
    nc = 2 * nx * ny
    sample = "riemann_2d numerics (TENO p=3 legendre, HLLC, SSPRK3, cfl 0.1) on cartesian_tri %dx%d = %d cells, %d steps" % (nx, ny, nc, n_steps)
    if os.path.exists(harness):
        with tempfile.TemporaryDirectory() as td:
            toml = os.path.join(td, "input.toml")
            write_ref_toml(toml, nx, ny, n_steps + n_warmup)
            env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="spread", OMP_PLACES="cores")
            out = subprocess.run([harness, "time", toml, str(n_steps), str(n_warmup)], env=env, capture_output=True, text=True, check=True)
            line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
            r = json.loads(line)
        return r["cell_updates_per_s_per_stage"], dict(kind="reference", cores=int(r["threads"]), sample=sample + " (reference Kokkos OpenMP build, oracle/_ref)",
                                                        ms_per_step=1e3 * r["seconds"] / n_steps, init_seconds=r["init_seconds"])
    import oracle
    oracle.build()
    om = oracle.Mesh.generate("cartesian_tri", nx, ny, 1.0, 1.0)
    so = oracle.Solver(om, "TENO", "HLLC", "SSPRK3", order=3, bcs=SYM4)
    U, P = riemann2d_state(om.get("cell_coords"))
    so.set_state(U, P)
    for _ in range(n_warmup):
        so.take_step(so.calc_dt(0.1) if np.isfinite(so.get("U")).all() else 1e-6)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        try:
            dt = so.calc_dt(0.1)
        except Exception:
            dt = 1e-6
        so.take_step(dt if dt > 0 else 1e-6)
    sec = time.perf_counter() - t0
    return nc * N_STAGES * n_steps / sec, dict(kind="port", cores=1, sample=sample + " (oracle port, scalar)", ms_per_step=1e3 * sec / n_steps)


def kernel_roofline_table(prof, nc, a, peak, _unused, solver):
    """{kernel: ms per launch, algorithmic bytes per cell (DESIGN.md kernel table; triangles: 1.5 faces per cell, Q face quadrature points),
    achieved GB/s and fraction of the measured HBM peak on those bytes, and the same on ncu's dram__bytes of the committed capture}"""
    Q = 2                                                   # TENO p = 3: two Gauss points per face (numerics/quadrature.h)
    alg = {"teno_stream": ALG_BYTES_RECON, "teno_recon": ALG_BYTES_RECON,
           "face_flux_teno": 1.5 * (48.0 + 2 * Q * 32.0),  # per face: 48 B of connectivity / geometry / product + both sides' Q face states
           "face_flux_fo": 1.5 * (48.0 + 2 * 32.0),        # per face: the same 48 B + the two cells' states
           "gather_stage": 72.0 + 1.5 * 32.0,              # state in / out, volume, stage vector + the face products of the cell's faces
           "cfl": 200.0}
    traffic, fp64 = {}, {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("n_cells") == nc and tj.get("workload") == a.workload and tj.get("fp_mode") == a.fp:
            traffic = tj.get("bytes_per_launch", {})
            fp64 = tj.get("fp64_pipe_pct_of_peak", {})
    except Exception:
        pass
    out = {}
    for k, (ms_total, launches) in prof.items():
        if not launches or k not in alg:
            continue
        ms = ms_total / launches
        row = {"ms_per_launch": ms, "algorithmic_bytes_per_cell": alg[k], "achieved": alg[k] * nc / (ms * 1e-3) / 1e9, "unit": "GB/s"}
        row["frac"] = row["achieved"] / peak
        if traffic.get(k):
            row["traffic"] = traffic[k]
            row["frac_traffic"] = traffic[k] / (ms * 1e-3) / 1e9 / peak
        if fp64.get(k) is not None:
            row["fp64_pipe_pct_of_peak_ncu"] = fp64[k]
        out[k] = row
    return out


STRONG_TAG = "STRONG_RECORD "
T_START = time.perf_counter()


def json_line(obj):
    """One line of STRICT JSON: a non-finite float (a record of a run that went wrong may hold one) becomes a string, not the bare NaN /
    Infinity token Python would print and other parsers reject."""
    import math

    def safe(x):
        if isinstance(x, np.generic):
            x = x.item()
        if isinstance(x, float) and not math.isfinite(x):
            return repr(x)
        if isinstance(x, dict):
            return {str(k): safe(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [safe(v) for v in x]
        if isinstance(x, np.ndarray):
            return [safe(v) for v in x.tolist()]
        return x
    try:
        return json.dumps(safe(obj), allow_nan=False)
    except (TypeError, ValueError):          # never the reason for a missing line
        return json.dumps(obj, default=str)


def strong_records_in_child(a, popen=subprocess.Popen, min_left=120.0, task="strong"):
    """N = 1: the strong-scaling records (a 16 M-cell mesh: ~100 GB of device memory, minutes of set-up) - and, task by task, the records of
    bench_multi.EXPERIMENTS - run in a CHILD process under a deadline, after the main line has been measured: whatever happens there - a
    crash, an out-of-memory kill, a hang, an illegal address in a kernel that has never run on hardware - the main line is printed.
    Records completed before a failure are kept (the child prints each one as it finishes)."""
    import signal
    deadline = float(os.environ.get("MLB_BENCH_DEADLINE", "760"))          # the driver kills a scaling run at 870 s
    elapsed = time.perf_counter() - T_START
    left = deadline - elapsed
    if left < min_left:
        return [{"workload": task, "skipped": "time budget of the bench run (%.0f s elapsed)" % elapsed}]
    cmd = [sys.executable, os.path.abspath(__file__), "--strong-child", "--child-task", task, "--gpus", "1", "--steps", str(a.steps), "--warmup", str(a.warmup),
           "--fp", a.fp]
    env = dict(os.environ, MLB_BENCH_ELAPSED="%.1f" % elapsed)
    recs, note, out = [], None, ""
    try:
        p = popen(cmd, stdout=subprocess.PIPE, text=True, env=env, start_new_session=True)
    except Exception as ex:
        return [{"workload": task, "error": "could not start the child process: %s" % str(ex)[:200]}]
    try:
        out, _ = p.communicate(timeout=left)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except Exception:
            p.kill()
        try:
            out, _ = p.communicate(timeout=30)
        except Exception:
            out = ""
        note = {"workload": task, "aborted": "deadline of %.0f s reached before the remaining records of this child finished" % deadline}
    for l in (out or "").splitlines():
        if l.startswith(STRONG_TAG):
            try:
                recs.append(json.loads(l[len(STRONG_TAG):]))
            except Exception:
                pass
    if note is None and p.returncode != 0:
        note = {"workload": task, "error": "the child process ended with code %s" % p.returncode}
    if note is not None:
        recs.append(note)
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mu", type=float, default=0.0, help="dynamic viscosity (vortex workloads): > 0 adds the Navier-Stokes terms (BASELINE configs[4])")
    ap.add_argument("--workload", default="riemann_2d", choices=["riemann_2d", "vortex", "vortex_mixed", "sod", "wedge"],
                    help="riemann_2d = BASELINE configs[1] (the metric's configuration); vortex = configs[3] family: isentropic vortex on a "
                         "jittered, id-shuffled triangulation (--nx 2828 --ny 2828 = 16 M cells)")
    ap.add_argument("--nx", type=int, default=1024)
    ap.add_argument("--ny", type=int, default=1024)
    ap.add_argument("--fp", default="fast", choices=["strict", "fast"])
    ap.add_argument("--recon", default="TENO", choices=["TENO", "FO"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU owns an nx x ny block (default, the driver's scaling run); strong = the nx x ny mesh is "
                         "split over the GPUs")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling records (16 M / 64 M-cell partitioned vortex mesh) appended to "
                                                              "the default riemann_2d line")
    ap.add_argument("--ref-full", action="store_true", help="--impl reference only: run the reference on the FULL --nx x --ny mesh instead of the bounded "
                                                            "sample (1024x1024: ~8 min of serial set-up, ~30 GB) and cache the number in profiles/reference_full_size.json")
    ap.add_argument("--strong-child", action="store_true", help="internal (N = 1): run the strong-scaling records only and print them as one JSON list; "
                                                                "bench.py starts this in a child process so that no failure there can cost the main line")
    ap.add_argument("--child-task", default="strong", help="internal: strong | one of bench_multi.EXPERIMENTS")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "examples/riemann_2d: cartesian_tri %dx%d, %s+HLLC+SSPRK3, cfl=0.1, four-quadrant IC" % (
        a.nx, a.ny, "TENO(legendre,p=3)" if a.recon == "TENO" else "first-order reconstruction")
    if a.workload in ("vortex", "vortex_mixed"):
        workload = ("synthetic isentropic vortex on a jittered (+-0.15 h, seed 12345), id-shuffled %s %dx%d of [0,10]^2, "
                    "TENO(legendre,p=3)+HLLC+SSPRK3, cfl=0.1, extrapolation BCs%s"
                    % ("triangulation" if a.workload == "vortex" else "mixed triangle / quadrilateral mesh (half the quads cut in two)", a.nx, a.ny,
                       ", viscous: mu = %g, Pr = 0.72 (BASELINE configs[4])" % a.mu if a.mu > 0 else ""))

    if a.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(a.steps, 50))
        nxr, nyr = (a.nx, a.ny) if a.ref_full else REF_SAMPLE
        warm = max(1, min(a.warmup, 10))        # the driver compares `warmup` with what it asked for: honour it (a step of the sample is ~60 ms)
        value, info = reference_cpu(steps, warm, nx=nxr, ny=nyr)
        if a.ref_full:          # the full-size run (serial set-up ~8 min at 1024^2, ~30 GB): cache the number next to the other evidence
            try:
                json.dump({"value": value, "unit": "cell-updates/s", "cores": info["cores"], "sample": info["sample"], "ms_per_step": info["ms_per_step"],
                           "init_seconds": info.get("init_seconds"), "steps": steps}, open(os.path.join(ROOT, "profiles", "reference_full_size.json"), "w"), indent=1)
            except Exception:
                pass
        line = {"impl": "reference", "metric": "cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s", "n_gpus": a.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": info["sample"]},
                "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
                "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json_line(line))
        return

    import torch
    import mallard_b200 as mb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    if a.strong_child and world > 1:      # N > 1, inside the per-rank child processes started by bench_multi.experiment_children()
        import bench_multi
        peak, peak_src = hbm_peak()
        child = {"native_weak": bench_multi.native_weak_child, "graph_strong": bench_multi.graph_strong_child}.get(a.child_task, bench_multi.viscous_strong_child)
        rec = child(a, rank, world, local_rank, peak, peak_src)
        if rank == 0:
            print(STRONG_TAG + json_line(rec), flush=True)
        return
    if world > 1:
        import bench_multi
        return bench_multi.run(a, rank, world, local_rank, workload)
    if a.strong_child:          # N = 1, inside the child process started by strong_records_in_child() below
        import bench_multi
        torch.cuda.set_device(0)
        mb.set_host_threads(host_cores())
        peak, peak_src = hbm_peak()
        emit = lambda r: print(STRONG_TAG + json_line(r), flush=True)      # noqa: E731
        if a.child_task == "strong":
            bench_multi.strong_records(a, 0, 1, 0, peak, peak_src, on_record=emit)
        else:
            emit(bench_multi.experiment_record(a.child_task, a, peak, peak_src))
        return

    torch.cuda.set_device(0)
    t_setup = time.perf_counter()
    cfl, integ = 0.1, "SSPRK3"
    if a.workload in ("sod", "wedge"):
        # BASELINE configs[0] / configs[2] verbatim (examples/sod, examples/wedge): first order + HLLC + SSPRK3, cfl 1.  A few
        # thousand cells: a step is launch-bound, mlb_run replays it as a CUDA graph.  Use --steps 2000 or more.
        a.recon, a.no_cpu_baseline, a.no_e2e, cfl = "FO", True, True, 1.0
        R = 101325.0 / (298.15 * 1.225)
        cv = R / 0.4
        if a.workload == "sod":
            mesh = mb.Mesh.generate("cartesian", 1000, 1, 1.0, 1.0e-3)
            x = mesh.arrays["cell_coords"][:, 0]
            rho, p = np.where(x < 0.5, 1.0, 0.125), np.where(x < 0.5, 1.0, 0.1)
            u = np.zeros_like(x)
            bcs = SYM4
            workload = "examples/sod: cartesian 1000x1, first order + HLLC + SSPRK3, cfl 1"
        else:
            mesh = mb.Mesh.generate("wedge", 150, 50, 4.0, 1.5)
            n = mesh.n_cells
            p, T, u = np.full(n, 101325.0), np.full(n, 300.0), np.full(n, 600.0)
            rho = p / (R * T)
            bcs = [dict(name="left", type="upt", u=[600.0, 0.0], p=101325.0, T=300.0), dict(name="right", type="p_out", p=101325.0),
                   dict(name="top", type="symmetry"), dict(name="bottom", type="symmetry")]
            workload = "examples/wedge: wedge 150x50 quads, upt / p_out / symmetry, first order + HLLC + SSPRK3, cfl 1"
        e = p / (0.4 * rho)
        U0, P0 = np.stack([rho, rho * u, 0.0 * rho, rho * (e + 0.5 * u * u)], 1), None
    elif a.workload in ("vortex", "vortex_mixed"):
        from mallard_b200 import synthetic as syn
        mb.set_host_threads(host_cores())
        mesh = syn.jittered_tri(a.nx, a.ny, 10.0, 10.0, seed=12345) if a.workload == "vortex" else syn.mixed_tri_quad(a.nx, a.ny, 10.0, 10.0, seed=12345)
        U0, P0, bcs = syn.isentropic_vortex(mesh.arrays["cell_coords"]), None, syn.EXTRAP4
        a.no_cpu_baseline = True      # the reference cannot read an unstructured mesh (mesh.cpp:41-43); its per-cell cost is mesh independent
    else:
        Lx = a.nx / float(a.ny)                          # square cells; 1024 x 1024 is the reference's unit square
        mesh = mb.Mesh.generate("cartesian_tri", a.nx, a.ny, Lx, 1.0)
        xy = mesh.arrays["cell_coords"]
        U0, P0 = riemann2d_state(np.stack([xy[:, 0] / Lx, xy[:, 1]], 1))
        bcs = SYM4
    nc = mesh.n_cells
    mesh_s = time.perf_counter() - t_setup
    gas = dict(mu=a.mu) if (a.mu > 0 and a.workload in ("vortex", "vortex_mixed")) else None
    s = mb.Solver(mesh, a.recon, "HLLC", integ, order=3, bcs=bcs, fp_mode=a.fp, keep_stage_rhs=False, gas=gas,
                  teno_fixed=(a.workload == "vortex_mixed" or a.mu > 0))     # viscous / mixed-mesh runs are meant to stay finite
    stats = s.get("stats")
    setup_s = time.perf_counter() - t_setup
    s.set_state(U0, P0)

    # ---- device-resident timing (the state is reset before every timed region: reference-faithful TENO spreads NaN from
    #      the initial discontinuities by about one stencil width per stage, exactly as the reference does)
    s.run(a.warmup, cfl=cfl)
    launches0 = s.launch_count
    clocks = ClockSampler(0)
    clocks.start()
    s.synchronize()
    s.event_record(0)
    s.run(a.steps, cfl=cfl)
    s.event_record(1)
    ms = s.event_elapsed_ms(0, 1)
    clk = clocks.stop()
    launches = s.launch_count - launches0
    value = nc * N_STAGES * a.steps / (ms * 1e-3)

    # ---- per-kernel device time over the same region (CUDA events on the library's stream) for the roofline
    s.set_state(U0, P0)
    s.run(a.warmup, cfl=cfl)
    s.profile(True)
    s.run(a.steps, cfl=cfl)
    prof = s.profile_read()
    s.profile(False)
    peak, peak_src = hbm_peak()
    top = max(prof, key=lambda k: prof[k][0]) if prof else None
    roof = None
    kernels = {k: {"ms_total": v[0], "launches": int(v[1]), "ms_per_launch": v[0] / max(1, v[1])} for k, v in prof.items()}
    if top:
        per_launch_ms = prof[top][0] / prof[top][1]
        alg = {"teno_recon": ALG_BYTES_RECON, "teno_stream": ALG_BYTES_RECON, "face_flux_teno": 48.0, "gather_stage": 72.0, "face_flux_fo": 80.0,
               "cfl": 200.0}.get(top, ALG_BYTES_STAGE)
        achieved = alg * nc / (per_launch_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this very
        # workload (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); null for any other workload / size
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("n_cells") == nc and tj.get("workload") == a.workload and tj.get("fp_mode") == a.fp:
                traffic = tj["bytes_per_launch"].get(top)
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                # the same launch duration against the bytes DRAM really moved (the compact tables are smaller than the reference layout)
                "achieved_traffic": (traffic / (per_launch_ms * 1e-3) / 1e9) if traffic else None,
                "frac_traffic": (traffic / (per_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_cell": alg, "share_of_step": prof[top][0] / sum(v[0] for v in prof.values())}
        stage_ms = sum(prof[k][0] for k in prof if k != "cfl") / (a.steps * N_STAGES)
        alg_stage = ALG_BYTES_STAGE if a.recon == "TENO" else 152.0      # first order on triangles = 152 B per cell-update
        roof["stage_algorithmic_bytes_per_cell"] = alg_stage
        roof["stage_achieved"] = alg_stage * nc / (stage_ms * 1e-3) / 1e9
        roof["stage_frac"] = roof["stage_achieved"] / peak

    # ---- every kernel of the step against ITS bound (north_star: "the fused TENO+Riemann face kernel and RK update reach >= 60 % of
    #      their roofline"): event-timed launch duration against the algorithmic bytes of DESIGN.md's kernel table, and - where the committed
    #      ncu capture is of this very workload - against the bytes DRAM really moved; the FP64-pipe figure is ncu's, from the same capture
    kernel_rooflines = None
    try:
        kernel_rooflines = kernel_roofline_table(prof, nc, a, peak, None, s)
    except Exception as ex:          # reporting only: never the reason for a missing line
        kernel_rooflines = {"error": str(ex)[:200]}

    # ---- end to end through the take_step seam with host buffers (pinned): H2D U, calc_dt + step, D2H U every step
    e2e = None
    if not a.no_e2e:
        pin = torch.empty((nc, 4), dtype=torch.float64, pin_memory=True)
        Uh = pin.numpy()
        Uh[:] = U0
        k_e2e = max(3, min(a.steps, 10))
        for _ in range(3):
            s.take_step_host(Uh, cfl=cfl)
        Uh[:] = U0
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            s.take_step_host(Uh, cfl=cfl)
        sec = time.perf_counter() - t0
        e2e = {"value": nc * N_STAGES * k_e2e / sec, "unit": "cell-updates/s", "h2d_bytes_per_step": nc * 32, "d2h_bytes_per_step": nc * 32 + 64,
               "steps": k_e2e, "ms_per_step": 1e3 * sec / k_e2e, "api": "mlb_take_step_host (C ABI; host buffers in reference layout)"}

    cpu = None
    if not a.no_cpu_baseline:
        try:
            v, info = reference_cpu(20, 2, nx=REF_SAMPLE[0], ny=REF_SAMPLE[1])      # ~12 s of serial reference set-up + a few seconds of steps on the host cores
            cpu = {"value": v, "unit": "cell-updates/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
        except Exception as ex:   # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "cell-updates/s", "cores": host_cores(), "kind": "unavailable", "sample": str(ex)[:200]}

    graph_replayed = int(s.get("stats")[11])
    # ---- the timed steps run on data that turns non-finite (reference-faithful weights, as the reference): the same loop on data that
    #      stays finite (normalised weights) backs the claim that the cost is data independent
    finite = None
    if a.workload == "riemann_2d" and a.recon == "TENO" and not a.no_e2e:
        try:
            s.close()
            sf = mb.Solver(mesh, a.recon, "HLLC", integ, order=3, bcs=bcs, fp_mode=a.fp, keep_stage_rhs=False, teno_fixed=True)
            sf.set_state(U0, P0)
            sf.run(a.warmup, cfl=cfl)
            sf.synchronize()
            sf.event_record(0)
            sf.run(a.steps, cfl=cfl)
            sf.event_record(1)
            msf = sf.event_elapsed_ms(0, 1)
            Uf = sf.get_state()
            finite = {"teno_fixed": 1, "value": nc * N_STAGES * a.steps / (msf * 1e-3), "unit": "cell-updates/s", "ms_per_step": msf / a.steps,
                      "finite_fraction_of_cells_after_the_run": float(np.isfinite(Uf).all(axis=1).mean()),
                      "note": "normalised TENO weights (SURVEY N2): same kernels, same launches; compare with the line's value"}
            sf.close()
        except Exception as ex:
            finite = {"error": str(ex)[:200]}

    # ---- strong-scaling records: this N's point of the 16 M-cell (BASELINE configs[3]) / 64 M-cell partitioned vortex meshes
    strong, experiments = None, []
    if a.workload == "riemann_2d" and a.recon == "TENO" and not a.no_strong and (a.nx, a.ny) == (1024, 1024):
        s.close()                                    # (a closed context ignores a second close): the 16 M-cell mesh needs the GPU's memory
        strong = strong_records_in_child(a)
        import bench_multi
        for task in bench_multi.EXPERIMENTS:       # configs[4] numerics and configs[3] as worded, one child each (see bench_multi.experiment_record)
            experiments += strong_records_in_child(a, task=task)

    line = {"metric": "cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_cells": nc, "fp_mode": a.fp, "recon": a.recon, "l2": (("inputs larger than L2 (TENO tables %.1f GB per stage)" if a.recon == "TENO" else "inputs larger than L2 (%.1f GB of state, connectivity and face products per stage)") % (stats[2] / 1e9))
                             if stats[2] > 252e6 else "working set %.1f MB fits in L2: a launch-bound configuration, steps replayed as a CUDA graph" % (stats[2] / 1e6),
                       "graph_replayed_steps": graph_replayed, "data_independence": finite,
                       "device_bytes": stats[2], "preprocess_seconds": stats[1], "setup_seconds": setup_s, "mesh_seconds": mesh_s,
                       "preprocess": {"host_total_s": stats[1], "host_stencil_search_s": stats[8], "host_matrices_s": stats[9],
                                      "device_table_build_s": stats[10]},
                       "note": ("reference-faithful TENO: like the reference, the state turns non-finite inside step 1 (SURVEY 0.2); cost is "
                                "data-independent") if a.recon == "TENO" else "first-order path (the numerics of examples/sod and examples/wedge)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "kernels": kernels, "kernel_rooflines": kernel_rooflines, "strong": strong,
            "experiments": experiments or None}
    print(json_line(line))


if __name__ == "__main__":
    main()
