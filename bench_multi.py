"""bench.py's multi-GPU legs (launched by torchrun, one rank per GPU) and the strong-scaling records of every line.

Main line, N > 1 — weak scaling of BASELINE.json configs[1]: every rank owns one nx x ny block (2*nx*ny triangles) of a
(nx*N) x ny cartesian_tri mesh, partitioned in x by cell centroid.

`strong` records (every N, appended to the same JSON line) — BASELINE.json configs[3] / [4] family: ONE jittered, id-shuffled
triangulation (2828^2 quads = 16.0 M cells; 5657^2 = 64.0 M cells where it fits) cut by the library's recursive coordinate
bisection; every rank builds ITS PART of the mesh only (synthetic.jittered_tri_local -> mlb_create_local), so no rank ever holds
the global mesh.  value = global cells x stages x steps / time; efficiency is value_N / (N x value_1) of the same mesh.

Driver of the partitioned step: by default the split-phase C ABI (mlb_halo_pack / mlb_stage_begin / mlb_stage ...) with
torch.distributed as the communicator - NCCL send/recv straight between the library's device buffers on the library's
communication stream under the reconstruction of the interior cells, all_reduce(max) of the device-resident dt - the path the
round-1 scaling run proved on 2, 4 and 8 GPUs.  MLB_BENCH_NATIVE=1 selects the library's own NCCL driver instead (mlb_comm_init /
mlb_run_distributed: the same schedule inside the library, the step replayed as a CUDA graph); measured on 2 GPUs
(profiles/r02b_bench_n2_native.json), its 8-GPU run did not complete inside this round's GPU budget, so it is not the default.
Timing: barrier + device synchronise on both sides, CUDA events on the library's compute stream, MAX over ranks.

A watchdog ends the process cleanly (rank 0 prints the line with whatever records are complete) if the run approaches the
driver's per-run limit: a strong-scaling record must never cost the main line.
"""
import json
import os
import resource
import time

import numpy as np

import sys
import threading

T_START = time.perf_counter() - float(os.environ.get("MLB_BENCH_ELAPSED", "0"))      # (a child of bench.py inherits its parent's clock)
NATIVE = os.environ.get("MLB_BENCH_NATIVE") == "1"
DEADLINE_S = float(os.environ.get("MLB_BENCH_DEADLINE", "760"))    # the driver kills a run at 870 s
_STATE = {"line": None, "strong": [], "rank": 0, "done": False}


def _watchdog():
    while not _STATE["done"]:
        time.sleep(1.0)
        if time.perf_counter() - T_START > DEADLINE_S:
            if _STATE["rank"] == 0 and _STATE["line"] is not None:
                line = dict(_STATE["line"])
                line["strong"] = list(_STATE["strong"]) + [{"aborted": "deadline of %.0f s reached before the remaining records finished" % DEADLINE_S}]
                import bench
                sys.stdout.write(bench.json_line(line) + "\n")
                sys.stdout.flush()
            for pid in _STATE.get("children", []):       # (per-rank child processes of experiment_children)
                try:
                    os.killpg(pid, 9)
                except Exception:
                    pass
            os._exit(0 if _STATE["line"] is not None or _STATE["rank"] != 0 else 3)


def start_watchdog(rank):
    _STATE["rank"] = rank
    threading.Thread(target=_watchdog, daemon=True).start()
STRONG_MESHES = (("vortex_16M", 2828), ("vortex_64M", 5657))      # BASELINE configs[3] (16 M cells) and the configs[4] mesh (64 M)
MAX_CELLS_PER_GPU = 24.0e6                                        # TENO p=3 tables: 6.5 kB per cell of 180 GB
TIME_BUDGET_S = 560.0                                             # do not start another strong record after this much wall time


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the end-to-end leg (and the
    threads that fill them) sit on the GPU's NUMA node instead of wherever the launcher started the process.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def _reduce(x, world, op="max"):
    if world == 1:
        return np.asarray(x, dtype=np.float64)
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.asarray(x, dtype=np.float64), device="cuda")
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return t.cpu().numpy()


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


class Run:
    """One partitioned (or single-GPU) solver and its timed loop."""

    def __init__(self, solver, ds, world):
        self.s, self.ds, self.world = solver, ds, world

    def steps(self, n, cfl):
        if self.ds is not None:
            self.ds.run(n, cfl)
        else:
            self.s.run(n, cfl=cfl)

    def timed(self, n, cfl):
        s = self.s
        _barrier(self.world)
        s.synchronize()
        s.event_record(0)
        self.steps(n, cfl)
        s.event_record(1)
        ms = s.event_elapsed_ms(0, 1)
        s.synchronize()
        _barrier(self.world)
        return float(_reduce([ms], self.world)[0])

    def recon_roofline(self, n_steps, cfl, n_stages, peak, peak_src, alg_bytes):
        """Per-kernel device time over n_steps (CUDA events on the compute stream, steps issued eagerly): the reconstruction
        kernel's time per stage on the SLOWEST rank (interior + rim launches) against its algorithmic bytes."""
        s = self.s
        s.profile(True)
        self.steps(n_steps, cfl)
        prof = s.profile_read()
        s.profile(False)
        tot = sum(v[0] for v in prof.values())
        rec = prof.get("teno_stream", prof.get("teno_recon", (0.0, 0)))
        n_recon = int(s.get("stats")[6])
        ms_stage = rec[0] / (n_steps * n_stages)
        ach = alg_bytes * n_recon / (ms_stage * 1e-3) / 1e9 if ms_stage > 0 else 0.0
        stage_all = sum(v[0] for k, v in prof.items() if k != "cfl") / (n_steps * n_stages)
        worst = _reduce([ms_stage, stage_all], self.world, "max")
        # no ncu capture exists for this workload: the bytes the device holds (tables >> everything else; within 6 % of ncu's dram__bytes per
        # launch on the riemann_2d workload, profiles/ncu_traffic.json) stand in for the real traffic of one reconstruction pass
        dev_bytes = float(s.get("stats")[2])
        ach_dev = dev_bytes / (ms_stage * 1e-3) / 1e9 if ms_stage > 0 else 0.0
        fr = _reduce([ach / peak, ach_dev / peak], self.world, "min")
        return {"bound": "hbm", "kernel": "teno_stream (interior + rim launches of a stage)", "achieved": float(fr[0] * peak), "peak": peak, "unit": "GB/s",
                "frac": float(fr[0]), "traffic": None, "traffic_proxy_device_bytes_this_rank": dev_bytes, "frac_on_device_bytes": float(fr[1]),
                "peak_source": peak_src, "algorithmic_bytes_per_cell": alg_bytes,
                "rank": "slowest", "ms_per_stage_slowest_rank": float(worst[0]), "all_kernels_ms_per_stage_slowest_rank": float(worst[1]),
                "cells_reconstructed_this_rank": n_recon, "share_of_step_this_rank": rec[0] / tot if tot else None,
                "kernels_this_rank": {k: {"ms_total": v[0], "launches": int(v[1])} for k, v in prof.items()}}


TRANSPORT = ("grouped ncclSend/ncclRecv between device buffers + ncclAllReduce(max) of dt inside the library (mlb_run_distributed), step replayed as a CUDA graph"
             if NATIVE else "NCCL send/recv between the library's device buffers on its communication stream + all_reduce(max) of the device-resident dt "
                            "(split-phase C ABI driven over torch.distributed)")


def strong_record(name, nq, a, rank, world, device, peak, peak_src, mu=0.0, graph=False):
    """Strong scaling: the nq x nq jittered, id-shuffled triangulation of [0,10]^2 split over `world` GPUs.  mu > 0: with the
    Navier-Stokes terms and normalised TENO weights (BASELINE configs[4]).  graph: cut by the library's graph partitioner
    (mlb_partition_graph_csr on the closed-form dual graph of the mesh, computed by every rank for itself) instead of the
    recursive coordinate bisection."""
    import bench
    import mallard_b200 as mb
    from mallard_b200 import synthetic as syn
    nc = 2 * nq * nq
    rec = {"workload": "%s: isentropic vortex, jittered (+-0.15 h, seed 12345), id-shuffled triangulation %dx%d of [0,10]^2 (%d cells), "
                       "TENO(legendre,p=3%s)+HLLC+SSPRK3, cfl 0.1, extrapolation BCs%s; %s over %d GPU(s), rank-local ingest"
                       % (name, nq, nq, nc, ", normalised weights" if mu > 0 else "", ", Navier-Stokes terms: mu = %g, Pr = 0.72" % mu if mu > 0 else "",
                          "graph-partitioned (mlb_partition_graph_csr: multilevel recursive bisection of the cell-face dual graph)" if graph
                          else "recursive coordinate bisection", world),
           "n_cells": nc, "n_gpus": world, "scaling": "strong", "partitioner": "graph" if graph else "coordinate bisection"}
    t0 = time.perf_counter()
    part_fn = None
    if graph:      # every rank cuts the global dual graph for itself (deterministic, thread-count independent: all ranks agree)
        xadj, adj = syn.jittered_tri_dual_graph(nq, nq, seed=12345)
        t1 = time.perf_counter()
        gpart = mb.partition_graph(None, world, xadj=xadj, adj=adj)
        rows = np.repeat(np.arange(nc, dtype=np.int64), np.diff(xadj).astype(np.int64))
        rec.update(partition_seconds=time.perf_counter() - t1, dual_graph_seconds=t1 - t0, edge_cut=int((gpart[rows] != gpart[adj]).sum() // 2))
        del xadj, adj, rows
        part_fn = lambda n: gpart      # noqa: E731
    layers, lp, ds, s = 8, None, None, None
    for attempt in range(3):
        lp = syn.jittered_tri_local(nq, nq, 10.0, 10.0, world, rank, seed=12345, layers=layers, part_fn=part_fn)
        ok = 1.0
        try:
            kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode=a.fp, keep_stage_rhs=False)
            if mu > 0:
                kw.update(gas=dict(mu=mu), teno_fixed=True)
            if world > 1:
                s = mb.Solver(lp.mesh, part=lp.part_local, rank=rank, n_ranks=world, device=device, local=lp.local, **kw)
            else:
                s = mb.Solver(lp.mesh, device=device, **kw)
        except mb.MallardError as ex:
            if "more ghost layers" not in str(ex):
                raise
            ok, s = 0.0, None
        if _reduce([ok], world, "min")[0] > 0:      # every rank must agree before anything collective happens
            break
        if s is not None:
            s.close()
        layers *= 2
    else:
        raise RuntimeError("rank-local mesh: ghost layers insufficient")
    if world > 1:
        from mallard_b200.parallel import DistributedSolver
        ds = DistributedSolver.from_solver(s, rank, world, device, local=lp.local, native=NATIVE)
    stats = s.get("stats")
    n_owned = int(stats[4])
    s.set_state(syn.isentropic_vortex(lp.mesh.arrays["cell_coords"]))
    setup_s = time.perf_counter() - t0
    run = Run(s, ds, world)
    run.timed(a.warmup, 0.1)
    clocks = bench.ClockSampler(device) if rank == 0 else None      # the base point of a strong-scaling series is only as good as its clocks
    if clocks:
        clocks.start()
    ms = run.timed(a.steps, 0.1)
    if clocks:
        rec["clocks"] = clocks.stop()
    rec.update(value=nc * bench.N_STAGES * a.steps / (ms * 1e-3), unit="cell-updates/s", ms_per_step=ms / a.steps, steps=a.steps, warmup=a.warmup)
    rec["roofline"] = run.recon_roofline(max(2, min(a.steps, 5)), 0.1, bench.N_STAGES, peak, peak_src, bench.ALG_BYTES_RECON)
    if world > 1:
        peers, sc, rc = s.halo_info()
        h = _reduce([float(sc.sum()), float(rc.sum()), float(len(peers))], world, "max")
        rec["halo"] = {"max_send_cells_per_stage": int(h[0]), "max_recv_cells_per_stage": int(h[1]), "max_peers": int(h[2]),
                       "transport": TRANSPORT}
        own = _reduce([float(n_owned)], world, "max")
        rec["cells_per_gpu_max"] = int(own[0])
    mem = _reduce([resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0, stats[2] / 1e9, float(lp.mesh.n_cells), setup_s, stats[1], lp.seconds], world, "max")
    rec.update(host_rss_gb_per_rank_max=float(mem[0]), device_gb_per_rank_max=float(mem[1]), local_mesh_cells_max=int(mem[2]), ghost_layers=layers,
               setup_seconds=float(mem[3]), preprocess_seconds=float(mem[4]), mesh_seconds=float(mem[5]), graph_replayed_steps=int(s.get("stats")[11]))
    s.close()
    if rank == 0:
        # (a graph-partitioned record is compared with the same base point as the coordinate-bisection record of its mesh - one GPU
        #  holds the whole mesh whoever cuts it - and leaves no point of its own)
        strong_efficiency(rec, name + ("_viscous" if mu > 0 else ""), world, store=not graph)
    return rec


LIVE_BASELINES = os.environ.get("MLB_STRONG_BASELINES", "/tmp/mlb_strong_baselines.json")


def strong_efficiency(rec, name, world, live_path=None, max_age_s=6 * 3600.0, store=True):
    """Strong-scaling efficiency of a record = value_N / (N / N0 x value_N0), N0 the smallest GPU count this mesh has been measured
    on (1 where the mesh fits one GPU; the 64 M-cell mesh starts at 4).  The base point is the one measured ON THIS BOX by an earlier
    bench.py run of the same series (the driver's scaling run goes N = 1, 2, 4, 8 on one lease: every run leaves its records in
    LIVE_BASELINES) - same clocks, same host - and only failing that the committed number of an earlier round (profiles/
    r02_strong_baselines.json, which says where it comes from)."""
    import bench
    live_path = live_path or LIVE_BASELINES
    live = {}
    try:
        if time.time() - os.path.getmtime(live_path) < max_age_s:
            live = json.load(open(live_path))
    except Exception:
        live = {}
    base, src = live.get(name), "measured on this box by the N = %d run of the same series"
    if not (base and base.get("n_gpus", 0) < world):
        base, src = None, "profiles/r02_strong_baselines.json (%s)"
        try:
            base = json.load(open(os.path.join(bench.ROOT, "profiles", "r02_strong_baselines.json"))).get(name)
        except Exception:
            pass
        if not (base and base.get("n_gpus", 0) < world):
            base = None
    if base and base.get("value"):
        n0 = int(base["n_gpus"])
        rec["efficiency"] = rec["value"] / (world / n0 * base["value"])
        rec["efficiency_base"] = "%s on %d GPU(s): %.4g cell-updates/s, %s" % (name, n0, base["value"], src % (base.get("source") or n0))
    elif world == 1:
        rec["efficiency"] = 1.0
    # leave this run's point for the larger runs that follow (keep the smallest GPU count per mesh)
    try:
        if store and (name not in live or live[name].get("n_gpus", 1 << 30) >= world):
            live[name] = {"n_gpus": world, "value": rec["value"], "ms_per_step": rec["ms_per_step"]}
            tmp = live_path + ".%d.tmp" % os.getpid()
            json.dump(live, open(tmp, "w"))
            os.replace(tmp, live_path)
    except Exception:
        pass


def strong_records(a, rank, world, device, peak, peak_src, on_record=None):
    out = []

    def add(r):
        out.append(r)
        _STATE["strong"] = list(out)
        if on_record is not None:
            on_record(r)
    for name, nq in STRONG_MESHES:
        nc = 2 * nq * nq
        if nc / world > MAX_CELLS_PER_GPU:
            add({"workload": name, "n_cells": nc, "skipped": "%.1f M cells per GPU do not fit 180 GB" % (nc / world / 1e6)})
            continue
        elapsed = float(_reduce([time.perf_counter() - T_START], world, "max")[0])
        if elapsed > TIME_BUDGET_S:
            add({"workload": name, "n_cells": nc, "skipped": "time budget of the bench run (%.0f s elapsed)" % elapsed})
            continue
        try:
            add(strong_record(name, nq, a, rank, world, device, peak, peak_src))
        except Exception as ex:      # a strong record never costs the main line
            add({"workload": name, "n_cells": nc, "error": str(ex)[:300]})
            if world > 1:            # the ranks may have diverged: nothing collective can follow
                break
    return out


# ---- N = 1 only, each in its own child process of bench.py: the two configurations of BASELINE.json that the reference itself cannot run
#      (configs[4]: viscous terms; configs[3] as literally worded: mixed triangles / quadrilaterals under TENO).  Their kernels were
#      written after the round-2 GPU budget had been spent and are validated by the host emulation of their source
#      (tests/test_kernel_emulation.py); these records are their measurement.
EXPERIMENTS = ("vortex_viscous", "vortex_mixed", "small_step", "strict_mode", "first_order_33M")


def measured_path_record(task, a, peak):
    """Two paths that HAVE run on hardware, measured by the builder only so far (profiles/r01j_bench_strict.json, r02a_bench_first_order_33M...):
    `strict_mode` = the main line's configuration in the bit-faithful floating-point mode; `first_order_33M` = the first-order numerics of
    examples/sod and examples/wedge on a 33.5 M-cell triangulation (152 B per cell-update: SURVEY 8d).  Same timing as the main line."""
    import bench
    import mallard_b200 as mb
    if task == "strict_mode":
        nx = ny = int(os.environ.get("MLB_EXPERIMENT_NQ", "1024"))
        recon, fp, alg, what = "TENO", "strict", bench.ALG_BYTES_STAGE, "examples/riemann_2d (cartesian_tri %dx%d, TENO(legendre,p=3)+HLLC+SSPRK3, cfl 0.1) in STRICT mode" % (nx, ny)
    else:
        nx = ny = int(os.environ.get("MLB_EXPERIMENT_NQ", "4096"))
        recon, fp, alg, what = "FO", a.fp, 152.0, "first order + HLLC + SSPRK3 on cartesian_tri %dx%d, four-quadrant data, cfl 0.1" % (nx, ny)
    t0 = time.perf_counter()
    mesh = mb.Mesh.generate("cartesian_tri", nx, ny, 1.0, 1.0)
    U0, P0 = bench.riemann2d_state(mesh.arrays["cell_coords"])
    s = mb.Solver(mesh, recon, "HLLC", "SSPRK3", order=3, bcs=bench.SYM4, fp_mode=fp, keep_stage_rhs=False)
    setup_s = time.perf_counter() - t0
    s.set_state(U0, P0)
    s.run(a.warmup, cfl=0.1)
    s.synchronize()
    s.event_record(0)
    s.run(a.steps, cfl=0.1)
    s.event_record(1)
    ms = s.event_elapsed_ms(0, 1)
    s.set_state(U0, P0)
    s.profile(True)
    s.run(max(2, min(a.steps, 5)), cfl=0.1)
    prof = s.profile_read()
    s.profile(False)
    nc = mesh.n_cells
    n_prof = max(2, min(a.steps, 5))
    stage_ms = sum(v[0] for k, v in prof.items() if k != "cfl") / (n_prof * bench.N_STAGES)
    rec = {"workload": task + ": " + what, "n_cells": nc, "n_gpus": 1, "fp_mode": fp, "value": nc * bench.N_STAGES * a.steps / (ms * 1e-3), "unit": "cell-updates/s",
           "ms_per_step": ms / a.steps, "steps": a.steps, "warmup": a.warmup, "setup_seconds": setup_s, "device_gb": s.get("stats")[2] / 1e9,
           "kernels": {k: {"ms_per_launch": v[0] / max(1, v[1]), "launches": int(v[1])} for k, v in prof.items()},
           "roofline": {"bound": "hbm", "stage_algorithmic_bytes_per_cell": alg, "stage_achieved": alg * nc / (stage_ms * 1e-3) / 1e9 if stage_ms > 0 else None,
                        "peak": peak, "unit": "GB/s", "stage_frac": alg * nc / (stage_ms * 1e-3) / 1e9 / peak if stage_ms > 0 else None}}
    s.close()
    return rec


def small_step_record(a):
    """BASELINE configs[0] / configs[2] verbatim (examples/sod: 1000 cells, examples/wedge: 7500 cells; first order + HLLC + SSPRK3, cfl 1):
    launch-bound meshes.  mlb_run as it ships (one step = 7 kernels, replayed as a CUDA graph) against MLB_SMALL_STEP=1 (all steps in one
    cooperative kernel, csrc/small_step.cuh), from the same state: microseconds per step both ways and whether the two final states are the
    same bits."""
    import mallard_b200 as mb
    n = int(os.environ.get("MLB_EXPERIMENT_STEPS", "2000"))
    R = 101325.0 / (298.15 * 1.225)
    rec = {"workload": "small_step: examples/sod (cartesian 1000x1) and examples/wedge (150x50), first order + HLLC + SSPRK3, cfl 1, %d steps per leg: "
                       "CUDA-graph replayed multi-kernel step vs one cooperative kernel for the whole run (MLB_SMALL_STEP=1, opt-in)" % n,
           "n_gpus": 1, "unit": "us per step", "fp_mode": a.fp,
           "verification": "cooperative kernel written after the round-2 GPU budget was spent; its phases run on the host against the oracle and the "
                           "reference's dumps (tests/test_kernel_emulation.py); first measured here"}
    for case in ("sod", "wedge"):
        if case == "sod":
            mesh = mb.Mesh.generate("cartesian", 1000, 1, 1.0, 1.0e-3)
            x = mesh.arrays["cell_coords"][:, 0]
            rho, p, u = np.where(x < 0.5, 1.0, 0.125), np.where(x < 0.5, 1.0, 0.1), np.zeros_like(x)
            bcs = [dict(name=z, type="symmetry") for z in ("left", "right", "top", "bottom")]
        else:
            mesh = mb.Mesh.generate("wedge", 150, 50, 4.0, 1.5)
            p, u = np.full(mesh.n_cells, 101325.0), np.full(mesh.n_cells, 600.0)
            rho = p / (R * 300.0)
            bcs = [dict(name="left", type="upt", u=[600.0, 0.0], p=101325.0, T=300.0), dict(name="right", type="p_out", p=101325.0),
                   dict(name="top", type="symmetry"), dict(name="bottom", type="symmetry")]
        U0 = np.stack([rho, rho * u, 0.0 * rho, p / 0.4 + 0.5 * rho * u * u], 1)
        out = {"n_cells": mesh.n_cells}
        states = {}
        for leg, env in (("graph_replayed_kernels", None), ("cooperative_kernel", {"MLB_SMALL_STEP": "1"}),
                         ("cooperative_kernel_8_blocks", {"MLB_SMALL_STEP": "1", "MLB_SMALL_STEP_BLOCKS": "8"})):
            for k in ("MLB_SMALL_STEP", "MLB_SMALL_STEP_BLOCKS"):
                os.environ.pop(k, None)
            os.environ.update(env or {})
            s = mb.Solver(mesh, "FO", "HLLC", "SSPRK3", bcs=bcs, fp_mode=a.fp, keep_stage_rhs=False)
            s.set_state(U0)
            s.run(64, cfl=1.0)                         # warm-up (graph instantiation / cooperative launch attributes)
            s.set_state(U0)
            s.synchronize()
            s.event_record(0)
            s.run(n, cfl=1.0)
            s.event_record(1)
            ms = s.event_elapsed_ms(0, 1)
            states[leg] = s.get_state()
            st = s.get("stats")
            out[leg] = {"us_per_step": 1e3 * ms / n, "cell_updates_per_s": mesh.n_cells * 3 * n / (ms * 1e-3),
                        "steps_in_the_cooperative_kernel": int(st[12]) if len(st) > 12 else None, "graph_replayed_steps": int(st[11])}
            s.close()
        for k in ("MLB_SMALL_STEP", "MLB_SMALL_STEP_BLOCKS"):
            os.environ.pop(k, None)
        ref = states["graph_replayed_kernels"]
        out["finite"] = bool(np.isfinite(ref).all())
        for leg in ("cooperative_kernel", "cooperative_kernel_8_blocks"):
            out[leg]["same_bits_as_the_multi_kernel_path"] = bool(np.array_equal(states[leg], ref))
            out[leg]["max_difference_of_the_field_scale"] = float(np.abs(states[leg] - ref).max() / max(np.abs(ref).max(), 1e-300))
        rec[case] = out
    return rec


def experiment_record(task, a, peak, peak_src):
    import bench
    import mallard_b200 as mb
    from mallard_b200 import synthetic as syn
    mb.set_host_threads(bench.host_cores())
    if task == "small_step":
        return small_step_record(a)
    if task in ("strict_mode", "first_order_33M"):
        return measured_path_record(task, a, peak)
    t0 = time.perf_counter()
    if task == "vortex_viscous":
        nq, mu = int(os.environ.get("MLB_EXPERIMENT_NQ", "1024")), 1.0e-3
        mesh = syn.jittered_tri(nq, nq, 10.0, 10.0, seed=12345)
        what = ("isentropic vortex on a jittered, id-shuffled triangulation %dx%d of [0,10]^2, TENO(legendre,p=3, normalised weights)+HLLC+SSPRK3, cfl 0.1, "
                "Navier-Stokes terms: mu = %g, Pr = 0.72 (BASELINE configs[4] numerics on one GPU)" % (nq, nq, mu))
        legs = (("viscous", dict(gas=dict(mu=mu))), ("inviscid_same_mesh", {}))
    elif task == "vortex_mixed":
        nq = int(os.environ.get("MLB_EXPERIMENT_NQ", "512"))
        mesh = syn.mixed_tri_quad(nq, nq, 10.0, 10.0, seed=12345)
        what = ("isentropic vortex on a jittered mixed triangle / quadrilateral mesh (%dx%d quads of [0,10]^2, half of them cut in two), TENO(legendre,p=3, "
                "normalised weights, five stencils per quadrilateral)+HLLC+SSPRK3, cfl 0.1 (BASELINE configs[3] as worded; non-streaming reconstruction kernel)" % (nq, nq))
        legs = (("mixed", {}),)
    else:
        raise ValueError("unknown experiment %r" % task)
    nc = mesh.n_cells
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    rec = {"workload": task + ": " + what, "n_cells": nc, "n_gpus": 1, "unit": "cell-updates/s", "steps": a.steps, "warmup": a.warmup,
           "mesh_seconds": time.perf_counter() - t0,
           "verification": "kernels written after the round-2 GPU budget was spent; validated by the host emulation of their source and analytically "
                           "(tests/test_kernel_emulation.py); first measured here"}
    for leg, extra in legs:
        t1 = time.perf_counter()
        s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode=a.fp, keep_stage_rhs=False, teno_fixed=True, **extra)
        setup_s = time.perf_counter() - t1
        s.set_state(U0)
        s.run(a.warmup, cfl=0.1)
        s.synchronize()
        s.event_record(0)
        s.run(a.steps, cfl=0.1)
        s.event_record(1)
        ms = s.event_elapsed_ms(0, 1)
        U = s.get_state()
        s.set_state(U0)
        s.profile(True)
        s.run(max(2, min(a.steps, 5)), cfl=0.1)
        prof = s.profile_read()
        s.profile(False)
        stats = s.get("stats")
        out = {"value": nc * bench.N_STAGES * a.steps / (ms * 1e-3), "ms_per_step": ms / a.steps, "setup_seconds": setup_s, "device_gb": stats[2] / 1e9,
               "finite_fraction_of_cells_after_the_run": float(np.isfinite(U).all(axis=1).mean()),
               "kernels": {k: {"ms_per_launch": v[0] / max(1, v[1]), "launches": int(v[1])} for k, v in prof.items()}}
        if leg != "inviscid_same_mesh":
            # the residual of the initial state through the FAST kernels (timed above) and through the bit-faithful STRICT ones, on the device:
            # both modes of kernels that have no reference to be compared with must at least agree with each other to rounding
            try:
                s.set_state(U0)
                rf = s.calc_rhs()
                ss = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode="strict" if a.fp == "fast" else "fast", teno_fixed=True, **extra)
                ss.set_state(U0)
                rs = ss.calc_rhs()
                ss.close()
                col = np.abs(rs).max(axis=0)
                scale = np.maximum(np.maximum(col, 1e-3 * col.max()), 1e-300)
                out["fast_vs_strict_residual"] = {"max_difference_of_the_field_scale": float((np.abs(rf - rs).max(axis=0) / scale).max()),
                                                  "finite": bool(np.isfinite(rf).all() and np.isfinite(rs).all())}
            except Exception as ex:
                out["fast_vs_strict_residual"] = {"error": str(ex)[:200]}
        s.close()
        if leg == "inviscid_same_mesh":
            rec[leg] = {k: out[k] for k in ("value", "ms_per_step", "kernels")}
        else:
            rec.update(out)
    return rec


VISCOUS_MU = 1.0e-3
EXPERIMENT_START_BY_S = 430.0       # a 64 M-cell viscous record costs ~4 min (mesh, preprocessing, run): do not start it later than this


def viscous_strong_child(a, rank, world, local_rank, peak, peak_src):
    """Body of the per-rank CHILD process (bench.py --strong-child --child-task viscous_strong under RANK / WORLD_SIZE > 1): its own
    process group, the largest strong-scaling mesh that fits `world` GPUs, viscous."""
    import datetime
    import torch
    import torch.distributed as dist
    import bench
    import mallard_b200 as mb
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    mb.set_host_threads(max(1, bench.host_cores() // world))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=1800))
    fits = [(n, q) for n, q in STRONG_MESHES if 2 * q * q / world <= MAX_CELLS_PER_GPU]
    name, nq = fits[-1]
    rec = strong_record(name, nq, a, rank, world, local_rank, peak, peak_src, mu=VISCOUS_MU)
    rec["verification"] = ("viscous kernels and the second ghost ring of viscous contexts were written after the round-2 GPU budget was spent; validated by "
                           "the host emulation of their source (tests/test_kernel_emulation.py); first measured here, in a process of its own")
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    return rec


def graph_strong_child(a, rank, world, local_rank, peak, peak_src):
    """Body of the per-rank child for task `graph_strong`: the 16 M-cell strong-scaling mesh (BASELINE configs[3]) cut by the library's
    GRAPH partitioner (north_star: "the mesh is graph-partitioned across the GPUs") instead of the coordinate bisection - same mesh,
    same numerics, same timing as the `strong` record next to it, so the two partitioners are compared on the device (ms per step,
    ghost cells per stage, peers)."""
    import datetime
    import torch
    import torch.distributed as dist
    import bench
    import mallard_b200 as mb
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    mb.set_host_threads(max(1, bench.host_cores() // world))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=1800))
    name, nq = STRONG_MESHES[0]
    rec = strong_record(name, nq, a, rank, world, local_rank, peak, peak_src, graph=True)
    rec["verification"] = ("the graph partitioner is host code (CPU suite: sizes, determinism, connectivity, plans and emulated rank contexts on graph partitions); "
                           "this is the first run of a graph-partitioned mesh on hardware, in a process of its own")
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    return rec


def weak_problem(a, rank, world):
    """The main line's configuration: every rank owns one nx x ny block of a (nx * world) x ny cartesian_tri mesh, cut in x."""
    import bench
    import mallard_b200 as mb
    gnx = a.nx * world
    Lx = gnx / float(a.ny)
    mesh = mb.Mesh.generate("cartesian_tri", gnx, a.ny, Lx, 1.0)
    xy = mesh.arrays["cell_coords"]
    part = np.minimum((xy[:, 0] * (world / Lx)).astype(np.int32), world - 1)
    U0, P0 = bench.riemann2d_state(np.stack([xy[:, 0] / Lx, xy[:, 1]], 1))
    return mesh, part, U0, P0


def native_weak_child(a, rank, world, local_rank, peak, peak_src):
    """Body of the per-rank child for task `native_weak`: the weak-scaling main line through the library's own NCCL driver (the path whose
    first 8-GPU run hung, profiles/r02d_8gpu_native_hang.txt), with MLB_COMM_TRACE=1 so that a run that stops says where."""
    import datetime
    import torch
    import torch.distributed as dist
    import bench
    import mallard_b200 as mb
    from mallard_b200.parallel import DistributedSolver
    os.environ["MLB_COMM_TRACE"] = "1"
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    mb.set_host_threads(max(1, bench.host_cores() // world))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=1800))
    mesh, part, U0, P0 = weak_problem(a, rank, world)
    ds = DistributedSolver(mesh, part, rank, world, local_rank, native=True, recon="TENO", riemann="HLLC", integrator="SSPRK3", order=3, bcs=bench.SYM4,
                           fp_mode=a.fp, keep_stage_rhs=False)
    ds.set_state(U0, P0)
    run_ = Run(ds.s, ds, world)
    run_.timed(a.warmup, 0.1)
    ms = run_.timed(a.steps, 0.1)
    rec = {"workload": "native_weak: the main line's configuration (cartesian_tri %dx%d per GPU, %d GPUs) through mlb_comm_init / mlb_run_distributed: NCCL inside "
                       "the library, the step replayed as a CUDA graph" % (a.nx, a.ny, world), "n_cells": mesh.n_cells, "n_gpus": world,
           "value": mesh.n_cells * bench.N_STAGES * a.steps / (ms * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms / a.steps,
           "graph_replayed_steps": int(ds.s.get("stats")[11]),
           "nccl_settings": {k: os.environ.get(k) for k in ("NCCL_GRAPH_REGISTER", "NCCL_NVLS_ENABLE") if os.environ.get(k) is not None}}
    ds.s.close()
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    return rec


def experiment_children(a, rank, world, popen=None, task="viscous_strong", start_by=None, time_limit=None, extra_env=None):
    """N > 1: one CHILD process per rank, with a process group of their own, for code that runs on hardware for the first time - it must not
    be able to take the main line (or the measured strong-scaling records) with it.  Tasks: `viscous_strong` = BASELINE configs[4] (viscous,
    the 64 M-cell mesh where it fits); `graph_strong` = the 16 M-cell mesh cut by the graph partitioner; `native_weak` = the main line's configuration through the library's own NCCL driver
    (mlb_comm_init / mlb_run_distributed) with MLB_COMM_TRACE=1.  Every rank decides the same way (collective), starts its child, waits for
    it under the time limit; rank 0 returns the records its child printed; the last `[mlb comm]` lines of EVERY rank's child are attached
    to a record that did not complete (which rank stopped where)."""
    import signal
    import subprocess
    import bench
    elapsed = float(_reduce([time.perf_counter() - T_START], world, "max")[0])
    _STATE["children_ok"] = None                # (the same on every rank: decided collectively below)
    if elapsed > (EXPERIMENT_START_BY_S if start_by is None else start_by):
        return [{"workload": task, "skipped": "time budget of the bench run (%.0f s elapsed)" % elapsed}]
    left = DEADLINE_S - elapsed - 40.0          # the parent's own watchdog fires at DEADLINE_S: be done before it
    if time_limit is not None:
        left = min(left, time_limit)
    env = dict(os.environ, MLB_BENCH_ELAPSED="%.1f" % elapsed, MASTER_PORT=str(int(os.environ.get("MASTER_PORT", "29500")) + 23 + 2 * len(_STATE.setdefault("tasks", []))))
    _STATE["tasks"].append(task)                # (every group of children gets a port of its own)
    env.update(extra_env or {})
    for k in ("TORCHELASTIC_USE_AGENT_STORE", "TORCHELASTIC_RUN_ID", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_MAX_RESTARTS"):
        env.pop(k, None)                        # the children rendezvous among themselves (rank 0's child hosts the store), not through torchrun's agent
    cmd = [sys.executable, os.path.join(bench.ROOT, "bench.py"), "--strong-child", "--child-task", task, "--gpus", str(world), "--steps", str(a.steps),
           "--warmup", str(a.warmup), "--fp", a.fp, "--nx", str(a.nx), "--ny", str(a.ny)]
    recs, note, out, err, p = [], None, "", "", None
    try:
        p = (popen or subprocess.Popen)(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    except Exception as ex:                     # (no early return: the collective below needs every rank)
        note = {"workload": task, "error": "rank %d could not start its child process: %s" % (rank, str(ex)[:200])}
    _STATE["children"] = [p.pid] if p else []   # the watchdog takes them along should it fire
    try:
        if p:
            out, err = p.communicate(timeout=left)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except Exception:
            p.kill()
        try:
            out, err = p.communicate(timeout=30)
        except Exception:
            out, err = "", ""
        note = {"workload": task, "aborted": "no result within %.0f s (children killed)" % left}
    _STATE["children"] = []
    sys.stderr.write(err or "")                 # the children's log lines belong in this run's log
    for l in (out or "").splitlines():
        if l.startswith(bench.STRONG_TAG):
            try:
                recs.append(json.loads(l[len(bench.STRONG_TAG):]))
            except Exception:
                pass
    if note is None and p.returncode != 0:
        note = {"workload": task, "error": "the child process of rank %d ended with code %s" % (rank, p.returncode)}
    code = p.returncode if p else None
    # did every rank's child finish?  (collective over the PARENTS' group, which is intact whatever the children did)
    bad = float(_reduce([0.0 if (note is None and (recs or rank != 0)) else 1.0], world, "max")[0])
    _STATE["children_ok"] = bad == 0
    if bad > 0:
        import torch.distributed as dist
        trace = [l for l in (err or "").splitlines() if l.startswith("[mlb comm]")][-3:] or [l for l in (err or "").splitlines() if l.strip()][-2:]
        traces = [None] * world
        dist.all_gather_object(traces, {"rank": rank, "exit_code": code, "last_lines": trace})
        if note is None:
            note = {"workload": task, "error": "the child process of another rank did not finish"}
        note["children"] = traces
    if note is not None:
        recs.append(note)
    return recs


def run(a, rank, world, local_rank, workload):
    import torch
    import torch.distributed as dist
    import bench
    import mallard_b200 as mb
    from mallard_b200.parallel import DistributedSolver

    start_watchdog(rank)
    torch.cuda.set_device(local_rank)
    total_cores = bench.host_cores()
    numa = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation: first touch then lands next to the GPU
    # host preprocessing is OpenMP-parallel inside every rank: share the cores instead of oversubscribing them
    # (torchrun presets OMP_NUM_THREADS=1, which would serialise the TENO table construction)
    mb.set_host_threads(max(1, min(total_cores // world, bench.host_cores())))
    # (collective timeout above this file's own deadline: a rank left waiting by a failed strong record leaves through the watchdog
    #  with exit code 0, not through torch's NCCL watchdog with SIGABRT)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=1800))
    peak, peak_src = bench.hbm_peak()

    t_setup = time.perf_counter()
    strong = getattr(a, "scaling", "weak") == "strong" or a.workload == "vortex"
    local = None
    if a.workload == "vortex":
        # BASELINE configs[3]: ONE jittered, id-shuffled triangulation (a.nx x a.ny quads) split over the GPUs by the library's
        # recursive coordinate bisection - an irregular cut through an unstructured numbering; rank-local ingest
        from mallard_b200 import synthetic as syn
        lp = syn.jittered_tri_local(a.nx, a.ny, 10.0, 10.0, world, rank, seed=12345, layers=10)
        mesh, part, local, nc = lp.mesh, lp.part_local, lp.local, lp.n_global
        U0, P0, bcs = syn.isentropic_vortex(mesh.arrays["cell_coords"]), None, syn.EXTRAP4
        gnx = a.nx
    else:
        gnx = a.nx if strong else a.nx * world                                   # strong: the given mesh is split; weak: one block per GPU
        Lx = gnx / float(a.ny)                                                    # square cells
        mesh = mb.Mesh.generate("cartesian_tri", gnx, a.ny, Lx, 1.0)
        nc = mesh.n_cells
        xy = mesh.arrays["cell_coords"]
        part = np.minimum((xy[:, 0] * (world / Lx)).astype(np.int32), world - 1)  # rank r owns the strip x in [r, r+1) Lx / world
        U0, P0 = bench.riemann2d_state(np.stack([xy[:, 0] / Lx, xy[:, 1]], 1))    # the four-quadrant IC stretched over the strip
        bcs = bench.SYM4
    visc = dict(gas=dict(mu=a.mu), teno_fixed=True) if (getattr(a, "mu", 0.0) > 0 and a.workload == "vortex") else {}
    ds = DistributedSolver(mesh, part, rank, world, local_rank, local=local, native=NATIVE, recon=a.recon, riemann="HLLC", integrator="SSPRK3",
                           order=3, bcs=bcs, fp_mode=a.fp, keep_stage_rhs=False, **visc)
    s = ds.s
    stats = s.get("stats")
    n_owned = int(stats[4])
    setup_s = time.perf_counter() - t_setup
    ds.set_state(U0, P0)
    owned_ids = ds.owned

    run_ = Run(s, ds, world)
    run_.timed(a.warmup, 0.1)
    launches0 = s.launch_count
    clocks = bench.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms = run_.timed(a.steps, 0.1)
    clk = clocks.stop() if rank == 0 else None
    launches = s.launch_count - launches0
    value = nc * bench.N_STAGES * a.steps / (ms * 1e-3)
    replays = int(s.get("stats")[11])
    ds.set_state(U0, P0)
    roof = None
    if a.recon == "TENO":       # (per-kernel events under the split-phase driver: never the reason for a missing line - every rank takes the same branch,
        try:                    #  the reductions inside are the only collectives and come after the last call that can throw)
            roof = run_.recon_roofline(max(2, min(a.steps, 5)), 0.1, bench.N_STAGES, peak, peak_src, bench.ALG_BYTES_RECON)
        except mb.MallardError as ex:
            raise               # a device-side failure: nothing after it can be trusted
        except Exception as ex:
            roof = {"error": str(ex)[:200]}

    # ---- end to end: per step H2D of the rank's own cells from pinned memory, the step (halo + all-reduce), D2H
    e2e = None
    if not a.no_e2e:
        pin = torch.empty((n_owned, 4), dtype=torch.float64, pin_memory=True)
        Uh = pin.numpy()
        Uh[:] = U0[owned_ids]
        k_e2e = max(3, min(a.steps, 10))
        for _ in range(2):
            ds.step_host(Uh, 0.1)
        Uh[:] = U0[owned_ids]
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            ds.step_host(Uh, 0.1)
        dist.barrier()
        sec = float(_reduce([time.perf_counter() - t0], world)[0])
        e2e = {"value": nc * bench.N_STAGES * k_e2e / sec, "unit": "cell-updates/s", "h2d_bytes_per_step": nc * 32, "d2h_bytes_per_step": nc * 32,
               "steps": k_e2e, "ms_per_step": 1e3 * sec / k_e2e,
               "api": ("mlb_take_step_distributed_host (C ABI: pinned host buffers of the rank's own cells -> device, the partitioned step incl. "
                       "NCCL halo exchange and dt all-reduce inside the library, device -> host)") if NATIVE else
                      "DistributedSolver.step_host (mlb_set_owned / split-phase stage API / mlb_get_owned; pinned host buffers of the rank's own cells)"}

    peers, sc, rc = ds.peers, ds.send_counts, ds.recv_counts
    halo = _reduce([float(sc.sum()), float(rc.sum()), float(len(peers))], world)
    s.close()
    line = None
    if rank == 0:
        line = {"metric": "cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload + ("; partitioned by recursive coordinate bisection over %d GPUs, rank-local ingest" % world if a.workload == "vortex" else
                                                  ("; the mesh is" if strong else " per GPU; global mesh %dx%d" % (gnx, a.ny)) + " partitioned in x over %d GPUs" % world),
                           "n_cells": nc, "cells_per_gpu": n_owned, "fp_mode": a.fp, "recon": a.recon,
                           "l2": "inputs larger than L2 (TENO tables %.1f GB per GPU per stage)" % (stats[2] / 1e9),
                           "driver": ("native (mlb_comm_init / mlb_run_distributed): NCCL inside the library, %d of %d timed steps replayed as a CUDA graph"
                                      % (min(replays, a.steps), a.steps)) if NATIVE else
                                     "split-phase C ABI (mlb_halo_pack / mlb_stage_begin / mlb_stage) over torch.distributed (NCCL)",
                           "halo": {"max_send_cells_per_stage": int(halo[0]), "max_recv_cells_per_stage": int(halo[1]),
                                    "max_peers": int(halo[2]), "transport": TRANSPORT},
                           "setup_seconds": setup_s, "cpus_bound_per_rank": numa},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": None, "strong": None}
    _STATE["line"] = line              # from here on the watchdog can deliver the main line on its own
    strong_recs = strong_records(a, rank, world, local_rank, peak, peak_src) if (a.workload == "riemann_2d" and not a.no_strong) else None
    experiments = None
    if strong_recs is not None and not any("error" in r for r in strong_recs):      # (after an error the ranks may have diverged: nothing collective)
        try:
            experiments = experiment_children(a, rank, world)
            # the 16 M-cell mesh once more, cut by the graph partitioner (~1 min of partitioning per rank on top of a strong record)
            experiments += experiment_children(a, rank, world, task="graph_strong", start_by=470.0)
            # the library's own NCCL driver on this many GPUs (diagnosis of the round-2 8-GPU hang): NCCL's defaults first; if that does not
            # come back, once more without graph-time buffer registration and NVLS (neither can matter for 230 kB exchanges and an 8-byte
            # all-reduce) - the record says which attempt, if any, completed
            limit = float(os.environ.get("MLB_NATIVE_TRIAL_LIMIT", "100"))       # global mesh + context + NCCL set-up take ~45 s of it
            nat = experiment_children(a, rank, world, task="native_weak", start_by=540.0, time_limit=limit)
            experiments += nat
            if _STATE.get("children_ok") is False:
                experiments += experiment_children(a, rank, world, task="native_weak", start_by=590.0, time_limit=limit,
                                                   extra_env={"NCCL_GRAPH_REGISTER": "0", "NCCL_NVLS_ENABLE": "0"})
        except Exception as ex:
            experiments = (experiments or []) + [{"workload": "experiments", "error": str(ex)[:300]}]
    _STATE["done"] = True
    if rank == 0:
        line["strong"] = strong_recs
        line["experiments"] = experiments
        print(bench.json_line(line), flush=True)
    # the line is out: nothing below may keep the job alive (a peer that failed inside a strong record never reaches the barrier)
    threading.Thread(target=lambda: (time.sleep(float(os.environ.get("MLB_BENCH_EXIT_GRACE", "30"))), os._exit(0)), daemon=True).start()
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
