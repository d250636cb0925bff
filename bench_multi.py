"""bench.py's N>1 leg (launched by torchrun, one rank per GPU).

Weak scaling of BASELINE.json configs[1]: every rank owns one nx x ny block (2*nx*ny triangles) of a (nx*N) x ny
cartesian_tri mesh; the global mesh is partitioned in x by cell centroid.  Per RK stage the ghost-cell conserved states are
exchanged peer to peer with NCCL send/recv straight between the library's device buffers, per step one double is
all-reduced (max) for dt (mallard_b200/parallel.py).  Timing: barrier + device synchronise on both sides, CUDA events on
the library's compute stream, MAX over ranks.
"""
import json
import os
import time

import numpy as np


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


def run(a, rank, world, local_rank, workload):
    import torch
    import torch.distributed as dist
    import bench
    import mallard_b200 as mb
    from mallard_b200.parallel import DistributedSolver

    torch.cuda.set_device(local_rank)
    total_cores = bench.host_cores()
    numa = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation: first touch then lands next to the GPU
    # host preprocessing is OpenMP-parallel inside every rank: share the cores instead of oversubscribing them
    # (torchrun presets OMP_NUM_THREADS=1, which would serialise the TENO table construction)
    mb.set_host_threads(max(1, min(total_cores // world, bench.host_cores())))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t_setup = time.perf_counter()
    strong = getattr(a, "scaling", "weak") == "strong" or a.workload == "vortex"
    if a.workload == "vortex":
        # BASELINE configs[3]: ONE jittered, id-shuffled triangulation (a.nx x a.ny quads) split over the GPUs by the library's
        # recursive coordinate bisection (mlb_partition) - an irregular cut through an unstructured numbering
        from mallard_b200 import synthetic as syn
        mesh = syn.jittered_tri(a.nx, a.ny, 10.0, 10.0, seed=12345)
        nc = mesh.n_cells
        part = mb.partition(mesh, world)
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
    ds = DistributedSolver(mesh, part, rank, world, local_rank, recon=a.recon, riemann="HLLC", integrator="SSPRK3", order=3,
                           bcs=bcs, fp_mode=a.fp, keep_stage_rhs=False)
    s = ds.s
    stats = s.get("stats")
    n_owned = int(stats[4])
    setup_s = time.perf_counter() - t_setup
    ds.set_state(U0, P0)

    def timed(n_steps):
        dist.barrier()
        s.synchronize()
        s.event_record(0)
        for _ in range(n_steps):
            ds.step(0.1)
        s.event_record(1)
        ms = s.event_elapsed_ms(0, 1)
        s.synchronize()
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    timed(a.warmup)
    launches0 = s.launch_count
    clocks = bench.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms = timed(a.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = s.launch_count - launches0
    value = nc * bench.N_STAGES * a.steps / (ms * 1e-3)

    # ---- end to end: per step H2D of the rank's own cells from pinned memory, the step (halo + all-reduce), D2H
    e2e = None
    if not a.no_e2e:
        pin = torch.empty((n_owned, 4), dtype=torch.float64, pin_memory=True)
        Uh = pin.numpy()
        Uh[:] = U0[ds.owned]
        k_e2e = max(3, min(a.steps, 10))
        for _ in range(2):
            ds.step_host(Uh, 0.1)
        Uh[:] = U0[ds.owned]
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            ds.step_host(Uh, 0.1)
        dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
        e2e = {"value": nc * bench.N_STAGES * k_e2e / sec, "unit": "cell-updates/s", "h2d_bytes_per_step": nc * 32, "d2h_bytes_per_step": nc * 32,
               "steps": k_e2e, "ms_per_step": 1e3 * sec / k_e2e,
               "api": "DistributedSolver.step_host (mlb_set_owned / split-phase stage API / mlb_get_owned; host buffers of the rank's own cells)"}

    peers, sc, rc = ds.peers, ds.send_counts, ds.recv_counts
    halo = torch.tensor([float(sc.sum()), float(rc.sum()), float(len(peers))], dtype=torch.float64, device="cuda")
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"metric": "cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload + ("; partitioned by recursive coordinate bisection over %d GPUs" % world if a.workload == "vortex" else
                                                  ("; the mesh is" if strong else " per GPU; global mesh %dx%d" % (gnx, a.ny)) + " partitioned in x over %d GPUs" % world),
                           "n_cells": nc, "cells_per_gpu": n_owned, "fp_mode": a.fp, "recon": a.recon,
                           "l2": "inputs larger than L2 (TENO tables %.1f GB per GPU per stage)" % (stats[2] / 1e9),
                           "halo": {"max_send_cells_per_stage": int(halo[0].item()), "max_recv_cells_per_stage": int(halo[1].item()),
                                    "max_peers": int(halo[2].item()), "transport": "NCCL send/recv between device buffers + all_reduce(max) of dt"},
                           "setup_seconds": setup_s, "cpus_bound_per_rank": numa},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
