#!/usr/bin/env python
"""Host-only comparison of the two partitioners behind the C ABI (mlb_partition: recursive coordinate bisection on centroids;
mlb_partition_graph: multilevel recursive bisection of the cell-face dual graph, csrc/partition_graph.cpp) on the meshes the
strong-scaling records use, scaled to what the preprocessor handles in a minute on a few host cores: edge cut, and - what a stage
really moves - the ghost cells each rank receives per stage (TENO p = 3 stencils, from the rank's own plan), its peer count, the balance.
No GPU involved.  Usage: python scripts/partition_compare.py [n_quads_per_side] [n_parts] > profiles/r02f_partition_graph_vs_rcb.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallard_b200 as mb  # noqa: E402
from mallard_b200 import synthetic as syn  # noqa: E402


def edge_cut(mesh, part):
    cof = mesh.arrays["cells_of_face"].reshape(-1, 2)
    m = (cof[:, 0] >= 0) & (cof[:, 1] >= 0)
    return int((part[cof[m, 0]] != part[cof[m, 1]]).sum())


def halo(mesh, part, n_parts, bcs):
    recv, peers = [], []
    for r in range(n_parts):
        plan = mb.Plan(mesh, "TENO", order=3, bcs=bcs, fp_mode="fast", part=part, rank=r, n_ranks=n_parts)
        rc = plan.get("halo_recv_counts")
        recv.append(int(rc.sum()))
        peers.append(int((rc > 0).sum()))
        plan.close() if hasattr(plan, "close") else None
    return recv, peers


def row(name, mesh, n_parts, bcs, with_halo=True):
    out = []
    for label, fn in (("coordinate bisection", mb.partition), ("graph (multilevel)", mb.partition_graph)):
        t = time.perf_counter()
        part = fn(mesh, n_parts)
        sec = time.perf_counter() - t
        sizes = np.bincount(part, minlength=n_parts)
        line = "%-34s %-22s cells %9d  parts %2d  sizes %d..%d  edge cut %7d  %.2f s" % (name, label, mesh.n_cells, n_parts, sizes.min(), sizes.max(), edge_cut(mesh, part), sec)
        if with_halo:
            recv, peers = halo(mesh, part, n_parts, bcs)
            line += "  ghost cells received per stage: max %d, total %d; peers: max %d" % (max(recv), sum(recv), max(peers))
        out.append(line)
        print(line, flush=True)
    return out


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 362
    n_parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    mb.set_host_threads(os.cpu_count() or 1)
    print("# partitioners compared on the host (%d threads); TENO p = 3 stencil halos from mlb_plan_create per rank" % (os.cpu_count() or 1))
    sym = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]
    row("jittered, id-shuffled %dx%d" % (nq, nq), syn.jittered_tri(nq, nq, 10.0, 10.0, seed=12345), n_parts, syn.EXTRAP4)
    row("jittered, id-shuffled %dx%d" % (nq, nq), syn.jittered_tri(nq, nq, 10.0, 10.0, seed=12345), 3, syn.EXTRAP4)
    row("jittered, id-shuffled %dx%d" % (nq, nq), syn.jittered_tri(nq, nq, 10.0, 10.0, seed=12345), 6, syn.EXTRAP4)
    row("cartesian_tri strip %dx%d" % (4 * nq, nq // 2), mb.Mesh.generate("cartesian_tri", 4 * nq, nq // 2, 8.0, 1.0), n_parts, sym)
    row("examples/wedge 150x50", mb.Mesh.generate("wedge", 150, 50, 4.0, 1.5), n_parts,
        [dict(name="left", type="upt"), dict(name="right", type="p_out"), dict(name="top", type="symmetry"), dict(name="bottom", type="symmetry")], with_halo=False)
    big = syn.jittered_tri(1024, 1024, 10.0, 10.0, seed=12345)
    row("jittered, id-shuffled 1024x1024", big, 8, syn.EXTRAP4, with_halo=False)


if __name__ == "__main__":
    main()
