"""CPU-side tests of the product's host code: C-ABI surface, host mesh generators, mesh preprocessor.
No compute entry point is called here (no GPU in this suite); parity of the host-built tables is checked against the
oracle and against the fixtures generated from the real reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_util as gu
import mallard_b200 as mb
from mallard_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mallard_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mlb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(_abi.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    assert b"sm_100a" in mb.lib().mlb_version()


def test_abi_revision_is_checked_on_both_sides_of_the_boundary(tmp_path):
    """Every struct crosses the boundary by pointer: a host built against an older header would pass structs of the wrong size.  The
    header carries a revision (MLB_ABI_VERSION), the ctypes mirror and the drop-in harness present theirs to mlb_check_abi, and a binary
    that predates the header it embeds is a build error (this is how a stale oracle/_ref/bin/mallard_dropin was caught in round 2)."""
    import re
    import subprocess
    from mallard_b200 import _abi
    header = open(os.path.join(ROOT, "include", "mallard_b200.h")).read()
    assert int(re.search(r"#define MLB_ABI_VERSION (\d+)", header).group(1)) == _abi.ABI_VERSION
    L = mb.lib()
    assert L.mlb_check_abi(_abi.ABI_VERSION) == 0
    assert L.mlb_check_abi(_abi.ABI_VERSION - 1) != 0 and "ABI mismatch" in L.mlb_last_error(None).decode()
    dropin = os.path.join(ROOT, "oracle", "_ref", "bin", "mallard_dropin")
    if os.path.exists(dropin):
        import __graft_entry__ as ge
        assert not ge.reference_binaries_stale(os.path.dirname(dropin)), "oracle/_ref/bin is older than the header / harness it was built from: run build()"
        toml = tmp_path / "input.toml"
        toml.write_text('[run]\nn_steps = 1\ncfl = 1.0\n[mesh]\ntype = "cartesian"\nNx = 8\nNy = 1\nLx = 1.0\nLy = 0.1\n[initialize]\ntype = "constant"\n'
                        'u = [0.0, 0.0]\np = 1.0\nT = 300.0\n[[boundaries]]\nname = "left"\ntype = "symmetry"\n[[boundaries]]\nname = "right"\ntype = "symmetry"\n'
                        '[[boundaries]]\nname = "top"\ntype = "symmetry"\n[[boundaries]]\nname = "bottom"\ntype = "symmetry"\n[numerics]\nriemann_solver = "HLLC"\n'
                        'time_integrator = "SSPRK3"\ncheck_nan = false\n[numerics.face_reconstruction]\ntype = "FO"\n[physics]\ntype = "euler"\ngamma = 1.4\n'
                        'p_ref = 101325.0\nT_ref = 298.15\nrho_ref = 1.225\n[output]\ncheck_interval = 1000000\n')
        import torch
        if not torch.cuda.is_available():      # the harness gets past the ABI check and then fails where it must: no device, no fallback
            p = subprocess.run([dropin, "-i", str(toml), "--quiet"], capture_output=True, text=True, cwd=tmp_path, env=dict(os.environ, OMP_PROC_BIND="false"))
            assert p.returncode != 0 and "ABI mismatch" not in p.stderr and "no CPU fallback" in p.stderr, p.stderr[-600:]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path needs a box without CUDA")
    m = mb.Mesh.generate("cartesian", 4, 2, 1.0, 0.5)
    with pytest.raises(mb.MallardError, match="no CPU fallback"):
        mb.Solver(m, "FO", "HLLC", "SSPRK3", bcs=[dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")])
    with pytest.raises(mb.MallardError, match="no CPU fallback"):
        mb.riemann_flux("HLLC", [(1.0, 0.0)], [[1, 0, 0, 1, 3.5]], [[1, 0, 0, 1, 3.5]])


def test_interface_errors_mirror_reference():
    m = mb.Mesh.generate("cartesian", 4, 2, 1.0, 0.5)
    with pytest.raises(mb.MallardError, match="Unknown mesh type"):       # mesh/mesh.cpp:35-36
        mb.Mesh.generate("file", 2, 2)
    with pytest.raises(mb.MallardError, match="Unknown Riemann solver type"):   # solver/solver.cpp:130-132
        mb.Solver(m, "FO", "Roe", "SSPRK3")
    with pytest.raises(mb.MallardError, match="Unknown time integrator type"):  # default "LSSSPRK3" is not in the map
        mb.Solver(m, "FO", "HLLC", "LSSSPRK3")
    with pytest.raises(mb.MallardError, match="not found in mesh"):             # solver/solver.cpp:211-213
        mb.Solver(m, "FO", "HLLC", "SSPRK3", bcs=[dict(name="inlet", type="symmetry")])
    with pytest.raises(mb.MallardError, match="Missing p for boundary"):        # boundary_p_out.cpp:38-40
        mb.Solver(m, "FO", "HLLC", "SSPRK3", bcs=[dict(name="right", type="p_out")])
    # where the reference throws "TENO has only been implemented for triangular cells" (face_reconstruction.cpp:485-487), quadrilaterals
    # are supported here: one central + four directional stencils per cell
    quads = mb.Plan(mb.Mesh.generate("cartesian", 8, 8), "TENO", order=2)
    assert (quads.S, quads.n_slots, quads.M) == (5, 4, 12)


@pytest.mark.parametrize("mtype,nx,ny,Lx,Ly", [("cartesian", 7, 5, 2.0, 1.0), ("cartesian", 1000, 1, 1.0, 0.001),
                                               ("cartesian_tri", 6, 9, 1.0, 1.5), ("cartesian_tri", 1, 1, 1.0, 1.0),
                                               ("wedge", 30, 10, 4.0, 1.5), ("wedge", 150, 50, 4.0, 1.5)])
def test_host_mesh_generators_bit_exact(oracle_mod, mtype, nx, ny, Lx, Ly):
    ref = oracle_mod.Mesh.generate(mtype, nx, ny, Lx, Ly)
    got = mb.Mesh.generate(mtype, nx, ny, Lx, Ly)
    for k, v in ref.arrays().items():
        a = got.arrays[k].reshape(v.shape)
        if k == "face_normals":
            assert np.array_equal(np.isnan(a), np.isnan(v))
            a, v = np.nan_to_num(a), np.nan_to_num(v)
        assert np.array_equal(a, v), k
    assert [n for n, _ in got.zones] == gu.ZONES
    for n, f in got.zones:
        assert np.array_equal(f, ref.zone(n)), n


@pytest.mark.parametrize("name", [n for n in gu.names() if n.startswith("teno")])
def test_teno_tables_bit_exact_against_reference(name):
    meta, g = gu.load(name)
    mm = meta["mesh"]
    mesh = mb.Mesh.generate(mm["type"], mm["Nx"], mm["Ny"], mm["Lx"], mm["Ly"])
    r = meta["recon"]
    for renumber in ("rcm", "none"):
        plan = mb.Plan(mesh, "TENO", basis=r["basis_type"], order=r["basis_order"], factor=r["max_stencil_size_factor"],
                       quad_cell_order=r.get("quadrature_order_cell", 0), quad_face_order=r.get("quadrature_order_face", 0),
                       bcs=meta["bcs"], renumber=renumber)
        for k in g:
            if k.startswith("teno:") and k not in ("teno:meta", "teno:quad_cell_points", "teno:quad_cell_weights", "teno:quad_face_points"):
                got = plan.get(k)
                assert np.array_equal(got.reshape(g[k].shape), g[k]), (k, renumber)


def test_teno_tables_bit_exact_against_oracle_larger_mesh(oracle_mod):
    om = oracle_mod.Mesh.generate("cartesian_tri", 14, 11, 1.0, 0.8)
    osol = oracle_mod.Solver(om, "TENO", order=3, bcs=[dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")])
    mesh = mb.Mesh.generate("cartesian_tri", 14, 11, 1.0, 0.8)
    plan = mb.Plan(mesh, "TENO", order=3)
    for k in ("teno:offsets_stencil_groups", "teno:offsets_stencils", "teno:stencils", "teno:offsets_reconstruction_matrices",
              "teno:reconstruction_matrices", "teno:transformed_areas", "teno:integral_psi_target", "teno:oscillation_indicator"):
        assert np.array_equal(plan.get(k), osol.get(k).reshape(-1)), k


@pytest.mark.parametrize("mtype,nx,ny", [("cartesian", 40, 30), ("cartesian_tri", 25, 20), ("wedge", 30, 10)])
def test_renumbering_is_a_valid_deterministic_permutation(mtype, nx, ny):
    mesh = mb.Mesh.generate(mtype, nx, ny, 2.0, 1.0)
    p1 = mb.Plan(mesh, "FO", renumber="rcm")
    p2 = mb.Plan(mesh, "FO", renumber="rcm")
    perm = p1.get("perm_cells")
    assert np.array_equal(perm, p2.get("perm_cells"))
    assert np.array_equal(np.sort(perm), np.arange(mesh.n_cells))
    # faces: every real face exactly once, ordered by owner (lower library cell id)
    pf = p1.get("perm_faces")
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    assert np.array_equal(np.sort(pf), np.nonzero(real)[0])
    inv = np.empty(mesh.n_cells, dtype=np.int64); inv[perm] = np.arange(mesh.n_cells)
    cof = mesh.arrays["cells_of_face"][pf]
    a = inv[cof[:, 0]]; b = np.where(cof[:, 1] >= 0, inv[np.maximum(cof[:, 1], 0)], np.iinfo(np.int64).max)
    owner = np.minimum(a, b)
    assert np.all(np.diff(owner) >= 0)
    # RCM does not blow up the bandwidth of the dual graph compared with the generator's natural order
    inter = cof[:, 1] >= 0
    bw_rcm = np.abs(a[inter] - b[inter]).max()
    bw_nat = np.abs(cof[inter, 0] - cof[inter, 1]).max()
    assert bw_rcm <= 2 * bw_nat + 2
    ident = mb.Plan(mesh, "FO", renumber="none").get("perm_cells")
    assert np.array_equal(ident, np.arange(mesh.n_cells))


def test_rhs_accumulation_order_matches_reference_zone_order():
    """Per cell, slots must be visited interior-zone order first, then boundaries in input order."""
    mesh = mb.Mesh.generate("cartesian", 6, 4, 1.0, 1.0)
    bcs = [dict(name=n, type="symmetry") for n in ("bottom", "left", "top", "right")]   # deliberately not the zone order
    plan = mb.Plan(mesh, "FO", bcs=bcs, renumber="rcm")
    perm, pf = plan.get("perm_cells"), plan.get("perm_faces")
    sf = plan.get("slot_face").reshape(plan.n_slots, plan.Npad)
    order = plan.get("rhs_order")
    key = {}
    for k, f in enumerate(mesh.get_face_zone("interior")):
        key[int(f)] = (0, k)
    for b, bc in enumerate(bcs):
        for k, f in enumerate(mesh.get_face_zone(bc["name"])):
            key[int(f)] = (1 + b, k)
    for i in range(plan.N):
        visited = [int(pf[sf[(order[i] >> (2 * j)) & 3, i] & 0x7FFFFFFF]) for j in range(4)]
        keys = [key[f] for f in visited]
        assert keys == sorted(keys), (i, visited)
        foc = mesh.arrays["faces_of_cell"][4 * perm[i]:4 * perm[i] + 4]
        assert sorted(visited) == sorted(int(x) for x in foc)


def test_partition_is_balanced_and_deterministic():
    mesh = mb.Mesh.generate("cartesian_tri", 20, 16, 2.0, 1.0)
    for n in (2, 3, 4, 8):
        p1, p2 = mb.partition(mesh, n), mb.partition(mesh, n)
        assert np.array_equal(p1, p2)
        counts = np.bincount(p1, minlength=n)
        assert counts.min() > 0 and counts.max() - counts.min() <= n
    assert np.all(mb.partition(mesh, 1) == 0)


def _edge_cut(mesh, part):
    cof = mesh.arrays["cells_of_face"].reshape(-1, 2)
    m = (cof[:, 0] >= 0) & (cof[:, 1] >= 0)
    return int((part[cof[m, 0]] != part[cof[m, 1]]).sum())


def _n_components(n, pairs):
    """connected components of an undirected graph given as an [m][2] edge list (union-find)"""
    parent = np.arange(n)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in pairs:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
    return len({find(i) for i in range(n)})


def test_graph_partition_is_balanced_deterministic_and_cuts_about_as_few_faces_as_the_coordinate_bisection():
    """mlb_partition_graph (csrc/partition_graph.cpp; north_star: "graph-partitioned"): multilevel recursive bisection of the cell-face dual
    graph.  The part sizes are exactly the coordinate bisection's, the result does not depend on the number of host threads, every part of
    these (convex) meshes is connected, and the cut is within 20 % of the straight cuts coordinate bisection makes on them - and shorter
    where a part count that is no power of two makes coordinate bisection cut 1 : 2."""
    from mallard_b200 import synthetic as syn
    meshes = [mb.Mesh.generate("cartesian_tri", 48, 32, 3.0, 2.0), syn.jittered_tri(40, 40, 10.0, 10.0, seed=7), mb.Mesh.generate("wedge", 60, 20, 4.0, 1.5)]
    for mesh in meshes:
        cof = mesh.arrays["cells_of_face"].reshape(-1, 2)
        inner = cof[(cof[:, 0] >= 0) & (cof[:, 1] >= 0)]
        for n in (2, 3, 4, 8):
            mb.set_host_threads(4)
            pg = mb.partition_graph(mesh, n)
            mb.set_host_threads(1)
            assert np.array_equal(pg, mb.partition_graph(mesh, n))
            pr = mb.partition(mesh, n)
            assert np.array_equal(np.bincount(pg, minlength=n), np.bincount(pr, minlength=n))
            assert _edge_cut(mesh, pg) <= 1.2 * _edge_cut(mesh, pr) + 4
            for r in range(n):
                ids = np.flatnonzero(pg == r)
                loc = -np.ones(mesh.n_cells, dtype=np.int64)
                loc[ids] = np.arange(len(ids))
                e = inner[(pg[inner[:, 0]] == r) & (pg[inner[:, 1]] == r)]
                assert _n_components(len(ids), loc[e]) == 1
        assert _edge_cut(mesh, mb.partition_graph(mesh, 3)) <= _edge_cut(mesh, mb.partition(mesh, 3))
    mb.set_host_threads(os.cpu_count() or 1)
    assert np.all(mb.partition_graph(meshes[0], 1) == 0)


def test_graph_partition_follows_the_connectivity_not_the_coordinates():
    """A domain folded like a U (two 40 x 10 strips of quadrilaterals joined at one end, lying on top of each other in the coordinates a
    bisection would look at): the graph partitioner, given the CSR graph alone, cuts the 80 x 10 channel across - 10 edges per cut -
    wherever it is folded; degenerate inputs (no edges, isolated vertices, more parts than vertices) keep the size contract."""
    nx, ny = 80, 10
    idx = np.arange(nx * ny).reshape(nx, ny)
    pairs = np.concatenate([np.stack([idx[:-1].ravel(), idx[1:].ravel()], 1), np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1)])
    both = np.concatenate([pairs, pairs[:, ::-1]])
    both = both[np.lexsort((both[:, 1], both[:, 0]))]
    xadj = np.concatenate([[0], np.cumsum(np.bincount(both[:, 0], minlength=nx * ny))]).astype(np.uint64)
    for n in (2, 4):
        part = mb.partition_graph(None, n, xadj=xadj, adj=both[:, 1])
        assert np.bincount(part, minlength=n).tolist() == [nx * ny // n] * n
        assert int((part[pairs[:, 0]] != part[pairs[:, 1]]).sum()) == ny * (n - 1)
    # no edges at all / more parts than vertices: sizes as the coordinate bisection would make them
    lonely = mb.partition_graph(None, 3, xadj=np.zeros(8, dtype=np.uint64), adj=np.zeros(0, dtype=np.uint32))
    assert sorted(np.bincount(lonely, minlength=3).tolist()) == [2, 2, 3]
    tiny = mb.partition_graph(None, 4, xadj=np.array([0, 1, 2], dtype=np.uint64), adj=np.array([1, 0], dtype=np.uint32))
    assert len(tiny) == 2 and set(tiny.tolist()) <= {0, 1, 2, 3}
    with pytest.raises(mb.MallardError):
        mb.partition_graph(None, 2, xadj=np.array([0, 1], dtype=np.uint64), adj=np.array([5], dtype=np.uint32))


def test_closed_form_dual_graph_of_the_jittered_mesh_is_the_mesh_s_own():
    """synthetic.jittered_tri_dual_graph (what a rank that never holds the 16 M-cell mesh hands mlb_partition_graph_csr): the same edges as the
    generated mesh's cells_of_face, hence the same partition as mlb_partition_graph of the mesh"""
    from mallard_b200 import synthetic as syn
    for nx, ny, seed in ((13, 9, 5), (8, 20, 12345)):
        m = syn.jittered_tri(nx, ny, 10.0, 10.0, seed=seed)
        xadj, adj = syn.jittered_tri_dual_graph(nx, ny, seed=seed)
        cof = m.arrays["cells_of_face"].reshape(-1, 2)
        inner = cof[(cof[:, 0] >= 0) & (cof[:, 1] >= 0)]
        rows = np.repeat(np.arange(m.n_cells), np.diff(xadj).astype(np.int64))
        assert set(zip(rows.tolist(), adj.tolist())) == set(map(tuple, np.concatenate([inner, inner[:, ::-1]]).tolist()))
        # the partition depends on the order of a vertex's neighbours only through ties: sizes and cut quality are those of the mesh's graph
        a, b = mb.partition_graph(None, 4, xadj=xadj, adj=adj), mb.partition_graph(m, 4)
        assert np.array_equal(np.bincount(a, minlength=4), np.bincount(b, minlength=4))
        assert abs(_edge_cut(m, a) - _edge_cut(m, b)) <= 0.25 * _edge_cut(m, b) + 4


def test_graph_partition_drives_a_partitioned_plan():
    """any partition vector is a valid input of the preprocessor: the plans of a graph-partitioned mesh own every cell exactly once and
    announce halos that are each other's mirror image"""
    mesh = mb.Mesh.generate("cartesian_tri", 30, 20, 3.0, 2.0)
    n = 3
    part = mb.partition_graph(mesh, n)
    owned = np.zeros(mesh.n_cells, dtype=np.int64)
    for r in range(n):
        plan = mb.Plan(mesh, "TENO", order=3, bcs=SYM4, fp_mode="fast", part=part, rank=r, n_ranks=n)
        perm = plan.get("perm_cells")
        owned[perm[:plan.N_owned]] += 1
        assert (part[perm[:plan.N_owned]] == r).all()
        peers, counts, ids = plan.get("halo_peers"), plan.get("halo_recv_counts"), plan.get("halo_recv_ids")
        off = 0
        for p, c in zip(peers, counts):
            assert (part[ids[off:off + int(c)]] == p).all()
            off += int(c)
    assert (owned == 1).all()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_streaming_tables_are_the_reference_tables_compacted(order):
    """FAST mode streams compact tables (mallard_b200/csrc/teno_stream.cuh): rows 1..K-1 x columns 1..M-1 of every
    reference matrix with the transformed areas folded into the columns.  What is dropped must be exact zeros and what is
    kept must be the reference value times the reference area, bit for bit."""
    mesh = mb.Mesh.generate("cartesian_tri", 9, 7, 1.0, 0.7)
    plan = mb.Plan(mesh, "TENO", order=order, fp_mode="fast")
    K, M, S, CT = plan.K, plan.M, 4, plan.stream_tile
    assert CT in (8, 32)
    KR, MC = K - 1, M - 1
    npair = MC // 2
    off_g, off_s = plan.get("teno:offsets_stencil_groups"), plan.get("teno:offsets_stencils")
    sten, mats, areas = plan.get("teno:stencils"), plan.get("teno:reconstruction_matrices"), plan.get("teno:transformed_areas")
    perm = plan.get("perm_cells")
    inv = np.empty(mesh.n_cells, np.int64); inv[perm] = np.arange(mesh.n_cells)
    n_ft = (plan.N_recon + CT - 1) // CT
    ids = plan.get("fm_ids").reshape(n_ft, S, MC, CT)
    mat = plan.get("fm_mat").reshape(n_ft, S, KR, (2 * npair + 1) * CT)
    a0 = plan.get("fm_area0")
    OI = plan.get("teno:oscillation_indicator").reshape(K, K)
    assert not OI[0].any() and not OI[:, 0].any()               # the constant mode carries no oscillation
    OIs = plan.get("OIs").reshape(KR, KR)
    O = OI[1:, 1:]
    assert np.array_equal(OIs, np.triu(O + O.T, 1) + np.diag(np.diag(O)))   # a^T OI a folded onto the upper triangle
    n_empty = 0
    for i in range(plan.N_recon):
        c = int(perm[i]); ft, fl = divmod(i, CT)
        for s in range(S):
            g = off_g[c] + s
            lo, hi = int(off_s[g]), int(off_s[g + 1])
            row = mat[ft, s]
            got = np.empty((KR, MC))
            for m in range(MC):
                got[:, m] = row[:, (m // 2 * CT + fl) * 2 + (m & 1)] if m < 2 * npair else row[:, 2 * npair * CT + fl]
            if hi == lo:   # an empty stencil lists the cell itself (zero right-hand side) and carries a zero matrix
                assert (ids[ft, s, :, fl] == i).all() and not got.any()
                n_empty += 1
                continue
            A = mats[K * lo:K * hi].reshape(K, M); ar = areas[lo:hi]
            assert not A[0].any() and not A[:, 0].any()
            assert np.array_equal(ids[ft, s, :, fl], inv[sten[lo + 1:hi]])
            if s == 0:
                assert a0[i] == ar[0]
            assert np.array_equal(got, A[1:, 1:] * ar[None, 1:])
    assert n_empty > 0    # boundary faces have no directional stencil
    for i in range(plan.N_recon, n_ft * CT):   # padding cells of the last tile
        ft, fl = divmod(i, CT)
        assert (ids[ft, :, :, fl] == i).all()


def test_synthetic_unstructured_mesh_matches_oracle_geometry_and_teno_tables(oracle_mod):
    """The jittered, id-shuffled triangulation (BASELINE configs[3] family) as plain arrays: geometry computed by the library's
    host code and the TENO tables of the host preprocessor are bit-identical to the oracle's on the same arrays."""
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(14, 11, 10.0, 10.0, seed=12345)
    a = mesh.arrays
    keys = ["node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell",
            "offsets_nodes_of_face", "nodes_of_face", "cells_of_face"]
    om = oracle_mod.Mesh.from_arrays({k: a[k] for k in keys}, mesh.zones)
    for k in ("cell_coords", "cell_volume", "face_area", "face_normals"):
        assert np.array_equal(om.get(k).reshape(-1), a[k].reshape(-1)), k
    assert abs(a["cell_volume"].sum() - 100.0) < 1e-10 and a["cell_volume"].min() > 0
    assert sorted(len(f) for n, f in mesh.zones if n != "interior") == [11, 11, 14, 14]
    # a second call with the same seed gives the same mesh; another seed a different one
    assert np.array_equal(syn.jittered_tri(14, 11, 10.0, 10.0, seed=12345).arrays["cells_of_face"], a["cells_of_face"])
    assert not np.array_equal(syn.jittered_tri(14, 11, 10.0, 10.0, seed=1).arrays["cells_of_face"], a["cells_of_face"])
    plan = mb.Plan(mesh, "TENO", order=3, bcs=syn.EXTRAP4, fp_mode="strict")
    so = oracle_mod.Solver(om, "TENO", "HLLC", "SSPRK3", order=3, bcs=syn.EXTRAP4)
    for k in ("teno:stencils", "teno:reconstruction_matrices", "teno:transformed_areas", "teno:offsets_stencils"):
        assert np.array_equal(plan.get(k), so.get(k)), k


SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_partitioned_plan_numbers_interior_cells_first(n_ranks):
    """Multi-GPU overlap: the owned cells whose TENO stencils hold owned cells only come first, so their
    reconstruction can run while the ghost states are in flight; each class keeps the unpartitioned plan's relative order."""
    mesh = mb.Mesh.generate("cartesian_tri", 30, 20, 3.0, 2.0)
    part = mb.partition(mesh, n_ranks)
    whole = mb.Plan(mesh, "TENO", order=3, bcs=SYM4, fp_mode="fast")
    for r in range(n_ranks):
        plan = mb.Plan(mesh, "TENO", order=3, bcs=SYM4, fp_mode="fast", part=part, rank=r, n_ranks=n_ranks)
        ni = int(plan.get("n_interior")[0])
        CT, S, MC = plan.stream_tile, 4, plan.M - 1
        n_ft = (plan.N_recon + CT - 1) // CT
        ids = plan.get("fm_ids").reshape(n_ft, S, MC, CT)
        touches_ghost = np.array([(ids[i // CT, :, :, i % CT] >= plan.N_owned).any() for i in range(plan.N_owned)])
        assert 0 < ni < plan.N_owned
        assert not touches_ghost[:ni].any() and touches_ghost[ni:].all()
        perm = plan.get("perm_cells")
        assert (part[perm[:plan.N_owned]] == r).all() and (part[perm[plan.N_owned:]] != r).all()
        assert int(whole.get("n_interior")[0]) == whole.N_owned


def test_bench_reference_arm_line_and_no_gpu_failure():
    """bench.py's contract on a box without a GPU: the reference arm (oracle/_ref, or the oracle port) prints one JSON line
    with the metric's keys; the product arm refuses to run (no CPU fallback)."""
    import json
    import subprocess
    import sys
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import bench
    value, info = bench.reference_cpu(2, 1, nx=20, ny=16)
    assert value > 0 and info["kind"] in ("reference", "port") and info["cores"] >= 1 and "640 cells" in info["sample"]
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True)
        assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_bench_strong_records_cannot_cost_the_main_line(monkeypatch):
    """At N = 1 the strong-scaling records (16 M cells, ~100 GB of device memory) run in a child process under a deadline: records
    finished before a crash or a hang are kept, the failure is reported as a record, and the parent always returns."""
    import json
    import subprocess
    import sys
    import types
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import bench
    a = types.SimpleNamespace(steps=2, warmup=1, fp="fast")

    def child(script):
        return lambda cmd, **kw: subprocess.Popen([sys.executable, "-c", script], **kw)
    rec = bench.STRONG_TAG + json.dumps({"workload": "w", "value": 1.0})
    monkeypatch.setenv("MLB_BENCH_DEADLINE", "100000")
    ok = bench.strong_records_in_child(a, popen=child("print('noise'); print(%r); print(%r)" % (rec, rec)))
    assert ok == [{"workload": "w", "value": 1.0}] * 2
    crash = bench.strong_records_in_child(a, popen=child("import os; print(%r, flush=True); os._exit(9)" % rec))
    assert crash[0]["value"] == 1.0 and "code 9" in crash[1]["error"]
    monkeypatch.setenv("MLB_BENCH_DEADLINE", "%f" % (time_since_bench_import(bench) + 3.0))
    hang = bench.strong_records_in_child(a, popen=child("import time; print(%r, flush=True); time.sleep(600)" % rec), min_left=0.0)
    assert hang[0]["value"] == 1.0 and "deadline" in hang[1]["aborted"]
    monkeypatch.setenv("MLB_BENCH_DEADLINE", "1")
    assert "time budget" in bench.strong_records_in_child(a)[0]["skipped"]


def test_strong_scaling_efficiency_uses_the_base_point_measured_on_the_same_box(tmp_path):
    """bench.py's `strong` records: efficiency = value_N / (N / N0 x value_N0) against the smallest GPU count the mesh was measured on
    by an earlier run of the same series on this box; the committed number of an earlier round only when there is none."""
    import json
    import bench_multi as bm
    live = str(tmp_path / "live.json")
    r1 = {"value": 8e8, "ms_per_step": 60.0}
    bm.strong_efficiency(r1, "vortex_16M", 1, live)
    assert r1["efficiency"] == 1.0
    r8 = {"value": 6e9, "ms_per_step": 8.0}
    bm.strong_efficiency(r8, "vortex_16M", 8, live)
    assert abs(r8["efficiency"] - 6e9 / (8 * 8e8)) < 1e-12 and "this box" in r8["efficiency_base"]
    r4 = {"value": 3e9, "ms_per_step": 64.0}
    bm.strong_efficiency(r4, "vortex_64M", 4, live)                     # first point of a mesh that does not fit fewer GPUs: no efficiency yet
    assert "efficiency" not in r4
    r8 = {"value": 5.7e9, "ms_per_step": 33.0}
    bm.strong_efficiency(r8, "vortex_64M", 8, live)
    assert abs(r8["efficiency"] - 0.95) < 1e-12 and "on 4 GPU(s)" in r8["efficiency_base"]
    assert json.load(open(live))["vortex_64M"]["n_gpus"] == 4           # the smallest count stays the base
    rf = {"value": 6e9, "ms_per_step": 8.0}
    bm.strong_efficiency(rf, "vortex_16M", 8, str(tmp_path / "absent.json"))
    assert "r02_strong_baselines.json" in rf["efficiency_base"] and rf["efficiency"] > 0


def time_since_bench_import(bench):
    import time
    return time.perf_counter() - bench.T_START


def test_degenerate_meshes_are_handled_on_the_host():
    """Edge cases: a one-cell mesh (empty interior zone, every face on a boundary), and TENO on meshes that cannot fill a
    stencil of M cells or that contain quadrilaterals (the reference throws for both)."""
    m = mb.Mesh.generate("cartesian", 1, 1, 1.0, 1.0)
    assert m.n_cells == 1 and dict((n, len(f)) for n, f in m.zones)["interior"] == 0
    p = mb.Plan(m, "FO", bcs=SYM4)
    assert (p.N, p.NF) == (1, 4)
    with pytest.raises(mb.MallardError, match="too small to fill a stencil"):                    # a single quadrilateral has no neighbours to reconstruct from
        mb.Plan(m, "TENO", order=3, bcs=SYM4)
    for n in (1, 2, 3):
        with pytest.raises(mb.MallardError, match="too small to fill a stencil"):
            mb.Plan(mb.Mesh.generate("cartesian_tri", n, n, 1.0, 1.0), "TENO", order=3, bcs=SYM4, fp_mode="fast")


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/mallard_b200.h must compile as C99 (no C++ or torch types) and a C program must link
    against libmallard_b200.so and call it."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "mallard_b200.h"\n'
                   'int main(void) {\n  mlb_host_mesh *m = 0;\n  mlb_mesh v;\n'
                   '  if (mlb_host_mesh_generate(&m, 1, 3, 2, 1.0, 1.0)) return 1;\n'
                   '  if (mlb_host_mesh_view(m, &v)) return 2;\n'
                   '  printf("%s %u\\n", mlb_version(), v.n_cells);\n  mlb_host_mesh_free(m);\n  return 0;\n}\n')
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "mallard_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-L", libdir, "-lmallard_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert out.startswith("mallard_b200") and out.split()[-1] == "12"      # cartesian_tri 3 x 2 = 12 triangles


@pytest.mark.parametrize("kind", ["tri", "quad"])
def test_gmsh_file_round_trip(tmp_path, kind):
    """a mesh written to Gmsh MSH 2.2 and read back holds the same nodes, cells, boundary zones and (bit for
    bit, being computed by the same host code from the same node order) cell volumes; faces are renumbered."""
    from mallard_b200 import meshio, synthetic as syn
    mesh = syn.jittered_tri(9, 7, 3.0, 2.0, seed=4) if kind == "tri" else mb.Mesh.generate("wedge", 12, 6, 4.0, 1.5)
    path = tmp_path / "mesh.msh"
    meshio.write_gmsh(mesh, str(path))
    back = meshio.read_gmsh(str(path))
    a, b = mesh.arrays, back.arrays
    assert back.n_cells == mesh.n_cells and back.n_nodes == mesh.n_nodes
    assert np.array_equal(a["node_coords"], b["node_coords"])
    assert np.array_equal(a["nodes_of_cell"], b["nodes_of_cell"]) and np.array_equal(a["offsets_nodes_of_cell"], b["offsets_nodes_of_cell"])
    assert np.array_equal(a["cell_volume"], b["cell_volume"]) and np.array_equal(a["cell_coords"], b["cell_coords"])
    real = gu.real_faces(a["cells_of_face"], a["nodes_of_face"])
    assert back.n_faces == int(real.sum())
    edge = lambda arr, f: frozenset(map(int, arr["nodes_of_face"].reshape(-1, 2)[f]))
    za, zb = dict(mesh.zones), dict(back.zones)
    assert set(za) == set(zb)
    for n in za:
        assert {edge(a, f) for f in za[n]} == {edge(b, f) for f in zb[n]}, n
    # every face: side 0 is the lower-numbered cell, normals point out of it, areas match the originals
    cof = b["cells_of_face"].reshape(-1, 2)
    assert ((cof[:, 1] < 0) | (cof[:, 0] < cof[:, 1])).all()
    area_a = {edge(a, f): a["face_area"][f] for f in np.nonzero(real)[0]}
    assert all(area_a[edge(b, f)] == b["face_area"][f] for f in range(back.n_faces))
    plan = mb.Plan(back, "TENO" if kind == "tri" else "FO", order=2, bcs=[dict(name=n, type="extrapolation") for n in zb if n != "interior"],
                   fp_mode="fast")
    assert plan.N == back.n_cells


def test_gmsh_reader_rejects_what_it_cannot_represent(tmp_path):
    from mallard_b200 import meshio
    p = tmp_path / "bad.msh"
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError, match="MSH 2.x ASCII"):
        meshio.read_gmsh(str(p))
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n1 0 0 0\n2 1 0 0\n3 0 1 0\n$EndNodes\n$Elements\n1\n1 1 2 1 1 1 2\n$EndElements\n")
    with pytest.raises(ValueError, match="no triangles or quadrilaterals"):
        meshio.read_gmsh(str(p))


# ------------------------------------------------------------------------------------------------------------------
# Rank-local ingest (mlb_create_local / mlb_plan_create_local): no rank holds the global mesh
# ------------------------------------------------------------------------------------------------------------------
def test_local_mesh_generator_equals_the_restriction_of_the_global_mesh():
    """synthetic.jittered_tri_local builds a rank's part of the jittered, id-shuffled triangulation from closed forms; it must be
    the very arrays local_mesh.extract_local cuts out of the global mesh (and the whole mesh when nothing is cut)."""
    from mallard_b200 import local_mesh as lm, synthetic as syn
    nx, ny = 14, 11
    g = syn.jittered_tri(nx, ny, 10.0, 8.0, seed=12345)
    whole = syn.jittered_tri_local(nx, ny, 10.0, 8.0, 1, 0, seed=12345)
    for k in g.arrays:
        assert np.array_equal(g.arrays[k], whole.mesh.arrays[k]), k
    assert [(n, f.tolist()) for n, f in g.zones] == [(n, f.tolist()) for n, f in whole.mesh.zones]
    part = mb.partition(g, 3)
    assert np.array_equal(part, mb.partition_coords(g.arrays["cell_coords"], 3))       # the partition needs the centroids only
    for r in range(3):
        lp = syn.jittered_tri_local(nx, ny, 10.0, 8.0, 3, r, seed=12345, layers=2)
        keep = np.zeros(g.n_cells, bool)
        keep[lp.global_cell_ids] = True
        assert keep[part == r].all() and lp.n_owned == int((part == r).sum())
        ex, gc, gf = lm.extract_local(g, keep)
        assert np.array_equal(gc, lp.global_cell_ids) and np.array_equal(part[gc], lp.part_local)
        for k in ex.arrays:
            assert np.array_equal(ex.arrays[k], lp.mesh.arrays[k]), (r, k)
        assert [(n, f.tolist()) for n, f in ex.zones] == [(n, f.tolist()) for n, f in lp.mesh.zones]
        assert np.array_equal(lm.local_info(g, gc)["cell0_nodes"], lp.local["cell0_nodes"])
        cof = lp.mesh.arrays["cells_of_face"]
        assert (cof[:, 1] == lm.CUT).any() and (cof[:, 0] >= 0).all()
    # extract_local + dilate on a mesh that is not a structured parent's child: quads of the wedge
    w = mb.Mesh.generate("wedge", 12, 8, 4.0, 1.5)
    keep = lm.dilate(w, mb.partition(w, 2) == 0, 2)
    ex, gc, gf = lm.extract_local(w, keep)
    assert ex.n_cells == int(keep.sum()) and np.array_equal(ex.arrays["cell_volume"], w.arrays["cell_volume"][gc])


@pytest.mark.parametrize("fp", ["strict", "fast"])
def test_plan_of_a_local_mesh_is_bit_identical_to_the_plan_of_the_partitioned_global_mesh(fp):
    """Stencil membership depends on the order of cell ids and integral_psi_target on the global mesh's cell 0
    (face_reconstruction.cpp:598-602): the local mesh keeps the global order and is told cell 0's nodes, so every table - ids,
    pseudo-inverses, slots, accumulation order - is bit-identical; a halo that is too thin is refused, never silently wrong."""
    from mallard_b200 import synthetic as syn
    nx, ny, W = 34, 30, 3
    g = syn.jittered_tri(nx, ny, 10.0, 8.0, seed=12345)
    part = mb.partition(g, W)
    for r in range(W):
        thin = syn.jittered_tri_local(nx, ny, 10.0, 8.0, W, r, seed=12345, layers=3)
        with pytest.raises(mb.MallardError, match="more ghost layers"):
            mb.Plan(thin.mesh, "TENO", order=3, bcs=syn.EXTRAP4, part=thin.part_local, rank=r, n_ranks=W, fp_mode=fp, local=thin.local)
        lp = syn.jittered_tri_local(nx, ny, 10.0, 8.0, W, r, seed=12345, layers=10)
        pl = mb.Plan(lp.mesh, "TENO", order=3, bcs=syn.EXTRAP4, part=lp.part_local, rank=r, n_ranks=W, fp_mode=fp, local=lp.local)
        pg = mb.Plan(g, "TENO", order=3, bcs=syn.EXTRAP4, part=part, rank=r, n_ranks=W, fp_mode=fp)
        assert (pl.N, pl.N_owned, pl.N_recon, pl.NF) == (pg.N, pg.N_owned, pg.N_recon, pg.NF) and lp.mesh.n_cells < g.n_cells
        assert np.array_equal(lp.global_cell_ids[pl.get("perm_cells")], pg.get("perm_cells"))
        names = ["slot_face", "slot_nbr", "rhs_order", "n_interior", "teno:integral_psi_target", "teno:oscillation_indicator"]
        names += ["fm_ids", "fm_mat", "fm_area0", "OIs"] if fp == "fast" else ["st_ids"]
        for name in names:
            a, b = pl.get(name), pg.get(name)
            assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), (r, name)
        assert np.array_equal(lp.global_cell_ids[pl.get("halo_recv_ids")], pg.get("halo_recv_ids"))
    # first order needs one layer only
    lp = syn.jittered_tri_local(nx, ny, 10.0, 8.0, W, 1, seed=12345, layers=1)
    pl = mb.Plan(lp.mesh, "FO", bcs=syn.EXTRAP4, part=lp.part_local, rank=1, n_ranks=W, local=lp.local)
    pg = mb.Plan(g, "FO", bcs=syn.EXTRAP4, part=part, rank=1, n_ranks=W)
    assert np.array_equal(pl.get("slot_nbr"), pg.get("slot_nbr")) and np.array_equal(pl.get("rhs_order"), pg.get("rhs_order"))
