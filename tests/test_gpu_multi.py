"""Partitioned (multi-GPU) execution on the device.  (1) several rank contexts on ONE GPU with the halo exchange done by
device-to-device copies between the library's buffers: every rank's result must equal the single-context result bit for
bit in STRICT mode; (2) with >= 2 GPUs, the real thing: one process per GPU, NCCL halo exchange + all-reduce of dt."""
import os
import socket

import numpy as np
import pytest
import torch

import mallard_b200 as mb
from mallard_b200.parallel import device_tensor

pytestmark = pytest.mark.gpu
SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]


def _state(xy):
    rho = 1.0 + 0.3 * np.sin(2 * np.pi * xy[:, 0]) * np.cos(2 * np.pi * 1.5 * xy[:, 1])
    u = 0.4 + 0.2 * np.cos(2 * np.pi * xy[:, 1]); v = -0.3 + 0.2 * np.sin(2 * np.pi * xy[:, 0])
    p = 1.0 + 0.2 * np.cos(2 * np.pi * (xy[:, 0] - xy[:, 1]))
    e = p / (0.4 * rho)
    return np.stack([rho, rho * u, rho * v, rho * (e + 0.5 * (u * u + v * v))], 1)


class Loopback:
    """n rank contexts on one device; exchange = copies between their send/receive buffers.
    async_copies=True: nothing synchronises the host inside a step - the copies are enqueued on the RECEIVER's communication
    stream (mlb_comm_stream), ordered against the sender's pack and the sender's next pack by events, exactly the ordering
    NCCL's send/recv pairs provide; the interior reconstruction (mlb_stage_begin) then really runs concurrently with them.
    locals_=[LocalPart...]: every context is created from its rank-local mesh (mlb_create_local)."""

    def __init__(self, mesh, part, n, begin=True, async_copies=False, locals_=None, **kw):
        self.begin = begin   # enqueue the interior reconstruction right after the exchange (mlb_stage_begin), or leave it to mlb_stage
        self.async_copies = async_copies
        self.locals_ = locals_
        if locals_ is None:
            self.s = [mb.Solver(mesh, part=part, rank=r, n_ranks=n, device=0, **kw) for r in range(n)]
        else:
            self.s = [mb.Solver(lp.mesh, part=lp.part_local, rank=r, n_ranks=n, device=0, local=lp.local, **kw) for r, lp in enumerate(locals_)]
        self.cs = [torch.cuda.ExternalStream(s.comm_stream, device=0) for s in self.s]
        self.ev_packed = [torch.cuda.Event() for _ in self.s]
        self.ev_copied = [None for _ in self.s]
        wants = []
        for s in self.s:
            peers, _, rc = s.halo_info()
            wants.append({int(p): s.halo_recv_ids(i, rc[i]) for i, p in enumerate(peers)})
        for r, s in enumerate(self.s):
            send = {q: wants[q][r] for q in range(n) if q != r and r in wants[q] and len(wants[q][r])}
            sp = sorted(send)
            s.halo_set_send_ids(sp, [send[p] for p in sp])
        self.info = [s.halo_info() for s in self.s]
        self.buf = []
        for s, (peers, sc, rc) in zip(self.s, self.info):
            a, b = s.halo_buffers()
            self.buf.append((device_tensor(a, 4 * int(sc.sum()), 0), device_tensor(b, 4 * int(rc.sum()), 0)))

    def exchange_async(self, stage):
        for r, s in enumerate(self.s):
            for e in self.ev_copied:                      # a send buffer is rewritten only after every reader of the last round is done
                if e is not None:
                    self.cs[r].wait_event(e)
            s.halo_pack(stage)
            self.ev_packed[r].record(self.cs[r])
        for r, (peers, sc, rc) in enumerate(self.info):
            ro = 0
            with torch.cuda.stream(self.cs[r]):
                for p, n_recv in zip(peers, rc):
                    pp, psc, _ = self.info[int(p)]
                    so = 4 * int(psc[:list(pp).index(r)].sum())
                    n = 4 * int(n_recv)
                    if n:
                        self.cs[r].wait_event(self.ev_packed[int(p)])
                        self.buf[r][1][ro:ro + n].copy_(self.buf[int(p)][0][so:so + n], non_blocking=True)
                    ro += n
            e = torch.cuda.Event()
            e.record(self.cs[r])
            self.ev_copied[r] = e
        for s in self.s:
            s.halo_unpack(stage)
            if self.begin:
                s.stage_begin(stage)

    def exchange(self, stage):
        if self.async_copies:
            return self.exchange_async(stage)
        for s in self.s:
            s.halo_pack(stage)
        for s in self.s:
            s.synchronize()
        for r, (peers, sc, rc) in enumerate(self.info):
            ro = 0
            for p, n_recv in zip(peers, rc):
                pp, psc, _ = self.info[int(p)]
                so = 4 * int(psc[:list(pp).index(r)].sum())
                n = 4 * int(n_recv)
                assert int(psc[list(pp).index(r)]) == int(n_recv)
                self.buf[r][1][ro:ro + n].copy_(self.buf[int(p)][0][so:so + n])
                ro += n
        torch.cuda.synchronize()
        for s in self.s:
            s.halo_unpack(stage)
            if self.begin:
                s.stage_begin(stage)

    def step(self, cfl):
        self.exchange(0)
        mx = max(s.local_max_spectral_radius() for s in self.s)      # the one host round trip per step of this test driver
        for s in self.s:
            s.apply_dt(cfl, mx)
            s.stage(0)
        for st in range(1, self.s[0].n_stages):
            self.exchange(st)
            for s in self.s:
                s.stage(st)
        for s in self.s:
            s.finish_step()

    def set_state(self, U0):
        for r, s in enumerate(self.s):
            s.set_state(U0 if self.locals_ is None else U0[self.locals_[r].global_cell_ids.astype(np.int64)])

    def state(self):
        if self.locals_ is None:
            return sum(s.get_state() for s in self.s)   # every rank exports zeros outside its own cells
        U = np.zeros((self.locals_[0].n_global, 4))
        for lp, s in zip(self.locals_, self.s):
            U[lp.global_cell_ids.astype(np.int64)] += s.get_state()
        return U


@pytest.mark.parametrize("recon,integ,n_ranks,fp", [("FO", "SSPRK3", 2, "strict"), ("FO", "RK4", 3, "strict"), ("TENO", "SSPRK3", 2, "strict"),
                                                    ("TENO", "SSPRK3", 4, "strict"), ("TENO", "SSPRK3", 3, "fast")])
def test_partitioned_ranks_reproduce_the_single_context_run(recon, integ, n_ranks, fp):
    mtype = "cartesian_tri" if recon == "TENO" else "wedge"
    mesh = mb.Mesh.generate(mtype, 36, 24, 4.0, 1.5)
    U0 = _state(mesh.arrays["cell_coords"])
    kw = dict(recon=recon, riemann="HLLC", integrator=integ, order=3, bcs=SYM4, fp_mode=fp, teno_fixed=True)
    one = mb.Solver(mesh, **kw)
    one.set_state(U0)
    part = mb.partition(mesh, n_ranks)
    many = Loopback(mesh, part, n_ranks, begin=(n_ranks != 3), **kw)
    many.set_state(U0)
    for _ in range(3):
        one.calc_dt(0.3)
        one.take_step()
        many.step(0.3)
    U1, Un = one.get_state(), many.state()
    assert np.isfinite(U1).all()
    if fp == "strict":
        assert np.array_equal(U1, Un)           # same per-cell operation order whatever the partition
    else:
        assert np.abs(Un - U1).max() <= 1e-12 * np.abs(U1).max()
    t1, n1 = one.time()
    for s in many.s:
        t, n = s.time()
        assert n == n1 and t == t1
    # owned-cell host seam: set_owned / get_owned round trip in owned_cells order
    s0 = many.s[0]
    own = s0.owned_cells()
    assert np.array_equal(s0.get_owned(), Un[own])


@pytest.mark.parametrize("n_ranks,fp", [(4, "strict"), (3, "fast")])
def test_partitioned_unstructured_mesh_reproduces_the_single_context_run(n_ranks, fp):
    """BASELINE configs[3] family: a jittered, id-shuffled triangulation cut by recursive coordinate bisection (ragged cuts
    through an unstructured numbering, up to three peers per rank) - bit-identical to the single-context run in STRICT mode."""
    from mallard_b200 import synthetic as syn
    mesh = syn.jittered_tri(40, 30, 10.0, 10.0, seed=11)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", order=3, bcs=syn.EXTRAP4, fp_mode=fp, teno_fixed=True)
    one = mb.Solver(mesh, **kw)
    one.set_state(U0)
    part = mb.partition(mesh, n_ranks)
    assert sorted(np.unique(part)) == list(range(n_ranks))
    many = Loopback(mesh, part, n_ranks, **kw)
    assert max(len(i[0]) for i in many.info) >= 2          # some rank talks to more than one peer
    many.set_state(U0)
    for _ in range(3):
        one.calc_dt(0.3)
        one.take_step()
        many.step(0.3)
    U1, Un = one.get_state(), many.state()
    assert np.isfinite(U1).all()
    if fp == "strict":
        assert np.array_equal(U1, Un)
    else:
        assert np.abs(Un - U1).max() <= 1e-12 * np.abs(U1).max()


@pytest.mark.parametrize("n_ranks,fp,integ", [(4, "strict", "SSPRK3"), (3, "fast", "RK4")])
def test_rank_local_ingest_reproduces_the_single_context_run(n_ranks, fp, integ):
    """mlb_create_local: every rank context is built from ITS PART of the mesh only (generated without the global mesh,
    synthetic.jittered_tri_local) and exchanges GLOBAL ids with its peers; the run must be bit-identical (STRICT) to the
    single-context run on the global mesh.  The copies between the ranks' buffers are asynchronous (no host synchronisation
    inside a step): the interior reconstruction overlaps them as it overlaps NCCL's transfers."""
    from mallard_b200 import synthetic as syn
    nx, ny = 48, 36
    mesh = syn.jittered_tri(nx, ny, 10.0, 10.0, seed=12345)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    kw = dict(recon="TENO", riemann="HLLC", integrator=integ, order=3, bcs=syn.EXTRAP4, fp_mode=fp, teno_fixed=True)
    one = mb.Solver(mesh, **kw)
    one.set_state(U0)
    with pytest.raises(mb.MallardError, match="more ghost layers"):
        lp = syn.jittered_tri_local(nx, ny, 10.0, 10.0, n_ranks, 0, seed=12345, layers=2)
        mb.Solver(lp.mesh, part=lp.part_local, rank=0, n_ranks=n_ranks, device=0, local=lp.local, **kw)
    locals_ = [syn.jittered_tri_local(nx, ny, 10.0, 10.0, n_ranks, r, seed=12345, layers=10) for r in range(n_ranks)]
    assert sum(lp.n_owned for lp in locals_) == mesh.n_cells and max(lp.mesh.n_cells for lp in locals_) < mesh.n_cells
    many = Loopback(None, None, n_ranks, async_copies=True, locals_=locals_, **kw)
    many.set_state(U0)
    for _ in range(4):
        one.calc_dt(0.3)
        one.take_step()
        many.step(0.3)
    U1, Un = one.get_state(), many.state()
    assert np.isfinite(U1).all()
    if fp == "strict":
        assert np.array_equal(U1, Un)
    else:
        assert np.abs(Un - U1).max() <= 1e-12 * np.abs(U1).max()


@pytest.mark.parametrize("fp,integ", [("fast", "SSPRK3"), ("fast", "RK4"), ("strict", "SSPRK3")])
def test_asynchronous_exchange_overlapping_the_interior_reconstruction(fp, integ):
    """The path that ships: FAST mode has interior tiles (interior_tiles() > 0), so mlb_stage_begin launches their
    reconstruction on the compute stream while pack / transfer / unpack (and the ghost-primitive refresh of stage 0) run on the
    communication stream, ordered by ev_state / ev_halo only.  No host synchronisation inside a step; many steps, so that a
    missing ordering edge has every chance to show."""
    mesh = mb.Mesh.generate("cartesian_tri", 96, 64, 3.0, 2.0)      # 12 288 cells: thousands of interior tiles per rank
    U0 = _state(mesh.arrays["cell_coords"] / 2.0)
    kw = dict(recon="TENO", riemann="HLLC", integrator=integ, order=3, bcs=SYM4, fp_mode=fp, teno_fixed=True, keep_stage_rhs=False)
    one = mb.Solver(mesh, **kw)
    one.set_state(U0)
    n_ranks = 3
    part = mb.partition(mesh, n_ranks)
    many = Loopback(mesh, part, n_ranks, async_copies=True, **kw)
    many.set_state(U0)
    for _ in range(12):
        one.calc_dt(0.3)
        one.take_step()
        many.step(0.3)
    U1, Un = one.get_state(), many.state()
    assert np.isfinite(U1).all()
    if fp == "strict":
        assert np.array_equal(U1, Un)
    else:
        assert np.abs(Un - U1).max() <= 12 * 1e-12 * np.abs(U1).max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from mallard_b200.parallel import DistributedSolver
    from mallard_b200 import synthetic as syn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="4")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        # (a) split-phase ABI driven from Python over torch.distributed, global mesh on every rank, STRICT
        mesh = mb.Mesh.generate("cartesian_tri", 48, 32, 2.0, 1.0)
        U0 = _state(mesh.arrays["cell_coords"])
        part = mb.partition(mesh, world)
        ds = DistributedSolver(mesh, part, rank, world, rank, recon="TENO", riemann="HLLC", integrator="SSPRK3", order=3, bcs=SYM4,
                               fp_mode="strict", teno_fixed=True)
        ds.set_state(U0)
        t, n = ds.run(3, cfl=0.3)
        U = ds.gather_state()
        if rank == 0:
            np.save(os.path.join(out_dir, "U.npy"), U)
            np.save(os.path.join(out_dir, "t.npy"), np.array([t, n]))
        # (b) the native driver (NCCL inside the library, CUDA-graph replayed steps) on rank-local meshes: FAST (interior tiles
        #     overlap the exchange) with SSPRK3 and RK4, and STRICT
        nx, ny = 64, 48
        for tag, fp, integ in (("fast_ssprk3", "fast", "SSPRK3"), ("fast_rk4", "fast", "RK4"), ("strict_ssprk3", "strict", "SSPRK3")):
            lp = syn.jittered_tri_local(nx, ny, 10.0, 10.0, world, rank, seed=12345, layers=10)
            dn = DistributedSolver(lp.mesh, lp.part_local, rank, world, rank, local=lp.local, native=True, recon="TENO", riemann="HLLC",
                                   integrator=integ, order=3, bcs=syn.EXTRAP4, fp_mode=fp, teno_fixed=True, keep_stage_rhs=False)
            dn.s.set_state(syn.isentropic_vortex(lp.mesh.arrays["cell_coords"]))
            t, n = dn.run(9, cfl=0.3)                  # 1 eager step + 8 graph replays
            replays = int(dn.s.get("stats")[11])
            Uh = dn.s.get_owned()
            Uh2 = dn.step_host(Uh.copy(), 0.3)         # the host-buffer seam on top: one more step
            U = dn.gather_state()
            if rank == 0:
                np.save(os.path.join(out_dir, "U_%s.npy" % tag), U)
                np.save(os.path.join(out_dir, "t_%s.npy" % tag), np.array([t, n, replays]))
            dn.s.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_halo_exchange_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from mallard_b200 import synthetic as syn
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mesh = mb.Mesh.generate("cartesian_tri", 48, 32, 2.0, 1.0)
    one = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", order=3, bcs=SYM4, fp_mode="strict", teno_fixed=True)
    one.set_state(_state(mesh.arrays["cell_coords"]))
    t, _ = one.run(3, cfl=0.3)
    assert np.array_equal(np.load(tmp_path / "U.npy"), one.get_state())
    assert np.load(tmp_path / "t.npy")[0] == t
    # native driver on rank-local meshes against the single-GPU run on the global mesh
    g = syn.jittered_tri(64, 48, 10.0, 10.0, seed=12345)
    U0 = syn.isentropic_vortex(g.arrays["cell_coords"])
    for tag, fp, integ in (("fast_ssprk3", "fast", "SSPRK3"), ("fast_rk4", "fast", "RK4"), ("strict_ssprk3", "strict", "SSPRK3")):
        one = mb.Solver(g, "TENO", "HLLC", integ, order=3, bcs=syn.EXTRAP4, fp_mode=fp, teno_fixed=True, keep_stage_rhs=False)
        one.set_state(U0)
        t, _ = one.run(10, cfl=0.3)
        U1, Un = one.get_state(), np.load(tmp_path / ("U_%s.npy" % tag))
        tn = np.load(tmp_path / ("t_%s.npy" % tag))
        assert tn[2] == 8, "the distributed steps were not replayed as a CUDA graph"
        assert np.isfinite(U1).all()
        if fp == "strict":
            assert np.array_equal(U1, Un), tag
        else:
            assert np.abs(Un - U1).max() <= 1e-12 * 10 * np.abs(U1).max(), tag
