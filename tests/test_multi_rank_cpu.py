"""world_size-2/3 tests of the multi-GPU host logic on CPU (gloo): partition -> per-rank plans (owned cells, ghost lists) ->
exchange plan -> the very exchange routine the NCCL path runs (mallard_b200.parallel.halo_exchange), on CPU tensors.
No device code runs here; the device side of the same path is covered by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mallard_b200 as mb
from mallard_b200.parallel import exchange_plan, halo_exchange

SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, recon, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = mb.Mesh.generate("cartesian_tri", 14, 10, 2.0, 1.0)
        nc = mesh.n_cells
        recon, _, partitioner = recon.partition(":")
        part = mb.partition_graph(mesh, world) if partitioner == "graph" else mb.partition(mesh, world)
        plan = mb.Plan(mesh, recon, order=2, bcs=SYM4, part=part, rank=rank, n_ranks=world)
        perm = plan.get("perm_cells")
        owned, ghosts = perm[:plan.N_owned], perm[plan.N_owned:plan.N]
        assert np.array_equal(np.sort(owned), np.nonzero(part == rank)[0])
        assert (part[ghosts] != rank).all()
        peers = plan.get("halo_peers"); rc = plan.get("halo_recv_counts"); flat = plan.get("halo_recv_ids")
        assert np.array_equal(np.sort(flat), np.sort(ghosts))  # every ghost is received exactly once ...
        assert (np.diff(part[flat]) >= 0).all()                # ... in a buffer grouped by ascending owner
        assert np.array_equal(np.unique(part[flat]), peers)
        recv = np.split(flat, np.cumsum(rc)[:-1].astype(np.int64)) if len(rc) else []
        send = exchange_plan(rank, world, peers, recv)
        for p, ids in send.items():
            assert (part[ids] == rank).all()                   # a rank is only asked for cells it owns
        # the union of peers (receive from / send to), ascending, as mlb_halo_set_send_ids lays the buffers out
        allp = sorted(set(int(p) for p in peers) | set(send))
        rcount = [int(rc[list(peers).index(p)]) if p in peers else 0 for p in allp]
        scount = [len(send.get(p, ())) for p in allp]
        # "device state": a global field known to every rank; owners publish theirs, ghosts must receive exactly it
        field = np.random.default_rng(5).normal(size=(nc, 4))
        send_ids = np.concatenate([send[p] for p in allp if p in send]) if send else np.zeros(0, np.int64)
        send_buf = torch.from_numpy(field[send_ids.astype(np.int64)].reshape(-1).copy())
        recv_buf = torch.full((4 * int(sum(rcount)),), np.nan, dtype=torch.float64)
        halo_exchange(send_buf, recv_buf, allp, scount, rcount)
        assert np.array_equal(recv_buf.numpy().reshape(-1, 4), field[flat.astype(np.int64)])
        # dt: all-reduce(max) of the rank-local maxima equals the global maximum
        local = torch.tensor([float(np.abs(field[owned.astype(np.int64)]).max())], dtype=torch.float64)
        dist.all_reduce(local, op=dist.ReduceOp.MAX)
        assert local.item() == np.abs(field).max()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("%d %d" % (plan.N_owned, plan.N - plan.N_owned))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,recon", [(2, "FO"), (2, "TENO"), (3, "TENO"), (3, "TENO:graph")])
def test_halo_plan_and_exchange_over_gloo(tmp_path, world, recon):
    mp.spawn(_worker, args=(world, _free_port(), recon, str(tmp_path)), nprocs=world, join=True)
    got = [tuple(int(x) for x in open(tmp_path / ("ok%d" % r)).read().split()) for r in range(world)]
    assert sum(n for n, _ in got) == 2 * 14 * 10 and all(g > 0 for _, g in got)
    if recon.startswith("TENO"):   # stencil halos are several rings deep, first-order halos one ring
        assert min(g for _, g in got) > 14
