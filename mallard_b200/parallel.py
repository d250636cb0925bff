"""Multi-GPU execution of the explicit residual path (new; the reference is single-process — SURVEY §8e).

One process per GPU.  The mesh is partitioned over the ranks (mallard_b200.partition, or any caller-supplied vector); each
rank's context owns its cells plus ghost copies of every remote cell its residual stencils read.  Per RK stage the ghost
conserved states (4 doubles per cell) are exchanged peer to peer, per step one double is all-reduced (max) for dt:

    communication stream:  pack -> send/recv -> unpack                       (per stage)
    compute stream:        reconstruction of the INTERIOR cells | wait for unpack | rim cells, fluxes, residual + RK update
                           (stage 0 also: local max spectral radius -> all_reduce(max) -> dt, between the two halves)

The preprocessor numbers the owned cells whose TENO stencils contain no ghost first, so the bulk of a stage (the
table-streaming reconstruction kernel) runs while the ghost states are in flight.

Two drivers of the same schedule:
  * native=True (what bench.py measures): the library's own NCCL driver (mlb_comm_init / mlb_run_distributed, csrc/api.cu) —
    grouped ncclSend/ncclRecv between the library's device buffers, ncclAllReduce(max) of the device-resident spectral radius,
    the whole step captured once and replayed as a CUDA graph: no host code between the stages of a step.  torch.distributed
    only ships the 128-byte NCCL unique id to the ranks.
  * native=False: the split-phase C ABI (mlb_halo_pack / mlb_stage_begin / mlb_stage ...) driven from here with
    torch.distributed as the communicator — the same calls a host with its own communicator (MPI) would make; the exchange
    *plan* and the packing order are backend independent and are exercised with gloo in the CPU tests.

`local=` takes a rank-local mesh (local_mesh.extract_local / synthetic.jittered_tri_local): no rank ever holds the global mesh.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import Solver


def exchange_plan(rank, world, peers, recv_ids, group=None):
    """Every rank announces, per peer, the reference ids of the ghost cells it receives from that peer; returns
    {peer: ids this rank must send to it, in the order the peer's receive buffer expects}."""
    wants = {int(p): np.ascontiguousarray(ids, dtype=np.uint32) for p, ids in zip(peers, recv_ids)}
    gathered = [None] * world
    dist.all_gather_object(gathered, wants, group=group)
    send = {}
    for r in range(world):
        if r == rank:
            continue
        ids = gathered[r].get(rank)
        if ids is not None and len(ids):
            send[r] = ids
    return send


def halo_exchange(send_buf, recv_buf, peers, send_counts, recv_counts, group=None, width=4):
    """Moves `send_buf` (peers in ascending order, send_counts[i]*width values each) to the peers and fills `recv_buf`.
    Works on CPU tensors (gloo) and CUDA tensors (NCCL: enqueued on the current stream, the host does not wait)."""
    ops, so, ro = [], 0, 0
    for p, sc, rc in zip(peers, send_counts, recv_counts):
        sc, rc = int(sc) * width, int(rc) * width
        if rc:
            ops.append(dist.P2POp(dist.irecv, recv_buf[ro:ro + rc], int(p), group))
        if sc:
            ops.append(dist.P2POp(dist.isend, send_buf[so:so + sc], int(p), group))
        so += sc
        ro += rc
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class _DeviceBuffer:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, n_doubles, device):
    if not ptr or n_doubles == 0:
        return torch.empty(0, dtype=torch.float64, device="cuda:%d" % device)
    return torch.as_tensor(_DeviceBuffer(ptr, n_doubles), device="cuda:%d" % device)


class DistributedSolver:
    """Solver::run's loop over a partitioned mesh.  `part` is the partition vector in reference numbering."""

    @classmethod
    def from_solver(cls, solver, rank, world, device, group=None, local=None, native=True):
        """Wraps a partitioned Solver that already exists (its creation may have to be retried collectively)."""
        return cls(None, None, rank, world, device, group, local, native, solver=solver)

    def __init__(self, mesh, part, rank=None, world=None, device=None, group=None, local=None, native=False, solver=None, **solver_kw):
        """mesh / part: the global mesh and partition vector, or (local=dict(global_ids, n_global, cell0_nodes)) this rank's
        part of the mesh and the owners of ITS cells."""
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = torch.cuda.current_device() if device is None else device
        self.group = group
        self.native = native
        self.local = local
        self.s = solver if solver is not None else Solver(mesh, part=part, rank=self.rank, n_ranks=self.world, device=self.device, local=local, **solver_kw)
        s = self.s
        if native:
            from . import comm_unique_id
            box = [comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)       # 128 bytes; everything after this is inside the library
            s.comm_init(box[0])
        else:
            peers, _, rc = s.halo_info()
            recv = [s.halo_recv_ids(i, rc[i]) for i in range(len(peers))]      # ids of the mesh, or global ids (local ingest)
            send = exchange_plan(self.rank, self.world, peers, recv, group)
            send_peers = sorted(send)
            s.halo_set_send_ids(send_peers, [send[p] for p in send_peers])
        self.peers, self.send_counts, self.recv_counts = s.halo_info()
        sp, rp = s.halo_buffers()
        self.send_t = device_tensor(sp, 4 * int(self.send_counts.sum()), self.device)
        self.recv_t = device_tensor(rp, 4 * int(self.recv_counts.sum()), self.device)
        self.scal_t = device_tensor(s.scalars_device(), 8, self.device)
        self.stream = torch.cuda.ExternalStream(s.stream, device=self.device)
        self.comm = torch.cuda.ExternalStream(s.comm_stream, device=self.device)
        self.n_stages = s.n_stages
        self.owned = s.owned_cells()

    # -- state
    def set_state(self, U, P=None):
        """Global arrays in reference numbering (every rank passes the same arrays; only held cells are kept)."""
        self.s.set_state(U, P)

    def exchange(self, stage):
        s = self.s
        s.halo_pack(stage)
        with torch.cuda.stream(self.comm):
            halo_exchange(self.send_t, self.recv_t, self.peers, self.send_counts, self.recv_counts, self.group)
        s.halo_unpack(stage)
        s.stage_begin(stage)       # interior cells: overlaps the exchange just enqueued

    def step(self, cfl=None):
        """One time step; cfl=None keeps the dt set by set_dt.  Fully asynchronous on the device."""
        s = self.s
        self.exchange(0)
        if cfl is not None and cfl > 0:
            s.local_max_spectral_radius(sync=False)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.scal_t[2:3], op=dist.ReduceOp.MAX, group=self.group)
            s.apply_dt_device(cfl)
        s.stage(0)
        for st in range(1, self.n_stages):
            self.exchange(st)
            s.stage(st)

    def run(self, n_steps, cfl=None):
        if self.native:
            self.s.run_distributed(n_steps, cfl if cfl is not None else 0.0)
            return self.s.time()
        for _ in range(n_steps):
            self.step(cfl)
        self.s.finish_step()   # synchronises the compute stream
        return self.s.time()

    def step_host(self, U_owned, cfl):
        """The take_step seam with HOST buffers for the rank's own cells: H2D of U_owned ([n_owned][4], owned_cells order),
        one step (halo exchanges and the dt all-reduce included), D2H of the result into U_owned."""
        if self.native:
            return self.s.take_step_distributed_host(U_owned, cfl)[0]
        self.s.set_owned(U_owned)
        self.step(cfl)
        return self.s.get_owned(U_owned)

    def gather_state(self):
        """Global state in reference numbering on every rank (owned cells of all ranks combined)."""
        U = self.s.get_state()                        # zeros outside the owned cells
        if self.local is not None:                    # local-mesh numbering -> global numbering
            G = np.zeros((int(self.local["n_global"]), 4))
            G[np.asarray(self.local["global_ids"], dtype=np.int64)] = U
            U = G
        U = torch.from_numpy(U)
        if self.world > 1:
            Ud = U.cuda(self.device)
            dist.all_reduce(Ud, op=dist.ReduceOp.SUM, group=self.group)
            U = Ud.cpu()
        return U.numpy()
