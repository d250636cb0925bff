"""TEST INFRASTRUCTURE - stand-ins that let bench.py / bench_multi.py run their whole control flow on a box without a GPU.

What the driver runs at the end of a round is `python bench.py` (and torchrun ... bench.py): a Python error on a line that has never
executed - a record added late, a key renamed - would cost the round's headline number.  The kernels cannot run here, but every
line of the bench scripts can: `install()` replaces mallard_b200.Solver (and, for N > 1, the NCCL pieces) by objects with the same
call signatures that compute nothing, so the scripts build their JSON lines exactly as they do on the GPU box.  The signatures are
checked against the real classes (inspect), so a call the real Solver would reject fails here too.  Nothing in the product imports
this file."""
import inspect
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class FakeSolver:
    """The part of mallard_b200.Solver the bench scripts use.  A step takes 0.1 ms of pretend device time."""

    MS_PER_STEP = 0.1

    def __init__(self, mesh, *args, **kw):
        import mallard_b200 as mb
        bound = inspect.signature(mb._RealSolver.__init__).bind(self, mesh, *args, **kw)
        bound.apply_defaults()
        p = bound.arguments
        self.mesh, self.nc = mesh, mesh.n_cells
        part, rank = p["part"], p["rank"]
        self.n_owned = int((np.asarray(part) == rank).sum()) if part is not None else self.nc
        self.recon = p["recon"]
        self.n_stages = 3
        self.launch_count, self._replays, self._t, self._ev, self._prof, self._profiling = 0, 0, 0.0, {}, {}, False
        self._clock = 0.0
        self._U = np.zeros((self.nc, 4))
        self.closed = False
        self.stream = self.comm_stream = 0

    def _alive(self):
        assert not self.closed, "call on a closed context"

    def get(self, name):
        self._alive()
        assert name == "stats", name
        return np.array([self.launch_count, 1.5, 3.0e8 if self.nc < 100000 else 6.5e3 * self.nc, self.nc, self.n_owned, 1.5 * self.nc, self.n_owned + 100,
                         self.n_stages, 1.0, 0.1, 0.3, self._replays, 0.0], dtype=np.float64)

    def set_state(self, U, P=None):
        self._alive()
        U = np.asarray(U)
        assert U.shape == (self.nc, 4), U.shape
        assert P is None or np.asarray(P).shape == (self.nc, 5)
        self._U = U.copy()

    def calc_rhs(self):
        self._alive()
        return np.ones((self.nc, 4))

    def get_state(self, prim=False):
        self._alive()
        return self._U.copy()

    def _advance(self, n):
        per_step = 10 if self.recon == "TENO" else 7
        self.launch_count += per_step * n
        self._clock += self.MS_PER_STEP * n
        if self._profiling:
            for k, share in (("teno_stream", 0.8), ("face_flux_teno", 0.12), ("gather_stage", 0.04), ("cfl", 0.04)) if self.recon == "TENO" else \
                    (("face_flux_fo", 0.7), ("gather_stage", 0.2), ("cfl", 0.1)):
                ms, ln = self._prof.get(k, (0.0, 0))
                self._prof[k] = (ms + share * self.MS_PER_STEP * n, ln + (n if k == "cfl" else 3 * n))
        else:
            self._replays += max(0, n - 1)

    def run(self, n_steps, cfl=None, dt=None):
        self._alive()
        self._advance(n_steps)
        self._t += 1e-3 * n_steps
        return self._t, n_steps

    def run_distributed(self, n_steps, cfl):
        return self.run(n_steps, cfl=cfl)

    def take_step_host(self, U, cfl=None, dt=None):
        self._alive()
        assert U.shape == (self.nc, 4) and U.dtype == np.float64
        self._advance(1)
        return 1e-3

    def synchronize(self):
        self._alive()

    def event_record(self, i):
        self._ev[i] = self._clock

    def event_elapsed_ms(self, a, b):
        return self._ev[b] - self._ev[a]

    def profile(self, on):
        self._profiling = bool(on)
        if on:
            self._prof = {}

    def profile_read(self):
        return dict(self._prof)

    def halo_info(self):
        return np.array([1], dtype=np.int32), np.array([1000], dtype=np.uint64), np.array([1000], dtype=np.uint64)

    def time(self):
        return self._t, 0

    def close(self):
        self.closed = True


class FakeDistributedSolver:
    """mallard_b200.parallel.DistributedSolver over a FakeSolver (no communicator, no device buffers)."""

    @classmethod
    def from_solver(cls, solver, rank, world, device, group=None, local=None, native=True):
        return cls(None, None, rank, world, device, group, local, native, solver=solver)

    def __init__(self, mesh, part, rank=None, world=None, device=None, group=None, local=None, native=False, solver=None, **solver_kw):
        from mallard_b200 import parallel
        inspect.signature(parallel._RealDistributedSolver.__init__).bind(self, mesh, part, rank, world, device, group, local, native, solver, **solver_kw)
        self.s = solver if solver is not None else FakeSolver(mesh, part=part, rank=rank, n_ranks=world, device=device, local=local, **solver_kw)
        self.rank, self.world = rank, world
        self.peers, self.send_counts, self.recv_counts = self.s.halo_info()
        self.owned = np.nonzero(np.asarray(part) == rank)[0] if part is not None else np.arange(self.s.nc)
        self.n_stages = 3

    def set_state(self, U, P=None):
        self.s.set_state(U, P)

    def run(self, n_steps, cfl=None):
        return self.s.run(n_steps, cfl=cfl)

    def step_host(self, U_owned, cfl):
        assert U_owned.shape == (len(self.owned), 4)
        self.s._advance(1)


def install(monkeypatch=None, world=1):
    """Patches the GPU-only pieces; with a pytest monkeypatch everything is undone at the end of the test."""
    import torch
    import mallard_b200 as mb
    from mallard_b200 import parallel
    import bench
    import bench_multi

    def setattr_(obj, name, value):
        if monkeypatch is not None:
            monkeypatch.setattr(obj, name, value, raising=False)
        else:
            setattr(obj, name, value)
    if not hasattr(mb, "_RealSolver"):
        setattr_(mb, "_RealSolver", mb.Solver)
        setattr_(parallel, "_RealDistributedSolver", parallel.DistributedSolver)
    setattr_(mb, "Solver", FakeSolver)
    setattr_(parallel, "DistributedSolver", FakeDistributedSolver)
    setattr_(torch.cuda, "is_available", lambda: True)
    setattr_(torch.cuda, "set_device", lambda *a, **k: None)
    real_empty = torch.empty
    setattr_(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: vv for kk, vv in k.items() if kk != "pin_memory"}))
    # the CPU legs are not what is under test: a constant instead of the 20 s reference run
    setattr_(bench, "reference_cpu", lambda n_steps, n_warmup, nx=96, ny=96: (2.0e6, dict(kind="reference", cores=8, sample="mock", ms_per_step=60.0, init_seconds=1.0)))
    # N = 1: the strong-scaling records run in a child process - here a child that carries the same stand-ins
    import functools
    import subprocess
    me = os.path.abspath(__file__)
    real_child = bench.strong_records_in_child
    if not isinstance(real_child, functools.partial):
        setattr_(bench, "strong_records_in_child",
                 functools.partial(real_child, popen=lambda cmd, **kw: subprocess.Popen([cmd[0], me] + list(cmd[2:]), **kw)))
    real_exp = bench_multi.experiment_children
    if not isinstance(real_exp, functools.partial):      # N > 1: the per-rank children of the viscous record carry the stand-ins too
        setattr_(bench_multi, "experiment_children",
                 functools.partial(real_exp, popen=lambda cmd, **kw: subprocess.Popen([cmd[0], me] + list(cmd[2:]), **kw)))
    if world > 1:
        import torch.distributed as dist
        real_init = dist.init_process_group
        setattr_(dist, "init_process_group", lambda backend, device_id=None, timeout=None: real_init("gloo", timeout=timeout))

        def reduce_cpu(x, world_, op="max"):
            t = torch.tensor(np.asarray(x, dtype=np.float64))
            if world_ > 1:
                dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
            return t.numpy()
        setattr_(bench_multi, "_reduce", reduce_cpu)
        setattr_(bench_multi, "bind_to_gpu_numa_node", lambda index: None)


if __name__ == "__main__":       # child of tests/test_bench_flow.py: `python tests/bench_mock.py <bench.py arguments>` under RANK / WORLD_SIZE
    install(world=int(os.environ.get("WORLD_SIZE", "1")))
    import bench
    import bench_multi
    small = os.environ.get("MLB_MOCK_STRONG")
    if small:
        bench_multi.STRONG_MESHES = tuple((n, int(q)) for n, q in (x.split(":") for x in small.split(",")))
    fail = os.environ.get("MLB_MOCK_FAIL_STRONG_ON_RANK")
    if fail is not None:
        real = bench_multi.strong_record

        def failing(name, nq, a, rank, world, *rest, **kw):
            if rank == int(fail):
                raise RuntimeError("mock failure of the strong record on rank %d" % rank)
            time.sleep(float(os.environ.get("MLB_MOCK_PEER_WAIT", "0")))
            return real(name, nq, a, rank, world, *rest, **kw)
        bench_multi.strong_record = failing
    hang = os.environ.get("MLB_MOCK_HANG_TASK")
    if hang and "--child-task" in sys.argv and sys.argv[sys.argv.index("--child-task") + 1] == hang:      # a child that stops inside NCCL
        sys.stderr.write("[mlb comm] rank %s/%s: run: capturing a step\n" % (os.environ.get("RANK"), os.environ.get("WORLD_SIZE")))
        sys.stderr.flush()
        time.sleep(600)
    sys.argv = ["bench.py"] + sys.argv[1:]
    bench.main()
