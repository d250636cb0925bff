"""TEST INFRASTRUCTURE - ctypes front end of tests/emul/kernel_emulation.cpp (the CUDA kernel source compiled for the host and
executed thread by thread; see that file).  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

import mallard_b200 as mb
from mallard_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = {"strict": os.path.join(HERE, "libkernel_emulation.so"), "fast": os.path.join(HERE, "libkernel_emulation_fast.so")}
# STRICT: the source as kernels_strict.cu compiles it, no contraction.  FAST: as kernels_fast.cu compiles it (MLB_STREAM_KERNELS: lean
# Riemann flux, reciprocals, monomials by multiplication) with the host compiler contracting a * b + c to FMA - not nvcc's choice of
# contractions instruction by instruction, but the same class of re-rounding: what it checks is that the FAST-mode tolerances of the
# tests hold for kernels that have not run on hardware yet.
FLAGS = {"strict": ["-ffp-contract=off"], "fast": ["-DMLB_STREAM_KERNELS=1", "-ffp-contract=fast", "-mfma"]}
_LIB = {}


def fast_available():
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


def build(variant="strict"):
    src = os.path.join(HERE, "kernel_emulation.cpp")
    deps = [src] + [os.path.join(ROOT, "mallard_b200", "csrc", f) for f in ("kernels_impl.cuh", "teno_generic.cuh", "small_step.cuh", "stage_plan.h",
                                                                             "kernel_args.h", "mlb_internal.h")]
    deps.append(os.path.join(ROOT, "mallard_b200", "libmallard_b200.so"))
    so = SO[variant]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return
    mb.build()
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2"] + FLAGS[variant] + ["-fPIC", "-shared", "-I", os.path.join(ROOT, "mallard_b200", "csrc"),
                           src, "-L", os.path.join(ROOT, "mallard_b200"), "-lmallard_b200", "-Wl,-rpath," + os.path.join(ROOT, "mallard_b200"), "-o", so])


def lib(variant="strict"):
    if variant not in _LIB:
        build(variant)
        mb.lib()
        L = C.CDLL(SO[variant])
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.POINTER(_abi.MeshView), C.POINTER(_abi.Numerics), C.POINTER(_abi.Physics), C.POINTER(_abi.Bc), C.c_int, C.c_void_p,
                                 C.c_int, C.c_int, C.c_void_p]
        L.emu_n_owned.argtypes = [C.c_void_p]
        L.emu_n_held.argtypes = [C.c_void_p]
        L.emu_last_error.restype = C.c_char_p
        for n in ("emu_set_state", "emu_face_values", "emu_rhs", "emu_gradients"):
            getattr(L, n).restype = C.c_int
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_force_generic.argtypes = [C.c_void_p, C.c_int]
        L.emu_n_quad.argtypes = [C.c_void_p]
        L.emu_run.argtypes = [C.c_void_p, C.c_uint, C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.emu_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_step_count.restype = C.c_ulonglong
        L.emu_step_count.argtypes = [C.c_void_p]
        L.emu_get_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.emu_time.restype = C.c_double
        L.emu_time.argtypes = [C.c_void_p]
        L.emu_calc_dt.restype = C.c_double
        L.emu_calc_dt.argtypes = [C.c_void_p, C.c_double]
        L.emu_get_primitives.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_set_primitives.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_primitives.argtypes = [C.POINTER(_abi.Physics), C.c_ulonglong, C.c_void_p, C.c_void_p]
        L.emu_riemann_flux.argtypes = [C.c_int, C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _LIB[variant] = L
    return _LIB[variant]


def riemann_flux(kind, n_unit, L, R, gamma=1.4, fp_mode="strict"):
    """mallard_b200.riemann_flux through the emulated kernel source: rows (rho, u, v, p, h), kind in Rusanov / HLL / HLLC."""
    n_unit = np.ascontiguousarray(n_unit, dtype=np.float64).reshape(-1, 2)
    L = np.ascontiguousarray(L, dtype=np.float64).reshape(-1, 5)
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 5)
    out = np.empty((L.shape[0], 4))
    p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    lib(fp_mode).emu_riemann_flux(mb.RIEMANN[kind], L.shape[0], p(n_unit), p(L), p(R), float(gamma), p(out))
    return out


def primitives(U, gas=None, fp_mode="strict"):
    """mallard_b200.compute_primitives through the emulated kernel source: U [n][4] -> (u, v, p, T, h) [n][5]"""
    U = np.ascontiguousarray(U, dtype=np.float64).reshape(-1, 4)
    P = np.empty((U.shape[0], 5))
    phys = mb._physics(gas)
    lib(fp_mode).emu_primitives(C.byref(phys), U.shape[0], U.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p))
    return P


class EmulatedSolver:
    """The calc_face_values / calc_rhs surface of mallard_b200.Solver, computed by the emulated kernels (fp_mode: see FLAGS)."""

    def __init__(self, mesh, recon="FO", riemann="HLLC", integrator="SSPRK3", gas=None, basis="legendre", order=3, factor=2.0, quad_cell_order=0,
                 quad_face_order=0, bcs=(), teno_fixed=False, renumber="rcm", part=None, rank=0, n_ranks=1, fp_mode="strict", local=None):
        """local: `mesh` is a rank-local mesh (local_mesh.extract_local / synthetic.jittered_tri_local) and `part` the owners of ITS cells."""
        self.mesh = mesh
        self._L = lib(fp_mode)
        num = mb._numerics(recon, riemann, integrator, basis, order, factor, quad_cell_order, quad_face_order, "strict", renumber, teno_fixed, True)
        phys = mb._physics(gas)
        self._keep = []
        cb = (_abi.Bc * max(1, len(bcs)))()
        for i, b in enumerate(bcs):
            nb = b["name"].encode()
            self._keep.append(nb)
            cb[i].zone_name, cb[i].type = nb, mb.BC[b["type"]]
            u = b.get("u", (0.0, 0.0))
            cb[i].u[0], cb[i].u[1] = float(u[0]), float(u[1])
            cb[i].p, cb[i].T = float(b.get("p", 0.0)), float(b.get("T", 0.0))
        v, keep = mesh.view()
        self._keep.append(keep)
        pp = None if part is None else np.ascontiguousarray(part, dtype=np.int32)
        self._keep.append(pp)
        c0 = None if local is None else np.ascontiguousarray(local["cell0_nodes"], dtype=np.float64)
        self._keep.append(c0)
        self._h = self._L.emu_create(C.byref(v), C.byref(num), C.byref(phys), cb, len(bcs), None if pp is None else pp.ctypes.data_as(C.c_void_p), rank, n_ranks,
                                     None if c0 is None else c0.ctypes.data_as(C.c_void_p))
        if not self._h:
            raise RuntimeError(self._L.emu_last_error().decode())
        self.n_quad = self._L.emu_n_quad(self._h)
        self.n_owned, self.n_held = self._L.emu_n_owned(self._h), self._L.emu_n_held(self._h)

    def _ok(self, rc):
        if rc:
            raise RuntimeError(self._L.emu_last_error().decode())

    def force_generic(self, on=True):
        self._L.emu_force_generic(self._h, int(on))

    def set_state(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.shape == (self.mesh.n_cells, 4)
        self._ok(self._L.emu_set_state(self._h, U.ctypes.data_as(C.c_void_p)))

    def set_primitives(self, P):
        """the P of mlb_set_state(U, P): primitives [n_cells][5] the stepping state starts from instead of those recomputed from U"""
        P = np.ascontiguousarray(P, dtype=np.float64)
        assert P.shape == (self.mesh.n_cells, 5)
        self._ok(self._L.emu_set_primitives(self._h, P.ctypes.data_as(C.c_void_p)))

    def calc_face_values(self):
        F = np.empty((self.mesh.n_faces, self.n_quad, 2, 4))
        self._ok(self._L.emu_face_values(self._h, F.ctypes.data_as(C.c_void_p)))
        return F

    def calc_rhs(self):
        r = np.zeros((self.mesh.n_cells, 4))
        self._ok(self._L.emu_rhs(self._h, r.ctypes.data_as(C.c_void_p)))
        return r

    def gradients(self):
        G = np.zeros((self.mesh.n_cells, 6))
        self._ok(self._L.emu_gradients(self._h, G.ctypes.data_as(C.c_void_p)))
        return G

    def run(self, n_steps, cfl=None, dt=None, small_blocks=0):
        """n_steps time steps (first-order contexts): the multi-kernel sequence of mlb_run, or - small_blocks > 0 - the phases of the
        cooperative small-mesh kernel on a grid of that many blocks.  Returns (t, dt of the last step)."""
        t, d = C.c_double(), C.c_double()
        self._ok(self._L.emu_run(self._h, int(n_steps), float(cfl or 0.0), float(dt or 0.0), int(small_blocks), C.byref(t), C.byref(d)))
        return t.value, d.value

    def get_state(self):
        U = np.zeros((self.mesh.n_cells, 4))
        self._ok(self._L.emu_get_state(self._h, U.ctypes.data_as(C.c_void_p)))
        return U

    def get_array(self, name):
        out = np.zeros(self.mesh.n_cells if name == "cfl_local" else (self.mesh.n_cells, 4))
        self._ok(self._L.emu_get_array(self._h, name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    def calc_dt(self, cfl):
        return float(self._L.emu_calc_dt(self._h, float(cfl)))

    def get_primitives(self):
        P = np.zeros((self.mesh.n_cells, 5))
        self._ok(self._L.emu_get_primitives(self._h, P.ctypes.data_as(C.c_void_p)))
        return P

    def time(self):
        return float(self._L.emu_time(self._h))

    def step_count(self):
        return int(self._L.emu_step_count(self._h))

    def close(self):
        if self._h:
            self._L.emu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EmulatedAsSolver:
    """mallard_b200.Solver's surface, as far as the gated GPU tests of tests/test_gpu_parity.py use it, computed by the emulated kernels: lets
    those tests - the very functions, with their tolerances - run on the host before they run on hardware (tests/
    test_gated_gpu_tests_on_the_emulator.py).  Reads MLB_TENO_GENERIC / MLB_SMALL_STEP / MLB_SMALL_STEP_BLOCKS per call like the library."""

    def __init__(self, mesh, recon="FO", riemann="HLLC", integrator="SSPRK3", gas=None, basis="legendre", order=3, factor=2.0, quad_cell_order=0,
                 quad_face_order=0, bcs=(), fp_mode="strict", renumber="rcm", teno_fixed=False, keep_stage_rhs=True, device=0, part=None, rank=0,
                 n_ranks=1, local=None):
        if part is not None or local is not None:
            raise NotImplementedError("the emulated Solver is a single context")
        self.mesh, self.fp_mode = mesh, fp_mode
        self._e = EmulatedSolver(mesh, recon, riemann, integrator, gas=gas, basis=basis, order=order, factor=factor, quad_cell_order=quad_cell_order,
                                 quad_face_order=quad_face_order, bcs=bcs, teno_fixed=teno_fixed, renumber=renumber, fp_mode=fp_mode)
        self.n_quad, self.n_stages = self._e.n_quad, {"FE": 1, "RK4": 4, "SSPRK3": 3}[integrator]
        self._small_ok = recon == "FO" and integrator != "FE" and not (gas or {}).get("mu", 0.0) > 0      # small_step_eligible (api.cu)
        self._dirty, self._dt, self._small_steps, self.launch_count = False, -1.0, 0, 0

    def _generic(self):
        self._e.force_generic(os.environ.get("MLB_TENO_GENERIC") == "1")

    def _current(self):                 # the residual / face-value entry points work on the state the last step left
        if self._dirty:
            self._e.set_state(self._e.get_state())
            self._dirty = False

    def set_state(self, U, P=None):
        self._e.set_state(U)
        if P is not None:
            self._e.set_primitives(P)
        self._dirty = False

    def calc_face_values(self):
        self._generic(); self._current()
        return self._e.calc_face_values()

    def calc_rhs(self):
        self._generic(); self._current()
        return self._e.calc_rhs()

    def calc_dt(self, cfl):
        self._dt = self._e.calc_dt(cfl)
        if self._dt < 0:
            raise mb.MallardError("dt negative: %f." % self._dt)
        return self._dt

    def take_step(self, dt=None):
        self._generic()
        self._e.run(1, dt=self._dt if dt is None else dt)
        self._dirty = True

    def run(self, n_steps, cfl=0.0):
        self._generic()
        small = os.environ.get("MLB_SMALL_STEP") == "1" and self._small_ok
        blocks = (int(os.environ.get("MLB_SMALL_STEP_BLOCKS", "0") or 0) or 4) if small else 0
        t, dt = self._e.run(n_steps, cfl=cfl if cfl > 0 else None, dt=None if cfl > 0 else self._dt, small_blocks=blocks)
        self._small_steps += n_steps if small else 0
        self._dirty, self._dt = True, dt
        if dt < 0:
            raise mb.MallardError("dt negative: %f." % dt)
        return t, dt

    def get_state(self, prim=False, cfl_local=False):
        U = self._e.get_state()
        if not prim and not cfl_local:
            return U
        out = [U]
        if prim:
            out.append(self._e.get_primitives())
        if cfl_local:
            out.append(self._e.get_array("cfl_local"))
        return tuple(out)

    def get(self, name):
        if name == "stats":
            return np.array([0.0] * 12 + [float(self._small_steps)])
        return self._e.get_array(name)

    def time(self):
        return self._e.time(), self._e.step_count()

    def synchronize(self):
        pass

    def close(self):
        self._e.close()
