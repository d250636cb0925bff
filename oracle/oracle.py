"""TEST INFRASTRUCTURE — ctypes front-end of oracle/libmallard_oracle.so (the CPU restatement of the
reference hot path).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MESH_TYPES = {"cartesian": 0, "cartesian_tri": 1, "wedge": 2}
RECON = {"FO": 0, "TENO": 1}
RIEMANN = {"Rusanov": 0, "HLL": 1, "HLLC": 2}
INTEGRATOR = {"FE": 0, "RK4": 1, "SSPRK3": 2}
BC = {"symmetry": 0, "extrapolation": 1, "wall_adiabatic": 2, "upt": 3, "p_out": 4}
BASIS = {"monomial": 0, "legendre": 1}

_DTYPES = {
    "nodes_of_cell": np.uint32, "offsets_nodes_of_cell": np.uint32, "faces_of_cell": np.uint32,
    "offsets_faces_of_cell": np.uint32, "nodes_of_face": np.uint32, "offsets_nodes_of_face": np.uint32,
    "cells_of_face": np.int32, "sizes": np.uint32, "teno:poly_indices": np.uint8,
    "teno:offsets_stencil_groups": np.uint32, "teno:offsets_stencils": np.uint32, "teno:stencils": np.uint32,
    "teno:offsets_reconstruction_matrices": np.uint32,
}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libmallard_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmallard_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_mesh_generate.restype = C.c_void_p
        L.orc_mesh_generate.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        L.orc_mesh_from_arrays.restype = C.c_void_p
        L.orc_mesh_from_arrays.argtypes = [C.c_uint32] * 3 + [C.c_void_p] * 8
        L.orc_mesh_add_zone.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32]
        L.orc_mesh_free.argtypes = [C.c_void_p]
        L.orc_solver_create.restype = C.c_void_p
        L.orc_solver_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_double, C.c_int, C.c_int]
        L.orc_solver_free.argtypes = [C.c_void_p]
        L.orc_solver_add_bc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        L.orc_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_rhs_override.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_teno_fixed.argtypes = [C.c_void_p, C.c_int]
        L.orc_calc_face_values.argtypes = [C.c_void_p]
        L.orc_calc_rhs.argtypes = [C.c_void_p]
        L.orc_calc_dt.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.orc_take_step.argtypes = [C.c_void_p, C.c_double]
        L.orc_get.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.orc_riemann_flux.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_prims.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc:
        raise RuntimeError(lib().orc_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _get(mh, sh, name):
    L = lib()
    nb = C.c_uint64(0)
    _check(L.orc_get(mh, sh, name.encode(), None, C.byref(nb)))
    dt = np.dtype(np.uint32 if name.startswith("zone:") else _DTYPES.get(name, np.float64))
    out = np.empty(nb.value // dt.itemsize, dtype=dt)
    if nb.value:
        _check(L.orc_get(mh, sh, name.encode(), _ptr(out), C.byref(nb)))
    return out


class Mesh:
    """Mirror of the reference Mesh data (src/mesh/mesh.h:228-253) held by the oracle."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(lib().orc_last_error().decode())
        self.h = handle
        self.zone_names = []
        self.n_cells, self.n_faces, self.n_nodes = (int(x) for x in _get(self.h, None, "sizes"))

    @classmethod
    def generate(cls, mtype, nx, ny, Lx=1.0, Ly=1.0):
        m = cls(lib().orc_mesh_generate(MESH_TYPES[mtype], nx, ny, Lx, Ly))
        m.zone_names = ["interior", "right", "top", "left", "bottom"]
        return m

    @classmethod
    def from_arrays(cls, d, zones):
        """d: dict with node_coords, offsets_nodes_of_cell, nodes_of_cell, offsets_faces_of_cell, faces_of_cell,
        offsets_nodes_of_face, nodes_of_face, cells_of_face; zones: list of (name, faces)."""
        a = {k: np.ascontiguousarray(d[k]) for k in d}
        nc = len(a["offsets_nodes_of_cell"]) - 1
        nf = len(a["offsets_nodes_of_face"]) - 1
        nn = a["node_coords"].shape[0]
        m = cls(lib().orc_mesh_from_arrays(
            nc, nf, nn, _ptr(a["node_coords"].astype(np.float64)),
            _ptr(a["offsets_nodes_of_cell"].astype(np.uint32)), _ptr(a["nodes_of_cell"].astype(np.uint32)),
            _ptr(a["offsets_faces_of_cell"].astype(np.uint32)), _ptr(a["faces_of_cell"].astype(np.uint32)),
            _ptr(a["offsets_nodes_of_face"].astype(np.uint32)), _ptr(a["nodes_of_face"].astype(np.uint32)),
            _ptr(a["cells_of_face"].astype(np.int32))))
        for name, faces in zones:
            f = np.ascontiguousarray(faces, dtype=np.uint32)
            lib().orc_mesh_add_zone(m.h, name.encode(), _ptr(f), len(f))
            m.zone_names.append(name)
        return m

    def get(self, name):
        a = _get(self.h, None, name)
        if name in ("node_coords", "cell_coords", "face_normals", "cells_of_face"):
            a = a.reshape(-1, 2)
        return a

    def zone(self, name):
        return _get(self.h, None, "zone:" + name)

    def arrays(self):
        keys = ["node_coords", "cell_coords", "cell_volume", "face_area", "face_normals", "nodes_of_cell",
                "offsets_nodes_of_cell", "faces_of_cell", "offsets_faces_of_cell", "nodes_of_face",
                "offsets_nodes_of_face", "cells_of_face"]
        return {k: self.get(k) for k in keys}

    def __del__(self):
        try:
            lib().orc_mesh_free(self.h)
        except Exception:
            pass


DEFAULT_GAS = dict(gamma=1.4, p_ref=101325.0, T_ref=298.15, rho_ref=1.225, p_min=-1e20, p_max=1e20)


def gas6(gas=None):
    g = dict(DEFAULT_GAS)
    g.update(gas or {})
    return np.array([g["gamma"], g["p_ref"], g["T_ref"], g["rho_ref"], g["p_min"], g["p_max"]], dtype=np.float64)


class Solver:
    """Mirror of the reference Solver's hot-path surface (src/solver/solver.h:47-108)."""

    def __init__(self, mesh, recon="FO", riemann="HLLC", integrator="SSPRK3", gas=None, basis="legendre", order=3,
                 factor=2.0, quad_cell_order=0, quad_face_order=0, bcs=(), teno_fixed=False):
        self.mesh = mesh
        g = gas6(gas)
        self.h = lib().orc_solver_create(mesh.h, RECON[recon], RIEMANN[riemann], INTEGRATOR[integrator], _ptr(g),
                                         BASIS[basis], order, factor, quad_cell_order, quad_face_order)
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        for bc in bcs:
            self.add_bc(**bc)
        if teno_fixed:
            lib().orc_set_teno_fixed(self.h, 1)

    def add_bc(self, name, type, u=(0.0, 0.0), p=0.0, T=0.0):
        data = np.zeros(4)
        if type == "upt":
            data[:] = [u[0], u[1], p, T]
        elif type == "p_out":
            data[0] = p
        _check(lib().orc_solver_add_bc(self.h, BC[type], name.encode(), _ptr(data)))

    def set_state(self, U, P=None):
        U = np.ascontiguousarray(U, dtype=np.float64)
        Pp = None if P is None else _ptr(np.ascontiguousarray(P, dtype=np.float64))
        _check(lib().orc_set_state(self.h, _ptr(U), Pp))

    def set_rhs_override(self, r):
        _check(lib().orc_set_rhs_override(self.h, None if r is None else _ptr(np.ascontiguousarray(r, dtype=np.float64))))

    def calc_face_values(self):
        _check(lib().orc_calc_face_values(self.h))
        return self.get("F")

    def calc_rhs(self):
        _check(lib().orc_calc_rhs(self.h))
        return self.get("rhs0")

    def calc_dt(self, cfl):
        dt = C.c_double(0)
        _check(lib().orc_calc_dt(self.h, cfl, C.byref(dt)))
        return dt.value

    def take_step(self, dt):
        _check(lib().orc_take_step(self.h, dt))

    def get(self, name):
        a = _get(None, self.h, name)
        nc, nf = self.mesh.n_cells, self.mesh.n_faces
        if name in ("U", "U_temp") or name.startswith("rhs"):
            a = a.reshape(nc, 4)
        elif name == "P":
            a = a.reshape(nc, 5)
        elif name == "F":
            a = a.reshape(nf, -1, 2, 4)
        elif name == "teno:poly_indices":
            a = a.reshape(-1, 2)
        return a

    def __del__(self):
        try:
            lib().orc_solver_free(self.h)
        except Exception:
            pass


def riemann_flux(kind, n_unit, L, R, gamma=1.4):
    """L, R: [n][5] rows (rho, u, v, p, h) — RiemannSolver::calc_flux arguments (riemann_solver.h:85-90)."""
    n_unit = np.ascontiguousarray(n_unit, dtype=np.float64).reshape(-1, 2)
    L = np.ascontiguousarray(L, dtype=np.float64).reshape(-1, 5)
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 5)
    out = np.empty((L.shape[0], 4))
    _check(lib().orc_riemann_flux(RIEMANN[kind], L.shape[0], _ptr(n_unit), _ptr(L), _ptr(R), gamma, _ptr(out)))
    return out


def prims(U, gas=None):
    U = np.ascontiguousarray(U, dtype=np.float64).reshape(-1, 4)
    P = np.empty((U.shape[0], 5))
    rc = np.empty(3)
    g = gas6(gas)
    _check(lib().orc_prims(_ptr(g), U.shape[0], _ptr(U), _ptr(P), _ptr(rc)))
    return P, rc
