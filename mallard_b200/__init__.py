"""mallard_b200 — host-side mirror of the reference's hot-path interface over the C ABI (include/mallard_b200.h).

`Mesh` mirrors the data surface of the reference Mesh (src/mesh/mesh.h:228-253; generators src/mesh/mesh.cpp:305-848) and
`Solver` the hot-path surface of the reference Solver (src/solver/solver.h:47-108): calc_rhs / calc_dt / take_step /
update state, with the same names and argument meaning.  All computation happens in libmallard_b200.so (CUDA, sm_100a);
nothing here falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import BASIS, BC, FP, INTEGRATOR, MESH, RECON, RENUMBER, RIEMANN, build, lib  # noqa: F401

__all__ = ["Mesh", "Solver", "Plan", "riemann_flux", "compute_primitives", "partition", "partition_coords", "partition_graph", "comm_unique_id",
           "set_host_threads", "build", "lib"]
COMM_ID_BYTES = 128

DEFAULT_GAS = dict(gamma=1.4, p_ref=101325.0, T_ref=298.15, rho_ref=1.225, p_min=-1e20, p_max=1e20, mu=0.0, Pr=0.72)
_MESH_KEYS = ["node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell",
              "offsets_nodes_of_face", "nodes_of_face", "cells_of_face", "cell_coords", "cell_volume", "face_area",
              "face_normals"]
_MESH_DTYPES = dict(node_coords=np.float64, cell_coords=np.float64, cell_volume=np.float64, face_area=np.float64,
                    face_normals=np.float64, cells_of_face=np.int32)


class MallardError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _last_error(h=None):
    return lib().mlb_last_error(h).decode()


class Mesh:
    """Mesh arrays in the reference's numbering."""

    def __init__(self):
        self._h = None
        self.arrays = {}
        self.zones = []   # [(name, uint32 faces)]

    # -- Mesh::init (src/mesh/mesh.cpp:32-66) for the generated types
    @classmethod
    def generate(cls, mtype, Nx=100, Ny=100, Lx=1.0, Ly=1.0):
        if mtype not in MESH:
            raise MallardError("Unknown mesh type: %s." % mtype)
        m = cls()
        h = C.c_void_p()
        if lib().mlb_host_mesh_generate(C.byref(h), MESH[mtype], Nx, Ny, Lx, Ly):
            raise MallardError(_last_error())
        m._take(h)
        return m

    def _take(self, h):
        """Copies the arrays of a library-side host mesh into numpy and frees it."""
        m = self
        m._h = h
        v = _abi.MeshView()
        if lib().mlb_host_mesh_view(h, C.byref(v)):
            raise MallardError(_last_error())
        nc, nf, nn = v.n_cells, v.n_faces, v.n_nodes

        def arr(p, n, dt):
            if not p or n == 0:                      # an empty zone (the interior zone of a one-cell mesh) has no storage
                return np.empty(0, dtype=dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
        onc = arr(v.offsets_nodes_of_cell, nc + 1, np.uint32)
        ofc = arr(v.offsets_faces_of_cell, nc + 1, np.uint32)
        onf = arr(v.offsets_nodes_of_face, nf + 1, np.uint32)
        m.arrays = dict(
            node_coords=arr(v.node_coords, 2 * nn, np.float64).reshape(nn, 2),
            offsets_nodes_of_cell=onc, nodes_of_cell=arr(v.nodes_of_cell, int(onc[-1]), np.uint32),
            offsets_faces_of_cell=ofc, faces_of_cell=arr(v.faces_of_cell, int(ofc[-1]), np.uint32),
            offsets_nodes_of_face=onf, nodes_of_face=arr(v.nodes_of_face, int(onf[-1]), np.uint32),
            cells_of_face=arr(v.cells_of_face, 2 * nf, np.int32).reshape(nf, 2),
            cell_coords=arr(v.cell_coords, 2 * nc, np.float64).reshape(nc, 2),
            cell_volume=arr(v.cell_volume, nc, np.float64), face_area=arr(v.face_area, nf, np.float64),
            face_normals=arr(v.face_normals, 2 * nf, np.float64).reshape(nf, 2))
        m.zones = []
        for i in range(v.n_zones):
            z = v.zones[i]
            m.zones.append((z.name.decode(), arr(z.faces, z.n_faces, np.uint32)))
        lib().mlb_host_mesh_free(h)
        m._h = None

    def compute_geometry(self):
        """Cell centroids / volumes, face areas / normals from connectivity and node coordinates
        (Mesh::compute_* , src/mesh/mesh.cpp:167-261), through the library's host code."""
        v, keep = self.view()
        h = C.c_void_p()
        if lib().mlb_host_mesh_from_arrays(C.byref(h), C.byref(v)):
            raise MallardError(_last_error())
        self._take(h)
        return self

    @classmethod
    def from_arrays(cls, arrays, zones):
        """arrays: connectivity (+ optional geometry) keyed as in the reference Mesh; zones: [(name, faces)]."""
        m = cls()
        for k in _MESH_KEYS:
            if k in arrays and arrays[k] is not None:
                m.arrays[k] = np.ascontiguousarray(arrays[k], dtype=_MESH_DTYPES.get(k, np.uint32))
        m.zones = [(n, np.ascontiguousarray(f, dtype=np.uint32)) for n, f in zones]
        return m

    @classmethod
    def from_cells(cls, node_xy, offsets_nodes_of_cell, nodes_of_cell, boundary_edges=None, edge_tags=None, names=None):
        """Mesh from cell-node lists (triangles / quadrilaterals, counter-clockwise): faces are the unique edges, tagged boundary
        edges ([n][2] node pairs + tags, names = {tag: zone name}) name the zones - mlb_host_mesh_from_cells."""
        xy = np.ascontiguousarray(node_xy, dtype=np.float64).reshape(-1, 2)
        onc = np.ascontiguousarray(offsets_nodes_of_cell, dtype=np.uint32)
        noc = np.ascontiguousarray(nodes_of_cell, dtype=np.uint32)
        be = np.ascontiguousarray(boundary_edges if boundary_edges is not None else np.zeros((0, 2)), dtype=np.uint32).reshape(-1, 2)
        bt = np.ascontiguousarray(edge_tags if edge_tags is not None else np.zeros(0), dtype=np.int32)
        names = dict(names or {})
        tags = np.array(sorted(names), dtype=np.int32)
        cn = (C.c_char_p * max(1, len(tags)))(*[names[int(t)].encode() for t in tags])
        h = C.c_void_p()
        if lib().mlb_host_mesh_from_cells(C.byref(h), len(xy), _ptr(xy), len(onc) - 1, _ptr(onc), _ptr(noc), len(be), _ptr(be), _ptr(bt),
                                          len(tags), _ptr(tags), cn):
            raise MallardError(_last_error())
        m = cls()
        m._take(h)
        return m

    @property
    def n_cells(self):
        return len(self.arrays["offsets_nodes_of_cell"]) - 1

    @property
    def n_faces(self):
        return len(self.arrays["offsets_nodes_of_face"]) - 1

    @property
    def n_nodes(self):
        return self.arrays["node_coords"].shape[0]

    def get_face_zone(self, name):
        for n, f in self.zones:
            if n == name:
                return f
        return None

    def view(self):
        """Returns (MeshView, keepalive) for the C ABI."""
        a = self.arrays
        zs = (_abi.Zone * max(1, len(self.zones)))()
        keep = [zs]
        for i, (n, f) in enumerate(self.zones):
            nb = n.encode()
            keep.append(nb)
            zs[i].name = nb
            zs[i].n_faces = len(f)
            zs[i].faces = _ptr(f)
        v = _abi.MeshView()
        v.n_cells, v.n_faces, v.n_nodes = self.n_cells, self.n_faces, self.n_nodes
        for k in _MESH_KEYS:
            setattr(v, k, _ptr(a.get(k)))
        v.n_zones = len(self.zones)
        v.zones = zs
        return v, keep


def set_host_threads(n):
    """OpenMP threads of the host preprocessor (launchers such as torchrun preset OMP_NUM_THREADS=1)."""
    return lib().mlb_set_host_threads(int(n))


def partition(mesh, n_parts):
    """Deterministic recursive-coordinate-bisection partition of the cells (reference numbering) — new, SURVEY §8e."""
    v, keep = mesh.view()
    part = np.empty(mesh.n_cells, dtype=np.int32)
    if lib().mlb_partition(C.byref(v), n_parts, _ptr(part)):
        raise MallardError(_last_error())
    return part


def partition_coords(cell_xy, n_parts):
    """The same partition from the cell centroids alone ([n][2]); what a rank that holds only its part of the mesh calls."""
    xy = np.ascontiguousarray(cell_xy, dtype=np.float64).reshape(-1, 2)
    part = np.empty(xy.shape[0], dtype=np.int32)
    if lib().mlb_partition_coords(xy.shape[0], _ptr(xy), n_parts, _ptr(part)):
        raise MallardError(_last_error())
    return part


def partition_graph(mesh, n_parts, xadj=None, adj=None):
    """Graph partition (multilevel recursive bisection of the cell-face dual graph, csrc/partition_graph.cpp): no coordinates involved,
    part sizes identical to `partition`'s.  mesh=None: partition the symmetric CSR graph (xadj [n+1] uint64, adj uint32)."""
    if mesh is None:
        xadj = np.ascontiguousarray(xadj, dtype=np.uint64)
        adj = np.ascontiguousarray(adj, dtype=np.uint32)
        part = np.empty(len(xadj) - 1, dtype=np.int32)
        if lib().mlb_partition_graph_csr(len(xadj) - 1, _ptr(xadj), _ptr(adj), n_parts, _ptr(part)):
            raise MallardError(_last_error())
        return part
    v, keep = mesh.view()
    part = np.empty(mesh.n_cells, dtype=np.int32)
    if lib().mlb_partition_graph(C.byref(v), n_parts, _ptr(part)):
        raise MallardError(_last_error())
    return part


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it and ships the bytes to every rank)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    if lib().mlb_comm_unique_id(buf):
        raise MallardError(_last_error())
    return bytes(buf.raw)


def _numerics(recon, riemann, integrator, basis, order, factor, quad_cell, quad_face, fp_mode, renumber, teno_fixed, keep_rhs):
    for table, key, what in ((RECON, recon, "face reconstruction"), (RIEMANN, riemann, "Riemann solver"),
                             (INTEGRATOR, integrator, "time integrator")):
        if key not in table:
            raise MallardError("Unknown %s type: %s." % (what, key))   # solver/solver.cpp:121-146
    n = _abi.Numerics()
    n.recon, n.riemann, n.integrator = RECON[recon], RIEMANN[riemann], INTEGRATOR[integrator]
    n.basis, n.basis_order, n.max_stencil_size_factor = BASIS[basis], order, factor
    n.quadrature_order_cell, n.quadrature_order_face = quad_cell, quad_face
    n.fp_mode, n.renumber, n.teno_fixed, n.keep_stage_rhs = FP[fp_mode], RENUMBER[renumber], int(teno_fixed), int(keep_rhs)
    return n


def _physics(gas):
    g = dict(DEFAULT_GAS)
    g.update(gas or {})
    return _abi.Physics(g["gamma"], g["p_ref"], g["T_ref"], g["rho_ref"], g["p_min"], g["p_max"], g["mu"], g["Pr"])


def _local_struct(local, mesh):
    """mlb_local_mesh from dict(global_ids, n_global, cell0_nodes); returns (struct, keepalive)."""
    gid = np.ascontiguousarray(local["global_ids"], dtype=np.uint32)
    assert len(gid) == mesh.n_cells
    lm = _abi.LocalMesh()
    lm.n_global_cells, lm.global_cell_ids = int(local["n_global"]), _ptr(gid)
    for i in range(6):
        lm.cell0_nodes[i] = float(local["cell0_nodes"][i])
    return lm, gid


class Plan:
    """The mesh preprocessor's output alone (host only; no device needed): renumbering, slot tables, TENO tables."""

    _DT = {"sizes": np.uint32, "n_interior": np.uint32, "perm_cells": np.uint32, "perm_faces": np.uint32, "slot_face": np.uint32, "slot_nbr": np.int32,
           "rhs_order": np.uint8, "st_ids": np.uint32, "ghost_owner": np.int32, "teno:poly_indices": np.uint8,
           "fm_ids": np.uint32, "halo_peers": np.int32, "halo_recv_counts": np.uint64, "halo_recv_ids": np.uint32,
           "teno:offsets_stencil_groups": np.uint32, "teno:offsets_stencils": np.uint32, "teno:stencils": np.uint32,
           "teno:offsets_reconstruction_matrices": np.uint32}

    def __init__(self, mesh, recon="FO", basis="legendre", order=3, factor=2.0, quad_cell_order=0, quad_face_order=0, bcs=(),
                 renumber="rcm", part=None, rank=0, n_ranks=1, fp_mode="strict", local=None):
        num = _numerics(recon, "HLLC", "SSPRK3", basis, order, factor, quad_cell_order, quad_face_order, fp_mode, renumber, False, False)
        keep = []
        cb = (_abi.Bc * max(1, len(bcs)))()
        for i, b in enumerate(bcs):
            nb = b["name"].encode()
            keep.append(nb)
            cb[i].zone_name = nb
            cb[i].type = BC[b["type"]]
        v, mkeep = mesh.view()
        par = _abi.Parallel(rank, n_ranks, 0)
        h = C.c_void_p()
        self._h = None
        pp = None if part is None else np.ascontiguousarray(part, dtype=np.int32)
        if local is not None:
            lm, lkeep = _local_struct(local, mesh)
            rc = lib().mlb_plan_create_local(C.byref(h), C.byref(v), C.byref(num), cb, len(bcs), _ptr(pp), C.byref(par), C.byref(lm))
        else:
            rc = lib().mlb_plan_create(C.byref(h), C.byref(v), C.byref(num), cb, len(bcs), _ptr(pp), C.byref(par))
        if rc:
            raise MallardError(_last_error())
        self._h = h
        s = self.get("sizes")
        (self.N, self.N_owned, self.N_recon, self.NF, self.n_slots, self.Q, self.K, self.M, self.Npad, self.S, self.Mp) = (int(x) for x in s[:11])
        self.stream_tile = int(s[11])    # cells per tile of the streaming (FAST mode) TENO tables

    def get(self, name):
        nb = C.c_uint64()
        if lib().mlb_plan_get(self._h, name.encode(), None, C.byref(nb)):
            raise MallardError(_last_error())
        dt = np.dtype(self._DT.get(name, np.float64))
        out = np.empty(nb.value // dt.itemsize, dtype=dt)
        if nb.value and lib().mlb_plan_get(self._h, name.encode(), _ptr(out), C.byref(nb)):
            raise MallardError(_last_error())
        return out.reshape(-1, 2) if name == "teno:poly_indices" else out

    def __del__(self):
        try:
            if self._h is not None:
                lib().mlb_plan_destroy(self._h)
        except Exception:
            pass


class Solver:
    """Hot-path surface of the reference Solver, executed on the GPU.

    bcs: list of dicts with the keys of a [[boundaries]] TOML table (name, type, u, p, T), in input order."""

    def __init__(self, mesh, recon="FO", riemann="HLLC", integrator="SSPRK3", gas=None, basis="legendre", order=3, factor=2.0,
                 quad_cell_order=0, quad_face_order=0, bcs=(), fp_mode="strict", renumber="rcm", teno_fixed=False,
                 keep_stage_rhs=True, device=0, part=None, rank=0, n_ranks=1, local=None):
        """part: partition vector over the cells of `mesh` (multi-GPU).  local: `mesh` is this rank's part of a larger mesh
        (local_mesh.extract_local / synthetic.jittered_tri_local): dict(global_ids, n_global, cell0_nodes) -> mlb_create_local."""
        self.mesh = mesh
        self.nc, self.nf = mesh.n_cells, mesh.n_faces
        self._h = None
        num = _numerics(recon, riemann, integrator, basis, order, factor, quad_cell_order, quad_face_order, fp_mode, renumber,
                        teno_fixed, keep_stage_rhs)
        phys = _physics(gas)
        keep = []
        cb = (_abi.Bc * max(1, len(bcs)))()
        for i, b in enumerate(bcs):
            if "name" not in b:
                raise MallardError("Boundary name not specified.")
            if "type" not in b:
                raise MallardError("Boundary type not specified.")
            if b["type"] not in BC:
                raise MallardError("Unknown boundary type: %s." % b["type"])
            if mesh.get_face_zone(b["name"]) is None:
                raise MallardError("Boundary name %s not found in mesh." % b["name"])
            for req in {"upt": ("u", "p", "T"), "p_out": ("p",)}.get(b["type"], ()):
                if req not in b:
                    raise MallardError("Missing %s for boundary: %s." % (req, b["name"]))
            nb = b["name"].encode()
            keep.append(nb)
            cb[i].zone_name = nb
            cb[i].type = BC[b["type"]]
            u = b.get("u", (0.0, 0.0))
            cb[i].u[0], cb[i].u[1] = float(u[0]), float(u[1])
            cb[i].p, cb[i].T = float(b.get("p", 0.0)), float(b.get("T", 0.0))
        v, mkeep = mesh.view()
        par = _abi.Parallel(rank, n_ranks, device)
        h = C.c_void_p()
        if part is None:
            rc = lib().mlb_create(C.byref(h), C.byref(v), C.byref(num), C.byref(phys), cb, len(bcs), C.byref(par))
        elif local is not None:
            part = np.ascontiguousarray(part, dtype=np.int32)
            assert len(part) == mesh.n_cells
            lm, lkeep = _local_struct(local, mesh)
            rc = lib().mlb_create_local(C.byref(h), C.byref(v), _ptr(part), C.byref(lm), C.byref(num), C.byref(phys), cb, len(bcs), C.byref(par))
        else:
            part = np.ascontiguousarray(part, dtype=np.int32)
            rc = lib().mlb_create_partitioned(C.byref(h), C.byref(v), _ptr(part), C.byref(num), C.byref(phys), cb, len(bcs), C.byref(par))
        if rc:
            raise MallardError(_last_error())
        self._h = h
        self.n_quad = lib().mlb_n_face_quadrature_points(h)
        self.n_stages = lib().mlb_n_stages(h)

    def _ok(self, rc):
        if rc:
            raise MallardError(_last_error(self._h))

    def close(self):
        if self._h is not None:
            lib().mlb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state (Solver::copy_host_to_device / copy_device_to_host, solver/solver.cpp:322-334)
    def set_state(self, U, P=None):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.shape == (self.nc, 4)
        P = None if P is None else np.ascontiguousarray(P, dtype=np.float64)
        self._ok(lib().mlb_set_state(self._h, _ptr(U), _ptr(P)))

    def get_state(self, prim=False, cfl_local=False):
        U = np.empty((self.nc, 4))
        P = np.empty((self.nc, 5)) if prim else None
        Cl = np.empty(self.nc) if cfl_local else None
        self._ok(lib().mlb_get_state(self._h, _ptr(U), _ptr(P), _ptr(Cl)))
        out = (U,) + ((P,) if prim else ()) + ((Cl,) if cfl_local else ())
        return out[0] if len(out) == 1 else out

    # -- Solver::do_checks / check_fields (solver/solver.cpp:422-498) as device reductions
    FIELD_NAMES = ("RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X", "U_Y", "P", "T", "H")

    def field_ranges(self):
        """({name: (min, max)}, number of NaN entries) of the conserved and primitive fields, reduced on the device."""
        mn, mx, nn = np.empty(9), np.empty(9), C.c_uint64()
        self._ok(lib().mlb_field_ranges(self._h, _ptr(mn), _ptr(mx), C.byref(nn)))
        return {n: (float(a), float(b)) for n, a, b in zip(self.FIELD_NAMES, mn, mx)}, int(nn.value)

    # -- DataWriter::write_vtu (io/data_writer.cpp:93-244)
    def write_vtu(self, prefix, step, variables=("CFL", "RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X", "U_Y", "P", "T", "H")):
        v, keep = self.mesh.view()
        names = (C.c_char_p * len(variables))(*[x.encode() for x in variables])
        self._ok(lib().mlb_write_vtu(self._h, C.byref(v), str(prefix).encode(), int(step), len(variables), names))
        return "%s_%06d.vtu" % (prefix, step)

    # -- FaceReconstruction::calc_face_values
    def calc_face_values(self):
        F = np.empty((self.nf, self.n_quad, 2, 4))
        self._ok(lib().mlb_calc_face_values(self._h, _ptr(F)))
        return F

    # -- Solver::calc_rhs (the rhs_func seam)
    def calc_rhs(self, U=None):
        rhs = np.empty((self.nc, 4))
        if U is None:
            self._ok(lib().mlb_calc_rhs(self._h, _ptr(rhs)))
        else:
            U = np.ascontiguousarray(U, dtype=np.float64)
            self._ok(lib().mlb_calc_rhs_host(self._h, _ptr(U), _ptr(rhs)))
        return rhs

    # -- Solver::calc_dt
    def calc_dt(self, cfl):
        dt = C.c_double()
        self._ok(lib().mlb_calc_dt(self._h, cfl, C.byref(dt)))
        return dt.value

    def set_dt(self, dt):
        self._ok(lib().mlb_set_dt(self._h, dt))

    # -- Solver::take_step
    def take_step(self, dt=None):
        if dt is not None:
            self.set_dt(dt)
        self._ok(lib().mlb_take_step(self._h))

    def take_step_host(self, U, cfl=0.0):
        U = np.ascontiguousarray(U, dtype=np.float64)
        dt = C.c_double()
        self._ok(lib().mlb_take_step_host(self._h, cfl, _ptr(U), C.byref(dt)))
        return U, dt.value

    # -- Solver::run's loop
    def run(self, n_steps, cfl=0.0):
        t, dt = C.c_double(), C.c_double()
        self._ok(lib().mlb_run(self._h, n_steps, cfl, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def time(self):
        t, s = C.c_double(), C.c_uint64()
        self._ok(lib().mlb_get_time(self._h, C.byref(t), C.byref(s)))
        return t.value, s.value

    def set_rhs_override(self, rhs):
        rhs = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
        self._ok(lib().mlb_set_rhs_override(self._h, _ptr(rhs)))

    _ARRAY_DTYPES = {"dev:fm_ids": np.uint32, "perm_cells": np.uint32, "perm_faces": np.uint32, "teno:poly_indices": np.uint8,
                     "teno:offsets_stencil_groups": np.uint32, "teno:offsets_stencils": np.uint32, "teno:stencils": np.uint32,
                     "teno:offsets_reconstruction_matrices": np.uint32}

    def get(self, name):
        nb = C.c_uint64()
        self._ok(lib().mlb_get_array(self._h, name.encode(), None, C.byref(nb)))
        dt = np.dtype(self._ARRAY_DTYPES.get(name, np.float64))
        out = np.empty(nb.value // dt.itemsize, dtype=dt)
        if nb.value:
            self._ok(lib().mlb_get_array(self._h, name.encode(), _ptr(out), C.byref(nb)))
        if name.startswith("rhs") or name == "U_temp":
            out = out.reshape(self.nc, 4)
        elif name == "teno:poly_indices":
            out = out.reshape(-1, 2)
        return out

    # -- native multi-GPU driver (NCCL inside the library)
    def comm_init(self, unique_id):
        self._ok(lib().mlb_comm_init(self._h, C.c_char_p(unique_id)))

    def run_distributed(self, n_steps, cfl=0.0):
        t, dt = C.c_double(), C.c_double()
        self._ok(lib().mlb_run_distributed(self._h, n_steps, cfl, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def take_step_distributed_host(self, U_owned, cfl=0.0):
        dt = C.c_double()
        self._ok(lib().mlb_take_step_distributed_host(self._h, cfl, _ptr(U_owned), C.byref(dt)))
        return U_owned, dt.value

    # -- measurement
    def event_record(self, slot):
        self._ok(lib().mlb_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ok(lib().mlb_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def profile(self, on):
        self._ok(lib().mlb_profile_enable(self._h, int(on)))

    def profile_read(self):
        n = 32
        names = (C.c_char_p * n)()
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        k = lib().mlb_profile_read(self._h, n, names, ms, cnt)
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(max(k, 0))}

    @property
    def launch_count(self):
        return int(lib().mlb_launch_count(self._h))

    def synchronize(self):
        self._ok(lib().mlb_synchronize(self._h))

    # -- multi-GPU split-phase surface
    def halo_info(self):
        n = C.c_int32()
        self._ok(lib().mlb_halo_info(self._h, C.byref(n), None, None, None))      # count first: the arrays are sized from it
        k = n.value
        peers = np.zeros(max(k, 1), dtype=np.int32)
        sc = np.zeros(max(k, 1), dtype=np.uint64)
        rc = np.zeros(max(k, 1), dtype=np.uint64)
        self._ok(lib().mlb_halo_info(self._h, C.byref(n), _ptr(peers), _ptr(sc), _ptr(rc)))
        return peers[:k].copy(), sc[:k].copy(), rc[:k].copy()

    def halo_recv_ids(self, peer_index, count):
        ids = np.empty(int(count), dtype=np.uint32)
        self._ok(lib().mlb_halo_recv_ids(self._h, peer_index, _ptr(ids)))
        return ids

    def halo_set_send_ids(self, peer_ranks, lists):
        pr = np.ascontiguousarray(peer_ranks, dtype=np.int32)
        counts = np.array([len(x) for x in lists], dtype=np.uint64)
        flat = np.ascontiguousarray(np.concatenate(lists) if len(lists) else np.zeros(0), dtype=np.uint32)
        self._ok(lib().mlb_halo_set_send_ids(self._h, len(lists), _ptr(pr), _ptr(counts), _ptr(flat)))

    def halo_buffers(self):
        s, r = C.c_void_p(), C.c_void_p()
        self._ok(lib().mlb_halo_buffers(self._h, C.byref(s), C.byref(r)))
        return s.value, r.value

    def halo_pack(self, stage):
        self._ok(lib().mlb_halo_pack(self._h, stage))

    def halo_unpack(self, stage):
        self._ok(lib().mlb_halo_unpack(self._h, stage))

    def stage_begin(self, s):
        """Enqueues the reconstruction of the interior cells of stage s (overlaps the halo exchange in flight)."""
        self._ok(lib().mlb_stage_begin(self._h, s))

    def stage(self, s):
        self._ok(lib().mlb_stage(self._h, s))

    def local_max_spectral_radius(self, sync=True):
        """Rank-local part of calc_dt.  sync=False leaves the value in the device scalar block (scalars_device()[2])."""
        if not sync:
            self._ok(lib().mlb_local_max_spectral_radius(self._h, None))
            return None
        m = C.c_double()
        self._ok(lib().mlb_local_max_spectral_radius(self._h, C.byref(m)))
        return m.value

    def scalars_device(self):
        """Device pointer of the scalar block: doubles {dt, t, max spectral radius, cfl, ...}."""
        return lib().mlb_scalars_device(self._h)

    def apply_dt_device(self, cfl):
        self._ok(lib().mlb_apply_dt_device(self._h, cfl))

    def set_owned(self, U_owned):
        U_owned = np.ascontiguousarray(U_owned, dtype=np.float64)
        self._ok(lib().mlb_set_owned(self._h, _ptr(U_owned)))

    def get_owned(self, out=None):
        if out is None:
            n = C.c_uint32()
            self._ok(lib().mlb_owned_cells(self._h, C.byref(n), None))
            out = np.empty((n.value, 4))
        self._ok(lib().mlb_get_owned(self._h, _ptr(out)))
        return out

    def apply_dt(self, cfl, global_max):
        self._ok(lib().mlb_apply_dt(self._h, cfl, global_max))

    def finish_step(self):
        self._ok(lib().mlb_finish_step(self._h))

    def owned_cells(self):
        n = C.c_uint32()
        self._ok(lib().mlb_owned_cells(self._h, C.byref(n), None))
        ids = np.empty(n.value, dtype=np.uint32)
        self._ok(lib().mlb_owned_cells(self._h, C.byref(n), _ptr(ids)))
        return ids

    @property
    def stream(self):
        return lib().mlb_stream(self._h)

    @property
    def comm_stream(self):
        """cudaStream_t the halo pack / transfers / unpack are ordered on (the compute stream when unpartitioned)."""
        return lib().mlb_comm_stream(self._h)


def riemann_flux(kind, n_unit, L, R, gamma=1.4, fp_mode="strict", device=0):
    """RiemannSolver::calc_flux (numerics/riemann_solver.h:85-90) on the GPU. L/R rows: rho,u,v,p,h."""
    if kind not in RIEMANN:
        raise MallardError("Unknown Riemann solver type: %s." % kind)
    n_unit = np.ascontiguousarray(n_unit, dtype=np.float64).reshape(-1, 2)
    L = np.ascontiguousarray(L, dtype=np.float64).reshape(-1, 5)
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 5)
    out = np.empty((L.shape[0], 4))
    if lib().mlb_riemann_flux(device, RIEMANN[kind], FP[fp_mode], L.shape[0], _ptr(n_unit), _ptr(L), _ptr(R), gamma, _ptr(out)):
        raise MallardError(_last_error())
    return out


def compute_primitives(U, gas=None, fp_mode="strict", device=0):
    """Physics::compute_primitives_from_conservatives on the GPU; also returns (R, cp, cv)."""
    U = np.ascontiguousarray(U, dtype=np.float64).reshape(-1, 4)
    P = np.empty((U.shape[0], 5))
    rc = np.empty(3)
    ph = _physics(gas)
    if lib().mlb_compute_primitives(device, FP[fp_mode], C.byref(ph), U.shape[0], _ptr(U), _ptr(P), _ptr(rc)):
        raise MallardError(_last_error())
    return P, rc
