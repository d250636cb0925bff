"""ctypes declarations of include/mallard_b200.h (the same stub a reference-side binding would use)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLB_LIB") or os.path.join(_HERE, "libmallard_b200.so")   # MLB_LIB: A/B kernel experiments

RECON = {"FO": 0, "TENO": 1}
RIEMANN = {"Rusanov": 0, "HLL": 1, "HLLC": 2}
INTEGRATOR = {"FE": 0, "RK4": 1, "SSPRK3": 2}
BC = {"symmetry": 0, "extrapolation": 1, "wall_adiabatic": 2, "upt": 3, "p_out": 4, "wall_noslip": 5}
BASIS = {"monomial": 0, "legendre": 1}
MESH = {"cartesian": 0, "cartesian_tri": 1, "wedge": 2}
RENUMBER = {"none": 0, "rcm": 1}
FP = {"strict": 0, "fast": 1}


class Zone(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n_faces", C.c_uint32), ("faces", C.c_void_p)]


class MeshView(C.Structure):
    _fields_ = [("n_cells", C.c_uint32), ("n_faces", C.c_uint32), ("n_nodes", C.c_uint32),
                ("node_coords", C.c_void_p), ("offsets_nodes_of_cell", C.c_void_p), ("nodes_of_cell", C.c_void_p),
                ("offsets_faces_of_cell", C.c_void_p), ("faces_of_cell", C.c_void_p),
                ("offsets_nodes_of_face", C.c_void_p), ("nodes_of_face", C.c_void_p), ("cells_of_face", C.c_void_p),
                ("cell_coords", C.c_void_p), ("cell_volume", C.c_void_p), ("face_area", C.c_void_p),
                ("face_normals", C.c_void_p), ("n_zones", C.c_uint32), ("zones", C.POINTER(Zone))]


class Numerics(C.Structure):
    _fields_ = [("recon", C.c_int32), ("riemann", C.c_int32), ("integrator", C.c_int32), ("basis", C.c_int32),
                ("basis_order", C.c_int32), ("max_stencil_size_factor", C.c_double),
                ("quadrature_order_cell", C.c_int32), ("quadrature_order_face", C.c_int32), ("fp_mode", C.c_int32),
                ("renumber", C.c_int32), ("teno_fixed", C.c_int32), ("keep_stage_rhs", C.c_int32)]


class Physics(C.Structure):
    _fields_ = [("gamma", C.c_double), ("p_ref", C.c_double), ("T_ref", C.c_double), ("rho_ref", C.c_double),
                ("p_min", C.c_double), ("p_max", C.c_double), ("mu", C.c_double), ("Pr", C.c_double)]


class Bc(C.Structure):
    _fields_ = [("zone_name", C.c_char_p), ("type", C.c_int32), ("u", C.c_double * 2), ("p", C.c_double), ("T", C.c_double)]


class Parallel(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("device", C.c_int32)]


class LocalMesh(C.Structure):
    _fields_ = [("n_global_cells", C.c_uint32), ("global_cell_ids", C.c_void_p), ("cell0_nodes", C.c_double * 6)]


ABI_VERSION = 4     # MLB_ABI_VERSION of the header revision these structures mirror (tests/test_host_side.py compares the two)

# every symbol include/mallard_b200.h declares: name -> (restype, argtypes)
VP, I32, U32, U64, DBL = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_double
SYMBOLS = {
    "mlb_version": (C.c_char_p, []),
    "mlb_check_abi": (C.c_int, [I32]),
    "mlb_set_host_threads": (C.c_int, [I32]),
    "mlb_last_error": (C.c_char_p, [VP]),
    "mlb_create": (C.c_int, [C.POINTER(VP), C.POINTER(MeshView), C.POINTER(Numerics), C.POINTER(Physics), C.POINTER(Bc), I32, C.POINTER(Parallel)]),
    "mlb_destroy": (None, [VP]),
    "mlb_set_state": (C.c_int, [VP, VP, VP]),
    "mlb_get_state": (C.c_int, [VP, VP, VP, VP]),
    "mlb_calc_face_values": (C.c_int, [VP, VP]),
    "mlb_n_face_quadrature_points": (C.c_int, [VP]),
    "mlb_calc_rhs": (C.c_int, [VP, VP]),
    "mlb_calc_rhs_host": (C.c_int, [VP, VP, VP]),
    "mlb_calc_dt": (C.c_int, [VP, DBL, C.POINTER(DBL)]),
    "mlb_set_dt": (C.c_int, [VP, DBL]),
    "mlb_take_step": (C.c_int, [VP]),
    "mlb_take_step_host": (C.c_int, [VP, DBL, VP, C.POINTER(DBL)]),
    "mlb_run": (C.c_int, [VP, U32, DBL, C.POINTER(DBL), C.POINTER(DBL)]),
    "mlb_get_time": (C.c_int, [VP, C.POINTER(DBL), C.POINTER(U64)]),
    "mlb_field_ranges": (C.c_int, [VP, VP, VP, C.POINTER(U64)]),
    "mlb_write_vtu": (C.c_int, [VP, C.POINTER(MeshView), C.c_char_p, C.c_uint32, C.c_int32, C.POINTER(C.c_char_p)]),
    "mlb_set_rhs_override": (C.c_int, [VP, VP]),
    "mlb_get_array": (C.c_int, [VP, C.c_char_p, VP, C.POINTER(U64)]),
    "mlb_event_record": (C.c_int, [VP, I32]),
    "mlb_event_elapsed_ms": (C.c_int, [VP, I32, I32, C.POINTER(C.c_float)]),
    "mlb_profile_enable": (C.c_int, [VP, I32]),
    "mlb_profile_read": (C.c_int, [VP, I32, C.POINTER(C.c_char_p), C.POINTER(DBL), C.POINTER(U64)]),
    "mlb_launch_count": (U64, [VP]),
    "mlb_synchronize": (C.c_int, [VP]),
    "mlb_stream": (VP, [VP]),
    "mlb_comm_stream": (VP, [VP]),
    "mlb_partition": (C.c_int, [C.POINTER(MeshView), I32, VP]),
    "mlb_create_partitioned": (C.c_int, [C.POINTER(VP), C.POINTER(MeshView), VP, C.POINTER(Numerics), C.POINTER(Physics), C.POINTER(Bc), I32, C.POINTER(Parallel)]),
    "mlb_partition_coords": (C.c_int, [U64, VP, I32, VP]),
    "mlb_partition_graph": (C.c_int, [C.POINTER(MeshView), I32, VP]),
    "mlb_partition_graph_csr": (C.c_int, [U32, VP, VP, I32, VP]),
    "mlb_create_local": (C.c_int, [C.POINTER(VP), C.POINTER(MeshView), VP, C.POINTER(LocalMesh), C.POINTER(Numerics), C.POINTER(Physics), C.POINTER(Bc), I32, C.POINTER(Parallel)]),
    "mlb_comm_unique_id": (C.c_int, [VP]),
    "mlb_comm_init": (C.c_int, [VP, VP]),
    "mlb_run_distributed": (C.c_int, [VP, U32, DBL, C.POINTER(DBL), C.POINTER(DBL)]),
    "mlb_take_step_distributed_host": (C.c_int, [VP, DBL, VP, C.POINTER(DBL)]),
    "mlb_halo_info": (C.c_int, [VP, C.POINTER(I32), VP, VP, VP]),
    "mlb_halo_recv_ids": (C.c_int, [VP, I32, VP]),
    "mlb_halo_set_send_ids": (C.c_int, [VP, I32, VP, VP, VP]),
    "mlb_halo_buffers": (C.c_int, [VP, C.POINTER(VP), C.POINTER(VP)]),
    "mlb_halo_pack": (C.c_int, [VP, I32]),
    "mlb_halo_unpack": (C.c_int, [VP, I32]),
    "mlb_n_stages": (C.c_int, [VP]),
    "mlb_stage": (C.c_int, [VP, I32]),
    "mlb_stage_begin": (C.c_int, [VP, I32]),
    "mlb_local_max_spectral_radius": (C.c_int, [VP, C.POINTER(DBL)]),
    "mlb_apply_dt": (C.c_int, [VP, DBL, DBL]),
    "mlb_scalars_device": (VP, [VP]),
    "mlb_apply_dt_device": (C.c_int, [VP, DBL]),
    "mlb_set_owned": (C.c_int, [VP, VP]),
    "mlb_get_owned": (C.c_int, [VP, VP]),
    "mlb_finish_step": (C.c_int, [VP]),
    "mlb_owned_cells": (C.c_int, [VP, C.POINTER(U32), VP]),
    "mlb_riemann_flux": (C.c_int, [I32, I32, I32, U64, VP, VP, VP, DBL, VP]),
    "mlb_compute_primitives": (C.c_int, [I32, I32, C.POINTER(Physics), U64, VP, VP, VP]),
    "mlb_plan_create": (C.c_int, [C.POINTER(VP), C.POINTER(MeshView), C.POINTER(Numerics), C.POINTER(Bc), I32, VP, C.POINTER(Parallel)]),
    "mlb_plan_create_local": (C.c_int, [C.POINTER(VP), C.POINTER(MeshView), C.POINTER(Numerics), C.POINTER(Bc), I32, VP, C.POINTER(Parallel), C.POINTER(LocalMesh)]),
    "mlb_plan_get": (C.c_int, [VP, C.c_char_p, VP, C.POINTER(U64)]),
    "mlb_plan_destroy": (None, [VP]),
    "mlb_host_mesh_generate": (C.c_int, [C.POINTER(VP), I32, U32, U32, DBL, DBL]),
    "mlb_host_mesh_from_arrays": (C.c_int, [C.POINTER(VP), C.POINTER(MeshView)]),
    "mlb_host_mesh_read_gmsh": (C.c_int, [C.POINTER(VP), C.c_char_p]),
    "mlb_host_mesh_from_cells": (C.c_int, [C.POINTER(VP), U32, VP, U32, VP, VP, U32, VP, VP, U32, VP, C.POINTER(C.c_char_p)]),
    "mlb_host_mesh_write_gmsh": (C.c_int, [C.POINTER(MeshView), C.c_char_p]),
    "mlb_host_mesh_view": (C.c_int, [VP, C.POINTER(MeshView)]),
    "mlb_host_mesh_free": (None, [VP]),
}

_LIB = None


def build(force=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-s", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    if force:
        subprocess.check_call(args + ["clean"])
    subprocess.check_call(args)


def lib():
    """Load libmallard_b200.so.  There is no fallback: a missing extension is an error."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("mallard_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.mlb_check_abi(ABI_VERSION):       # the structs above mirror ONE revision of include/mallard_b200.h
            raise RuntimeError("mallard_b200: " + L.mlb_last_error(None).decode() + " - the built library and mallard_b200/_abi.py disagree (rebuild)")
        _LIB = L
    return _LIB
