"""Mesh file reader / writer (the reference's `type = "file"` is a stub that throws, src/mesh/mesh.cpp:41-43).

Thin wrapper over the C ABI (mlb_host_mesh_read_gmsh / mlb_host_mesh_write_gmsh, csrc/mesh_io.cpp), which a C++ Mallard host
calls directly.  Gmsh MSH 2.2 ASCII, 2-D: triangles (element type 2) and quadrilaterals (type 3) become cells, 2-node lines
(type 1) carrying a physical tag name the boundary zones (`$PhysicalNames`).  The result is a `Mesh` holding the arrays of the
reference's Mesh (src/mesh/mesh.h:228-253) — faces are the unique cell edges, `cells_of_face[f][0]` is the lower-numbered
cell, face j of a cell is the edge (node j, node j+1), the zone "interior" lists the two-cell faces in ascending order — with
the geometry computed by the library's host code exactly as for the generated meshes.  No reference behaviour exists to be
faithful to; parity is pinned by round trips (tests/test_host_side.py) and by running the same mesh from a file and from
arrays (tests/test_gpu_parity.py).
"""
import ctypes as C

from . import Mesh, lib, _last_error


def read_gmsh(path):
    """Gmsh MSH 2.2 ASCII -> Mesh."""
    h = C.c_void_p()
    if lib().mlb_host_mesh_read_gmsh(C.byref(h), str(path).encode()):
        raise ValueError(_last_error())
    m = Mesh()
    m._take(h)
    return m


def write_gmsh(mesh, path):
    """Mesh -> Gmsh MSH 2.2 ASCII (cells + one line element per boundary face, one physical group per boundary zone)."""
    v, keep = mesh.view()
    if lib().mlb_host_mesh_write_gmsh(C.byref(v), str(path).encode()):
        raise ValueError(_last_error())
