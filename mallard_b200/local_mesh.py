"""Rank-local meshes (new): the part of a mesh one rank needs — its own cells plus a few layers of ghost cells —
in the form mlb_create_local expects (include/mallard_b200.h):

  * cells and faces keep the relative order of their global ids (stencil membership depends on that order);
  * a face whose other cell is not part of the local mesh is "cut": cells_of_face = (present cell, -2);
  * zones are the global zones restricted to the kept faces, in the global zone's order.

`extract_local` restricts a mesh that is held in full (any unstructured mesh; used by the tests and by hosts that read a mesh
file on one rank and ship the parts); synthetic.jittered_tri_local generates the same arrays without ever building the
global mesh.  Host-side numpy only; nothing here computes on the hot path.
"""
import numpy as np

from . import Mesh

CUT = -2


def _csr_take(offsets, values, rows):
    """Concatenated rows `rows` of a CSR array, and the new offsets."""
    offsets = offsets.astype(np.int64)
    lens = offsets[rows + 1] - offsets[rows]
    new_off = np.concatenate([[0], np.cumsum(lens)])
    idx = np.arange(int(new_off[-1]), dtype=np.int64) - np.repeat(new_off[:-1], lens) + np.repeat(offsets[rows], lens)
    return values[idx], new_off


def dilate(mesh, keep, layers):
    """Adds `layers` rings of face neighbours to the boolean cell mask `keep`."""
    cof = mesh.arrays["cells_of_face"]
    inner = cof[:, 1] >= 0
    a, b = cof[inner, 0], cof[inner, 1]
    keep = keep.copy()
    for _ in range(layers):
        grow = keep.copy()
        grow[a] |= keep[b]
        grow[b] |= keep[a]
        keep = grow
    return keep


def extract_local(mesh, keep):
    """Restriction of `mesh` to the cells with keep[c] == True.  Returns (local Mesh with geometry, global cell ids, global
    face ids), ids ascending."""
    A = mesh.arrays
    cells = np.nonzero(keep)[0].astype(np.int64)
    cmap = np.full(mesh.n_cells, -1, dtype=np.int64)
    cmap[cells] = np.arange(len(cells))
    noc, onc = _csr_take(A["offsets_nodes_of_cell"], A["nodes_of_cell"].astype(np.int64), cells)
    foc, ofc = _csr_take(A["offsets_faces_of_cell"], A["faces_of_cell"].astype(np.int64), cells)
    faces = np.unique(foc)
    fmap = np.full(mesh.n_faces, -1, dtype=np.int64)
    fmap[faces] = np.arange(len(faces))
    nodes = np.unique(noc)
    nmap = np.full(mesh.n_nodes, -1, dtype=np.int64)
    nmap[nodes] = np.arange(len(nodes))
    nof, onf = _csr_take(A["offsets_nodes_of_face"], A["nodes_of_face"].astype(np.int64), faces)
    cof = A["cells_of_face"][faces].astype(np.int64)
    la = cmap[cof[:, 0]]
    lb = np.where(cof[:, 1] >= 0, cmap[np.maximum(cof[:, 1], 0)], -1)
    boundary = cof[:, 1] < 0
    c0 = np.where(la >= 0, la, lb)                                   # the present cell (side 0's if both are)
    c1 = np.where(boundary, -1, np.where((la >= 0) & (lb >= 0), lb, CUT))
    assert (c0 >= 0).all()
    arrays = dict(node_coords=A["node_coords"][nodes],
                  offsets_nodes_of_cell=onc.astype(np.uint32), nodes_of_cell=nmap[noc].astype(np.uint32),
                  offsets_faces_of_cell=ofc.astype(np.uint32), faces_of_cell=fmap[foc].astype(np.uint32),
                  offsets_nodes_of_face=onf.astype(np.uint32), nodes_of_face=nmap[nof].astype(np.uint32),
                  cells_of_face=np.stack([c0, c1], 1).astype(np.int32))
    zones = []
    for name, zf in mesh.zones:
        loc = fmap[zf.astype(np.int64)]
        zones.append((name, loc[loc >= 0].astype(np.uint32)))
    m = Mesh.from_arrays(arrays, zones).compute_geometry()
    return m, cells.astype(np.uint32), faces.astype(np.uint32)


def local_info(mesh, global_cell_ids):
    """The `local` argument of Solver for a restriction of `mesh` (node coordinates of global cell 0 included)."""
    A = mesh.arrays
    n0 = A["nodes_of_cell"][int(A["offsets_nodes_of_cell"][0]):int(A["offsets_nodes_of_cell"][0]) + 3]
    return dict(global_ids=np.ascontiguousarray(global_cell_ids, dtype=np.uint32), n_global=mesh.n_cells,
                cell0_nodes=np.ascontiguousarray(A["node_coords"][n0], dtype=np.float64).reshape(6))
