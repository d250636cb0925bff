"""Mesh file reader / writer (the reference's `type = "file"` is a stub that throws, src/mesh/mesh.cpp:41-43).

Gmsh MSH 2.2 ASCII, 2-D: triangles (element type 2) and quadrilaterals (type 3) become cells, 2-node lines (type 1) carrying
a physical tag name the boundary zones (`$PhysicalNames`).  The result is a `Mesh` holding the arrays of the reference's
Mesh (src/mesh/mesh.h:228-253) — faces are the unique cell edges, `cells_of_face[f][0]` is the lower-numbered cell, face j
of a cell is the edge (node j, node j+1), the zone "interior" lists the two-cell faces in ascending order — with the
geometry computed by the library's host code exactly as for the generated meshes.  No reference behaviour exists to be
faithful to; parity is pinned by round trips (tests/test_host_side.py) and by running the same mesh from a file and from
arrays (tests/test_gpu_parity.py).  Host-side numpy only.
"""
import numpy as np

from . import Mesh

_NODES_OF = {1: 2, 2: 3, 3: 4, 15: 1}


def build_from_cells(node_xy, cells, boundary_edges=(), boundary_names=None):
    """node_xy [nn][2]; cells: list of node-index tuples (3 or 4 nodes, any mix); boundary_edges: iterable of (a, b, tag);
    boundary_names: {tag: zone name}.  Boundary faces no line element claims go to the zone "boundary"."""
    node_xy = np.ascontiguousarray(node_xy, dtype=np.float64)
    nc = len(cells)
    nn_cell = np.fromiter((len(c) for c in cells), dtype=np.int64, count=nc)
    onc = np.zeros(nc + 1, dtype=np.int64)
    np.cumsum(nn_cell, out=onc[1:])
    noc = np.fromiter((n for c in cells for n in c), dtype=np.int64, count=int(onc[-1]))
    # edge j of a cell = (node j, node j+1 mod n)
    cell_of_edge = np.repeat(np.arange(nc), nn_cell)
    local = np.arange(len(noc)) - onc[cell_of_edge]
    nxt = onc[cell_of_edge] + (local + 1) % nn_cell[cell_of_edge]
    a, b = noc, noc[nxt]
    key = np.minimum(a, b) * (len(node_xy) + 1) + np.maximum(a, b)
    ukey, face_of_edge, counts = np.unique(key, return_inverse=True, return_counts=True)
    nf = len(ukey)
    if counts.max() > 2:
        raise ValueError("mesh file: an edge is shared by more than two cells")
    # cells of each face: lower-numbered cell on side 0 (the edge list is already in ascending cell order)
    order = np.argsort(face_of_edge, kind="stable")
    first = np.zeros(nf, dtype=np.int64)
    np.cumsum(counts[:-1], out=first[1:])
    cof = np.full((nf, 2), -1, dtype=np.int64)
    cof[:, 0] = cell_of_edge[order[first]]
    two = counts == 2
    cof[two, 1] = cell_of_edge[order[first[two] + 1]]
    # nodes of a face in the orientation of its side-0 cell
    e0 = order[first]
    nof = np.stack([a[e0], b[e0]], 1)
    zones = [("interior", np.nonzero(two)[0])]
    names = dict(boundary_names or {})
    claimed = np.zeros(nf, dtype=bool)
    by_tag = {}
    lookup = {int(k): i for i, k in enumerate(ukey)} if len(boundary_edges) else {}
    for ea, eb, tag in boundary_edges:
        k = min(ea, eb) * (len(node_xy) + 1) + max(ea, eb)
        f = lookup.get(int(k))
        if f is None or two[f]:
            raise ValueError("mesh file: boundary line (%d, %d) is not a boundary edge of any cell" % (ea, eb))
        by_tag.setdefault(tag, []).append(f)
        claimed[f] = True
    for tag in sorted(by_tag):
        zones.append((names.get(tag, "zone_%d" % tag), np.array(sorted(by_tag[tag]), dtype=np.int64)))
    rest = np.nonzero(~two & ~claimed)[0]
    if len(rest):
        zones.append(("boundary", rest))
    arrays = dict(node_coords=node_xy, offsets_nodes_of_cell=onc.astype(np.uint32), nodes_of_cell=noc.astype(np.uint32),
                  offsets_faces_of_cell=onc.astype(np.uint32), faces_of_cell=face_of_edge.astype(np.uint32),
                  offsets_nodes_of_face=np.arange(0, 2 * nf + 1, 2, dtype=np.uint32), nodes_of_face=nof.reshape(-1).astype(np.uint32),
                  cells_of_face=cof.astype(np.int32))
    m = Mesh.from_arrays(arrays, [(n, f.astype(np.uint32)) for n, f in zones])
    return m.compute_geometry()


def read_gmsh(path):
    """Gmsh MSH 2.2 ASCII -> Mesh."""
    with open(path) as fh:
        lines = [l.strip() for l in fh]
    sections, i = {}, 0
    while i < len(lines):
        if lines[i].startswith("$") and not lines[i].startswith("$End"):
            name = lines[i][1:]
            j = lines.index("$End" + name, i)
            sections[name] = lines[i + 1:j]
            i = j
        i += 1
    fmt = sections.get("MeshFormat", ["?"])[0].split()
    if not fmt or not fmt[0].startswith("2") or (len(fmt) > 1 and fmt[1] != "0"):
        raise ValueError("mesh file: only Gmsh MSH 2.x ASCII is supported (got %r)" % " ".join(fmt))
    names = {}
    for l in sections.get("PhysicalNames", ["0"])[1:]:
        dim, tag, name = l.split(None, 2)
        names[(int(dim), int(tag))] = name.strip().strip('"')
    nl = sections["Nodes"]
    nn = int(nl[0])
    ids = np.empty(nn, dtype=np.int64)
    xy = np.empty((nn, 2))
    for k, l in enumerate(nl[1:nn + 1]):
        p = l.split()
        ids[k] = int(p[0]); xy[k, 0] = float(p[1]); xy[k, 1] = float(p[2])
    index = {int(t): k for k, t in enumerate(ids)}
    cells, edges = [], []
    for l in sections["Elements"][1:]:
        p = l.split()
        typ, ntags = int(p[1]), int(p[2])
        if typ not in _NODES_OF:
            raise ValueError("mesh file: unsupported element type %d" % typ)
        nodes = [index[int(t)] for t in p[3 + ntags:3 + ntags + _NODES_OF[typ]]]
        tag = int(p[3]) if ntags else 0
        if typ in (2, 3):
            cells.append(tuple(nodes))
        elif typ == 1:
            edges.append((nodes[0], nodes[1], tag))
    if not cells:
        raise ValueError("mesh file: no triangles or quadrilaterals")
    return build_from_cells(xy, cells, edges, {t: n for (d, t), n in names.items() if d == 1})


def write_gmsh(mesh, path):
    """Mesh -> Gmsh MSH 2.2 ASCII (cells + one line element per boundary face, one physical group per boundary zone)."""
    a = mesh.arrays
    onc, noc = a["offsets_nodes_of_cell"].astype(np.int64), a["nodes_of_cell"].astype(np.int64)
    nof = a["nodes_of_face"].reshape(-1, 2).astype(np.int64)
    bz = [(n, f) for n, f in mesh.zones if n != "interior"]
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % (len(bz) + 1))
        for t, (n, _) in enumerate(bz):
            fh.write('1 %d "%s"\n' % (t + 1, n))
        fh.write('2 %d "fluid"\n$EndPhysicalNames\n$Nodes\n%d\n' % (len(bz) + 1, mesh.n_nodes))
        for k, (x, y) in enumerate(a["node_coords"]):
            fh.write("%d %.17g %.17g 0\n" % (k + 1, x, y))
        n_el = sum(len(f) for _, f in bz) + mesh.n_cells
        fh.write("$EndNodes\n$Elements\n%d\n" % n_el)
        e = 1
        for t, (_, faces) in enumerate(bz):
            for f in faces:
                fh.write("%d 1 2 %d %d %d %d\n" % (e, t + 1, t + 1, nof[f, 0] + 1, nof[f, 1] + 1))
                e += 1
        for c in range(mesh.n_cells):
            nodes = noc[onc[c]:onc[c + 1]] + 1
            fh.write("%d %d 2 %d %d %s\n" % (e, 2 if len(nodes) == 3 else 3, len(bz) + 1, len(bz) + 1, " ".join(map(str, nodes))))
            e += 1
        fh.write("$EndElements\n")
