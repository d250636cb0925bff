"""Synthetic unstructured workloads (BASELINE.json configs[3]; SURVEY §8d "Config 4").

The reference can only generate structured meshes (src/mesh/mesh.cpp:305-848) and cannot read mesh files
(src/mesh/mesh.cpp:41-43), so an "unstructured" mesh is emulated the way SURVEY §8d prescribes: a triangulated Cartesian
grid whose interior nodes are jittered (seeded) and whose cell and face ids are randomly permuted, handed over as plain
arrays in the reference's Mesh layout (src/mesh/mesh.h:228-253).  The same arrays feed the oracle (oracle.Mesh.from_arrays),
so parity is pinned on exactly this mesh family.  Host-side numpy only; nothing here computes on the hot path.
"""
import numpy as np

from . import Mesh

ZONES = ("interior", "right", "top", "left", "bottom")


def jittered_tri(nx, ny, Lx=1.0, Ly=1.0, seed=12345, amp=0.15, shuffle=True):
    """Triangulated nx x ny grid on [0,Lx]x[0,Ly] (2*nx*ny triangles): phantom faces of cartesian_tri dropped,
    interior nodes moved by a uniform +-amp*h in x and y, cell ids and face ids randomly permuted."""
    base = Mesh.generate("cartesian_tri", nx, ny, Lx, Ly)
    a = base.arrays
    rng = np.random.default_rng(seed)
    nc = base.n_cells
    xy = a["node_coords"].copy()
    noc = a["nodes_of_cell"].reshape(nc, 3).astype(np.int64)
    foc = a["faces_of_cell"].reshape(nc, 3).astype(np.int64)
    nof = a["nodes_of_face"].reshape(-1, 2).astype(np.int64)
    cof = a["cells_of_face"].astype(np.int64)

    # ---- drop phantom faces
    real = nof[:, 0] != nof[:, 1]
    fmap = np.full(len(nof), -1, dtype=np.int64)
    fmap[real] = np.arange(int(real.sum()))
    nof, cof = nof[real], cof[real]
    foc = fmap[foc]
    assert (foc >= 0).all()
    zones = [(n, fmap[f.astype(np.int64)]) for n, f in base.zones]
    nf = len(nof)

    # ---- jitter interior nodes
    if amp > 0:
        h = min(Lx / nx, Ly / ny)
        eps = 1e-9 * max(Lx, Ly)
        inner = (xy[:, 0] > eps) & (xy[:, 0] < Lx - eps) & (xy[:, 1] > eps) & (xy[:, 1] < Ly - eps)
        d = rng.uniform(-amp * h, amp * h, size=xy.shape)
        xy[inner] += d[inner]

    # ---- permute cell and face ids
    if shuffle:
        pc = rng.permutation(nc)            # new cell i = old cell pc[i]
        ipc = np.empty(nc, dtype=np.int64); ipc[pc] = np.arange(nc)
        noc, foc = noc[pc], foc[pc]
        cof = np.where(cof >= 0, ipc[np.maximum(cof, 0)], -1)
        pf = rng.permutation(nf)
        ipf = np.empty(nf, dtype=np.int64); ipf[pf] = np.arange(nf)
        nof, cof = nof[pf], cof[pf]
        foc = ipf[foc]
        zones = [(n, np.sort(ipf[f])) for n, f in zones]

    arrays = dict(node_coords=xy,
                  offsets_nodes_of_cell=np.arange(0, 3 * nc + 1, 3, dtype=np.uint32), nodes_of_cell=noc.reshape(-1).astype(np.uint32),
                  offsets_faces_of_cell=np.arange(0, 3 * nc + 1, 3, dtype=np.uint32), faces_of_cell=foc.reshape(-1).astype(np.uint32),
                  offsets_nodes_of_face=np.arange(0, 2 * nf + 1, 2, dtype=np.uint32), nodes_of_face=nof.reshape(-1).astype(np.uint32),
                  cells_of_face=cof.astype(np.int32))
    m = Mesh.from_arrays(arrays, [(n, f.astype(np.uint32)) for n, f in zones])
    return m.compute_geometry()


def mixed_tri_quad(nx, ny, Lx=1.0, Ly=1.0, seed=12345, amp=0.15, tri_fraction=0.5, shuffle=True):
    """Mixed-element mesh (BASELINE configs[3] as worded: "mixed tri/quad unstructured mesh"): an nx x ny grid on [0,Lx]x[0,Ly] with
    jittered interior nodes in which a seeded random `tri_fraction` of the quads is cut into two triangles (along either diagonal)
    and the rest stay quadrilaterals; cell order randomly permuted; zones left / right / top / bottom / interior.  The
    reference's TENO throws on anything but triangles (face_reconstruction.cpp:485-487): there is no oracle for this mesh family."""
    rng = np.random.default_rng(seed)
    ii, jj = np.divmod(np.arange((nx + 1) * (ny + 1), dtype=np.int64), ny + 1)
    X = np.stack([ii * (Lx / nx), jj * (Ly / ny)], 1)
    if amp > 0:
        h = min(Lx / nx, Ly / ny)
        d = rng.uniform(-amp * h, amp * h, size=X.shape)
        inner = (ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)
        X[inner] += d[inner]
    node = lambda i, j: i * (ny + 1) + j
    ic, jc = np.divmod(np.arange(nx * ny, dtype=np.int64), ny)
    bl, br, tl, tr = node(ic, jc), node(ic + 1, jc), node(ic, jc + 1), node(ic + 1, jc + 1)
    kind = np.where(rng.random(nx * ny) < tri_fraction, 1 + (rng.random(nx * ny) < 0.5).astype(np.int64), 0)   # 0 quad, 1 / 2 the two diagonals
    # cells in quad order: a quadrilateral (bl, br, tr, tl) or the two triangles of one of its diagonals; then the cell permutation
    is_quad = kind == 0
    n_per_quad = np.where(is_quad, 1, 2)
    first = np.concatenate([[0], np.cumsum(n_per_quad)])[:-1]
    n_cells = int(n_per_quad.sum())
    size = np.full(n_cells, 3, dtype=np.int64)
    size[first[is_quad]] = 4
    onc = np.concatenate([[0], np.cumsum(size)])
    noc = np.empty(int(onc[-1]), dtype=np.int64)
    q4 = first[is_quad]
    for k, nd in enumerate((bl, br, tr, tl)):
        noc[onc[q4] + k] = nd[is_quad]
    d1, d2 = kind == 1, kind == 2
    for sel, ta, tb in ((d1, (bl, br, tr), (bl, tr, tl)), (d2, (bl, br, tl), (br, tr, tl))):
        c0 = first[sel]
        for k in range(3):
            noc[onc[c0] + k] = ta[k][sel]
            noc[onc[c0 + 1] + k] = tb[k][sel]
    if shuffle:
        from .local_mesh import _csr_take
        noc, onc = _csr_take(onc, noc, rng.permutation(n_cells))
    jr, ir = np.arange(ny), np.arange(nx)
    edges = np.concatenate([np.stack([node(0, jr), node(0, jr + 1)], 1), np.stack([node(nx, jr), node(nx, jr + 1)], 1),
                            np.stack([node(ir, ny), node(ir + 1, ny)], 1), np.stack([node(ir, 0), node(ir + 1, 0)], 1)])
    tags = np.concatenate([np.full(ny, 1), np.full(ny, 2), np.full(nx, 3), np.full(nx, 4)])
    return Mesh.from_cells(X, onc, noc, edges, tags, {1: "left", 2: "right", 3: "top", 4: "bottom"})


class LocalPart:
    """What jittered_tri_local returns: the rank-local mesh and how it sits in the global one."""

    def __init__(self, mesh, global_cell_ids, part_local, local, n_global, n_owned, seconds):
        self.mesh, self.global_cell_ids, self.part_local, self.local = mesh, global_cell_ids, part_local, local
        self.n_global, self.n_owned, self.seconds = n_global, n_owned, seconds


def jittered_tri_dual_graph(nx, ny, seed=12345, amp=0.15):
    """Cell-face dual graph (CSR: xadj uint64 [nc + 1], adj uint32) of jittered_tri(nx, ny, ..., seed, amp) in ITS cell numbering, from
    closed forms - which triangles a triangle of the structured parent touches (csrc/mesh_host.cpp gen_tris: cr = 2 q, cl = 2 q + 1 of
    quad q = ic * ny + jc) and the generator's cell permutation - without building the mesh: what mlb_partition_graph_csr needs to
    cut a 16 M- or 64 M-cell mesh on a rank that never holds it (tests/test_host_side.py compares with the mesh's own cells_of_face)."""
    rng = np.random.default_rng(seed)
    nc = 2 * nx * ny
    if amp > 0:
        rng.uniform(-1.0, 1.0, size=((nx + 1) * (ny + 1), 2))      # (the node jitter comes first in the generator's random stream)
    pc = rng.permutation(nc)                                        # new cell i = old cell pc[i]
    ipc = np.empty(nc, dtype=np.int64)
    ipc[pc] = np.arange(nc)
    q = np.arange(nx * ny, dtype=np.int64)
    ic, jc = q // ny, q % ny
    nbr = np.full((nc, 3), -1, dtype=np.int64)                      # old numbering
    nbr[0::2, 0] = 2 * q + 1                                        # cr: the diagonal, the quad below (its cl), the quad to the right (its cl)
    nbr[0::2, 1] = np.where(jc > 0, 2 * (q - 1) + 1, -1)
    nbr[0::2, 2] = np.where(ic < nx - 1, 2 * (q + ny) + 1, -1)
    nbr[1::2, 0] = 2 * q                                            # cl: the diagonal, the quad above (its cr), the quad to the left (its cr)
    nbr[1::2, 1] = np.where(jc < ny - 1, 2 * (q + 1), -1)
    nbr[1::2, 2] = np.where(ic > 0, 2 * (q - ny), -1)
    nbr = nbr[pc]                                                   # rows in the new numbering
    ok = nbr >= 0
    xadj = np.concatenate([[0], np.cumsum(ok.sum(axis=1))]).astype(np.uint64)
    adj = ipc[nbr[ok]].astype(np.uint32)                            # row-major: every row's neighbours are contiguous
    return xadj, adj


def jittered_tri_local(nx, ny, Lx, Ly, n_parts, rank, seed=12345, amp=0.15, layers=10, part_fn=None):
    """The part of jittered_tri(nx, ny, Lx, Ly, seed, amp) that rank `rank` of `n_parts` needs: the cells the library's recursive
    coordinate bisection (mlb_partition_coords on the global centroids) assigns to it plus `layers` layers of quads around them,
    WITHOUT building the global mesh (a 64 M-cell mesh is 8 GB of connectivity per rank; this needs the node coordinates, the two
    id permutations and the centroids).  The arrays are those local_mesh.extract_local would cut out of the global mesh
    (tests/test_host_side.py asserts the equality), because every quantity of the structured parent - which cells a face
    separates, its nodes, its zone - has a closed form in the face's id (csrc/mesh_host.cpp gen_tris)."""
    import time
    from . import partition_coords
    from .local_mesh import CUT
    t0 = time.perf_counter()
    rng = np.random.default_rng(seed)
    nq, nc, col = nx * ny, 2 * nx * ny, 3 * ny + 1
    nf, nn = col * nx + ny, (nx + 1) * (ny + 1)
    # ---- the same random stream as jittered_tri: node jitter, cell permutation, face permutation
    ii, jj = np.divmod(np.arange(nn, dtype=np.int64), ny + 1)
    X = np.stack([ii * (Lx / nx), jj * (Ly / ny)], 1)
    if amp > 0:
        h = min(Lx / nx, Ly / ny)
        d = rng.uniform(-amp * h, amp * h, size=X.shape)
        inner = (ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)
        X[inner] += d[inner]
        del d, inner
    del ii, jj
    pc = rng.permutation(nc)                        # new cell i = old cell pc[i]
    pf = rng.permutation(nf)
    ipc = np.empty(nc, dtype=np.int64); ipc[pc] = np.arange(nc)
    ipf = np.empty(nf, dtype=np.int64); ipf[pf] = np.arange(nf)
    del pf

    node = lambda i, j: i * (ny + 1) + j

    def cell_nodes(old):                            # nodes_of_cell of old cell ids (gen_tris: cr = br,tr,bl ; cl = tl,bl,tr)
        q, upper = old // 2, (old % 2).astype(bool)
        ic, jc = q // ny, q % ny
        bl, br, tl, tr = node(ic, jc), node(ic + 1, jc), node(ic, jc + 1), node(ic + 1, jc + 1)
        return np.stack([np.where(upper, tl, br), np.where(upper, bl, tr), np.where(upper, tr, bl)], 1)

    # ---- global partition from the centroids alone (mean of the three nodes in nodes_of_cell order, mesh/mesh.cpp:167-261)
    if part_fn is None:
        if n_parts > 1:
            Xg = X.reshape(nx + 1, ny + 1, 2)
            bl, br, tl, tr = Xg[:-1, :-1], Xg[1:, :-1], Xg[:-1, 1:], Xg[1:, 1:]
            cold = np.empty((nx, ny, 2, 2))                           # [ic][jc][cr | cl][x, y]: structured numbering, no gathers
            cold[:, :, 0] = ((br + tr) + bl) / 3
            cold[:, :, 1] = ((tl + bl) + tr) / 3
            cxy = cold.reshape(nc, 2)[pc]                             # new cell i = old cell pc[i]
            del cold
            part = partition_coords(cxy, n_parts)
            del cxy
        else:
            part = np.zeros(nc, dtype=np.int32)
    else:
        part = part_fn(nc)
    owned_new = np.nonzero(part == rank)[0]
    # ---- quads to keep: those of the owned cells, grown by `layers` quads in the four grid directions
    keep = np.zeros((nx, ny), dtype=bool)
    keep.reshape(-1)[pc[owned_new] // 2] = True
    for _ in range(layers if n_parts > 1 else 0):
        g = keep.copy()
        g[1:, :] |= keep[:-1, :]; g[:-1, :] |= keep[1:, :]; g[:, 1:] |= keep[:, :-1]; g[:, :-1] |= keep[:, 1:]
        keep = g
    q = np.nonzero(keep.reshape(-1))[0].astype(np.int64)
    old_cells = np.stack([2 * q, 2 * q + 1], 1).reshape(-1)
    new_ids = ipc[old_cells]
    order = np.argsort(new_ids)
    old_cells, gids = old_cells[order], new_ids[order]           # local cell i = global (new) cell gids[i], ascending
    del new_ids, order

    cell_lut = np.full(nc + 1, -1, dtype=np.int32)               # old cell id -> local index, -1 if not kept (slot nc: "no cell")
    cell_lut[old_cells] = np.arange(len(old_cells), dtype=np.int32)

    def local_cell(old):
        return cell_lut[np.where(old >= 0, old, nc)].astype(np.int64)

    # ---- faces of the kept cells (gen_tris: cr = fB,fR,fD ; cl = fT,fL,fD)
    qq, upper = old_cells // 2, (old_cells % 2).astype(bool)
    ic, jc = qq // ny, qq % ny
    fL, fR = col * ic + jc, col * (ic + 1) + jc
    fB = col * ic + 2 * jc + ny
    foc_old = np.stack([np.where(upper, fB + 2, fB), np.where(upper, fL, fR), fB + 1], 1)
    fmask = np.zeros(nf, dtype=bool)
    fmask[foc_old.reshape(-1)] = True
    f_old = np.nonzero(fmask)[0]
    del fmask
    f_new = ipf[f_old]
    forder = np.argsort(f_new)
    f_old, fgids = f_old[forder], f_new[forder]                   # local face i = global (new) face fgids[i], ascending
    del f_new, forder
    face_lut = np.full(nf, -1, dtype=np.int32)
    face_lut[f_old] = np.arange(len(f_old), dtype=np.int32)
    foc = face_lut[foc_old]
    del face_lut
    # every face from its id: column, position in the column
    fi, r = f_old // col, f_old % col
    last = fi == nx                                 # right faces of the last column
    fi_c = np.minimum(fi, nx - 1)
    is_L = ~last & (r < ny)
    rr = r - ny
    is_T = ~last & ~is_L & (rr == 2 * ny)
    is_D = ~last & ~is_L & ~is_T & (rr % 2 == 1)
    is_B = ~last & ~is_L & ~is_T & ~is_D
    fj = np.where(last | is_L, r, np.where(is_T, ny - 1, rr // 2))
    fq = fi_c * ny + fj
    cr_, cl_ = 2 * fq, 2 * fq + 1
    c0 = np.where(is_L | is_T, cl_, cr_)
    c1 = np.where(is_D, cl_, np.where(is_L, np.where(fi_c == 0, -1, cr_ - 2 * ny), np.where(is_B, np.where(fj == 0, -1, cr_ - 1), -1)))
    bl, br, tl, tr = node(fi_c, fj), node(fi_c + 1, fj), node(fi_c, fj + 1), node(fi_c + 1, fj + 1)
    n0 = np.where(is_L, tl, np.where(is_T, tr, np.where(last, br, bl)))
    n1 = np.where(is_L, bl, np.where(is_T, tl, np.where(last, tr, np.where(is_D, tr, br))))
    la, lb = local_cell(c0), local_cell(c1)
    boundary = c1 < 0
    cof0 = np.where(la >= 0, la, lb)
    cof1 = np.where(boundary, -1, np.where((la >= 0) & (lb >= 0), lb, CUT))
    assert (cof0 >= 0).all()
    # ---- nodes
    noc_g = cell_nodes(old_cells)
    nmask = np.zeros(nn, dtype=bool)
    nmask[noc_g.reshape(-1)] = True
    nodes = np.nonzero(nmask)[0]
    node_lut = np.full(nn, -1, dtype=np.int32)
    node_lut[nodes] = np.arange(len(nodes), dtype=np.int32)
    del nmask
    nloc = len(old_cells)
    arrays = dict(node_coords=X[nodes],
                  offsets_nodes_of_cell=np.arange(0, 3 * nloc + 1, 3, dtype=np.uint32), nodes_of_cell=node_lut[noc_g].reshape(-1).astype(np.uint32),
                  offsets_faces_of_cell=np.arange(0, 3 * nloc + 1, 3, dtype=np.uint32), faces_of_cell=foc.reshape(-1).astype(np.uint32),
                  offsets_nodes_of_face=np.arange(0, 2 * len(f_old) + 1, 2, dtype=np.uint32),
                  nodes_of_face=node_lut[np.stack([n0, n1], 1)].reshape(-1).astype(np.uint32),
                  cells_of_face=np.stack([cof0, cof1], 1).astype(np.int32))
    lf = np.arange(len(f_old), dtype=np.uint32)
    zones = [("interior", lf[~boundary]), ("right", lf[last]), ("top", lf[is_T]), ("left", lf[is_L & (fi_c == 0)]), ("bottom", lf[is_B & (fj == 0)])]
    mesh = Mesh.from_arrays(arrays, zones).compute_geometry()
    cell0 = X[cell_nodes(pc[:1])[0]].reshape(6).copy()
    local = dict(global_ids=gids.astype(np.uint32), n_global=nc, cell0_nodes=cell0)
    return LocalPart(mesh, gids.astype(np.uint32), part[gids].astype(np.int32), local, nc, len(owned_new), time.perf_counter() - t0)


def isentropic_vortex(xy, gamma=1.4, u_inf=(0.5, 0.5), beta=5.0, centre=(5.0, 5.0)):
    """Conserved state [nc][4] of the isentropic vortex (rho_inf = p_inf = 1) at the given points."""
    x, y = xy[:, 0] - centre[0], xy[:, 1] - centre[1]
    r2 = x * x + y * y
    f = beta / (2.0 * np.pi) * np.exp(0.5 * (1.0 - r2))
    u = u_inf[0] - f * y
    v = u_inf[1] + f * x
    T = 1.0 - (gamma - 1.0) * beta * beta / (8.0 * gamma * np.pi ** 2) * np.exp(1.0 - r2)
    rho = T ** (1.0 / (gamma - 1.0))
    p = rho ** gamma
    E = p / ((gamma - 1.0) * rho) + 0.5 * (u * u + v * v)
    return np.stack([rho, rho * u, rho * v, rho * E], 1)


EXTRAP4 = [dict(name=n, type="extrapolation") for n in ("left", "right", "top", "bottom")]
