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
