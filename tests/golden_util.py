"""Helpers shared by the parity tests: load tests/golden/*.npz (made by oracle/make_golden.py from the real
reference) and rebuild the same configuration through the oracle."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESH_KEYS = ["node_coords", "cell_coords", "cell_volume", "face_area", "face_normals", "nodes_of_cell",
             "offsets_nodes_of_cell", "faces_of_cell", "offsets_faces_of_cell", "nodes_of_face", "offsets_nodes_of_face",
             "cells_of_face"]
ZONES = ["interior", "right", "top", "left", "bottom"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    meta = json.loads(bytes(d.pop("meta")).decode())
    return meta, d


def names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def oracle_mesh(oracle, meta):
    m = meta["mesh"]
    return oracle.Mesh.generate(m["type"], m["Nx"], m["Ny"], m["Lx"], m["Ly"])


def solver_kwargs(meta):
    r = meta["recon"]
    kw = dict(recon=r["type"], riemann=meta["riemann"], integrator=meta["integrator"], bcs=meta["bcs"])
    if meta.get("physics"):
        kw["gas"] = dict(meta["physics"])
    if r["type"] == "TENO":
        kw.update(basis=r.get("basis_type", "monomial"), order=r["basis_order"], factor=r.get("max_stencil_size_factor", 2.0),
                  quad_cell_order=r.get("quadrature_order_cell", 0), quad_face_order=r.get("quadrature_order_face", 0))
    return kw


def oracle_solver(oracle, meta, mesh=None):
    mesh = mesh or oracle_mesh(oracle, meta)
    return oracle.Solver(mesh, **solver_kwargs(meta))


def real_faces(cells_of_face, nodes_of_face):
    """cartesian_tri over-allocates faces: phantom faces have nodes (0,0)."""
    nof = np.asarray(nodes_of_face).reshape(-1, 2)
    return nof[:, 0] != nof[:, 1]


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over finite entries; non-finite patterns must coincide."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    fa, fb = np.isfinite(a), np.isfinite(b)
    if not np.array_equal(fa, fb):
        return np.inf
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return np.inf
    inf_ok = np.array_equal(a[~fa & ~np.isnan(a)], b[~fb & ~np.isnan(b)])
    if not inf_ok:
        return np.inf
    if not fa.any():
        return 0.0
    scale = np.maximum(np.abs(b[fb]), 1e-300)
    return float(np.max(np.abs(a[fa] - b[fb]) / scale))


def field_err(a, b):
    """Error relative to the FIELD scale: max over variables (last axis) of max|a-b| / max|b|.  Non-finite patterns must
    coincide (NaN with NaN, +-Inf with the same Inf).  This is the metric for FMA-contracted (FAST) results: entries that
    are exact zeros or heavy cancellations in the reference have no meaningful element-wise relative error."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return np.inf
    fa, fb = np.isfinite(a), np.isfinite(b)
    if not np.array_equal(fa, fb) or not np.array_equal(np.isnan(a), np.isnan(b)):
        return np.inf
    if not np.array_equal(a[~fa & ~np.isnan(a)], b[~fb & ~np.isnan(b)]):
        return np.inf
    a2 = np.where(fa, a, 0.0).reshape(-1, a.shape[-1]) if a.ndim > 1 else np.where(fa, a, 0.0).reshape(-1, 1)
    b2 = np.where(fb, b, 0.0).reshape(-1, b.shape[-1]) if b.ndim > 1 else np.where(fb, b, 0.0).reshape(-1, 1)
    col = np.abs(b2).max(axis=0)
    # a component that is (numerically) zero in the reference, e.g. the cross-flow momentum of a 1-D problem, is measured
    # against the largest component: FMA contraction leaves ~1e-17 there where the reference cancels exactly
    scale = np.maximum(np.maximum(col, 1e-3 * col.max()), 1e-300)
    return float(np.max(np.abs(a2 - b2).max(axis=0) / scale))


def elem_err(a, b, floor=1e-9):
    """ELEMENT-wise relative error max |a-b| / max(|b|, floor * column scale) over the entries finite in both (printed and
    bounded next to field_err for FAST-mode results: the floor keeps exact zeros / complete cancellations of the reference
    from dividing by nothing, everything above it is measured against its own magnitude)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    a2 = a.reshape(-1, a.shape[-1]) if a.ndim > 1 else a.reshape(-1, 1)
    b2 = b.reshape(-1, b.shape[-1]) if b.ndim > 1 else b.reshape(-1, 1)
    ok = np.isfinite(a2) & np.isfinite(b2)
    if not ok.any():
        return 0.0
    col = np.where(ok, np.abs(b2), 0.0).max(axis=0)
    den = np.maximum(np.abs(b2), floor * np.maximum(col, 1e-300)[None, :])
    return float(np.max(np.where(ok, np.abs(a2 - b2) / den, 0.0)))


def pattern_mismatch(a, b):
    """Fraction of entries whose finite / NaN / +-Inf class differs."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    cls = lambda x: np.where(np.isnan(x), 3, np.where(np.isposinf(x), 1, np.where(np.isneginf(x), 2, 0)))
    return float((cls(a) != cls(b)).mean())
