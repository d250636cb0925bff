import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import mallard_b200 as mb, oracle
bcs = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]
mesh = mb.Mesh.generate("cartesian_tri", 12, 10, 1.0, 1.0)
xy = mesh.arrays["cell_coords"]
rho = 1.0 + 0.2 * np.sin(2 * np.pi * xy[:, 0]) * np.cos(2 * np.pi * xy[:, 1])
u, v, p = 0.5 + 0.1 * np.cos(2 * np.pi * xy[:, 0]), 0.3 + 0.1 * np.sin(2 * np.pi * xy[:, 1]), 1.0 + 0.1 * np.cos(2 * np.pi * (xy[:, 0] + xy[:, 1]))
E = p / (0.4 * rho) + 0.5 * (u * u + v * v)
U0 = np.stack([rho, rho * u, rho * v, rho * E], 1)
for ren in ("none", "rcm"):
    s = mb.Solver(mesh, "TENO", "HLLC", "SSPRK3", bcs=bcs, fp_mode="strict", renumber=ren)
    s.set_state(U0)
    F = s.calc_face_values()
    om = oracle.Mesh.generate("cartesian_tri", 12, 10, 1.0, 1.0)
    so = oracle.Solver(om, "TENO", "HLLC", "SSPRK3", bcs=bcs)
    so.set_state(U0)
    Fo = so.calc_face_values()
    cof = mesh.arrays["cells_of_face"]; nof = mesh.arrays["nodes_of_face"].reshape(-1, 2)
    real = nof[:, 0] != nof[:, 1]
    print(ren, "nan count F", np.isnan(F[real]).sum(), "of", F[real].size)
    d = np.abs(F - Fo)
    d[~real] = 0; d[real & (cof[:, 1] < 0), :, 1] = 0
    print(" max abs diff", np.nanmax(d), "argmax", np.unravel_index(np.nanargmax(d), d.shape))
    bad = np.argwhere(np.isnan(F) & real[:, None, None, None])
    print(" first bad", bad[:10].tolist())
    rhs = s.calc_rhs(); rhso = so.calc_rhs()
    print(" rhs nan", np.isnan(rhs).sum(), "max diff", np.nanmax(np.abs(rhs - rhso)))
    print(" F sample", F[real][0].ravel()[:8], Fo[real][0].ravel()[:8])
