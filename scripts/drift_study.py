#!/usr/bin/env python
"""Drift over a run (north_star: "<= 1e-12 per step, with drift reported over the run").

Vortex advection on a jittered unstructured mesh, normalised TENO weights (the reference-faithful ones turn non-finite within
a step), cfl 0.1, N steps.  Four solutions of the same problem:
    S   STRICT (bit-faithful mode)
    S'  STRICT started from the initial state perturbed by +-1 ulp per entry (seeded signs)
    F   FAST (FMA contraction, compact tables, re-associated sums)
    F'  FAST from the perturbed state
and the distances |F - S|, |S' - S|, |F' - F| relative to the field scale, sampled every `every` steps.  If |F - S| tracks
|S' - S| (same growth rate, same order of magnitude), the growth is the scheme's own sensitivity to rounding-level
perturbations (the non-linear TENO weights switch stencils), not an error the FAST kernels accumulate.
Usage: python scripts/drift_study.py [n_steps=2000] [every=100] [nx=96] [mesh=jittered|cartesian] [cfl=0.1] [order=3] > profiles/...
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
import mallard_b200 as mb  # noqa: E402
from mallard_b200 import synthetic as syn  # noqa: E402


def main():
    n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    every = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    nx = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    kind = sys.argv[4] if len(sys.argv) > 4 else "jittered"
    cfl = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
    order = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    mesh = syn.jittered_tri(nx, nx, 10.0, 10.0, seed=5) if kind == "jittered" else syn.jittered_tri(nx, nx, 10.0, 10.0, seed=5, amp=0.0, shuffle=False)
    xy = mesh.arrays["cell_coords"]
    U0 = syn.isentropic_vortex(xy)
    rng = np.random.default_rng(2026)
    U0p = U0 * (1.0 + np.where(rng.random(U0.shape) < 0.5, -1.0, 1.0) * 2.0 ** -52)      # +-1 ulp per entry
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", order=order, bcs=syn.EXTRAP4, teno_fixed=True, keep_stage_rhs=False)
    sol = {"S": mb.Solver(mesh, fp_mode="strict", **kw), "S'": mb.Solver(mesh, fp_mode="strict", **kw),
           "F": mb.Solver(mesh, fp_mode="fast", **kw), "F'": mb.Solver(mesh, fp_mode="fast", **kw)}
    sol["S"].set_state(U0); sol["F"].set_state(U0); sol["S'"].set_state(U0p); sol["F'"].set_state(U0p)
    print("# drift study: %s %dx%d triangulation (%d cells), isentropic vortex, TENO p=%d (normalised weights) + HLLC + SSPRK3, cfl %g"
          % (kind, nx, nx, mesh.n_cells, order, cfl))
    print("# columns: step, t, |F-S| (FAST vs STRICT), |S'-S| (STRICT, 1-ulp perturbed start), |F'-F| (FAST, 1-ulp perturbed start), "
          "growth of |F-S| and |S'-S| per 100 steps")
    prev = None
    rows = []
    for k in range(n_steps // every):
        try:
            for s in sol.values():
                s.run(every, cfl=cfl)
        except mb.MallardError as ex:
            print("# run ended before step %d: %s (the solution itself left the finite range)" % (every * (k + 1), ex))
            break
        U = {n: s.get_state() for n, s in sol.items()}
        if not np.isfinite(U["S"]).all():
            print("# STRICT solution non-finite at step %d" % (every * (k + 1)))
            break
        d_fs, d_ss, d_ff = gu.field_err(U["F"], U["S"]), gu.field_err(U["S'"], U["S"]), gu.field_err(U["F'"], U["F"])
        g = "" if prev is None else "  x%.2f  x%.2f" % ((d_fs / prev[0]) ** (100.0 / every), (d_ss / prev[1]) ** (100.0 / every))
        prev = (d_fs, d_ss)
        rows.append((every * (k + 1), d_fs, d_ss, d_ff))
        w = int(np.argmax(np.abs(U["S'"] - U["S"]).max(axis=1)))
        print("%6d  t=%.4f  %.3e  %.3e  %.3e%s   max|S'-S| at (%.2f, %.2f), rho range [%.4f, %.4f]"
              % (every * (k + 1), sol["S"].time()[0], d_fs, d_ss, d_ff, g, xy[w, 0], xy[w, 1], U["S"][:, 0].min(), U["S"][:, 0].max()), flush=True)
    if not rows:
        return
    r = np.array(rows)
    ratio = r[:, 1] / r[:, 2]
    print("# |F-S| / |S'-S| over the run: min %.2f median %.2f max %.2f" % (ratio.min(), np.median(ratio), ratio.max()))
    print("# per-step average of |F-S| over the first %d steps: %.2e" % (every, r[0, 1] / every))


if __name__ == "__main__":
    main()
