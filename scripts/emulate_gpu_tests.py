#!/usr/bin/env python
"""TEST INFRASTRUCTURE - the UNGATED parity tests of tests/test_gpu_parity.py (the ones that have run on a B200) executed on the host with the
emulated kernels in place of mallard_b200.Solver: a regression check of the host code (preprocessor, stage schedule) and of the kernel source
at HEAD against the oracle and the reference's dumps, for boxes without a GPU.  Covers what tests/emul can run (no PTX streaming kernels).  Usage: python scripts/emulate_gpu_tests.py  -> "ok N bad 0"."""
import os, sys, itertools, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('', 'tests', os.path.join('tests', 'emul'), 'oracle'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import emulation, mallard_b200 as mb, oracle
oracle.build()
mb.Solver = emulation.EmulatedAsSolver
import test_gpu_parity as gp
ok=bad=0
def run(name, fn, *a):
    global ok,bad
    try:
        fn(*a); ok+=1; print("PASS", name, a if len(str(a))<90 else '', flush=True)
    except Exception as ex:
        bad+=1; print("FAIL", name, a, type(ex).__name__, str(ex)[:300].replace("\n"," "), flush=True)
seeded=[("cartesian", 96, 64, "FO", "HLLC", "SSPRK3", False), ("cartesian", 50, 70, "FO", "HLL", "RK4", False),
    ("cartesian_tri", 40, 30, "FO", "Rusanov", "FE", False), ("wedge", 60, 20, "FO", "HLLC", "SSPRK3", False),
    ("cartesian_tri", 24, 20, "TENO", "HLLC", "SSPRK3", False), ("cartesian_tri", 16, 18, "TENO", "Rusanov", "RK4", False),
    ("cartesian_tri", 24, 20, "TENO", "HLLC", "SSPRK3", True), ("cartesian_tri", 18, 16, "TENO", "HLL", "FE", True)]
for fp in ("strict","fast"):
    for c in seeded:
        run("oracle_seeded", gp.test_against_oracle_seeded, oracle, *c, fp)
    for c in [("FO", "HLLC", "SSPRK3", False), ("TENO", "HLLC", "SSPRK3", True), ("TENO", "HLLC", "SSPRK3", False), ("TENO", "HLL", "RK4", True)]:
        run("jittered", gp.test_jittered_unstructured_mesh_vs_oracle, oracle, *c, fp)
import golden_util as gu
os.environ["MLB_RUN_UNVERIFIED"]="1"
for fp in ("strict","fast"):
    for name in gu.names():
        meta,_=gu.load(name)
        run("dumps", gp.test_against_reference_dumps, name, fp)
print("ok",ok,"bad",bad)
