"""End-to-end drop-in check: the UNMODIFIED reference host (TOML input, mesh generator, exprtk initial condition, checks,
VTU writer; oracle/_ref/bin/mallard_dropin, see oracle/dropin_harness.cpp and INTEGRATION.md) with Solver::run's three
hot-path seams rerouted through libmallard_b200.so, against the unmodified reference itself (oracle/_ref/bin/Mallard,
Kokkos on the host cores), on the reference's own example inputs.  Compared artefact: the VTU files both runs write.
Both binaries are built where /root/reference exists (oracle/build_ref.sh) and travel to the GPU box prebuilt."""
import hashlib
import json
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "Mallard")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "bin", "mallard_dropin")

SOD = """
[run]
t_stop = 0.2
cfl = 1.0
[mesh]
type = "cartesian"
Nx = 1000
Ny = 1
Lx = 1.0
Ly = 0.001
[initialize]
type = "analytical"
rho = "var l := x <  0.5; var r := x >= 0.5; 1.0 * l + 0.125 * r"
u = ["0.0", "0.0"]
p = "var l := x <  0.5; var r := x >= 0.5; 1.0 * l + 0.1 * r"
%(bcs)s
[numerics]
riemann_solver = "HLLC"
time_integrator = "SSPRK3"
check_nan = true
[numerics.face_reconstruction]
type = "FO"
[physics]
type = "euler"
gamma = 1.4
p_ref = 101325.0
T_ref = 298.15
rho_ref = 1.225
[output]
check_interval = 100
[[write_data]]
prefix = "./solut/all/sod_all"
format = "vtu"
geometry = "all"
interval = 100
variables = ["CFL", "RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X",  "U_Y", "P", "T", "H"]
""" % dict(bcs="\n".join('[[boundaries]]\nname = "%s"\ntype = "symmetry"' % n for n in ("left", "right", "top", "bottom")))

WEDGE = """
[run]
n_steps = 400
cfl = 1.0
[mesh]
type = "wedge"
Nx = 150
Ny = 50
Lx = 4.0
Ly = 1.5
[initialize]
type = "constant"
u = [600.0, 0.0]
p = 101325.0
T = 300.0
[[boundaries]]
name = "left"
type = "upt"
u = [600.0, 0.0]
p = 101325.0
T = 300.0
[[boundaries]]
name = "right"
type = "p_out"
p = 101325.0
[[boundaries]]
name = "top"
type = "symmetry"
[[boundaries]]
name = "bottom"
type = "symmetry"
[numerics]
riemann_solver = "HLLC"
time_integrator = "SSPRK3"
check_nan = true
[numerics.face_reconstruction]
type = "FO"
[physics]
type = "euler"
gamma = 1.4
p_ref = 101325.0
T_ref = 298.15
rho_ref = 1.225
[output]
check_interval = 100
[[write_data]]
prefix = "./solut/all/wedge_all"
format = "vtu"
geometry = "all"
interval = 200
variables = ["CFL", "RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X",  "U_Y", "P", "T", "H"]
"""


def read_vtu(path):
    """Arrays of a reference-written VTU (raw appended data, 4-byte length headers; src/io/data_writer.cpp:93-244)."""
    raw = open(path, "rb").read()
    head, _, tail = raw.partition(b'<AppendedData encoding="raw">')
    payload = tail[tail.index(b"_") + 1:]
    out = {}
    dt = {"Float64": np.float64, "Float32": np.float32, "UInt32": np.uint32, "UInt8": np.uint8, "Int32": np.int32}
    for i, m in enumerate(re.finditer(rb'<DataArray type="(\w+)"(?: Name="(\w+)")?[^>]*offset="(\d+)"', head)):
        typ, name, off = m.group(1).decode(), (m.group(2) or b"points%d" % i).decode(), int(m.group(3))
        n = int(np.frombuffer(payload, np.uint32, 1, off)[0])
        out[name] = np.frombuffer(payload, dt[typ], n // np.dtype(dt[typ]).itemsize, off + 4)
    return out


def run(binary, toml, workdir, extra=()):
    os.makedirs(os.path.join(workdir, "solut", "all"))
    open(os.path.join(workdir, "input.toml"), "w").write(toml)
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")   # serial reference: deterministic atomics
    p = subprocess.run([binary, "-i", "input.toml", *extra], cwd=workdir, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN)), reason="oracle/_ref binaries not built (need /root/reference)")
@pytest.mark.parametrize("name,toml,fp,tol", [("sod", SOD, "strict", 1e-9), ("sod", SOD, "fast", 1e-9), ("wedge", WEDGE, "strict", 0.0),
                                              ("wedge", WEDGE, "fast", 1e-10)], ids=["sod-strict", "sod-fast", "wedge-strict", "wedge-fast"])
def test_reference_host_with_b200_hot_path_writes_the_reference_solution(tmp_path, name, toml, fp, tol):
    a, b = str(tmp_path / "ref"), str(tmp_path / "b200")
    run(REF, toml, a)
    out = run(DROPIN, toml, b, ("--fp", fp, "--quiet"))
    info = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    fa, fb = sorted(os.listdir(os.path.join(a, "solut", "all"))), sorted(os.listdir(os.path.join(b, "solut", "all")))
    assert fa == fb and len(fa) >= 3                       # same output schedule: same number of steps, same stop condition
    worst, identical = 0.0, 0
    for f in fa:
        pa, pb = os.path.join(a, "solut", "all", f), os.path.join(b, "solut", "all", f)
        identical += hashlib.md5(open(pa, "rb").read()).hexdigest() == hashlib.md5(open(pb, "rb").read()).hexdigest()
        va, vb = read_vtu(pa), read_vtu(pb)
        assert va.keys() == vb.keys()
        scale = {k: np.abs(v).max() for k, v in va.items()}
        for k in va:
            if va[k].dtype.kind != "f":
                assert np.array_equal(va[k], vb[k]), (f, k)
                continue
            # vector components share one scale (the cross-flow component of a 1-D problem is exactly 0 in the reference)
            group = [g for g in (("RHOU_X", "RHOU_Y"), ("U_X", "U_Y")) if k in g]
            ref_scale = max([scale[k]] + [scale[x] for g in group for x in g] + [1e-300])
            worst = max(worst, float(np.abs(va[k] - vb[k]).max() / ref_scale))
    print("%s[%s]: %d steps, %d/%d VTU files byte-identical to the reference's, worst field difference %.2e; %.3g cell-updates/s/stage "
          "through the reference host loop" % (name, fp, info["steps"], identical, len(fa), worst, info["cell_updates_per_s_per_stage"]))
    assert worst <= tol
    if tol == 0.0:
        assert identical == len(fa)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN)), reason="oracle/_ref binaries not built (need /root/reference)")
@pytest.mark.parametrize("name,toml,tol", [("sod", SOD, 1e-9), ("wedge", WEDGE, 0.0)], ids=["sod", "wedge"])
def test_native_vtu_writer_and_device_side_checks(tmp_path, name, toml, tol):
    """--native-io replaces the reference's per-step copy_device_to_host + do_checks + check_fields + VTU writer
    with mlb_field_ranges (device reduction) and mlb_write_vtu (fed from the device-resident fields).  The files must be
    the reference writer's byte for byte wherever the fields are (wedge, STRICT mode), and the "Scalar range" lines of the
    log must read the same."""
    a, b = str(tmp_path / "ref"), str(tmp_path / "native")
    out_ref = run(REF, toml, a)
    out_nat = run(DROPIN, toml, b, ("--fp", "strict", "--native-io"))
    fa = sorted(os.listdir(os.path.join(a, "solut", "all")))
    fb = sorted(os.listdir(os.path.join(b, "solut", "all")))
    assert [f.replace("_native", "") for f in fb] == fa and len(fa) >= 3
    for f, g in zip(fa, fb):
        ra, rb = open(os.path.join(a, "solut", "all", f), "rb").read(), open(os.path.join(b, "solut", "all", g), "rb").read()
        assert len(ra) == len(rb)
        # the XML part (header, offsets, footer) is identical whatever the field values
        cut = ra.index(b'<AppendedData encoding="raw">')
        assert ra[:cut] == rb[:cut] and ra[-40:] == rb[-40:]
        if tol == 0.0:
            assert ra == rb, f
        else:
            va, vb = read_vtu(os.path.join(a, "solut", "all", f)), read_vtu(os.path.join(b, "solut", "all", g))
            for k in va:
                if va[k].dtype.kind != "f":
                    assert np.array_equal(va[k], vb[k])
                else:
                    assert np.abs(va[k] - vb[k]).max() <= tol * max(np.abs(va[k]).max(), 1.0), (f, k)
    ranges = lambda txt: [l for l in txt.splitlines() if l.startswith("> Scalar range:")]
    ra, rb = ranges(out_ref), ranges(out_nat)
    assert len(ra) == len(rb) and len(ra) >= 9
    if tol == 0.0:
        assert ra == rb


# ---- examples/wedge to its own stop time (t_stop = 0.1, 7938 steps; the parametrised case above stops after 400) -----------
WEDGE_FULL = WEDGE.replace("n_steps = 400", "t_stop = 0.1").replace("interval = 200", "interval = 2000")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN)), reason="oracle/_ref binaries not built (need /root/reference)")
def test_wedge_to_t_stop_through_the_reference_host(tmp_path, capsys):
    """BASELINE configs[2] verbatim (examples/wedge/input.toml: t_stop = 0.1): the reference host loop decides when to stop
    from the dt the library returns, so the step count (7938) is itself a parity result.  Once the oblique shock has formed HLLC's
    star-pressure estimate leaves the PVRS branch for TRRS, whose pow() is libm's on the host and CUDA's on the device (last-ulp
    differences), so byte-identity holds for the files written before that (0, 2000) and a tolerance after it;
    the drift over the ~8000 steps of the run is printed."""
    a = str(tmp_path / "ref")
    run(REF, WEDGE_FULL, a)
    fa = sorted(os.listdir(os.path.join(a, "solut", "all")))
    assert len(fa) >= 5
    for fp, tol in (("strict", 1e-11), ("fast", 1e-9)):
        b = str(tmp_path / fp)
        out = run(DROPIN, WEDGE_FULL, b, ("--fp", fp, "--quiet"))
        info = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        fb = sorted(os.listdir(os.path.join(b, "solut", "all")))
        assert fa == fb, (fa, fb)                              # same number of steps to t_stop: the last file carries the step count
        worst, identical, per_file = 0.0, 0, []
        for f in fa:
            ra, rb = open(os.path.join(a, "solut", "all", f), "rb").read(), open(os.path.join(b, "solut", "all", f), "rb").read()
            identical += ra == rb
            va, vb = read_vtu(os.path.join(a, "solut", "all", f)), read_vtu(os.path.join(b, "solut", "all", f))
            wf = 0.0
            for k in va:
                if va[k].dtype.kind == "f":
                    group = [g for g in (("RHOU_X", "RHOU_Y"), ("U_X", "U_Y")) if k in g]
                    sc = max([np.abs(va[k]).max()] + [np.abs(va[x]).max() for g in group for x in g] + [1e-300])
                    wf = max(wf, float(np.abs(va[k] - vb[k]).max() / sc))
                else:
                    assert np.array_equal(va[k], vb[k]), (f, k)
            per_file.append("%s %.1e" % (f[-10:-4], wf))
            worst = max(worst, wf)
        with capsys.disabled():
            print("\nwedge to t_stop [%s]: %d steps, %d/%d files byte-identical, field difference per file: %s" % (fp, info["steps"], identical, len(fa), ", ".join(per_file)))
        assert worst <= tol
        if fp == "strict":
            assert identical >= 2


# ---- TENO through the reference host ------------------------------------------------------------------------------------
def _teno_toml(ic, n_steps, basis):
    rho, ux, uy, pp = ic
    return """
[run]
n_steps = %d
cfl = 0.1
[mesh]
type = "cartesian_tri"
Nx = 24
Ny = 20
Lx = 1.0
Ly = 1.0
[initialize]
type = "analytical"
rho = "%s"
u = ["%s", "%s"]
p = "%s"
%s
[numerics]
riemann_solver = "HLLC"
time_integrator = "SSPRK3"
check_nan = false
[numerics.face_reconstruction]
type = "TENO"
%sbasis_order = 3
max_stencil_size_factor = 2.0
[physics]
type = "euler"
gamma = 1.4
p_ref = 101325.0
T_ref = 298.15
rho_ref = 1.225
[output]
check_interval = 1
[[write_data]]
prefix = "./solut/all/teno_all"
format = "vtu"
geometry = "all"
interval = 1
variables = ["CFL", "RHO", "RHOU_X", "RHOU_Y", "RHOE", "U_X",  "U_Y", "P", "T", "H"]
""" % (n_steps, rho, ux, uy, pp, "\n".join('[[boundaries]]\nname = "%s"\ntype = "symmetry"' % n for n in ("left", "right", "top", "bottom")),
       ('basis_type = "%s"\n' % basis) if basis else "")


_Q = "var l := x <  0.8; var r := x >= 0.8; var b := y <  0.8; var t := y >= 0.8; "
TENO_RIEMANN = (_Q + "1.5 * r * t + 0.532258064516129 * l * t + 0.137992831541219 * l * b + 0.532258064516129 * r * b",
                _Q + "0.0 * r * t + 1.206045378311055 * l * t + 1.206045378311055 * l * b + 0.0 * r * b",
                _Q + "0.0 * r * t + 0.0 * l * t + 1.206045378311055 * l * b + 1.206045378311055 * r * b",
                _Q + "1.5 * r * t + 0.3 * l * t + 0.029032258064516 * l * b + 0.3 * r * b")
TENO_SMOOTH = ("1.0 + 0.2 * sin(2 * pi * x) * cos(2 * pi * y)", "0.5 + 0.1 * cos(2 * pi * x)", "0.3 + 0.1 * sin(2 * pi * y)",
               "1.0 + 0.1 * cos(2 * pi * (x + y))")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DROPIN)), reason="oracle/_ref binaries not built (need /root/reference)")
@pytest.mark.parametrize("ic,n_steps,basis", [(TENO_RIEMANN, 1, "legendre"), (TENO_SMOOTH, 1, "legendre"), (TENO_SMOOTH, 1, None)],
                         ids=["riemann_2d-IC", "smooth-IC", "smooth-IC-default-basis(monomial)"])
@pytest.mark.parametrize("fp", ["strict", "fast"])
def test_teno_toml_through_the_reference_host(tmp_path, ic, n_steps, basis, fp, capsys):
    """examples/riemann_2d's numerics block (TENO + HLLC + SSPRK3, cfl 0.1) on a small cartesian_tri mesh through the
    UNMODIFIED reference host with the hot path rerouted, against the stock binary, after the first step.  On the
    four-quadrant IC the reference turns non-finite inside that step: the non-finite pattern of every written
    field must coincide and the finite entries agree (NaN-pattern-aware comparison)."""
    toml = _teno_toml(ic, n_steps, basis)
    a, b = str(tmp_path / "ref"), str(tmp_path / "b200")
    run(REF, toml, a)
    run(DROPIN, toml, b, ("--fp", fp, "--quiet"))
    fa, fb = sorted(os.listdir(os.path.join(a, "solut", "all"))), sorted(os.listdir(os.path.join(b, "solut", "all")))
    assert fa == fb and len(fa) == n_steps + 1
    worst, n_bad = 0.0, 0
    for f in fa:
        va, vb = read_vtu(os.path.join(a, "solut", "all", f)), read_vtu(os.path.join(b, "solut", "all", f))
        for k in va:
            if va[k].dtype.kind != "f":
                assert np.array_equal(va[k], vb[k]), (f, k)
                continue
            x, y = va[k].astype(np.float64), vb[k].astype(np.float64)
            assert np.array_equal(np.isnan(x), np.isnan(y)), (f, k, int(np.isnan(x).sum()), int(np.isnan(y).sum()))
            assert np.array_equal(np.isinf(x), np.isinf(y)) and np.array_equal(x[np.isinf(x)], y[np.isinf(y)]), (f, k)
            ok = np.isfinite(x)
            n_bad += int((~ok).sum())
            if ok.any():
                # each finite entry against its own magnitude, floored at 1e-6 of the field's largest finite entry (the
                # discontinuity branch leaves entries many orders of magnitude apart in one field)
                den = np.maximum(np.abs(x[ok]), 1e-6 * np.abs(x[ok]).max() + 1e-300)
                worst = max(worst, float(np.max(np.abs(x[ok] - y[ok]) / den)))
    with capsys.disabled():
        print("\nTENO through the reference host [%s]: %d non-finite entries (coinciding), worst finite difference %.2e" % (fp, n_bad, worst))
    assert worst <= (1e-11 if fp == "strict" else 1e-9)
