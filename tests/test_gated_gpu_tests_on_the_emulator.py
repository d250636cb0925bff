"""The GPU tests of the code paths that have not run on hardware yet (tests/test_gpu_parity.py, behind MLB_RUN_UNVERIFIED) executed on the
HOST: the very test functions, with their meshes, states and tolerances, with mallard_b200.Solver replaced by the emulated kernels
(tests/emul/emulation.py: EmulatedAsSolver - the kernel source compiled for the host, STRICT and FAST builds).  What passes here can still
fail on a B200 for the reasons an emulation cannot see (launch configuration, nvcc's own FMA contractions, barriers), but not because an
assertion or a tolerance was wrong: that is how three assertions that WERE wrong got fixed before their first run on hardware.

The CPU suite runs a cross-section; MLB_EMULATE_GATED=full runs every parametrisation (~3 min; its output is committed as
profiles/r02e_gated_gpu_tests_on_the_emulator.txt)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emulation  # noqa: E402
import mallard_b200 as mb  # noqa: E402
import test_gpu_parity as gp  # noqa: E402

FULL = os.environ.get("MLB_EMULATE_GATED") == "full"
FP = ["strict", "fast"] if emulation.fast_available() else ["strict"]


@pytest.fixture
def emulated(monkeypatch):
    monkeypatch.setattr(mb, "Solver", emulation.EmulatedAsSolver)
    for k in ("MLB_TENO_GENERIC", "MLB_SMALL_STEP", "MLB_SMALL_STEP_BLOCKS"):
        monkeypatch.delenv(k, raising=False)


def _some(cases, keep):
    return cases if FULL else [c for i, c in enumerate(cases) if i in keep]


def test_generic_kernel_forced_onto_specialised_configurations(emulated, monkeypatch):
    gp.test_generic_teno_kernel_is_bit_identical_to_the_specialised_one(monkeypatch)


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("order,factor,qc,basis,fixed", _some([(5, 2.0, 5, "legendre", True), (6, 2.0, 5, "legendre", True), (7, 2.0, 5, "monomial", True),
                                                               (9, 2.0, 5, "legendre", True), (2, 1.5, 0, "legendre", False), (3, 3.0, 0, "monomial", True),
                                                               (4, 2.5, 0, "legendre", True)], {0, 4, 5}))
def test_orders_5_to_9_and_other_stencil_factors(emulated, oracle_mod, order, factor, qc, basis, fixed, fp):
    gp.test_teno_orders_5_to_9_and_other_stencil_factors_vs_oracle(oracle_mod, order, factor, qc, basis, fixed, fp)


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("order,tri_fraction", _some([(1, 0.5), (2, 0.5), (3, 0.5), (3, 0.0), (2, 1.0), (4, 0.6)], {2, 3}))
def test_quadrilateral_and_mixed_meshes(emulated, order, tri_fraction, fp):
    gp.test_teno_on_quadrilateral_and_mixed_meshes_is_k_exact(order, tri_fraction, fp)


def test_first_order_on_a_mixed_mesh(emulated, oracle_mod):
    gp.test_first_order_on_a_mixed_mesh_matches_oracle(oracle_mod)


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("mtype", _some(["cartesian", "cartesian_tri", "mixed"], {0, 2}))
def test_viscous_couette(emulated, mtype, fp):
    gp.test_viscous_residual_of_couette_flow(mtype, fp)


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("mtype", ["cartesian", "mixed"])
def test_isothermal_walls(emulated, mtype, fp):
    gp.test_steady_heat_conduction_between_isothermal_walls(mtype, fp)


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("order,mtype", [(2, "cartesian_tri"), (3, "cartesian"), (5, "cartesian_tri")])
def test_couette_under_teno(emulated, order, mtype, fp):
    gp.test_couette_flow_under_teno_is_a_steady_state_up_to_viscous_heating(order, mtype, fp)


@pytest.mark.skipif(not FULL, reason="~1 min on the host (MLB_EMULATE_GATED=full); the CPU suite integrates the same problem with a fixed dt in test_kernel_emulation.py")
def test_decaying_shear_layer(emulated):
    gp.test_decaying_shear_layer_follows_the_diffusion_equation()


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("case", _some(["sod", "wedge", "tri_hll_rk4"], {0, 2}))
def test_cooperative_small_mesh_kernel(emulated, oracle_mod, monkeypatch, case, fp):
    gp.test_cooperative_small_mesh_kernel_equals_the_multi_kernel_path(oracle_mod, monkeypatch, case, fp)


@pytest.mark.parametrize("fp", FP)
@pytest.mark.parametrize("name", _some(sorted(gp.GENERIC_KERNEL_FIXTURES | gp.LATE_FIXTURES), {0, 2}))
def test_reference_dumps_behind_the_gate(emulated, monkeypatch, name, fp):
    monkeypatch.setenv("MLB_RUN_UNVERIFIED", "1")
    gp.test_against_reference_dumps(name, fp)


def test_injected_jittered_mesh_against_the_unmodified_reference(emulated, tmp_path, capsys):
    if not os.path.exists(gp.REF_HARNESS):
        pytest.skip("the unmodified reference is not built here")
    gp.test_multi_tile_streaming_path_vs_unmodified_reference_on_an_injected_jittered_mesh(tmp_path, capsys)


def test_run_from_a_gmsh_file(emulated, tmp_path):
    """Not gated, but the reader moved behind the C ABI (csrc/mesh_io.cpp) after the GPU suite last ran: the same check with the emulated kernels."""
    gp.test_mesh_read_from_a_gmsh_file_runs_like_the_same_mesh_from_arrays(tmp_path)
