"""The control flow of bench.py / bench_multi.py, end to end, on a box without a GPU (tests/bench_mock.py stands in for the Solver and
for NCCL): what the driver runs at the end of a round must print its one JSON line whatever the late additions to it do."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "bench_mock.py")
CONTRACT = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
            "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]


def _line(stdout):
    lines = [l for l in stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout[-2000:]            # ONE JSON line
    return json.loads(lines[0])


def test_default_single_gpu_line_with_its_strong_records(tmp_path):
    """`python bench.py --steps 20 --warmup 5` as the driver runs it: the configs[1] mesh at full size (mesh generation and the
    initial state are the real code), the end-to-end leg, the finite-data leg, the strong-scaling child process."""
    env = dict(os.environ, MLB_MOCK_STRONG="vortex_16M:40,vortex_64M:5657", MLB_STRONG_BASELINES=str(tmp_path / "live.json"), MLB_EXPERIMENT_NQ="24", MLB_EXPERIMENT_STEPS="10")
    p = subprocess.run([sys.executable, MOCK, "--gpus", "1", "--steps", "20", "--warmup", "5"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _line(p.stdout)
    assert all(k in d for k in CONTRACT), sorted(set(CONTRACT) - set(d))
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 5 and d["config"]["n_cells"] == 2 * 1024 * 1024
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 1024 * 1024 * 32 and d["gpu_launches"] == 200
    assert d["roofline"]["kernel"] == "teno_stream" and 0 < d["roofline"]["frac"] and d["roofline"]["peak"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["config"]["data_independence"]["teno_fixed"] == 1
    kr = d["kernel_rooflines"]                                       # every kernel of the step against its own bound
    assert set(kr) == {"teno_stream", "face_flux_teno", "gather_stage", "cfl"} and all(0 < r["frac"] and r["ms_per_launch"] > 0 for r in kr.values())
    assert kr["face_flux_teno"]["traffic"] > 0 and kr["face_flux_teno"]["frac_traffic"] > 0 and kr["face_flux_teno"]["fp64_pipe_pct_of_peak_ncu"] > 0
    strong = d["strong"]
    assert [r.get("n_cells") for r in strong] == [3200, 2 * 5657 * 5657]
    assert strong[0]["efficiency"] == 1.0 and strong[0]["roofline"]["frac"] > 0 and strong[0]["n_gpus"] == 1
    assert "do not fit" in strong[1]["skipped"]                      # 64 M cells on one GPU
    assert json.load(open(tmp_path / "live.json"))["vortex_16M"]["n_gpus"] == 1
    ex = d["experiments"]                                            # configs[4] numerics and configs[3] as worded, one child process each
    assert [r["workload"].split(":")[0] for r in ex] == ["vortex_viscous", "vortex_mixed", "small_step", "strict_mode", "first_order_33M"]
    assert ex[3]["fp_mode"] == "strict" and ex[3]["roofline"]["stage_frac"] > 0 and ex[4]["n_cells"] == 2 * 24 * 24 and ex[4]["kernels"]["face_flux_fo"]["launches"] > 0
    assert ex[2]["sod"]["n_cells"] == 1000 and ex[2]["wedge"]["cooperative_kernel"]["same_bits_as_the_multi_kernel_path"] is True
    assert ex[0]["n_cells"] == 2 * 24 * 24 and ex[0]["value"] > 0 and ex[0]["inviscid_same_mesh"]["value"] > 0 and ex[0]["finite_fraction_of_cells_after_the_run"] == 1.0
    assert 24 * 24 < ex[1]["n_cells"] < 2 * 24 * 24 and ex[1]["value"] > 0 and "kernels" in ex[1]
    assert ex[0]["fast_vs_strict_residual"] == {"max_difference_of_the_field_scale": 0.0, "finite": True} and "fast_vs_strict_residual" in ex[1]


def test_reference_arm_line_matches_the_contract():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "5"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, MLB_REF_SAMPLE="48x40"))
    assert p.returncode == 0, p.stderr[-2000:]
    d = _line(p.stdout)
    assert d["impl"] == "reference" and d["warmup"] == 5 and d["steps"] == 2 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1 and "48x40" in d["cpu_baseline"]["sample"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(world, args, extra_env):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **extra_env)
        procs.append(subprocess.Popen([sys.executable, MOCK] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    return [p.communicate(timeout=600) + (p.returncode,) for p in procs]


@pytest.mark.parametrize("world", [3])
def test_multi_gpu_line_with_its_strong_records(world, tmp_path):
    """The N > 1 leg as torchrun starts it (one process per rank; gloo instead of NCCL): the weak main line, then the strong records on
    rank-local parts of ONE mesh, efficiency against the base point an earlier run of the series left on the box."""
    live = tmp_path / "live.json"
    json.dump({"vortex_16M": {"n_gpus": 1, "value": 1.0e8, "ms_per_step": 1.0}}, open(live, "w"))
    out = _launch(world, ["--gpus", str(world), "--steps", "6", "--warmup", "3", "--nx", "48", "--ny", "32"],
                  dict(MLB_MOCK_STRONG="vortex_16M:60,vortex_64M:80", MLB_STRONG_BASELINES=str(live), MLB_BENCH_EXIT_GRACE="5"))
    assert [o[2] for o in out] == [0] * world, [o[1][-1500:] for o in out]
    assert all(not [l for l in o[0].splitlines() if l.startswith("{")] for o in out[1:])          # rank 0 alone prints
    d = _line(out[0][0])
    assert all(k in d for k in CONTRACT), sorted(set(CONTRACT) - set(d))
    assert d["n_gpus"] == world and d["scaling"] == "weak" and d["config"]["n_cells"] == 2 * 48 * world * 32 and d["roofline"]["frac"] > 0
    assert d["e2e"]["value"] > 0 and "split-phase" in d["config"]["driver"]
    strong = d["strong"]
    assert [r["n_cells"] for r in strong] == [7200, 12800] and all(r["scaling"] == "strong" and r["n_gpus"] == world for r in strong)
    assert strong[0]["efficiency"] > 0 and "this box" in strong[0]["efficiency_base"] and "efficiency" not in strong[1]
    assert strong[0]["halo"]["max_peers"] >= 1 and strong[0]["ghost_layers"] >= 8 and strong[0]["roofline"]["rank"] == "slowest"
    assert json.load(open(live))["vortex_64M"]["n_gpus"] == world
    ex = d["experiments"]                       # BASELINE configs[4]: the largest mesh that fits, viscous, in per-rank child processes with their own group
    assert len(ex) == 3 and ex[0]["n_cells"] == 12800 and ex[0]["n_gpus"] == world and "Navier-Stokes" in ex[0]["workload"] and ex[0]["value"] > 0
    assert ex[0]["halo"]["max_peers"] >= 1 and "verification" in ex[0]
    # the 16 M-cell mesh of the first strong record once more, cut by the graph partitioner: same base point, no point of its own left behind
    assert ex[1]["partitioner"] == "graph" and ex[1]["n_cells"] == strong[0]["n_cells"] and ex[1]["value"] > 0 and ex[1]["edge_cut"] > 0
    assert strong[0]["partitioner"] == "coordinate bisection" and ex[1]["efficiency_base"] == strong[0]["efficiency_base"] and ex[1]["partition_seconds"] > 0
    assert ex[2]["workload"].startswith("native_weak") and ex[2]["value"] > 0 and ex[2]["n_cells"] == d["config"]["n_cells"]   # the library's own NCCL driver, diagnosed in children


def test_a_failing_strong_record_does_not_cost_the_main_line(tmp_path):
    """One rank fails inside a strong record: rank 0 still prints the main line (with the failure as a record) and every process ends
    with exit code 0 inside the deadline."""
    out = _launch(2, ["--gpus", "2", "--steps", "4", "--warmup", "3", "--nx", "32", "--ny", "32"],
                  dict(MLB_MOCK_STRONG="vortex_16M:60", MLB_STRONG_BASELINES=str(tmp_path / "live.json"), MLB_BENCH_EXIT_GRACE="3",
                       MLB_MOCK_FAIL_STRONG_ON_RANK="1", MLB_BENCH_DEADLINE="60"))
    assert [o[2] for o in out] == [0, 0], [o[1][-1500:] for o in out]
    d = _line(out[0][0])
    assert d["value"] > 0 and d["n_gpus"] == 2
    assert any("error" in r or "aborted" in r for r in d["strong"]), d["strong"]


def test_a_hanging_native_driver_child_is_reported_with_every_ranks_last_trace_line(tmp_path):
    """The diagnostic children of the library's own NCCL driver: a child that does not come back is killed at its time limit, the record
    carries the last `[mlb comm]` line of EVERY rank's child, the second attempt (NCCL settings changed) follows, and the main line and the
    records before it are intact."""
    out = _launch(2, ["--gpus", "2", "--steps", "4", "--warmup", "3", "--nx", "32", "--ny", "32"],
                  dict(MLB_MOCK_STRONG="vortex_16M:60", MLB_STRONG_BASELINES=str(tmp_path / "live.json"), MLB_BENCH_EXIT_GRACE="3",
                       MLB_MOCK_HANG_TASK="native_weak", MLB_NATIVE_TRIAL_LIMIT="10"))
    assert [o[2] for o in out] == [0, 0], [o[1][-1500:] for o in out]
    d = _line(out[0][0])
    assert d["value"] > 0 and d["strong"][0]["value"] > 0
    ex = d["experiments"]
    assert "Navier-Stokes" in ex[0]["workload"] and ex[0]["value"] > 0
    hung = [r for r in ex if r.get("workload") == "native_weak"]
    assert len(hung) == 2 and all("aborted" in r for r in hung)                  # NCCL's defaults, then without graph registration / NVLS
    assert sorted(c["rank"] for c in hung[0]["children"]) == [0, 1]
    assert all("capturing a step" in c["last_lines"][-1] for c in hung[0]["children"])
