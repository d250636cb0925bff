#!/usr/bin/env python
"""Summarise an ncu report (raw page) and a launch list into text for profiles/.  Usage: ncu_summary.py prof.ncu-rep [launches.csv]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio"]


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
BENCH_NAME = [("teno_stream", "teno_stream"), ("teno_recon", "teno_recon"), ("face_flux", "face_flux_teno"), ("gather_stage", "gather_stage"), ("cfl_kernel", "cfl")]


def write_traffic(rows, idx, units, path, meta):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (first captured launch of each kernel) -> JSON for bench.py"""
    import json
    out = {}
    for r in rows:
        name = r[idx["Kernel Name"]]
        key = next((b for pat, b in BENCH_NAME if pat in name), None)
        if key is None or key in out:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]].replace(",", "")) * UNIT[units[idx[m]]]
        out[key] = tot
    meta = dict(kv.split("=", 1) for kv in meta)
    if "n_cells" in meta:
        meta["n_cells"] = int(meta["n_cells"])
    json.dump(dict(meta, bytes_per_launch=out, source="ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum)"),
              open(path, "w"), indent=1)


def main():
    if "--traffic" in sys.argv:      # ncu_summary.py prof.ncu-rep --traffic out.json key=value ...
        i = sys.argv.index("--traffic")
        raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        write_traffic(rows[2:], {h: j for j, h in enumerate(rows[0])}, rows[1], sys.argv[i + 1], sys.argv[i + 2:])
        return
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        print("==== %s  (grid %s, block %s)" % (name, r[idx.get("launch__grid_size", 0)], r[idx.get("launch__block_size", 0)]))
        for w in WANT:
            if w in idx and r[idx[w]] not in ("", "0", "n/a"):
                print("  %-86s %s %s" % (w, r[idx[w]], units[idx[w]]))
    if len(sys.argv) > 2:
        rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
        h = rows[0]
        tot = defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            d = dict(zip(h, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                k = d["Kernel Name"][:70]
                tot[k][0] += 1
                tot[k][1] += float(d["Metric Value"].replace(",", ""))
        all_t = sum(t for _, t in tot.values())
        print("==== launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print("  %-70s launches %4d  total %10.3f ms  avg %9.3f ms  share %5.1f%%" % (k, n, t / 1e6, t / n / 1e6, 100 * t / all_t))


if __name__ == "__main__":
    main()
