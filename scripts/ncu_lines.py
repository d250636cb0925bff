#!/usr/bin/env python
"""Per-CUDA-source-line stall samples from an ncu report (needs -lineinfo).  Usage: ncu_lines.py rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = None; recs = []; total = 0.0
for r in rows:
    if r and r[0] == "Line No":
        h = r; ix = {}
        for i, n in enumerate(h): ix.setdefault(n, i)
        stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if h is None or len(r) < len(h) or r[ix["Address"]] != "-": continue     # keep the per-source-line summary rows only
    try: ns = float(r[ix["# Samples"]] or 0)
    except ValueError: continue
    total += ns
    st = sorted(((float(r[ix[c]] or 0), c[6:]) for c in stall), reverse=True)[:3]
    recs.append((ns, r[0], r[1].strip()[:95], " ".join("%s:%d" % (n, v) for v, n in st if v), r[ix["Instructions Executed"]]))
print("total samples", total)
for ns, ln, src, why, ne in sorted(recs, reverse=True)[:top]:
    print("%5.1f%% L%-4s %-95s | %s | inst %s" % (100 * ns / max(total, 1), ln, src, why, ne))
