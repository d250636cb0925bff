#!/usr/bin/env python
"""Per-instruction stall attribution from an ncu report's source page.  Usage: ncu_stalls.py rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"], capture_output=True, text=True).stdout
lines = raw.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = defaultdict(float); per_op = defaultdict(float); recs = []
total_samples = 0
seen = set()
for r in rows[1:]:
    if len(r) < len(h) or r[ix["Address"]] in seen: continue
    seen.add(r[ix["Address"]])
    try: ns = float(r[ix["# Samples"]] or 0)
    except ValueError: continue
    total_samples += ns
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    per_op[op.split(".")[0]] += ns
    st = {c: float(r[ix[c]] or 0) for c in stall_cols}
    for c, v in st.items(): tot[c] += v
    why = " ".join("%s:%d" % (c[6:], v) for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    recs.append((ns, r[ix["Address"]][-5:], r[ix["Source"]][:60], why))
print("total samples", total_samples)
print("by stall reason:", ", ".join("%s %.1f%%" % (c[6:], 100 * v / max(1, sum(tot.values()))) for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
print("by opcode:", ", ".join("%s %.1f%%" % (o, 100 * v / max(1, total_samples)) for o, v in sorted(per_op.items(), key=lambda kv: -kv[1])[:10]))
for ns, addr, src, why in sorted(recs, reverse=True)[:top]:
    print("%6.2f%%  %s  %-60s %s" % (100 * ns / max(1, total_samples), addr, src, why))
