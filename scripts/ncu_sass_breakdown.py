#!/usr/bin/env python
"""Instruction mix and stall attribution of one kernel from an `ncu --set full --import-source on` report, per SASS opcode class:
    ncu -i gpurun_out/prof.ncu-rep --page source --csv --kernel-name regex:face_flux --launch-count 1 > ff.csv
    python scripts/ncu_sass_breakdown.py ff.csv > profiles/r02i_face_flux_sass_breakdown.txt
Answers "what does this kernel issue, and what do its warps wait for" without a GPU (the report is read, nothing is measured here)."""
import csv
import re
import sys
from collections import defaultdict

CLASSES = [("FP64 fused multiply-add / mul / add", r"^(DFMA|DMUL|DADD)"), ("FP64 compare / min / max / set", r"^(DSETP|DMNMX|DSET)"),
           ("special function unit (reciprocal / rsqrt seeds)", r"^MUFU"), ("global / local loads", r"^(LDG|LDL|LD\b|LDC|ULDC)"),
           ("global / local stores", r"^(STG|STL|ST\b)"), ("shuffles", r"^SHFL"), ("integer / address arithmetic", r"^(IADD|IMAD|LEA|SHF|LOP|IABS|ISCADD|UIADD|UIMAD|ULEA|USHF|ULOP|I2F|F2I|I2I|SEL|USEL|PRMT|MOV|UMOV|S2R|CS2R|R2UR|S2UR|IMNMX|VIMNMX|UISETP|FSEL)"),
           ("predicates / integer compares", r"^(ISETP|PLOP|P2R|R2P|FSETP|UPLOP|VOTE)"), ("branches / control", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|WARPSYNC|NOP|BAR|JMP|BRX|YIELD|DEPBAR)"),
           ("FP32 (division / sqrt seeds and corrections)", r"^(FFMA|FMUL|FADD|FCHK|F2F|FMNMX)")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    kernel = rows[0][1]
    head = rows[1]
    col = {n: i for i, n in enumerate(head)}
    stall_cols = [n for n in head if n.startswith("stall_") and "Not Issued" not in n]
    mix = defaultdict(lambda: [0, 0, 0, defaultdict(int)])       # class -> [static instructions, executed, samples, stalls]
    total_exec = total_samples = 0
    top = []
    n_sass = 0
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        if r[0] == "Address" or r[0] == "Kernel Name":      # the next launch of the kernel in the same report: one launch is enough
            break
        n_sass += 1
        src = r[col["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else "?"
        cls = next((name for name, pat in CLASSES if re.match(pat, op)), "other (" + op.split(".")[0] + ")")
        ex = int(float(r[col["Instructions Executed"]] or 0))
        sm = int(float(r[col["# Samples"]] or 0))
        m = mix[cls]
        m[0] += 1; m[1] += ex; m[2] += sm
        for s in stall_cols:
            v = int(float(r[col[s]] or 0))
            if v:
                m[3][s] += v
        total_exec += ex; total_samples += sm
        top.append((sm, src[:70], {s: int(float(r[col[s]] or 0)) for s in stall_cols if float(r[col[s]] or 0) > 0}))
    print("# %s" % kernel)
    print("# %d SASS instructions, %d warp-level instructions executed, %d stall samples" % (n_sass, total_exec, total_samples))
    print("%-52s %8s %14s %7s %9s %7s   %s" % ("class", "static", "executed", "share", "samples", "share", "what the warps that sat on these instructions were waiting for"))
    for cls, (st, ex, sm, stalls) in sorted(mix.items(), key=lambda kv: -kv[1][2]):
        why = ", ".join("%s %.0f %%" % (k.replace("stall_", ""), 100.0 * v / max(1, sm)) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        print("%-52s %8d %14d %6.1f%% %9d %6.1f%%   %s" % (cls, st, ex, 100.0 * ex / max(1, total_exec), sm, 100.0 * sm / max(1, total_samples), why))
    allst = defaultdict(int)
    for _, (_, _, _, stalls) in mix.items():
        for k, v in stalls.items():
            allst[k] += v
    print("# all samples by reason: " + ", ".join("%s %.1f %%" % (k.replace("stall_", ""), 100.0 * v / max(1, total_samples)) for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
    print("# the ten instructions with the most samples:")
    for sm, src, st in sorted(top, key=lambda t: -t[0])[:10]:
        print("#   %6d (%4.1f %%)  %-70s %s" % (sm, 100.0 * sm / max(1, total_samples), src, ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])))


if __name__ == "__main__":
    main()
