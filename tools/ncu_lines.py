#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) of one kernel from an .ncu-rep.
Usage: python tools/ncu_lines.py <rep> <kernel-regex> [top_n]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + k],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fpath = None
lines = []
tot_inst = tot_samp = 0
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    inst = num(r[hdr.index("Instructions Executed")])
    samp = num(r[hdr.index("# Samples")])
    lines.append((inst, samp, fpath, ln, r[1].strip()[:110]))
    tot_inst += inst; tot_samp += samp
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
for inst, samp, f, ln, src in sorted(lines, key=lambda x: -x[1])[:top]:
    print(f"{100*inst/tot_inst:5.1f}% inst {100*samp/max(1,tot_samp):5.1f}% samp  {f}:{ln}  {src}")
