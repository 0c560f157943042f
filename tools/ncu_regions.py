#!/usr/bin/env python
"""Instruction / stall-sample share of source-line regions of one kernel from an .ncu-rep.
Usage: python tools/ncu_regions.py <rep> <kernel-regex> file:lo-hi=name [file:lo-hi=name ...]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    loc, name = a.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    regions.append((f, int(lo), int(hi), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + k],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fpath = None; hdr = None
tot = {}; ti = ts = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    def num(x):
        try: return int(x)
        except ValueError: return 0
    inst = num(r[hdr.index("Instructions Executed")]); samp = num(r[hdr.index("# Samples")])
    name = "other"
    for f, lo, hi, n in regions:
        if fpath == f and lo <= ln <= hi: name = n; break
    a = tot.setdefault(name, [0, 0]); a[0] += inst; a[1] += samp
    ti += inst; ts += samp
print(f"total warp instructions {ti}, samples {ts}")
for n, (i, s) in sorted(tot.items(), key=lambda x: -x[1][0]):
    print(f"{n:24s} {100*i/ti:5.1f}% inst  {100*s/max(1,ts):5.1f}% samples")
