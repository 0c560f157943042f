#!/usr/bin/env python
"""Hot SASS instructions of one kernel from an .ncu-rep (executed count and stall samples per instruction).
Usage: python tools/ncu_sass_hot.py <rep> <kernel-regex> [min_share_of_executed=0.004] [lo hi]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "-k", "regex:" + k],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if "# Samples" in r)
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body); ts = sum(int(r[si]) for r in body)
print(len(body), "instructions,", tot, "executed,", ts, "samples")
if len(sys.argv) > 5:
    lo, hi = int(sys.argv[4]), int(sys.argv[5])
    for i in range(lo, min(hi, len(body))):
        print(i, body[i][1][:80].ljust(80), body[i][si], body[i][ie])
else:
    for i, r in enumerate(body):
        if int(r[ie]) > tot * share or int(r[si]) > ts * share * 2:
            st = sorted([(h.replace("stall_", ""), int(v)) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not" not in h and v.isdigit() and int(v)], key=lambda x: -x[1])
            print(i, r[1][:70].ljust(70), r[si], r[ie], st[:2])
