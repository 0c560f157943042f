#!/usr/bin/env python
"""Static SASS instruction count per source line (and per region) of one kernel, no GPU needed.
Usage: python tools/sass_by_line.py <lib.so> <kernel-substring> [file:lo-hi=name ...]"""
import subprocess, sys, tempfile, os, re, glob, collections
so, kern = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    loc, name = a.split("="); f, r = loc.split(":"); lo, hi = r.split("-"); regions.append((f, int(lo), int(hi), name))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
txt = ""
for c in glob.glob(os.path.join(d, "*.cubin")):
    txt += subprocess.run(["nvdisasm", "--print-line-info", "-c", c], capture_output=True, text=True).stdout
on = False; cur = None
per_line = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for ln in txt.splitlines():
    if ln.startswith("//----") and ".text." in ln:
        on = kern in ln; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        per_line[cur] += 1; ops[cur][m.group(2).split(".")[0]] += 1
tot = sum(per_line.values())
print("total static instructions", tot)
if regions:
    agg = collections.Counter(); aops = collections.defaultdict(collections.Counter)
    for (f, l), n in per_line.items():
        name = "other"
        for rf, lo, hi, rn in regions:
            if f == rf and lo <= l <= hi: name = rn; break
        agg[name] += n; aops[name].update(ops[(f, l)])
    for name, n in agg.most_common():
        print(f"{name:24s} {n:5d}  " + " ".join(f"{k}:{v}" for k, v in aops[name].most_common(8)))
else:
    for (f, l), n in sorted(per_line.items(), key=lambda x: -x[1])[:40]:
        print(f"{n:5d} {f}:{l}  " + " ".join(f"{k}:{v}" for k, v in ops[(f, l)].most_common(6)))
