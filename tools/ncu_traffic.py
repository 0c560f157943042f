#!/usr/bin/env python
"""DRAM traffic per frame of every kernel, from the launch table tools/ncu_summary.py wrote (one `ncu --set full`
capture of bench.py at --frames N): (dram__bytes_read.sum + dram__bytes_write.sum) / N  ->  profiles/traffic.json,
which bench.py scales by its sub-batch to fill roofline.traffic.
Usage: python tools/ncu_traffic.py profiles/<name>_launches.csv <frames in the profiled launch> [out.json]"""
import csv, json, sys

src, frames = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/traffic.json"
rows = list(csv.DictReader(open(src)))
res = {"source": src, "frames_per_launch": frames, "kernels": {}}
for r in rows:
    rd = float(r["dram__bytes_read.sum [Mbyte]"]) * 1e6
    wr = float(r["dram__bytes_write.sum [Mbyte]"]) * 1e6
    res["kernels"][r["kernel"]] = {"dram_read_bytes_per_frame": rd / frames, "dram_write_bytes_per_frame": wr / frames,
                                   "dram_bytes_per_frame": (rd + wr) / frames, "duration_us_under_ncu": float(r["gpu__time_duration.sum [us]"])}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res["kernels"], indent=1))
