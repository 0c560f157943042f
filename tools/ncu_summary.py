#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the two text summaries kept under profiles/:
  <out>_launches.csv : one row per profiled launch (duration, DRAM bytes, issue/occupancy figures)
  <out>_details.txt  : the Speed-of-Light / scheduler / occupancy sections of every launch
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1_x
"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
       "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
SECTIONS = ("GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis", "Scheduler Statistics",
            "Warp State Statistics", "Occupancy", "Launch Statistics", "Instruction Statistics")


def run(args):
    return subprocess.run(["ncu"] + args, check=True, capture_output=True, text=True).stdout


def kname(full):
    """'void h2j::fdct_quant_kernel<(bool)0>(const unsigned char *, ...)' -> 'fdct_quant_kernel'"""
    n = full.split("(")[0].split("<")[0].strip()
    if n.startswith("void "):
        n = n[5:]
    return n.split("::")[-1]


def main(rep, out):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units, body = rows[0], rows[1], rows[2:]
    cols = [c for c in RAW if c in head]
    with open(out + "_launches.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [f"{c} [{units[head.index(c)]}]" for c in cols])
        for r in body:
            w.writerow([r[head.index("ID")], kname(r[head.index("Kernel Name")])] + [r[head.index(c)] for c in cols])
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "details", "--csv"]))))
    head = rows[0]
    ki, idi, si, mi, ui, vi = (head.index(n) for n in ("Kernel Name", "ID", "Section Name", "Metric Name", "Metric Unit", "Metric Value"))
    with open(out + "_details.txt", "w") as f:
        last = None
        for r in rows[1:]:
            if r[si] not in SECTIONS or not r[mi]:
                continue
            key = (r[idi], kname(r[ki]))
            if key != last:
                f.write(f"\n==== launch {key[0]}: {key[1]}\n")
                last = key
            f.write(f"{r[si][:28]:28s} | {r[mi]:48s} | {r[vi]} {r[ui]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
