#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + per-bin instruction mix of the first kernel.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep <warp_bins> [out.md]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
warp_bins = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
r = rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
out = ["kernel: %s" % r[hdr.index("Kernel Name")], ""]
for k in keys:
    if k in hdr:
        out.append("%-70s %s %s" % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(r[hdr.index(k)])
        if v > 0.05:
            out.append("%-70s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "stall/"), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = None
cnt = collections.Counter()
tot = 0
nk = 0
for row in srows:
    if row and row[0] == "Kernel Name":
        nk += 1
        continue
    if row and row[0] == "Address":
        h = row
        continue
    if h is None or nk != 1 or len(row) < len(h):
        continue
    try:
        c = int(row[h.index("Instructions Executed")])
    except ValueError:
        continue
    s = row[1].strip()
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    cnt[op.split(".")[0]] += c
    tot += c
out.append("")
out.append("warp instructions per warp-bin (total %.1f):" % (tot / warp_bins))
for k, v in cnt.most_common(24):
    out.append("  %-12s %.2f" % (k, v / warp_bins))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt + "\n")
