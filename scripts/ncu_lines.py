#!/usr/bin/env python
"""Executed warp instructions / stall samples per CUDA source line of the first kernel in an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_lines.py rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fpath, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= ie or r[0] == "" or r[0] == "Function Name":
        continue
    try:
        out.append((int(r[ie]), int(r[ss]), fpath, int(r[0]), r[1].strip()))
    except ValueError:
        pass
ti, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print("instructions %d  samples %d" % (ti, ts))
byfile = {}
for o in out:
    byfile[o[2]] = byfile.get(o[2], 0) + o[0]
print("by file:", {k: "%.1f%%" % (100.0 * v / ti) for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])})
for i, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% instr %5.1f%% smp  %s:%d  %s" % (100.0 * i / ti, 100.0 * s / ts, f, ln, src[:110]))
