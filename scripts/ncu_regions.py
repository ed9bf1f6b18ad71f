#!/usr/bin/env python
"""Executed warp instructions and stall samples of the first kernel in an .ncu-rep, grouped by SASS address range
(2 KiB chunks) -- shows which part of a long kernel the time goes to.  usage: ncu_regions.py rep [chunk_hex]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
chunk_sz = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0x800
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        h, start = r, i
        break
ie, ss = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[start + 1:]:
    if len(r) < len(h):
        continue
    try:
        data.append((int(r[0], 16), r[1], int(r[ie]), int(r[ss])))
    except ValueError:
        pass
base = data[0][0]
ti, ts = sum(d[2] for d in data), sum(d[3] for d in data)
print("instructions %d  samples %d" % (ti, ts))
chunks = collections.OrderedDict()
for a, s, i, smp in data:
    c = chunks.setdefault((a - base) // chunk_sz, [0, 0])
    c[0] += i
    c[1] += smp
for k, (i, smp) in chunks.items():
    if i / ti > 0.01 or smp / ts > 0.01:
        print("%#8x  instr %5.1f%%  samples %5.1f%%" % (k * chunk_sz, 100 * i / ti, 100 * smp / ts))
if len(sys.argv) > 3:   # dump one range: lo hi
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    for a, s, i, smp in data:
        if lo <= a - base < hi:
            print("%#7x %9d %5d  %s" % (a - base, i, smp, s[:80]))
