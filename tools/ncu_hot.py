#!/usr/bin/env python
"""Top stall sites of one kernel in an ncu report (SASS page): python tools/ncu_hot.py x.ncu-rep [n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
iS, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in data)
print(rows[0][1][:100], "| samples", tot, "| SASS instructions", len(data), "| warp instr executed", sum(int(r[iex] or 0) for r in data))
agg = {}
for r in data:
    for i in stall:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:n]:
    st = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:2]
    print(f"{r[0][-5:]} {int(r[iS] or 0):5d} x{r[iex]:>8s}  {r[isrc][:72]:72s} {st}")
