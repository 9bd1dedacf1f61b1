#!/usr/bin/env python
"""python profiles/hotspots.py report.ncu-rep [kernel-id] [top] -- SASS lines with the most warp-stall samples."""
import csv
import subprocess
import sys

rep, kid, top = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "0"), int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{int(kid) + 1}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[0][1] if rows and len(rows[0]) > 1 else "")
H, D = rows[h], [r for r in rows[h + 1:] if len(r) > 5]
si, ii = H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = sum(num(d[si]) for d in D)
print("total samples", tot, "SASS instructions", len(D), "executed warp-insts", sum(num(d[ii]) for d in D))
best = sorted(range(len(D)), key=lambda i: -num(D[i][si]))[:top]
for i in sorted(best):
    print(f"{i:5d} {D[i][1][:80]:80s} {100.0 * num(D[i][si]) / max(tot, 1):5.1f}%  exec={D[i][ii]}")
