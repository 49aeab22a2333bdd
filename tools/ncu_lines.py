#!/usr/bin/env python
"""Top source lines of an .ncu-rep by executed warp instructions (and stall samples).
    python tools/ncu_lines.py rep.ncu-rep [n] [per_unit_divisor]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
div = float(sys.argv[3]) if len(sys.argv) > 3 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, h2, agg = None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        h2 = r; continue
    if h2 is None or len(r) < 10:
        continue
    if r[2] == "-" and r[0].isdigit():
        d = dict(zip(h2[4:], r[4:]))
        try:
            agg.append((int(d["Instructions Executed"]), int(d["# Samples"]), cur, int(r[0]), r[1].strip()[:100]))
        except (KeyError, ValueError):
            pass
ti = sum(a[0] for a in agg) or 1
ts = sum(a[1] for a in agg) or 1
print(f"total warp instructions {ti:.4g}" + (f" = {ti / div:.1f} per unit" if div else ""))
for a in sorted(agg, reverse=True)[:n]:
    per = f" {a[0] / div:6.2f}/unit" if div else ""
    print(f"{100 * a[0] / ti:5.1f}% inst{per} {100 * a[1] / ts:5.1f}% smp  {a[2]}:{a[3]:<4d} {a[4]}")
