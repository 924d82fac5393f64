#!/usr/bin/env python3
"""Executed warp-instructions and stall samples per CUDA SOURCE LINE of one kernel, straight from an .ncu-rep captured with
--import-source on (ncu --page source --print-source cuda,sass).  Usage: python tools/ncu_srclines.py report.ncu-rep kernel_regex [min_percent]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
H = rows[hdr[0]]
end = hdr[1] - 2 if len(hdr) > 1 else len(rows)             # first launch only
iex, ismp = H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows[hdr[0] + 1:end]:
    if len(r) == len(H) and r[0].isdigit():
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[iex]), int(r[ismp])))
        except ValueError:
            pass
tot = sum(l[2] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
print(f"{rows[1][1][:80] if len(rows) > 1 else ''}: {tot / 1e6:.1f} M warp-instructions, {ts} stall samples")
for ln, src, ex, sm in lines:
    if ex / tot * 100 >= minpct or sm / ts * 100 >= minpct:
        print(f"{ln:5d}  inst {ex / tot * 100:5.1f}%  stall {sm / ts * 100:5.1f}%   {src[:150]}")
