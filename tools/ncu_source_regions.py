#!/usr/bin/env python3
"""Aggregate the ncu source page (SASS view) of one kernel into regions: share of executed instructions and of stall samples.
Usage: python tools/ncu_source_regions.py report.ncu-rep kernel_regex [region_size]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may be concatenated; take the first block
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
H = rows[hdr[0]]
end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
data = [r for r in rows[hdr[0] + 1:end] if len(r) == len(H)]
isrc, iex, ismp = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[iex]) for r in data) or 1
ts = sum(int(r[ismp]) for r in data) or 1
print(f"kernel {rows[0][1] if rows[0] else ''}: {len(data)} SASS instructions, {tot} warp-instructions executed, {ts} stall samples")
for i0 in range(0, len(data), step):
    blk = data[i0:i0 + step]
    a = sum(int(r[iex]) for r in blk)
    sm = sum(int(r[ismp]) for r in blk)
    ops = {}
    for r in blk:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[iex])
    top = ", ".join(f"{k}:{v * 100 // max(a, 1)}%" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
    print(f"  [{i0:4d}-{i0 + len(blk) - 1:4d}] inst {a / tot * 100:5.1f}%  stall-samples {sm / ts * 100:5.1f}%   {top}")
