#!/usr/bin/env python3
"""Executed warp-instructions and stall samples per SOURCE LINE of one kernel: the SASS page of an .ncu-rep joined (by instruction
address) with the line table `nvdisasm -g` prints for the cubin inside liborb_b200.so.  The .so must be the build the report was taken from.
Usage: python tools/ncu_lines.py report.ncu-rep kernel_regex mangled_name_substring [min_percent]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, rx, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
H = rows[hdr[0]]
end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
data = [r for r in rows[hdr[0] + 1:end] if len(r) == len(H)]
iex, ismp = H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "orbslam_mapsave_b200", "liborb_b200.so")], cwd=d, capture_output=True)
    cub = [c for c in glob.glob(os.path.join(d, "orb_extract.*cubin")) + glob.glob(os.path.join(d, "*.cubin"))][0]
    for c in glob.glob(os.path.join(d, "*.cubin")):
        if mangled.split("_")[0] in open(c, "rb").read().decode("latin1"):
            pass
    dis = ""
    for c in sorted(glob.glob(os.path.join(d, "*.cubin")), key=len):
        t = subprocess.run(["nvdisasm", "-g", c], capture_output=True, text=True).stdout
        if f".text.{mangled}" in t or mangled in t:
            dis = t
            break
sec, line, lines = False, None, []
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        sec = mangled in l
        continue
    if not sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(line)
assert len(lines) >= len(data), (len(lines), len(data))
tot = sum(int(r[iex]) for r in data) or 1
ts = sum(int(r[ismp]) for r in data) or 1
agg = {}
for r, ln in zip(data, lines):
    a = agg.setdefault(ln, [0, 0, 0])
    a[0] += int(r[iex]); a[1] += int(r[ismp]); a[2] += 1
src = {}
print(f"{len(data)} SASS instructions, {tot} warp-instructions, {ts} stall samples")
for ln, (a, s, n) in sorted(agg.items(), key=lambda kv: (kv[0][0] != "orb_extract.cu", kv[0][1])):
    if a / tot * 100 < minpct and s / ts * 100 < minpct:
        continue
    f = os.path.join(root, "orbslam_mapsave_b200", "csrc", ln[0])
    if ln[0] not in src:
        src[ln[0]] = open(f).read().splitlines() if os.path.exists(f) else []
    text = src[ln[0]][ln[1] - 1].strip()[:110] if ln[1] - 1 < len(src[ln[0]]) else ""
    print(f"{ln[0]}:{ln[1]:5d}  inst {a / tot * 100:5.1f}%  stall {s / ts * 100:5.1f}%  sass {n:3d}   {text}")
