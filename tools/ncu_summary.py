#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small table: one row per profiled launch.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.md]"""
import csv
import subprocess
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("warps_act_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("regs", "launch__registers_per_thread"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("inst_M", "smsp__inst_executed.sum"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    print("| kernel | grid | " + " | ".join(n for n, _ in WANT) + " |")
    print("|---|---|" + "---|" * len(WANT))
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else ""
        cells = []
        for n, m in WANT:
            if m not in idx:
                cells.append("-")
                continue
            v, u = r[idx[m]].replace(",", ""), U[idx[m]]
            try:
                x = float(v)
            except ValueError:
                cells.append(v)
                continue
            if n == "time_us":
                x = x * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            if n.endswith("_MB"):
                x = x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            if n == "inst_M":
                x = x / 1e6
            cells.append(f"{x:.1f}" if abs(x) < 1e5 else f"{x:.3g}")
        print(f"| {name} | {grid} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
