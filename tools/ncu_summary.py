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


def source_sha():
    """SHA-256 over the CUDA sources the constants were measured on (bench.py recomputes it and flags stale constants)."""
    import glob
    import hashlib
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbslam_mapsave_b200", "csrc")
    h = hashlib.sha256()
    # (orb_probe.cu holds the copy / POPC probes of bench.py, measurement aids with no kernel of the product path: not part of the hash)
    for f in sorted(glob.glob(os.path.join(root, "*.cu")) + glob.glob(os.path.join(root, "*.cuh")) + glob.glob(os.path.join(root, "*.inc"))):
        if os.path.basename(f) == "orb_probe.cu":
            continue
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def emit_json(rep, out_path, units, label, command):
    """profiles/ncu_constants.json: per kernel of the capture the per-launch DRAM bytes, executed warp instructions and pipe shares,
    with the number of work units (frames / descriptor pairs) one launch processed and the SHA of the sources.  Several captures
    (extractor pass, matching) are merged under their labels."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}

    def val(r, m, scale=None):
        if m not in idx:
            return None
        try:
            x = float(r[idx[m]].replace(",", ""))
        except ValueError:
            return None
        u = U[idx[m]]
        if scale == "bytes":
            x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if scale == "us":
            x *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        return x
    kernels = []
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
        kernels.append({"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", ""), "grid": r[idx["Grid Size"]],
                        "time_us_under_ncu": val(r, "gpu__time_duration.sum", "us"),
                        "dram_bytes": (rd or 0) + (wr or 0), "warp_inst": val(r, "smsp__inst_executed.sum"),
                        "issue_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "alu_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                        "xu_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                        "alu_inst": val(r, "sm__inst_executed_pipe_alu.sum"), "xu_inst": val(r, "sm__inst_executed_pipe_xu.sum"),
                        "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
    doc = {}
    if os.path.exists(out_path):
        try:
            doc = json.load(open(out_path))
        except ValueError:
            doc = {}
    doc["source_sha256"] = source_sha()
    doc.setdefault("captures", {})[label] = {"report": os.path.basename(rep), "command": command, "units_per_launch": units, "kernels": kernels}
    json.dump(doc, open(out_path, "w"), indent=1)
    print(f"{out_path}: {len(kernels)} kernels under '{label}', sha {doc['source_sha256'][:12]}")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        # ncu_summary.py --json out.json report.ncu-rep label units_per_launch "command line of the capture"
        emit_json(sys.argv[3], sys.argv[2], float(sys.argv[5]), sys.argv[4], sys.argv[6] if len(sys.argv) > 6 else "")
        return
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
