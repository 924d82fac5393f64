#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 ORB front-end (BASELINE.json configs[1]).

A "step" = ORB extraction (pyramid -> FAST cells -> octree -> blur -> orientation + rBRIEF) of one batch of
`--frames` (default 4096) synthetic 640x480 frames per GPU, nFeatures=1000, 8 levels, scale 1.2, FAST 20/7.
  value      frames/s over all GPUs with the frames already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e        the same through the C-ABI call with HOST buffers (pinned): H2D of the frames and D2H of keypoints +
             descriptors inside the timed region
  roofline   dominant kernel (by measured share of the step) against the measured HBM copy peak
  cpu_baseline  the CPU oracle (port of the reference path) on the host cores, bounded sample, rank 0, N=1
  matching   secondary metric of BASELINE.json: Hamming top-2 + ratio pairs/s (all-pairs keyframe matching, sharded by
             query keyframe, NCCL all-gather of the per-rank match tables when N>1) against the measured POPC peak

`--impl reference` times the reference's own CPU implementation of the path — oracle/_ref: src/ORBextractor.cc compiled unmodified
over a stand-in for the OpenCV API slice it uses; the oracle port when oracle/_ref has not been built — on all host
threads, on a bounded sample of the same workload per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, NLEVELS, SCALE, INI_TH, MIN_TH = 640, 480, 1000, 8, 1.2, 20, 7
B_FRAME = 1961064          # algorithmic bytes per frame, SURVEY.md §8(d)
METRIC = "ORB frames/s (640x480, 1000 feat, 8 lvl)"


def _gen_one(seed):
    from orbslam_mapsave_b200.synth import synth
    return synth(W, H, seed)


def make_frames(n, seed0, unique):
    """n frames; `unique` distinct synth() seeds (seed0...), repeated cyclically.  Cached under /tmp."""
    unique = min(unique, n)
    cache = f"/tmp/orb_synth_{W}x{H}_s{seed0}_u{unique}.npy"
    if os.path.exists(cache):
        base = np.load(cache)
    else:
        import multiprocessing as mp
        nproc = max(1, min(32, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        with mp.get_context("fork").Pool(nproc) as pool:
            base = np.stack(pool.map(_gen_one, range(seed0, seed0 + unique), chunksize=8))
        try:
            np.save(cache + f".{os.getpid()}.tmp.npy", base)
            os.replace(cache + f".{os.getpid()}.tmp.npy", cache)
        except OSError:
            pass
    reps = (n + unique - 1) // unique
    return np.concatenate([base] * reps)[:n] if reps > 1 else base[:n]


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(torch, local_rank):
    """Best effort: run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers (first touch)
    are local to the GPU's PCIe root port.  Matters only when several ranks stream frames from host memory at once."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except (OSError, ValueError, AttributeError):
        pass
    return None


def load_ncu_constants():
    """profiles/ncu_constants.json (tools/ncu_summary.py --json): per-launch DRAM bytes / executed warp instructions of the kernels, taken
    from a committed ncu capture, with the SHA-256 of the CUDA sources they were measured on.  Returns (captures or None, info): the
    constants are dropped (info["stale"] = True) when the sources have changed since."""
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    try:
        doc = json.load(open(path))
    except (OSError, ValueError):
        return None, {"stale": True, "reason": "profiles/ncu_constants.json missing"}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from ncu_summary import source_sha
    sha = source_sha()
    if sha != doc.get("source_sha256"):
        return None, {"stale": True, "reason": "csrc changed since the capture", "captured_sha256": doc.get("source_sha256"), "current_sha256": sha}
    return doc["captures"], {"stale": False, "source_sha256": sha}


def ncu_kernel(captures, label, name):
    """(kernel record, units per launch) of the first launch whose name starts with `name` in capture `label`."""
    if not captures or label not in captures:
        return None, None
    for k in captures[label]["kernels"]:
        if k["kernel"].startswith(name):
            return k, captures[label]["units_per_launch"]
    return None, None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _ref_available():
    try:
        from oracle import ref_py
        return ref_py.available()
    except Exception:
        return False


def ref_extract_mt(frames, threads):
    """The reference's own ORBextractor (oracle/_ref: src/ORBextractor.cc compiled unmodified over the OpenCV stand-in) on `threads`
    host threads, one extractor instance per thread, frames dealt round-robin; ctypes releases the GIL in the call.  Returns seconds."""
    from oracle import ref_py
    exs = [ref_py.RefExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH) for _ in range(threads)]

    def work(t):
        for i in range(t, len(frames), threads):
            exs[t].extract(frames[i])

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return time.perf_counter() - t0


REF_NOTE = ("the reference's own src/ORBextractor.cc compiled unmodified (oracle/_ref); the image has no OpenCV C++, so cv::FAST / resize / "
            "GaussianBlur / copyMakeBorder / fastAtan2 behind it are the oracle's scalar restatements (bit-exact against cv2 4.13), slower than "
            "an optimised OpenCV build")


def cpu_extract_sample(frames, threads, budget_s):
    """CPU extraction on `threads` host threads over a bounded sample; returns (frames/s, n_frames, seconds, kind)."""
    use_ref = _ref_available()
    chunk = max(threads * 16, 64) if use_ref else max(threads * 4, 8)      # (python threads: amortise their start-up)
    done, secs = 0, 0.0
    from oracle import orb_oracle_py as orc
    while secs < budget_s and done + chunk <= len(frames):
        if use_ref:
            s = ref_extract_mt(frames[done:done + chunk], threads)
        else:
            s, _ = orc.extract_batch_mt(frames[done:done + chunk], NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, threads)
        secs += s
        done += chunk
    return done / secs, done, secs, ("reference" if use_ref else "port")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    frames = make_frames(max(threads * 8, 64), 0, 4096)
    per_step = len(frames)
    use_ref = _ref_available()
    from oracle import orb_oracle_py as orc

    def one(fr):
        if use_ref:
            return ref_extract_mt(fr, threads)
        return orc.extract_batch_mt(fr, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, threads)[0]

    for _ in range(args.warmup):
        one(frames[:threads])
    total_s = 0.0
    for _ in range(args.steps):
        total_s += one(frames)
    v = per_step * args.steps / total_s
    sample = f"{per_step} of the 4096 frames per step, {threads} host threads, one extractor instance per thread"
    _emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: batch of 640x480 synthetic frames, nFeatures=1000, 8 levels, scale 1.2, FAST 20/7",
                   "frames_per_step": per_step,
                   "note": REF_NOTE if use_ref else "CPU oracle port of src/ORBextractor.cc (oracle/_ref has not been built)"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def _emit(obj):
    """Print the ONE JSON line on the real stdout (fd 1 is pointed at stderr while the bench runs so that library
    chatter such as 'NCCL version ...' cannot pollute it)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--unique", type=int, default=4096, help="distinct synthetic frames per GPU (others repeat them)")
    ap.add_argument("--chunk", type=int, default=1024, help="frames per device pass of the device-resident measurement (workspace size)")
    ap.add_argument("--e2e-chunk", type=int, default=128, help="frames per pipelined chunk of the host-buffer (e2e) path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-match", action="store_true", help="skip the matching sub-benchmark")
    ap.add_argument("--match-q", type=int, default=32, help="query keyframes per GPU in the matching sub-benchmark")
    ap.add_argument("--match-db", type=int, default=512, help="database keyframes in the matching sub-benchmark")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C5 blocks (`configs` key; N=1 only)")
    ap.add_argument("--only-config", default=None, choices=["C1", "C3", "C5"],
                    help="measure just this BASELINE configuration and print its block (profiling aid: ncu runs, A/B)")
    ap.add_argument("--c3-frames", type=int, default=256)
    ap.add_argument("--c5-frames", type=int, default=32)
    ap.add_argument("--c5-batch", type=int, default=16, help="4K frames per device pass")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # synthetic frames first (fork pool) — before CUDA is touched
    import bench_configs as bc
    want_cfg = (world == 1 and not args.no_configs) or args.only_config
    c3_frames = c5_frames = None
    if want_cfg and args.only_config in (None, "C3"):
        c3_frames = bc.moving_window_frames(1280, 720, args.c3_frames, 16, 100000)
    if want_cfg and args.only_config in (None, "C5"):
        c5_frames = bc.make_frames(3840, 2160, args.c5_frames, 8, 200000)
    if args.only_config:
        args.frames = 64
    frames = make_frames(args.frames, rank * args.frames, args.unique)

    import torch
    import torch.distributed as dist
    import orbslam_mapsave_b200 as orb
    from orbslam_mapsave_b200 import capi
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure_configs(stream, popc):
        """BASELINE configs C1 / C3 / C5 (bench_configs.py), N=1."""
        peaks_ = {}
        try:
            peaks_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm = float(peaks_.get("hbm_gbs", 6650.0))
        out = {"hbm_peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks_ else "fallback (B200_PROFILING.md)"}
        if args.only_config in (None, "C1"):
            ref_ms = None if args.no_cpu else bc.reference_ms_per_frame(frames[0])
            out["C1"] = bc.c1_latency(torch, orb, capi, local_rank, frames[0], ref_ms=ref_ms)
        if c3_frames is not None:
            blk, ctx = bc.batch_config("C3", torch, orb, capi, local_rank, hbm, c3_frames, len(c3_frames), 128, min(args.steps, 5), stream)
            blk["matching"] = bc.consecutive_keyframe_matching(torch, capi, ctx, min(args.steps, 5), stream, popc)
            out["C3"] = blk
            del ctx
        if c5_frames is not None:
            caps, info = load_ncu_constants()
            c5ncu = dict(info)
            if caps and "c5_pass" in caps:                             # per-frame DRAM traffic / warp instructions of the 4K pass's kernels
                u = caps["c5_pass"]["units_per_launch"]
                agg = {}
                for k in caps["c5_pass"]["kernels"]:
                    a = agg.setdefault(k["kernel"].split("<")[0], {"dram_bytes_per_frame": 0.0, "warp_inst_per_frame": 0.0, "launches": 0})
                    a["dram_bytes_per_frame"] += k["dram_bytes"] / u
                    a["warp_inst_per_frame"] += (k["warp_inst"] or 0) / u
                    a["launches"] += 1
                c5ncu["kernels"] = agg
                c5ncu["dram_bytes_per_frame"] = sum(a["dram_bytes_per_frame"] for a in agg.values())
            blk, ctx = bc.batch_config("C5", torch, orb, capi, local_rank, hbm, c5_frames, 8, args.c5_batch, min(args.steps, 5), stream,
                                       ncu_constants=c5ncu)
            out["C5"] = blk
            del ctx
        torch.cuda.empty_cache()
        return out

    if args.only_config:
        tstream = torch.cuda.Stream(device=local_rank)
        torch.cuda.set_stream(tstream)
        popc, _ = orb.popc_peak(local_rank)
        _emit(measure_configs(tstream.cuda_stream, popc))
        return

    nF = args.frames
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None     # pinned staging memory local to this GPU's root port
    h_frames = torch.from_numpy(frames).pin_memory()
    d_frames = h_frames.cuda(non_blocking=True)
    ex = orb.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, max_batch=args.chunk, device=local_rank)
    cap = ex.max_keypoints()
    d_kp = torch.zeros((nF, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((nF, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nF, dtype=torch.int32, device="cuda")
    # a real (non-default) stream: the C ABI treats NULL as "the handle's own stream", and CUDA events must be recorded on
    # the stream the kernels are launched on
    tstream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step(stages=capi.STAGE_ALL):
        ex.extract_batch_device(d_frames, d_kp, d_desc, d_n, cap, stream=stream, stages=stages)

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    # ---- device-resident throughput (`value`)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    ex.check_status()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ex.launch_count()
    secs = timed(step, args.steps)
    launches = ex.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ex.check_status()
    value = world * nF * args.steps / secs
    total_kp = int(d_n.sum().item())

    # ---- per-stage device times (live, CUDA events) -> dominant kernel and its roofline
    stage_names = [("pyramid", capi.STAGE_PYRAMID), ("fast", capi.STAGE_FAST), ("octree", capi.STAGE_OCTREE),
                   ("blur", capi.STAGE_BLUR), ("describe", capi.STAGE_DESCRIBE)]
    stage_s = {}
    for nm, bit in stage_names:
        step(bit)
        stage_s[nm] = timed(lambda b=bit: step(b), max(1, min(args.steps, 3))) / max(1, min(args.steps, 3))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    kp_per_frame = total_kp / nF
    pyr_px = 950532
    # algorithmic bytes per frame of each stage (DESIGN.md "Kernels"): the reads + writes the stage cannot avoid
    cand_per_frame = 7100
    stage_bytes = {
        "pyramid": W * H + W * H + 2 * (pyr_px - W * H),               # read input, write level 0, write levels >= 1, read each source level once
        "fast": pyr_px + cand_per_frame * 8,                            # read every level once, write ~7.1k candidates (8 B)
        "octree": cand_per_frame * 8 + kp_per_frame * 8,                # read candidates, write selected keypoints
        "blur": 2 * pyr_px,                                             # read every level, write every blurred level
        "describe": kp_per_frame * (749 + 37 * 37 + 60),                # per keypoint: 749-px disc + 37x37 blurred patch + 28 B + 32 B out
    }
    dom = max(stage_s, key=stage_s.get)
    launches_per_step = launches / args.steps
    passes = (nF + args.chunk - 1) // args.chunk
    dom_launches = {"pyramid": NLEVELS, "fast": 1, "octree": 1, "blur": 1, "describe": 1}[dom] * passes
    dom_launch_s = stage_s[dom] / dom_launches
    achieved = stage_bytes[dom] * nF / stage_s[dom] / 1e9
    kname = {"pyramid": "k_resize", "fast": "k_fast_tma", "octree": "k_octree", "blur": "k_blur", "describe": "k_describe"}[dom]
    captures, ncu_info = load_ncu_constants()
    krec, kunits = ncu_kernel(captures, "c2_pass", kname)
    sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    traffic = issue = None
    if krec and dom != "pyramid":                                     # (the pyramid stage is several launches: no single-kernel constants)
        scale = args.chunk / kunits                                    # the capture's launch processed `kunits` frames
        traffic = krec["dram_bytes"] * scale
        winst = krec["warp_inst"] * scale
        # what actually bounds these kernels: warp-instruction issue, against 148 SMs x 4 schedulers x the SM clock seen in this run
        issue = {"warp_inst_per_launch": winst, "achieved_ginst_s": winst / dom_launch_s / 1e9, "peak_ginst_s": 148 * 4 * sm_hz / 1e9,
                 "frac": winst / dom_launch_s / (148 * 4 * sm_hz), "issue_pct_under_ncu": krec["issue_pct"]}
    roofline = {"bound": "hbm", "kernel": {"pyramid": "k_level0+k_resize"}.get(dom, kname),
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed capture (profiles/ncu_constants.json)
                "traffic": traffic, "ncu_constants": ncu_info,
                "peak_source": peak_src, "avg_launch_ms": dom_launch_s * 1e3, "algorithmic_bytes_per_frame": stage_bytes[dom],
                "stage_ms_per_step": {k: v * 1e3 for k, v in stage_s.items()},
                "issue": issue,
                "step_hbm_frac": (B_FRAME * nF / (secs / args.steps) / 1e9) / hbm_peak,
                "note": "640x480 pyramids are L2-resident and this stage is integer-issue bound, not HBM bound (SURVEY.md §7)"}

    # ---- end to end through the C ABI with host buffers
    h_kp = torch.zeros((nF, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.zeros((nF, cap, 32), dtype=torch.uint8).pin_memory()
    h_n = np.zeros(nF, np.int32)

    # the host-buffer path pipelines H2D / compute / D2H in chunks of its own handle's pass size (smaller chunks start earlier)
    ex2 = ex if args.e2e_chunk == args.chunk else orb.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, max_batch=args.e2e_chunk,
                                                                     device=local_rank)

    def e2e_step():
        capi.check(capi.lib().orbx_extract_batch(ex2.handle, capi._p(h_frames), nF, W, H, W, W * H, None, 0, 0,
                                                 capi._p(h_kp), capi._p(h_desc), cap, capi._p(h_n)))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * nF * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(nF * W * H),
           "d2h_bytes_per_step": int(h_n.sum()) * 60 + nF * 4, "ms_per_step": 1e3 * e2e_s / args.steps}
    # the same call from PAGEABLE caller buffers (what a caller holding cv::Mat frames hands over): the copies go through the driver's
    # staging and cannot overlap the way DMA from pinned memory does.  Reported next to the pinned figure, N=1 only.
    if world == 1:
        p_frames = frames if frames.flags.c_contiguous else np.ascontiguousarray(frames)
        p_kp = np.zeros((nF, cap, 7), np.float32)
        p_desc = np.zeros((nF, cap, 32), np.uint8)
        p_n = np.zeros(nF, np.int32)

        def pageable_step():
            capi.check(capi.lib().orbx_extract_batch(ex2.handle, capi._p(p_frames), nF, W, H, W, W * H, None, 0, 0,
                                                     capi._p(p_kp), capi._p(p_desc), cap, capi._p(p_n)))
        pageable_step()
        t0 = time.perf_counter()
        kp_ = max(2, min(args.steps, 5))
        for _ in range(kp_):
            pageable_step()
        ps = (time.perf_counter() - t0) / kp_
        e2e["pageable_caller_buffers"] = {"value": nF / ps, "unit": "frames/s", "ms_per_step": 1e3 * ps,
                                          "note": "orbx_extract_batch from ordinary (unpinned) host arrays: the handle's host-thread pool stages "
                                                  "them through pinned memory (non-temporal copies); 28 k frames/s before the pool existed"}
        # the same arrays after the caller registered them once (orbx_host_register = cudaHostRegister)
        t0 = time.perf_counter()
        for a_ in (p_frames, p_kp, p_desc):
            capi.check(capi.lib().orbx_host_register(capi._p(a_), a_.nbytes))
        reg_s = time.perf_counter() - t0
        pageable_step()
        t0 = time.perf_counter()
        for _ in range(kp_):
            pageable_step()
        rs = (time.perf_counter() - t0) / kp_
        for a_ in (p_frames, p_kp, p_desc):
            capi.check(capi.lib().orbx_host_unregister(capi._p(a_)))
        e2e["registered_caller_buffers"] = {"value": nF / rs, "unit": "frames/s", "ms_per_step": 1e3 * rs, "register_once_ms": 1e3 * reg_s,
                                            "note": "the same arrays after orbx_host_register (one-off page-locking by the caller)"}
        del p_kp, p_desc
    # the floor the platform sets for e2e: the same bytes moved with no compute at all — orb_h2d_probe (csrc/orb_probe.cu): one thread
    # per GPU, plain cudaMemcpyAsync per chunk from pinned memory, all GPUs of this run streaming at once, 20 % of the volume coming
    # back device->host at the same time (keypoints + descriptors are ~20 % of the frame bytes).  Rank 0 runs it for all N GPUs while
    # the other ranks wait on a CPU barrier.
    import ctypes as C
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    def cpu_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)
    cpu_barrier()
    if rank == 0:
        devs = np.arange(world, dtype=np.int32)
        probe = {}
        for label, flags in (("pinned", 2), ("pinned_numa_bound", 2 | 4), ("write_combined", 2 | 1)):
            each, tot, nodes = np.zeros(world, np.float64), C.c_double(), np.zeros(world, np.int32)
            rc = capi.lib().orb_h2d_probe(world, capi._p(devs), nF * W * H, args.e2e_chunk * W * H, 3, flags, capi._p(each), C.byref(tot), capi._p(nodes))
            if rc != 0:
                probe[label] = {"error": capi.lib().orb_last_error().decode()}
                continue
            probe[label] = {"h2d_gbs_total": tot.value, "h2d_gbs_per_gpu_min": float(each.min()), "ms_per_step": 1e3 * nF * W * H / (float(each.min()) * 1e9)}
            probe["gpu_numa_nodes"] = nodes.tolist()
        best = max((v for k, v in probe.items() if isinstance(v, dict) and "h2d_gbs_total" in v), key=lambda v: v["h2d_gbs_total"], default=None)
        e2e["copy_floor_probe"] = probe
        if best:
            e2e["copy_only_ms_per_step"] = best["ms_per_step"]
            e2e["copy_only_h2d_gbs_per_gpu"] = best["h2d_gbs_per_gpu_min"]
            e2e["e2e_over_copy_floor"] = e2e["ms_per_step"] / best["ms_per_step"]
        e2e["copy_floor_note"] = ("gpu_numa_nodes: -1 = the kernel exposes no NUMA node for the GPU (the reason bind_to_gpu_numa_node returns "
                                  "None in this container), -3 = sysfs entry not visible")
    cpu_barrier()
    # the same probe the way the e2e step itself is laid out: ONE PROCESS PER GPU, every rank streaming its own device at the same time
    # (each process's pinned buffers are first-touched by that process, like the e2e buffers; the single-process form above allocates all
    # of them from rank 0).  The ranks agree on a wall-clock start (rank 0's clock + 6 s, broadcast), allocate and warm up on their own
    # and begin their timed copies together; a rank that misses the start invalidates the figure.  The slowest rank sets the time.
    if world > 1:
        t_start = torch.tensor([time.time_ns() + 6_000_000_000], dtype=torch.int64)
        dist.broadcast(t_start, src=0, group=cpu_group)
        each1, tot1, node1, late1 = np.zeros(1, np.float64), C.c_double(), np.zeros(1, np.int32), C.c_double()
        dev1 = np.array([local_rank], np.int32)
        rc = capi.lib().orb_h2d_probe_at(1, capi._p(dev1), nF * W * H, args.e2e_chunk * W * H, 6, 2, int(t_start.item()), capi._p(each1),
                                         C.byref(tot1), capi._p(node1), C.byref(late1))
        mine = torch.tensor([float(each1[0]) if rc == 0 else 0.0], dtype=torch.float64)
        lo = mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=cpu_group)
        sm_ = mine.clone()
        dist.all_reduce(sm_, op=dist.ReduceOp.SUM, group=cpu_group)
        late = torch.tensor([late1.value], dtype=torch.float64)
        dist.all_reduce(late, op=dist.ReduceOp.MAX, group=cpu_group)
        if rank == 0 and float(lo) > 0:
            ms = 1e3 * nF * W * H / (float(lo) * 1e9)
            ok_sync = float(late) == 0.0
            e2e["copy_floor_probe"]["one_process_per_gpu"] = {"h2d_gbs_total": float(sm_), "h2d_gbs_per_gpu_min": float(lo), "ms_per_step": ms,
                                                              "common_start": ok_sync, "max_late_ms": float(late)}
            if ok_sync and ms < e2e.get("copy_only_ms_per_step", 1e30):
                e2e["copy_only_ms_per_step"] = ms
                e2e["copy_only_h2d_gbs_per_gpu"] = float(lo)
                e2e["e2e_over_copy_floor"] = e2e["ms_per_step"] / ms
        cpu_barrier()

    # ---- secondary metric: Hamming matches/s (all-pairs keyframe matching, sharded by query keyframe)
    matching = None
    if not args.no_match:
        from orbslam_mapsave_b200.synth import synth_descriptors
        per = 2000
        nq, ndb = args.match_q, args.match_db
        base = synth_descriptors(per, 7)
        rng = np.random.default_rng(11)
        db = rng.integers(0, 256, (ndb, per, 32), dtype=np.uint8)
        flips = rng.integers(0, 256, (ndb, per // 4, 32), dtype=np.uint8) * (rng.random((ndb, per // 4, 32)) < 0.05)
        db[:, : per // 4] = base[None, : per // 4] ^ flips.astype(np.uint8)      # planted near-duplicates across keyframes
        d_db = torch.from_numpy(db).cuda()
        q0 = (rank * nq) % ndb
        q1 = min(q0 + nq, ndb)
        cnt = torch.zeros(((q1 - q0) * ndb + 1) // 2 * 2, dtype=torch.uint16, device="cuda")

        from orbslam_mapsave_b200.sharding import gather_match_tables

        def mstep():
            capi.check(capi.lib().orbm_allpairs_device(capi._p(d_db), ndb, per, q0, q1, 50, 0.75, capi._p(cnt), None, None, stream))
            if world > 1:       # the one exchange step of the path: assemble the match table on every rank (NCCL over NVLink)
                gather_match_tables(cnt[: (q1 - q0) * ndb].view(q1 - q0, ndb), world * (q1 - q0), rank, world)
        for _ in range(3):
            mstep()
        msteps = max(2, args.steps)
        msecs = timed(mstep, msteps)
        pairs = world * (q1 - q0) * (ndb - 1) * per * per * msteps
        popc, clk = orb.popc_peak(local_rank)
        matching = {"metric": "Hamming top-2+ratio descriptor pairs/s", "value": pairs / msecs, "unit": "pairs/s",
                    "config": {"workload": f"all-pairs: {q1 - q0} query keyframes/GPU x {ndb} keyframes x {per} descriptors",
                               "collective": "all_gather(match-count table)" if world > 1 else "none"},
                    "roofline": {"bound": "popc", "achieved": pairs / msecs * 8 / 1e12, "peak": popc / 1e12 * world, "unit": "TPOPC/s",
                                 "frac": pairs / msecs * 8 / (popc * world), "peak_source": "orbm_popc_peak microbenchmark, this run",
                                 "note": "achieved = ALGORITHMIC POPC (8 per pair, SURVEY 8d); the kernel EXECUTES 4 POPC per pair (carry-save "
                                         "compression of the XOR words, orb_match.cu), so frac > 1 is expected",
                                 "executed": {"popc_per_pair": 4, "popc_pipe_frac": pairs / msecs * 4 / (popc * world)}}}
        mrec, munits = ncu_kernel(load_ncu_constants()[0], "matching", "k_allpairs")
        if mrec:                                                       # pipe shares as ncu measured them (committed capture, same sources)
            matching["roofline"]["executed"].update({"warp_inst_per_pair": mrec["warp_inst"] / munits, "alu_pipe_pct_under_ncu": mrec["alu_pct"],
                                                     "xu_pipe_pct_under_ncu": mrec["xu_pct"], "issue_pct_under_ncu": mrec["issue_pct"]})
        if rank == 0 and not args.no_cpu:
            from oracle import orb_oracle_py as orc
            thr = host_threads()
            qs = db[0][: max(thr * 16, 64)]
            _, _, _, s = orc.hamming_top2(qs, db[1], nthreads=thr)
            matching["cpu_baseline"] = {"value": len(qs) * per / s, "unit": "pairs/s", "cores": thr, "kind": "port",
                                        "sample": f"{len(qs)} queries x {per} descriptors, reference bit-hack popcount"}

    # ---- "next" row of the scope table: BoW assignment (DBoW2 transform) feeding the matcher, ORBvoc-shaped synthetic tree
    bow = None
    if not args.no_match:
        k, L = 10, 6
        n_nodes = (k ** (L + 1) - 1) // (k - 1)
        rng = np.random.default_rng(21)
        parent = ((np.arange(n_nodes, dtype=np.int64) - 1) // k).astype(np.int32)
        parent[0] = 0
        ndesc = rng.integers(0, 256, (n_nodes, 32), dtype=np.uint8)
        is_leaf = (np.arange(n_nodes) >= (k ** L - 1) // (k - 1)).astype(np.uint8)
        weights = np.where(is_leaf == 1, rng.uniform(0.1, 9.0, n_nodes), 0.0)
        voc = orb.ORBVocabulary.from_arrays(k, L, parent, ndesc, weights, is_leaf, device=local_rank)
        nd = 1 << 21
        d_feat = torch.randint(0, 256, (nd, 32), dtype=torch.uint8, device="cuda")
        d_w = torch.zeros(nd, dtype=torch.int32, device="cuda")
        d_nid = torch.zeros(nd, dtype=torch.int32, device="cuda")
        d_wt = torch.zeros(nd, dtype=torch.float64, device="cuda")

        def bstep():
            capi.check(capi.lib().orbv_transform_device(voc.handle, capi._p(d_feat), nd, 4, capi._p(d_w), capi._p(d_wt), capi._p(d_nid), stream))
        for _ in range(3):
            bstep()
        bsteps = max(3, min(args.steps, 10))
        bsecs = timed(bstep, bsteps)
        bow = {"metric": "BoW assignment descriptors/s (k=10, L=6 vocabulary, levelsup=4)", "value": world * nd * bsteps / bsecs,
               "unit": "descriptors/s", "hamming_distances_per_s": world * nd * bsteps * k * L / bsecs,
               "config": {"workload": f"{nd} random descriptors per GPU per step through a synthetic {n_nodes}-node tree (35 MB, L2-resident)"}}

    # ---- "next" rows: the tracker's window searches and the stereo matcher, single-call latency through the C ABI from host
    # arrays (rank 0, N=1 only: they are per-frame latency paths, "replicas only" under multi-GPU)
    tracking = None
    if rank == 0 and world == 1 and not args.no_match:
        import time as _t
        rng = np.random.default_rng(5)
        n = 2000
        sc = (1.2 ** np.arange(8)).astype(np.float32)
        pz = 1.0 / sc.astype(np.float64)
        fx_, fy_, cx_, cy_ = 520.0, 520.0, 320.0, 240.0
        x = rng.uniform(0, 640, n).astype(np.float32)
        y = rng.uniform(0, 480, n).astype(np.float32)
        octv = rng.choice(8, n, p=pz / pz.sum()).astype(np.int32)
        ang = rng.uniform(0, 360, n).astype(np.float32)
        dsc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        ur = np.where(rng.random(n) < 0.7, x - 20, -1).astype(np.float32)
        grid = orb.GridView(dsc, x, y, octv, sc, (0.0, 0.0, 640.0, 480.0), angle=ang, uright=ur, blocked=np.zeros(n, np.uint8))
        Tcw = np.concatenate([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)], 1)
        z = rng.uniform(1, 8, n)
        world_pts = np.stack([(x + rng.normal(0, 2, n) - cx_) / fx_ * z, (y + rng.normal(0, 2, n) - cy_) / fy_ * z, z], 1).astype(np.float32)
        flips = (rng.random((n, 32, 8)) < 0.05)
        mdesc = dsc ^ np.packbits(flips, axis=2).reshape(n, 32)
        args_f = (Tcw, Tcw, fx_, fy_, cx_, cy_, 40.0, 40.0 / fx_, np.ones(n, np.uint8), world_pts, octv, ang, mdesc, np.ones(n, np.uint8), 7.0, False)
        m = orb.ORBmatcher(0.9, True, device=local_rank)

        def lat(f, reps=30):
            f()
            t0 = _t.perf_counter()
            for _ in range(reps):
                r = f()
            return (_t.perf_counter() - t0) / reps, r
        t_gpu, (nm, _) = lat(lambda: m.SearchByProjectionFrame(grid, *args_f))
        tracking = {"search_by_projection_last_frame": {"ms_per_call": 1e3 * t_gpu, "features": n, "matches": int(nm),
                                                        "note": "C ABI from pageable host arrays: upload + 3 kernels + download"}}
        # many-frame form: 16 independent (CurrentFrame, LastFrame) pairs (several cameras / sessions) in ONE call
        jobs16 = [(grid,) + args_f] * 16
        t_fb, rb = lat(lambda: m.SearchByProjectionFrameBatch(jobs16), 10)
        tracking["search_by_projection_frame_batch"] = {
            "frames": 16, "features": n, "ms_per_batch_call": 1e3 * t_fb, "ms_looping_single_calls": 16 * 1e3 * t_gpu,
            "frames_per_s": 16 / t_fb, "matches": int(sum(r[0] for r in rb)),
            "note": "16 frame pairs: one upload, 3 launches (one resolve CTA per pair), one download"}
        left = frames[0]
        right = np.roll(left, -12, axis=1)
        exL, exR = orb.ORBextractor(1000, 1.2, 8, 20, 7, device=local_rank), orb.ORBextractor(1000, 1.2, 8, 20, 7, device=local_rank)
        kL, dL = exL(left, download_pyramid=False)
        kR, dR = exR(right, download_pyramid=False)
        t_st, (u_r, _) = lat(lambda: exL.ComputeStereoMatches(exR, kL, dL, kR, dR, 40.0, 0.08))
        tracking["compute_stereo_matches"] = {"ms_per_call": 1e3 * t_st, "left_keypoints": int(len(kL)), "matched": int((u_r >= 0).sum()),
                                              "note": "640x480, on the device-resident pyramids of two extractor handles"}
        # batched node-constrained searches: one frame against 10 candidate keyframes (Tracking::Relocalization's loop,
        # src/Tracking.cc:1621-1643) and one keyframe against 20 neighbours (LocalMapping::CreateNewMapPoints, src/LocalMapping.cc:215-268),
        # each as ONE call against the same work as a loop of single calls
        nodes_of = lambda m: rng.integers(0, 100, m) * 7 + 3

        def mk(m, nodes=None):
            nodes = nodes_of(m) if nodes is None else nodes
            v = orb.View(rng.integers(0, 256, (m, 32), dtype=np.uint8), orb.FeatureVector(nodes), rng.uniform(0, 360, m).astype(np.float32),
                         flag=(rng.random(m) < 0.7).astype(np.uint8), x=rng.uniform(0, 640, m).astype(np.float32),
                         y=rng.uniform(0, 480, m).astype(np.float32), octave=rng.integers(0, 8, m).astype(np.int32),
                         uright=np.full(m, -1, np.float32))
            v.nodes = nodes
            return v
        frame_v = mk(n)
        cands = []
        for _ in range(20):                                            # candidates share a third of the frame's features (true matches)
            nd, src = nodes_of(n), rng.choice(n, n // 3, replace=False)
            nd[src] = frame_v.nodes[src]
            c_ = mk(n, nd)
            c_.desc[src] = frame_v.desc[src] ^ np.packbits(rng.random((len(src), 32, 8)) < 0.04, axis=2).reshape(len(src), 32)
            cands.append(c_)
        mb_ = orb.ORBmatcher(0.75, True, device=local_rank)
        t_b, (nmb, _) = lat(lambda: mb_.SearchByBoWBatch(frame_v, cands[:10], kf_kf=False), 20)
        t_l, _ = lat(lambda: [mb_.SearchByBoW(c_, frame_v) for c_ in cands[:10]], 5)
        cnt_f = np.bincount(frame_v.nodes, minlength=1024).astype(np.float64)
        pairs_bow = float(sum(np.dot(np.bincount(c_.nodes[c_.flag == 1], minlength=1024).astype(np.float64), cnt_f) for c_ in cands[:10]))
        tracking["search_by_bow_batch"] = {"candidates": 10, "features": n, "ms_per_batch_call": 1e3 * t_b, "ms_looping_single_calls": 1e3 * t_l,
                                           "searches_per_s": 10 / t_b, "descriptor_pairs_per_s": pairs_bow / t_b, "matches": int(nmb.sum()),
                                           "note": "one frame against 10 candidate keyframes, ~100 vocabulary nodes: 2 launches, 1 upload, 1 download"}
        F12s = (rng.normal(0, 1, (20, 3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-3], [1e-3, 1e-3, 1e-1]])).astype(np.float32)
        eps = rng.uniform(100, 500, (20, 2)).astype(np.float32)
        sig = (sc * sc * 5000).astype(np.float32)
        for c_ in cands + [frame_v]:
            c_.flag = (rng.random(n) < 0.3).astype(np.uint8)           # triangulation looks at features WITHOUT map points
        t_tb, (nmt, _) = lat(lambda: mb_.SearchForTriangulationBatch(frame_v, cands, F12s, eps, [sc] * 20, [sig] * 20, False), 10)
        t_tl, _ = lat(lambda: [mb_.SearchForTriangulation(frame_v, c_, F12s[i], eps[i, 0], eps[i, 1], sc, sig, False) for i, c_ in enumerate(cands)], 3)
        tracking["search_for_triangulation_batch"] = {"neighbours": 20, "features": n, "ms_per_batch_call": 1e3 * t_tb,
                                                      "ms_looping_single_calls": 1e3 * t_tl, "searches_per_s": 20 / t_tb, "matches": int(nmt.sum())}
        if not args.no_cpu:
            from oracle import orb_oracle_py as orc
            og = orc.Grid(dsc, x, y, octv, sc, (0.0, 0.0, 640.0, 480.0), angle=ang, uright=ur, blocked=np.zeros(n, np.uint8))
            t_cpu, _ = lat(lambda: orc.search_projection_frame(og, *args_f, True), 5)
            tracking["search_by_projection_last_frame"]["cpu_port_ms_per_call"] = 1e3 * t_cpu
            oL, oR = orc.Extractor(1000, 1.2, 8, 20, 7), orc.Extractor(1000, 1.2, 8, 20, 7)
            okL, odL = oL.extract(left)
            okR, odR = oR.extract(right)
            tb = oL.tables()
            lv = ([oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)])
            t_cpu, _ = lat(lambda: orc.stereo_matches(lv[0], lv[1], tb["scale"], tb["inv_scale"], okL, odL, okR, odR, 40.0, 0.08), 5)
            tracking["compute_stereo_matches"]["cpu_port_ms_per_call"] = 1e3 * t_cpu

    # ---- the other BASELINE configurations (C1 latency, C3 extract + consecutive-keyframe matching, C5 4K stress), N=1 only
    configs = None
    if rank == 0 and want_cfg:
        popc_c, _ = orb.popc_peak(local_rank)
        configs = measure_configs(stream, popc_c)

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        thr = host_threads()
        v, n_done, s, kind = cpu_extract_sample(frames, thr, 12.0)
        cpu = {"value": v, "unit": "frames/s", "cores": thr, "kind": kind,
               "sample": f"first {n_done} of the {nF} frames ({s:.1f} s), one extractor instance per thread, frames dealt round-robin",
               "note": REF_NOTE if kind == "reference" else "CPU oracle port of src/ORBextractor.cc (oracle/_ref has not been built)"}

    if rank == 0:
        _emit({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: batch of 640x480 synthetic frames, nFeatures=1000, 8 levels, scale 1.2, FAST 20/7",
                       "frames_per_gpu_per_step": nF, "unique_frames_per_gpu": min(args.unique, nF), "frames_per_device_pass": args.chunk, "frames_per_e2e_chunk": args.e2e_chunk,
                       "keypoints_per_frame": kp_per_frame, "l2": "inputs_exceed_l2 (1.26 GB of frames per step per GPU)",
                       "parallelism": f"frame-sharded x{world}, no data-path collective", "numa_node_rank0": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "matching": matching, "bow": bow, "tracking": tracking, "configs": configs,
        })
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
