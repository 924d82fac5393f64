"""bench_configs.py — the other BASELINE.json configurations, measured inside the driver-run bench line (bench.py `configs` key).

  C1  one 640x480 frame through orbx_extract (the reference's per-frame call, Frame::ExtractORB -> ORBextractor::operator(),
      src/ORBextractor.cc:1042-1108): single-call latency from pageable and from pinned caller buffers, the reference's own
      CPU extractor (oracle/_ref) per-frame time beside it
  C3  1280x720 / 2000 features / 8 levels: frames/s of a batch pass + Hamming top-2 + ratio test between consecutive keyframes
      (orbm_hamming_top2_batch_device on the descriptors the pass just produced; the inner loop of src/ORBmatcher.cc:205-233)
  C5  3840x2160 / 8000 features / 12 levels: frames/s of a batch pass, per-stage device times and the HBM roofline on SURVEY §8(d)'s
      54 089 102 algorithmic bytes per frame (the one configuration large enough to lean on HBM)

Every number is measured live (CUDA events on the launching stream, inputs larger than L2 per pass); `frac` values are
recomputable from the fields printed next to them.
"""
import os
import time

import numpy as np

CONFIGS = {
    "C1": dict(W=640, H=480, nfeat=1000, nlevels=8),
    "C3": dict(W=1280, H=720, nfeat=2000, nlevels=8),
    "C5": dict(W=3840, H=2160, nfeat=8000, nlevels=12),
}
SCALE, INI_TH, MIN_TH = 1.2, 20, 7
STAGES = ("pyramid", "fast", "octree", "blur", "describe")


def _gen(args):
    from orbslam_mapsave_b200.synth import synth
    W, H, seed = args
    return synth(W, H, seed)


def make_frames(W, H, n, unique, seed0, extra_w=0):
    """`unique` distinct synth() frames of (W + extra_w) x H, cached under /tmp, repeated cyclically to n."""
    unique = min(unique, n)
    cache = f"/tmp/orb_synth_{W + extra_w}x{H}_s{seed0}_u{unique}.npy"
    if os.path.exists(cache):
        base = np.load(cache)
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(max(1, min(16, os.cpu_count() or 4))) as pool:
            base = np.stack(pool.map(_gen, [(W + extra_w, H, seed0 + i) for i in range(unique)]))
        try:
            np.save(cache + f".{os.getpid()}.tmp.npy", base)
            os.replace(cache + f".{os.getpid()}.tmp.npy", cache)
        except OSError:
            pass
    reps = (n + unique - 1) // unique
    return np.concatenate([base] * reps)[:n] if reps > 1 else base[:n]


def algorithmic_bytes(ex, W, H, nlevels, kp_per_frame, cand_per_frame):
    """B_frame = W*H (read input) + sum_{l>=1} w_l*h_l (write levels) + sum_{l>=0} w_l*h_l (read every level once)
    + nkp*(32+28), and the same split per stage (what each stage cannot avoid reading / writing)."""
    px = [ex.level_size(l) for l in range(nlevels)]
    pyr = sum(w * h for w, h in px)
    upper = pyr - W * H
    b_frame = W * H + upper + pyr + kp_per_frame * 60
    stage = {
        "pyramid": W * H + W * H + 2 * upper,                # read input, write level 0, read each source level once, write levels >= 1
        "fast": pyr + cand_per_frame * 8,                    # read every level once, write the candidates (8 B each)
        "octree": cand_per_frame * 8 + kp_per_frame * 8,     # read candidates, write the selection
        "blur": 2 * pyr,                                     # read every level, write every blurred level
        "describe": kp_per_frame * (749 + 37 * 37 + 60),     # per keypoint: 749-px disc + 37x37 blurred patch + 28 B + 32 B out
    }
    return b_frame, stage, pyr


def _timed(torch, fn, k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / k


def batch_config(name, torch, orb, capi, device, hbm_peak, frames, unique, batch, steps, stream, e2e=True, ncu_constants=None):
    """Device-resident frames/s, per-stage ms and roofline of one configuration; optional host-buffer (e2e) pass."""
    cfg = CONFIGS[name]
    W, H, nfeat, nl = cfg["W"], cfg["H"], cfg["nfeat"], cfg["nlevels"]
    nF = len(frames)
    h_frames = torch.from_numpy(frames).pin_memory()
    d_frames = h_frames.cuda(non_blocking=True)
    ex = orb.ORBextractor(nfeat, SCALE, nl, INI_TH, MIN_TH, W, H, max_batch=batch, device=device)
    cap = ex.max_keypoints()
    d_kp = torch.zeros((nF, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((nF, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nF, dtype=torch.int32, device="cuda")

    def step(stages=capi.STAGE_ALL):
        ex.extract_batch_device(d_frames, d_kp, d_desc, d_n, cap, stream=stream, stages=stages)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ex.check_status()
    l0 = ex.launch_count()
    secs = _timed(torch, step, steps)
    launches = (ex.launch_count() - l0) / steps
    ex.check_status()
    kp_per_frame = float(d_n.sum().item()) / nF
    # FAST candidates per frame (for the stage byte counts): read back from frame 0 of the last pass
    cand0 = sum(len(ex.candidates(0, l)) for l in range(nl))
    b_frame, stage_b, pyr_px = algorithmic_bytes(ex, W, H, nl, kp_per_frame, cand0)
    stage_s = {}
    bits = dict(zip(STAGES, (capi.STAGE_PYRAMID, capi.STAGE_FAST, capi.STAGE_OCTREE, capi.STAGE_BLUR, capi.STAGE_DESCRIBE)))
    for nm in STAGES:
        step(bits[nm])
        stage_s[nm] = _timed(torch, lambda b=bits[nm]: step(b), max(1, min(steps, 3)))
    step()                                                    # leave a complete pass behind
    torch.cuda.synchronize()
    dom = max(stage_s, key=stage_s.get)
    fps = nF / secs
    out = {
        "workload": f"{name}: batch of {nF} synthetic {W}x{H} frames ({min(unique, nF)} distinct), "
                    f"nFeatures={nfeat}, {nl} levels, scale 1.2, FAST 20/7; {batch} frames per device pass",
        "frames_per_s": fps, "ms_per_frame": 1e3 * secs / nF, "keypoints_per_frame": kp_per_frame, "fast_candidates_frame0": cand0,
        "pyramid_pixels": pyr_px, "gpu_launches_per_step": launches,
        "l2": f"inputs_exceed_l2 ({nF * W * H / 1e6:.0f} MB of frames + {nF * 2 * pyr_px / 1e6:.0f} MB of pyramid per step)",
        "stage_ms_per_frame": {k: 1e3 * v / nF for k, v in stage_s.items()},
        "roofline": {
            "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "algorithmic_bytes_per_frame": b_frame,
            "achieved": b_frame * fps / 1e9, "frac": b_frame * fps / 1e9 / hbm_peak,
            "dominant_stage": dom,
            "stages": {k: {"algorithmic_bytes_per_frame": stage_b[k], "achieved": stage_b[k] * nF / stage_s[k] / 1e9,
                           "frac": stage_b[k] * nF / stage_s[k] / 1e9 / hbm_peak} for k in STAGES},
        },
    }
    if ncu_constants:
        out["roofline"]["ncu"] = ncu_constants
    if e2e:
        h_kp = torch.zeros((nF, cap, 7), dtype=torch.float32).pin_memory()
        h_desc = torch.zeros((nF, cap, 32), dtype=torch.uint8).pin_memory()
        h_n = np.zeros(nF, np.int32)
        chunk = max(1, batch // 4)
        ex2 = orb.ORBextractor(nfeat, SCALE, nl, INI_TH, MIN_TH, W, H, max_batch=chunk, device=device)

        def e2e_step():
            capi.check(capi.lib().orbx_extract_batch(ex2.handle, capi._p(h_frames), nF, W, H, W, W * H, None, 0, 0,
                                                     capi._p(h_kp), capi._p(h_desc), cap, capi._p(h_n)))
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = max(2, min(steps, 5))
        for _ in range(k):
            e2e_step()
        s = (time.perf_counter() - t0) / k
        out["e2e"] = {"frames_per_s": nF / s, "ms_per_frame": 1e3 * s / nF, "h2d_bytes_per_step": int(nF * W * H),
                      "d2h_bytes_per_step": int(h_n.sum()) * 60 + nF * 4, "frames_per_chunk": chunk,
                      "note": "orbx_extract_batch from pinned host buffers, H2D + D2H inside the timed region"}
        del ex2
    return out, (ex, d_desc, d_n, cap, nF)


def consecutive_keyframe_matching(torch, capi, ctx, steps, stream, popc_peak):
    """C3's second half: top-2 + ratio between consecutive keyframes (k, k+1), all pairs of the batch in ONE launch of
    orbm_hamming_top2_batch_device, on the descriptors the extraction pass left in HBM."""
    ex, d_desc, d_n, cap, nF = ctx
    n = d_n.to(torch.int32)
    npairs = nF - 1
    q_off = (torch.arange(npairs, device="cuda", dtype=torch.int32) * cap).contiguous()
    db_off = (q_off + cap).contiguous()
    q_cnt = n[:-1].contiguous()
    db_cnt = n[1:].contiguous()
    best_idx = torch.zeros(nF * cap, dtype=torch.int32, device="cuda")
    best = torch.zeros_like(best_idx)
    second = torch.zeros_like(best_idx)
    max_q = int(q_cnt.max().item())

    def mstep():
        capi.check(capi.lib().orbm_hamming_top2_batch_device(capi._p(d_desc), capi._p(q_off), capi._p(q_cnt), capi._p(d_desc), capi._p(db_off),
                                                             capi._p(db_cnt), npairs, max_q, capi._p(best_idx), capi._p(best), capi._p(second), stream))
    for _ in range(3):
        mstep()
    s = _timed(torch, mstep, max(3, steps))
    pairs = float((q_cnt.double() * db_cnt.double()).sum().item())
    # the ratio test on the result (what SearchByBoW does with best / second, src/ORBmatcher.cc:231-233), for the reported match count
    ok = torch.zeros(nF * cap, dtype=torch.bool, device="cuda")
    for p in range(0, npairs, max(1, npairs // 8)):
        o, c = p * cap, int(q_cnt[p].item())
        b, s2 = best[o:o + c].float(), second[o:o + c].float()
        ok[o:o + c] = (b <= 50) & (b < 0.75 * s2)
    sampled = len(range(0, npairs, max(1, npairs // 8)))
    return {"metric": "Hamming top-2 + ratio between consecutive keyframes", "keyframe_pairs_per_s": npairs / s,
            "descriptor_pairs_per_s": pairs / s, "query_descriptors_per_s": float(q_cnt.sum().item()) / s,
            "ms_per_step": 1e3 * s, "keyframe_pairs": npairs, "matches_per_pair_sampled": float(ok.sum().item()) / sampled,
            "roofline": {"bound": "popc", "unit": "TPOPC/s", "achieved": pairs / s * 8 / 1e12, "peak": popc_peak / 1e12,
                         "frac": pairs / s * 8 / popc_peak, "note": "8 algorithmic POPC per descriptor pair (SURVEY 8d)"}}


def moving_window_frames(W, H, n, scenes, seed0, step_px=4):
    """Consecutive 'keyframes' with true matches: each of `scenes` wide synthetic scenes is viewed through a W x H window that
    moves `step_px` to the right per frame."""
    per = (n + scenes - 1) // scenes
    wide = make_frames(W, H, scenes, scenes, seed0, extra_w=per * step_px)
    out = np.empty((n, H, W), np.uint8)
    for i in range(n):
        s, k = divmod(i, per)
        out[i] = wide[s][:, k * step_px:k * step_px + W]
    return out


def c1_latency(torch, orb, capi, device, frame, reps=200, ref_ms=None):
    """Single-frame operator() latency through orbx_extract (blocking call: upload, 13 kernels, download)."""
    W, H = frame.shape[1], frame.shape[0]
    ex = orb.ORBextractor(1000, SCALE, 8, INI_TH, MIN_TH, W, H, max_batch=1, device=device)
    cap = ex.max_keypoints()
    import ctypes as C
    n = C.c_int()

    def run(img, kp, desc):
        capi.check(capi.lib().orbx_extract(ex.handle, capi._p(img), W, H, W, None, 0, capi._p(kp), capi._p(desc), cap, C.byref(n)))

    def lat(img, kp, desc):
        for _ in range(10):
            run(img, kp, desc)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            run(img, kp, desc)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        return {"median_ms": float(np.median(ts)), "p10_ms": float(np.percentile(ts, 10)), "p90_ms": float(np.percentile(ts, 90))}
    pageable = lat(np.ascontiguousarray(frame), np.zeros((cap, 7), np.float32), np.zeros((cap, 32), np.uint8))
    h_img = torch.from_numpy(np.ascontiguousarray(frame)).pin_memory()
    h_kp = torch.zeros((cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.zeros((cap, 32), dtype=torch.uint8).pin_memory()
    pinned = lat(h_img, h_kp, h_desc)
    out = {"workload": "C1: one synthetic 640x480 frame, nFeatures=1000, 8 levels, scale 1.2, FAST 20/7; orbx_extract (blocking), "
                       f"{reps} calls", "keypoints": int(n.value), "pageable_caller_buffers": pageable, "pinned_caller_buffers": pinned,
           "h2d_bytes": W * H, "d2h_bytes": int(n.value) * 60 + 8}
    if ref_ms is not None:
        out["reference_cpu_ms_per_frame"] = ref_ms
    return out


def reference_ms_per_frame(frame, nfeat=1000, nlevels=8, reps=5):
    """The reference's own extractor (oracle/_ref) on ONE host thread, per-frame time (test infrastructure, CPU baseline leg only)."""
    try:
        from oracle import ref_py
        if not ref_py.available():
            return None
        rex = ref_py.RefExtractor(nfeat, SCALE, nlevels, INI_TH, MIN_TH)
        rex.extract(frame)
        t0 = time.perf_counter()
        for _ in range(reps):
            rex.extract(frame)
        return 1e3 * (time.perf_counter() - t0) / reps
    except Exception:
        return None
