"""Randomised soak of the window searches against the oracle (many seeds, heavy claim contention, odd sizes).
usage: python tools/window_soak.py [n_seeds]"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import orbslam_mapsave_b200 as orb
from oracle import orb_oracle_py as orc
import proj_util as pu

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = 0
for seed in range(n_seeds):
    rng = np.random.default_rng(10_000 + seed)
    n = int(rng.choice([37, 300, 1000, 2000, 3000]))
    cluster = bool(rng.random() < 0.5)
    stereo = bool(rng.random() < 0.6)
    fa = pu.frame_arrays(n, rng, stereo=stereo, cluster=cluster)
    blocked = (rng.random(n) < rng.choice([0.0, 0.1, 0.5])).astype(np.uint8)
    g, og = pu.make_grids(fa, blocked, orb, orc)
    npts = int(rng.choice([5, 500, 2500, 5000]))
    mp = pu.map_points_for(fa, npts, rng, contention=float(rng.choice([0.0, 0.3, 0.9])))
    mp["claims"] = (rng.random(npts) < rng.choice([0.0, 0.5, 1.0])).astype(np.uint8)
    th = float(rng.choice([1.0, 3.0, 8.0]))
    ratio = float(rng.choice([0.6, 0.8, 1.0]))
    a = orc.search_projection_map(og, th=th, nnratio=ratio, **mp)
    b = orb.ORBmatcher(ratio, True).SearchByProjectionMapPoints(g, th=th, **mp)
    ok1 = a[0] == b[0] and np.array_equal(a[1], b[1])
    lf = pu.last_frame_for(fa, npts, rng, tz=float(rng.choice([0.0, 0.5, -0.5])))
    lf["claims"] = (rng.random(npts) < rng.choice([0.0, 0.5, 1.0])).astype(np.uint8)
    mono = not stereo
    ori = bool(rng.random() < 0.7)
    args = (lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], 40.0, 40.0 / lf["fx"], lf["has_point"], lf["world"], lf["octave"],
            lf["angle"], lf["desc"], lf["claims"], float(rng.choice([7.0, 15.0, 30.0])), mono)
    a = orc.search_projection_frame(og, *args, ori)
    b = orb.ORBmatcher(0.9, ori).SearchByProjectionFrame(g, *args)
    ok2 = a[0] == b[0] and np.array_equal(a[1], b[1])
    fa2 = pu.frame_arrays(n, rng, stereo=False, cluster=cluster)
    fa2["octave"][rng.random(n) < 0.6] = 0
    g2, og2 = pu.make_grids(fa2, None, orb, orc)
    f1 = pu.init_frame1_for(fa2, max(npts // 2, 3), rng)
    w = int(rng.choice([5, 30, 100, 200]))
    p1, p2 = f1["prev"].copy(), f1["prev"].copy()
    a = orc.search_initialization(og2, f1["desc1"], f1["octave1"], f1["angle1"], p1, w, ratio, ori)
    b = orb.ORBmatcher(ratio, ori).SearchForInitialization(g2, f1["desc1"], f1["octave1"], f1["angle1"], p2, w)
    ok3 = a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(p1, p2)
    if not (ok1 and ok2 and ok3):
        bad += 1
        print("MISMATCH seed", seed, ok1, ok2, ok3, flush=True)
print(f"window soak: {n_seeds} seeds, {bad} mismatching", flush=True)
