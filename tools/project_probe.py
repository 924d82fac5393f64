"""Latency / round-count probe of the window searches on realistic sizes (experiment helper, not a test)."""
import sys
import time

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from oracle import orb_oracle_py as orc
import proj_util as pu


def bench(name, f, of, reps=20):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    t = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(3):
        of()
    to = (time.perf_counter() - t0) / 3
    print(f"{name}: B200 {t * 1e3:.3f} ms/call  oracle(CPU, 1 thread) {to * 1e3:.3f} ms/call", flush=True)


for n, cluster in [(2000, False), (2000, True), (8000, False)]:
    rng = np.random.default_rng(1)
    fa = pu.frame_arrays(n, rng, stereo=True, cluster=cluster)
    blocked = (rng.random(n) < 0.1).astype(np.uint8)
    g, og = pu.make_grids(fa, blocked, orb, orc)
    mp = pu.map_points_for(fa, n, rng)
    m = orb.ORBmatcher(0.8, True)
    bench(f"SearchByProjection(F, MapPoints) n={n} cluster={cluster}", lambda: m.SearchByProjectionMapPoints(g, th=3.0, **mp),
          lambda: orc.search_projection_map(og, th=3.0, nnratio=0.8, **mp))
    lf = pu.last_frame_for(fa, n, rng)
    args = (lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], 40.0, 40.0 / lf["fx"], lf["has_point"], lf["world"], lf["octave"],
            lf["angle"], lf["desc"], lf["claims"], 7.0, False)
    bench(f"SearchByProjection(Cur, Last) n={n} cluster={cluster}", lambda: m.SearchByProjectionFrame(g, *args),
          lambda: orc.search_projection_frame(og, *args, True))
    fa2 = pu.frame_arrays(n, rng, stereo=False)
    fa2["octave"][rng.random(n) < 0.5] = 0
    g2, og2 = pu.make_grids(fa2, None, orb, orc)
    f1 = pu.init_frame1_for(fa2, n, rng)
    bench(f"SearchForInitialization n={n} window=100", lambda: m.SearchForInitialization(g2, f1["desc1"], f1["octave1"], f1["angle1"], f1["prev"].copy(), 100),
          lambda: orc.search_initialization(og2, f1["desc1"], f1["octave1"], f1["angle1"], f1["prev"].copy(), 100, 0.8, True))
