"""Latency of orbx_stereo_matches (experiment helper)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
for (W, H, nf) in [(640, 480, 1000), (1280, 720, 2000)]:
    left = synth(W, H, 0); right = np.roll(left, -12, axis=1)
    exL, exR = orb.ORBextractor(nf, 1.2, 8, 20, 7), orb.ORBextractor(nf, 1.2, 8, 20, 7)
    kL, dL = exL(left, download_pyramid=False); kR, dR = exR(right, download_pyramid=False)
    f = lambda: exL.ComputeStereoMatches(exR, kL, dL, kR, dR, 40.0, 0.08)
    f()
    t0 = time.perf_counter()
    for _ in range(50): u, d = f()
    print(W, H, nf, f"{(time.perf_counter()-t0)/50*1e3:.3f} ms/call, matched {(u>=0).sum()} of {len(kL)}", flush=True)
