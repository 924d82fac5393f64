#!/usr/bin/env python3
"""Single-call latencies through the C ABI (the real-time use of the drop-in: one frame / one keyframe pair per call)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth, synth_descriptors

def bench(fn, n=200, warm=20):
    for _ in range(warm): fn()
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n * 1e3

for (W, H, nf, nl) in [(640, 480, 1000, 8), (1280, 720, 2000, 8), (3840, 2160, 8000, 12)]:
    img = synth(W, H, 0)
    ex = orb.ORBextractor(nf, 1.2, nl, 20, 7)
    ex(img)
    print(f"orbx_extract {W}x{H} nf={nf}: {bench(lambda: ex(img, download_pyramid=False), 100 if W < 3000 else 20):.3f} ms/frame "
          f"(with mvImagePyramid download: {bench(lambda: ex(img), 50 if W < 3000 else 10):.3f} ms)")
rng = np.random.default_rng(0)
n = 2000
d1 = synth_descriptors(n, 1); d2 = synth_descriptors(n, 2, dup_of=d1)
node1 = rng.integers(0, 100, n); node2 = rng.integers(0, 100, n)
ang = rng.uniform(0, 360, n).astype(np.float32)
flag = np.ones(n, np.uint8)
kf = orb.View(d1, orb.FeatureVector(node1), ang, flag=flag)
fr = orb.View(d2, orb.FeatureVector(node2), ang, flag=flag)
m = orb.ORBmatcher(0.7, True)
print(f"SearchByBoW(KF,Frame) 2000x2000, 100 nodes: {bench(lambda: m.SearchByBoW(kf, fr)):.3f} ms/call")
print(f"hamming_top2 2000x2000: {bench(lambda: m.hamming_top2(d1, d2)):.3f} ms/call")
print(f"DescriptorDistance (1 pair): {bench(lambda: orb.ORBmatcher.DescriptorDistance(d1[0], d2[0])):.4f} ms/call")
