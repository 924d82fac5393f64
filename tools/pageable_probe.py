"""orbx_extract_batch from ordinary (pageable) host arrays: frames/s against the size of the host staging pool (ORBX_HOST_THREADS is read
once per process, so every pool size runs in its own process).  usage: python tools/pageable_probe.py [threads ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np
    import orbslam_mapsave_b200 as orb
    from orbslam_mapsave_b200 import capi
    from orbslam_mapsave_b200.synth import synth
    nF, W, H = 4096, 640, 480
    base = np.stack([synth(W, H, s) for s in range(64)])
    frames = np.ascontiguousarray(np.tile(base, (nF // 64, 1, 1)))
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, W, H, max_batch=128)
    cap = ex.max_keypoints()
    kp = np.zeros((nF, cap), capi.KP_DTYPE)
    desc = np.zeros((nF, cap, 32), np.uint8)
    n = np.zeros(nF, np.int32)

    def step():
        capi.check(capi.lib().orbx_extract_batch(ex.handle, capi._p(frames), nF, W, H, W, W * H, None, 0, 0, capi._p(kp), capi._p(desc), cap,
                                                 capi._p(n)))
    step()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    s = (time.perf_counter() - t0) / 5
    print(f"ORBX_HOST_THREADS={os.environ.get('ORBX_HOST_THREADS', 'default')}: {nF / s:9.0f} frames/s  {1e3 * s:7.2f} ms per 4096 frames "
          f"({(frames.nbytes + int(n.sum()) * 60) / s / 1e9:.1f} GB/s through host memcpy)", flush=True)
else:
    ths = sys.argv[1:] or ["1", "2", "4", "8", "16"]
    print(f"host threads available: {len(os.sched_getaffinity(0))}", flush=True)
    for t in ths:
        env = dict(os.environ, ORBX_HOST_THREADS=t)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)
