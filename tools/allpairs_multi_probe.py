"""orbm_allpairs_multi (one process, several GPUs, host pointers): pairs/s on 1 .. N GPUs and equality of the tables.
usage: python tools/allpairs_multi_probe.py [n_kf] [per_kf]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orbslam_mapsave_b200 as orb                                  # noqa: E402
from orbslam_mapsave_b200.matcher import allpairs_multi             # noqa: E402

n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 512
per = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(0)
desc = rng.integers(0, 256, (n_kf, per, 32), dtype=np.uint8)
ndev = orb.device_count()
out = {"n_kf": n_kf, "per_kf": per, "descriptor_pairs": float(n_kf) * (n_kf - 1) * per * per, "gpus_visible": ndev, "runs": []}
ref = None
n = 1
while n <= ndev:
    devs = tuple(range(n))
    allpairs_multi(desc[:8], devices=devs)                          # contexts, first-launch costs
    t0 = time.perf_counter()
    cnt = allpairs_multi(desc, devices=devs)
    s = time.perf_counter() - t0
    if ref is None:
        ref = cnt
    out["runs"].append({"gpus": n, "seconds": s, "pairs_per_s": out["descriptor_pairs"] / s, "table_equals_1gpu": bool(np.array_equal(cnt, ref))})
    print(out["runs"][-1], flush=True)
    n *= 2
print(json.dumps(out))
