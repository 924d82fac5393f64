"""Builds a synthetic saved map at keyframe-database scale with the orbmap_* builder, writes and reloads it, and times the
archive -> GPU path: parse, gather of the observation sets, batched ComputeDistinctiveDescriptors on the device.
usage: python tools/map_archive_probe.py [n_keyframes] [features_per_keyframe] [n_mappoints]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orbslam_mapsave_b200 as orb  # noqa: E402


def main():
    n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    n_mp = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
    rng = np.random.default_rng(0)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    ar = orb.MapArchive.create()
    slots = np.full((n_kf, n_feat), -1, np.int64)
    obs = [[] for _ in range(n_mp)]
    for j in range(n_mp):
        kfs = rng.choice(n_kf, size=int(rng.integers(2, 12)), replace=False)
        fs = rng.integers(0, n_feat, len(kfs))
        for k, f in zip(kfs, fs):
            if slots[k, f] < 0:
                slots[k, f] = j
                obs[j].append((int(k), int(f)))
    for i in range(n_kf):
        kp = np.zeros(n_feat, orb.KP_DTYPE)
        kp["x"], kp["y"] = rng.uniform(0, 640, n_feat), rng.uniform(0, 480, n_feat)
        kp["octave"] = rng.integers(0, 8, n_feat)
        info = dict(id=i, frame_id=i, timestamp=0.05 * i, scale_factor=1.2, log_scale_factor=float(np.log(np.float32(1.2))), fx=500., fy=500.,
                    cx=320., cy=240., bf=40., b=0.08, th_depth=3.2, min_x=0, min_y=0, max_x=640, max_y=480, first_connection=int(i == 0),
                    has_parent=int(i > 0), parent_id=max(i - 1, 0))
        ar.add_keyframe(info, kp, kp, rng.integers(0, 256, (n_feat, 32), dtype=np.uint8), slots[i], sf, sf * sf, 1 / (sf * sf),
                        np.eye(4, dtype=np.float32), np.eye(3, dtype=np.float32))
    for j in range(n_mp):
        ar.add_mappoint(j, rng.normal(size=3), rng.normal(size=3), np.zeros(32, np.uint8), obs[j][0][0] if obs[j] else -1,
                        [o[0] for o in obs[j]], [o[1] for o in obs[j]])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "map.bin")
        t0 = time.perf_counter()
        ar.save(path)
        t_save = time.perf_counter() - t0
        size = os.path.getsize(path)
        t0 = time.perf_counter()
        back = orb.MapArchive.load(path)
        t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    desc, off = back.observed_descriptors()
    t_gather = time.perf_counter() - t0
    orb.distinctive_descriptors(desc, off)                     # warm-up (context, allocation)
    t0 = time.perf_counter()
    best = orb.distinctive_descriptors(desc, off)
    t_gpu = time.perf_counter() - t0
    print(json.dumps(dict(n_keyframes=n_kf, features_per_keyframe=n_feat, n_mappoints=n_mp, file_bytes=size,
                          save_s=round(t_save, 3), load_s=round(t_load, 3), load_gb_s=round(size / t_load / 1e9, 2),
                          gather_s=round(t_gather, 4), observed_descriptors=int(off[-1]),
                          distinctive_descriptors_call_s=round(t_gpu, 4), mappoints_per_s=round(n_mp / t_gpu),
                          chosen_nonempty=int((best >= 0).sum()))))


if __name__ == "__main__":
    main()
