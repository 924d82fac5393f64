"""GPU parity: a saved map read through orbmap_* feeds the device entry points — the distinctive descriptor of
every map point from its observation set, and keyframe-to-keyframe top-2 matching on the stored mDescriptors — and the results
equal the oracle's on the in-memory arrays the archive was written from."""
import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
from oracle import orb_oracle_py as orc
from map_archive_writer import make_random_map, serialize_map

pytestmark = pytest.mark.gpu


def test_archive_feeds_distinctive_descriptors_and_matching(tmp_path):
    m = make_random_map(11, n_kf=6, n_feat=0, n_mp=0)
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    rng = np.random.default_rng(3)
    base = synth(640, 480, 5)
    # keyframes = the extractor's output on shifted views of one scene, so that real matches exist
    for i, k in enumerate(m["keyframes"]):
        kp, desc = ex(np.roll(base, (2 * i, -5 * i), axis=(0, 1)))
        n = len(kp)
        keys = [dict(x=kp["x"][j], y=kp["y"][j], angle=kp["angle"][j], response=kp["response"][j], octave=int(kp["octave"][j]), class_id=-1)
                for j in range(n)]
        k.update(n=n, keys=keys, keys_un=keys, desc=desc, uright=np.full(n, -1, np.float32), depth=np.full(n, -1, np.float32),
                 mappoint_ids=[-1] * n, grid=[[[] for _ in range(48)] for _ in range(64)])
    m["origins"] = [m["keyframes"][0]]
    # map points: observation sets of 0..6 keyframes; mDescriptor = the oracle's distinctive choice
    mps = []
    for j in range(700):
        obs = []
        for k in m["keyframes"]:
            if rng.random() < 0.6:
                f = int(rng.integers(0, k["n"]))
                if k["mappoint_ids"][f] < 0:
                    k["mappoint_ids"][f] = j
                    obs.append((k["id"], f))
        by_id = {k["id"]: k for k in m["keyframes"]}
        rows = np.stack([by_id[kf]["desc"][f] for kf, f in obs]) if obs else np.zeros((0, 32), np.uint8)
        best = orc.distinctive(rows)
        mps.append(dict(id=j, next_id=700, first_kf=0, first_frame=0, n_obs=len(obs), world_pos=rng.normal(size=(3, 1)).astype(np.float32),
                        obs=obs, normal=rng.normal(size=(3, 1)).astype(np.float32),
                        desc=(rows[best:best + 1] if best >= 0 else np.zeros((1, 32), np.uint8)),
                        ref_kf=(obs[0][0] if obs else None), visible=1, found=1, bad=False, min_dist=0.5, max_dist=9.0, _best=best))
    m["mappoints"] = mps
    path = tmp_path / "map.bin"
    path.write_bytes(serialize_map(m))

    ar = orb.MapArchive.load(path)
    got_desc, got_best = ar.distinctive_descriptors()
    assert np.array_equal(got_best, np.array([p["_best"] for p in mps], np.int32))
    stored = ar.mappoints()["desc"]
    has = got_best >= 0
    assert has.sum() > 600 and np.array_equal(got_desc[has], stored[has])

    # loop-closing style matching between stored keyframes: device top-2 on the archive's descriptors == oracle on the originals
    a, b = ar.keyframe(0), ar.keyframe(3)
    got = orb.ORBmatcher().hamming_top2(a["desc"], b["desc"])
    want = orc.hamming_top2(m["keyframes"][0]["desc"], m["keyframes"][3]["desc"])
    assert all(np.array_equal(x, y) for x, y in zip(got, want))
    best_dist = got[1]
    assert (best_dist <= 50).sum() > 100          # shifted views of one scene do match
