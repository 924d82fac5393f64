"""GPU parity tests of the ORBmatcher Hamming searches (through the C ABI) against the CPU oracle.  Integer/index
work: everything must be bit-exact."""
import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth_descriptors
from oracle import orb_oracle_py as orc

pytestmark = pytest.mark.gpu


def _flip(d, nbits, rng):
    d = d.copy()
    for b in rng.choice(256, nbits, replace=False):
        d[b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def test_descriptor_distance_matches_bit_hack():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b[:10] = a[:10]
    b[10] = ~a[10]
    got = orb.ORBmatcher.DescriptorDistance(a, b)
    want = np.array([orc.descriptor_distance(x, y) for x, y in zip(a, b)], np.int32)
    assert np.array_equal(got, want)
    assert got[0] == 0 and got[10] == 256
    assert orb.ORBmatcher.DescriptorDistance(a[20], b[20]) == want[20]


@pytest.mark.parametrize("nq,ndb", [(1, 1), (7, 0), (3, 1), (100, 2), (257, 513), (2000, 2000), (5000, 300), (129, 4097)])
def test_hamming_top2_vs_oracle(nq, ndb):
    db = synth_descriptors(ndb, 1)
    q = synth_descriptors(nq, 2, dup_of=db if ndb else None)
    if ndb >= 4 and nq >= 4:
        db[3] = db[1]                     # exact duplicate rows: first index must win, second == best
        q[0] = db[1]
    bi, bd, sd = orb.ORBmatcher().hamming_top2(q, db)
    obi, obd, osd = orc.hamming_top2(q, db)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(sd, osd)
    if ndb >= 4 and nq >= 4:
        assert bi[0] == 1 and bd[0] == 0 and sd[0] == 0


def _ratio_edge_scene():
    """Constructed ties of the ratio test: per vocabulary node one keyframe feature and two frame features at Hamming
    distances (best, second); `float(best) < nnratio * float(second)` (src/ORBmatcher.cc:231-233) is an equality for
    (25, 50) at ratio 0.5 and for (30, 40) at ratio 0.75, true one bit below and false one bit above; TH_LOW = 50 is hit exactly."""
    rng = np.random.default_rng(77)
    cases = [(25, 50), (24, 50), (26, 50), (30, 40), (29, 40), (31, 40), (50, 100), (49, 99), (50, 101), (0, 0), (0, 1), (10, 10)]
    n1 = len(cases)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = np.zeros((2 * n1, 32), np.uint8)
    for i, (b, s2) in enumerate(cases):
        d2[2 * i] = _flip(d1[i], b, rng)
        # the second candidate at distance s2 from d1[i]: flip s2 bits chosen independently of the first candidate's
        d2[2 * i + 1] = _flip(d1[i], s2, rng)
    node1 = np.arange(n1) * 2 + 3
    node2 = np.repeat(node1, 2)
    ang1 = np.full(n1, 10.0, np.float32)
    return dict(d1=d1, d2=d2, node1=node1, node2=node2, ang1=ang1, ang2=np.full(2 * n1, 10.0, np.float32), flag1=np.ones(n1, np.uint8),
                flag2=np.ones(2 * n1, np.uint8), cases=cases)


@pytest.mark.parametrize("ratio", [0.5, 0.75, 0.6, 1.0])
def test_ratio_test_equality_edge(ratio):
    s = _ratio_edge_scene()
    f1, f2 = orc.FeatVec(s["node1"]), orc.FeatVec(s["node2"])
    on, om = orc.search_bow_kf_f(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["ang2"], f2, ratio, False)
    k = orb.View(s["d1"], orb.FeatureVector(s["node1"]), s["ang1"], flag=s["flag1"])
    f = orb.View(s["d2"], orb.FeatureVector(s["node2"]), s["ang2"])
    gn, gm = orb.ORBmatcher(ratio, False).SearchByBoW(k, f)
    assert gn == on and np.array_equal(gm, om)
    # the expected outcome, from the definition: accepted iff best <= 50 and float32(best) < float32(ratio) * float32(second)
    for i, (b, s2) in enumerate(s["cases"]):
        lo, hi = min(b, s2), max(b, s2)
        want = lo <= 50 and np.float32(lo) < np.float32(ratio) * np.float32(hi)
        assert (om[2 * i] == i or om[2 * i + 1] == i) == bool(want), (i, b, s2, ratio)


def _scene(n1, n2, seed, nnodes=40, tri=False):
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    node1 = rng.integers(0, nnodes, n1) * 3 + 5
    node2 = rng.integers(0, nnodes + 6, n2) * 3 + 5          # some nodes exist on one side only
    # plant true matches at controlled distances inside the same node, incl. the TH_LOW edge (49/50/51) and ties
    k = min(n1, n2) // 2
    src = rng.choice(n1, k, replace=False)
    dst = rng.choice(n2, k, replace=False)
    for s, t in zip(src, dst):
        nb = int(rng.choice([0, 1, 5, 20, 35, 49, 50, 51, 60]))
        d2[t] = _flip(d1[s], nb, rng)
        node2[t] = node1[s]
    # competing near-duplicates so that the ratio test and the greedy claim matter
    for s, t in list(zip(src, dst))[: k // 3]:
        u = int(rng.integers(0, n2))
        d2[u] = _flip(d2[t], int(rng.choice([0, 1, 2, 30])), rng)
        node2[u] = node2[t]
    ang1 = rng.uniform(0, 360, n1).astype(np.float32)
    ang2 = (ang1[rng.integers(0, n1, n2)] if n1 else np.zeros(n2, np.float32)) + rng.choice([0.0, 2.0, 14.9, 15.0, 45.0, 200.0], n2).astype(np.float32)
    ang2 = np.mod(ang2, 360).astype(np.float32)
    flag1 = (rng.random(n1) < 0.7).astype(np.uint8)
    flag2 = (rng.random(n2) < 0.7).astype(np.uint8)
    out = dict(d1=d1, d2=d2, node1=node1, node2=node2, ang1=ang1, ang2=ang2, flag1=flag1, flag2=flag2)
    if tri:
        out.update(x1=rng.uniform(0, 640, n1).astype(np.float32), y1=rng.uniform(0, 480, n1).astype(np.float32),
                   x2=rng.uniform(0, 640, n2).astype(np.float32), y2=rng.uniform(0, 480, n2).astype(np.float32),
                   oct2=rng.integers(0, 8, n2).astype(np.int32),
                   ur1=np.where(rng.random(n1) < 0.3, 100.0, -1.0).astype(np.float32),
                   ur2=np.where(rng.random(n2) < 0.3, 100.0, -1.0).astype(np.float32))
    return out


@pytest.mark.parametrize("n1,n2,seed", [(300, 280, 0), (2000, 2000, 1), (1000, 40, 2), (50, 900, 3), (0, 10, 4), (10, 0, 5)])
@pytest.mark.parametrize("ratio,ori", [(0.7, True), (0.75, True), (0.6, False), (1.0, True)])
def test_search_by_bow_kf_frame(n1, n2, seed, ratio, ori):
    s = _scene(n1, n2, seed)
    fv1o, fv2o = orc.FeatVec(s["node1"]), orc.FeatVec(s["node2"])
    on, om = orc.search_bow_kf_f(s["d1"], s["flag1"], s["ang1"], fv1o, s["d2"], s["ang2"], fv2o, ratio, ori)
    kf = orb.View(s["d1"], orb.FeatureVector(s["node1"]), s["ang1"], flag=s["flag1"])
    fr = orb.View(s["d2"], orb.FeatureVector(s["node2"]), s["ang2"])
    n, m = orb.ORBmatcher(ratio, ori).SearchByBoW(kf, fr)
    assert n == on and np.array_equal(m, om)
    if n1 >= 300 and n2 >= 280:
        assert n > 10


@pytest.mark.parametrize("n1,n2,seed", [(300, 280, 10), (2000, 2000, 11), (700, 64, 12), (0, 0, 13)])
@pytest.mark.parametrize("ratio,ori", [(0.75, True), (0.6, False)])
def test_search_by_bow_kf_kf(n1, n2, seed, ratio, ori):
    s = _scene(n1, n2, seed)
    on, om = orc.search_bow_kf_kf(s["d1"], s["flag1"], s["ang1"], orc.FeatVec(s["node1"]), s["d2"], s["flag2"], s["ang2"],
                                  orc.FeatVec(s["node2"]), ratio, ori)
    k1 = orb.View(s["d1"], orb.FeatureVector(s["node1"]), s["ang1"], flag=s["flag1"])
    k2 = orb.View(s["d2"], orb.FeatureVector(s["node2"]), s["ang2"], flag=s["flag2"])
    n, m = orb.ORBmatcher(ratio, ori).SearchByBoW(k1, k2, kf_kf=True)
    assert n == on and np.array_equal(m, om)


@pytest.mark.parametrize("n1,n2,seed", [(300, 280, 20), (2000, 2000, 21), (64, 900, 22), (5, 0, 23)])
@pytest.mark.parametrize("only_stereo,ori", [(False, False), (True, False), (False, True)])
def test_search_for_triangulation(n1, n2, seed, only_stereo, ori):
    s = _scene(n1, n2, seed, tri=True)
    rng = np.random.default_rng(seed)
    # a fundamental matrix that makes a good fraction of planted pairs pass: lines through random points
    F12 = (rng.normal(0, 1, (3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-3], [1e-3, 1e-3, 1e-1]])).astype(np.float32)
    sf2 = (1.2 ** np.arange(8)).astype(np.float32)
    sig2 = (sf2 * sf2 * 5000).astype(np.float32)       # generous sigma so the epipolar gate passes often but not always
    ex, ey = 320.0, 240.0
    on, op = orc.search_triangulation(s["d1"], s["flag1"], s["ur1"], s["x1"], s["y1"], s["ang1"], orc.FeatVec(s["node1"]),
                                      s["d2"], s["flag2"], s["ur2"], s["x2"], s["y2"], s["ang2"], s["oct2"], orc.FeatVec(s["node2"]),
                                      F12, ex, ey, sf2, sig2, only_stereo, ori)
    k1 = orb.View(s["d1"], orb.FeatureVector(s["node1"]), s["ang1"], flag=s["flag1"], x=s["x1"], y=s["y1"],
                  octave=np.zeros(n1, np.int32), uright=s["ur1"])
    k2 = orb.View(s["d2"], orb.FeatureVector(s["node2"]), s["ang2"], flag=s["flag2"], x=s["x2"], y=s["y2"], octave=s["oct2"],
                  uright=s["ur2"])
    n, p = orb.ORBmatcher(0.6, ori).SearchForTriangulation(k1, k2, F12, ex, ey, sf2, sig2, only_stereo)
    assert n == on and np.array_equal(p, op)
    if n1 >= 300 and n2 >= 280 and not only_stereo:
        assert len(p) >= 1


def test_three_maxima():
    rng = np.random.default_rng(3)
    cases = [rng.integers(0, 50, 30) for _ in range(20)] + [np.zeros(30, int), np.full(30, 4), np.eye(30, dtype=int)[7] * 9,
                                                            np.array([100, 9, 10, 11] + [0] * 26)]
    for h in cases:
        assert orb.ORBmatcher.ComputeThreeMaxima(h) == orc.three_maxima(h)


def test_allpairs_counts_and_best_vs_oracle():
    import torch
    n_kf, per = 6, 300
    base = synth_descriptors(per, 100)
    kfs = [base] + [synth_descriptors(per, 101 + i, dup_of=base, dup_rate=0.4) for i in range(n_kf - 1)]
    desc = np.stack(kfs)
    d_desc = torch.from_numpy(desc).cuda()
    q0, q1 = 1, 5
    nq = q1 - q0
    cnt = torch.zeros(nq * n_kf + 1, dtype=torch.int16, device="cuda")
    bkf = torch.zeros(nq * per, dtype=torch.int32, device="cuda")
    bd = torch.zeros(nq * per, dtype=torch.int32, device="cuda")
    from orbslam_mapsave_b200 import capi
    capi.check(capi.lib().orbm_allpairs_device(capi._p(d_desc), n_kf, per, q0, q1, 50, 0.75, capi._p(cnt), capi._p(bkf), capi._p(bd),
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    cnt = cnt.cpu().numpy()[:nq * n_kf].astype(np.int64).reshape(nq, n_kf)
    bkf, bd = bkf.cpu().numpy().reshape(nq, per), bd.cpu().numpy().reshape(nq, per)
    for qi, q in enumerate(range(q0, q1)):
        best_d = np.full(per, 1 << 30)
        best_k = np.full(per, -1)
        for k in range(n_kf):
            if k == q:
                assert cnt[qi, k] == 0
                continue
            _, b1, b2 = orc.hamming_top2(desc[q], desc[k])
            ok = (b1 <= 50) & (b1.astype(np.float32) < np.float32(0.75) * b2.astype(np.float32))
            assert cnt[qi, k] == int(ok.sum()), (q, k)
            upd = b1 < best_d
            best_k[upd] = k
            best_d[upd] = b1[upd]
        assert np.array_equal(bd[qi], best_d) and np.array_equal(bkf[qi], best_k)


@pytest.mark.parametrize("devices", [(0,), (0, 0, 0), (0, 0, 0, 0, 0, 0, 0, 0, 0)])
def test_allpairs_multi_shards_and_gathers_like_the_oracle(devices):
    """orbm_allpairs_multi: query keyframes sharded over the listed devices (a device listed several times = several shards, more
    shards than keyframes = empty shards), one host thread per shard, rows gathered into the caller's table."""
    import torch
    from orbslam_mapsave_b200.matcher import allpairs_multi
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("not enough GPUs")
    n_kf, per = 7, 260
    base = synth_descriptors(per, 300)
    desc = np.stack([base] + [synth_descriptors(per, 301 + i, dup_of=base, dup_rate=0.4) for i in range(n_kf - 1)])
    cnt, bk, bd = allpairs_multi(desc, 50, 0.75, devices, want_best=True)
    for q in range(n_kf):
        best_d = np.full(per, 1 << 30)
        best_k = np.full(per, -1)
        for k in range(n_kf):
            if k == q:
                assert cnt[q, k] == 0
                continue
            _, b1, b2 = orc.hamming_top2(desc[q], desc[k])
            ok = (b1 <= 50) & (b1.astype(np.float32) < np.float32(0.75) * b2.astype(np.float32))
            assert cnt[q, k] == int(ok.sum()), (q, k)
            upd = b1 < best_d
            best_k[upd] = k
            best_d[upd] = b1[upd]
        assert np.array_equal(bd[q], best_d) and np.array_equal(bk[q], best_k)
    assert np.array_equal(allpairs_multi(desc, 50, 0.75, devices), cnt)


def test_allpairs_multi_every_visible_gpu():
    """With N GPUs visible the table from all of them equals the single-GPU table (runs as a 1-GPU identity check on a 1-GPU box)."""
    import torch
    from orbslam_mapsave_b200.matcher import allpairs_multi
    n_kf, per = 24, 500
    base = synth_descriptors(per, 400)
    desc = np.stack([base] + [synth_descriptors(per, 401 + i, dup_of=base, dup_rate=0.3) for i in range(n_kf - 1)])
    one = allpairs_multi(desc, 50, 0.75, (0,))
    every = allpairs_multi(desc, 50, 0.75, tuple(range(torch.cuda.device_count())))
    assert np.array_equal(one, every) and one.sum() > 0
    with pytest.raises(orb.OrbError):
        allpairs_multi(desc, 50, 0.75, (torch.cuda.device_count(),))


def test_popc_peak_is_sane():
    v, clk = orb.popc_peak()
    print(f"POPC peak {v / 1e12:.3f} T/s at nominal {clk / 1e9:.3f} GHz -> {v / clk / 148:.2f} POPC/clk/SM")
    assert v > 1e11


def test_config3_extract_then_match_consecutive_keyframes():
    """BASELINE.json config 3: 1280x720, nFeatures=2000, 8 levels; top-2 + ratio matching between consecutive keyframes
    (the second frame is a shifted, re-noised view of the same scene, so true matches exist).  GPU extraction + GPU matching
    must reproduce the oracle's extraction + matching exactly."""
    from orbslam_mapsave_b200.synth import synth
    a = synth(1280, 720, 77)
    rng = np.random.default_rng(5)
    b = np.roll(a, (4, 7), axis=(0, 1))
    b = np.clip(b.astype(np.int16) + rng.integers(-2, 3, b.shape), 0, 255).astype(np.uint8)
    ex = orb.ORBextractor(2000, 1.2, 8, 20, 7, max_batch=2)
    kp, desc, n = ex.extract_batch(np.stack([a, b]))
    oex = orc.Extractor(2000, 1.2, 8, 20, 7)
    (okp_a, od_a), (okp_b, od_b) = oex.extract(a), oex.extract(b)
    assert n[0] == len(okp_a) and n[1] == len(okp_b)
    da, db = desc[0, :n[0]], desc[1, :n[1]]
    assert np.unpackbits(da ^ od_a).sum() <= 1e-4 * da.size * 8 and np.unpackbits(db ^ od_b).sum() <= 1e-4 * db.size * 8
    bi, bd, sd = orb.ORBmatcher().hamming_top2(da, db)
    obi, obd, osd = orc.hamming_top2(da, db)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(sd, osd)
    good = (bd <= 50) & (bd.astype(np.float32) < np.float32(0.75) * sd.astype(np.float32))
    # matched keypoints must agree with the known shift (x + 7, y + 4) at level-0 scale for the vast majority
    ka, kb = kp[0, :n[0]][good], kp[1, :n[1]][bi[good]]
    dx, dy = kb["x"] - ka["x"], kb["y"] - ka["y"]
    ok = (np.abs(dx - 7) <= 2.5 * ka["size"] / 31) & (np.abs(dy - 4) <= 2.5 * ka["size"] / 31)
    assert good.sum() > 300 and ok.mean() > 0.9, (int(good.sum()), float(ok.mean()))


def test_distinctive_descriptors_vs_oracle():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:483-548) for a ragged batch of map points."""
    rng = np.random.default_rng(8)
    sizes = [1, 2, 3, 0, 5, 17, 32, 33, 64, 100, 257, 7, 2, 1000] + [int(v) for v in rng.integers(2, 40, 200)]
    sets = []
    for i, n in enumerate(sizes):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.stack([_flip(base, int(rng.integers(0, 60)), rng) for _ in range(n)]) if n else np.zeros((0, 32), np.uint8)
        if n >= 4 and i % 3 == 0:
            d[2] = d[1]                      # exact duplicates: ties in the medians, first index must win
        sets.append(d)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    got = orb.distinctive_descriptors(np.concatenate(sets), offsets)
    want = np.array([orc.distinctive(d) for d in sets], np.int32)
    assert np.array_equal(got, want)
    with pytest.raises(orb.OrbError):
        orb.distinctive_descriptors(np.zeros((1025, 32), np.uint8), np.array([0, 1025], np.int32))


def test_distances_and_top2_match_cv2_golden():
    """The device path against real-OpenCV Hamming distances (tests/golden/prim_hamming.npz: cv2.norm / BFMatcher.knnMatch)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "prim_hamming.npz"))
    assert np.array_equal(orb.ORBmatcher.DescriptorDistance(g["a"], g["b"]), g["dist"])
    bi, b1, b2 = orb.ORBmatcher().hamming_top2(g["q"], g["db"])
    assert np.array_equal(b1, g["best"]) and np.array_equal(b2, g["second"])
    uniq = g["best"] < g["second"]
    assert np.array_equal(bi[uniq], g["best_idx"][uniq])


@pytest.mark.parametrize("kf_kf,ori,ratio", [(False, True, 0.75), (True, True, 0.75), (False, False, 0.6), (True, False, 0.9)])
def test_search_by_bow_batch_equals_single_calls(kf_kf, ori, ratio):
    """orbm_search_by_bow_batch (one frame / keyframe against N candidates, two launches): identical to the oracle's single searches
    in a loop — candidates of different sizes, an empty one, one without common nodes."""
    rng = np.random.default_rng(5)
    sizes = [(1500, 1400), (1500, 300), (1500, 2000), (1500, 0), (1500, 50), (1500, 1500), (1500, 900)]
    base = _scene(1500, 1500, 100)
    anchor_d, anchor_node, anchor_ang, anchor_flag = base["d1"], base["node1"], base["ang1"], base["flag1"]
    others, want = [], []
    for i, (_, n2) in enumerate(sizes):
        s = _scene(1500, n2, 200 + i)
        # candidates share descriptors with the anchor so that real matches exist
        d2, node2 = s["d2"].copy(), s["node2"].copy()
        k = min(n2, 1500) // 2
        src = rng.choice(1500, k, replace=False) if k else np.zeros(0, int)
        for t, sidx in enumerate(src):
            d2[t] = _flip(anchor_d[sidx], int(rng.choice([0, 3, 20, 49, 50, 51])), rng)
            node2[t] = anchor_node[sidx] if i != 4 else anchor_node[sidx] + 1000        # candidate 4: no common nodes
        others.append((d2, node2, s["ang2"], s["flag2"]))
    A = orb.View(anchor_d, orb.FeatureVector(anchor_node), anchor_ang, flag=anchor_flag)
    O = [orb.View(d, orb.FeatureVector(nd), a, flag=f) for d, nd, a, f in others]
    nm, m = orb.ORBmatcher(ratio, ori).SearchByBoWBatch(A, O, kf_kf=kf_kf)
    fa = orc.FeatVec(anchor_node)
    total = 0
    for i, (d, nd, a, f) in enumerate(others):
        fo = orc.FeatVec(nd)
        if kf_kf:
            on, om = orc.search_bow_kf_kf(anchor_d, anchor_flag, anchor_ang, fa, d, f, a, fo, ratio, ori)
        else:
            on, om = orc.search_bow_kf_f(d, f, a, fo, anchor_d, anchor_ang, fa, ratio, ori)
        assert nm[i] == on and np.array_equal(m[i], om), i
        total += on
    assert total > 200


@pytest.mark.parametrize("only_stereo,ori", [(False, False), (True, False), (False, True)])
def test_search_for_triangulation_batch_equals_single_calls(only_stereo, ori):
    rng = np.random.default_rng(9)
    sf2 = (1.2 ** np.arange(8)).astype(np.float32)
    n1 = 1200
    base = _scene(n1, n1, 300, tri=True)
    views, args = [], []
    for i, n2 in enumerate([1100, 0, 1300, 64, 900]):
        s = _scene(n1, n2, 310 + i, tri=True)
        d2, node2 = s["d2"].copy(), s["node2"].copy()
        k = min(n2, n1) // 2
        src = rng.choice(n1, k, replace=False) if k else np.zeros(0, int)
        for t, sidx in enumerate(src):
            d2[t] = _flip(base["d1"][sidx], int(rng.choice([0, 3, 20, 49, 50, 51])), rng)
            node2[t] = base["node1"][sidx]
        F12 = (rng.normal(0, 1, (3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-3], [1e-3, 1e-3, 1e-1]])).astype(np.float32)
        ep = rng.uniform(100, 500, 2).astype(np.float32)
        sig = (sf2 * sf2 * rng.uniform(2000, 8000)).astype(np.float32)
        views.append(orb.View(d2, orb.FeatureVector(node2), s["ang2"], flag=s["flag2"], x=s["x2"], y=s["y2"], octave=s["oct2"], uright=s["ur2"]))
        args.append((d2, s["flag2"], s["ur2"], s["x2"], s["y2"], s["ang2"], s["oct2"], node2, F12, ep, sig))
    K1 = orb.View(base["d1"], orb.FeatureVector(base["node1"]), base["ang1"], flag=base["flag1"], x=base["x1"], y=base["y1"],
                  octave=np.zeros(n1, np.int32), uright=base["ur1"])
    nm, pairs = orb.ORBmatcher(0.6, ori).SearchForTriangulationBatch(K1, views, [a[8] for a in args], [a[9] for a in args],
                                                                     [sf2] * len(args), [a[10] for a in args], only_stereo)
    total = 0
    for i, (d2, f2, ur2, x2, y2, a2, o2, node2, F12, ep, sig) in enumerate(args):
        on, op = orc.search_triangulation(base["d1"], base["flag1"], base["ur1"], base["x1"], base["y1"], base["ang1"], orc.FeatVec(base["node1"]),
                                          d2, f2, ur2, x2, y2, a2, o2, orc.FeatVec(node2), F12, float(ep[0]), float(ep[1]), sf2, sig,
                                          only_stereo, ori)
        assert len(pairs[i]) == len(op) and np.array_equal(pairs[i], op), i
        total += len(op)
    assert only_stereo or total > 20
