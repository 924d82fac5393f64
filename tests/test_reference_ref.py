"""The oracle against the REFERENCE ITSELF: oracle/_ref holds the reference's own src/ORBextractor.cc, compiled unmodified against
a stand-in for the slice of the OpenCV API it uses (oracle/ref_shim/cvshim.hpp; the image primitives behind that API are the
oracle's restatements, pinned to cv2 4.13 by tests/golden/prim_*.npz).  CPU-only; skipped when oracle/_ref has not been built
(it is built by `make -C oracle` wherever /root/reference exists and travels with the repo snapshot)."""
import glob
import os

import numpy as np
import pytest

from oracle import orb_oracle_py as orc
from oracle import ref_py
from orbslam_mapsave_b200.synth import synth

pytestmark = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built (needs the reference checkout at build time)")
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "chain_*.npz")))


def _case(path):
    g = np.load(path)
    W, H, seed, nf, nl, ini, mn, use_mask = (int(v) for v in g["params"])
    img = g["image"] if "image" in g.files else synth(W, H, seed)
    mask = None
    if use_mask:
        mask = np.full(img.shape, 255, np.uint8)
        mask[60:220, 150:260] = 0
    return g, img, mask, (nf, float(g["scaleFactor"]), nl, ini, mn)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[6:-4])
def test_reference_with_monotonic_heap_equals_golden_bit_for_bit(path):
    """With a heap that never reuses an address the reference's pointer tie-break (src/ORBextractor.cc:683) is "created later
    first" — the canonical rule of the oracle and the GPU path — and its output equals the cv2-chain fixtures (which the oracle and
    the GPU path reproduce) in everything: order, coordinates, size, angle bits, response, octave, class_id, descriptors."""
    g, img, mask, (nf, sf, nl, ini, mn) = _case(path)
    kp, desc = ref_py.extract_monotonic_heap(img, nf, sf, nl, ini, mn, mask)
    want = g["kp"]
    assert len(kp) == len(want)
    for i, f in enumerate(ref_py.KP_FIELDS):
        assert np.array_equal(kp[:, i].view(np.uint32), want[f].astype(np.float32).view(np.uint32)), f
    assert np.array_equal(desc, g["desc"])


@pytest.mark.parametrize("seed", [0, 7])
def test_reference_with_stock_heap_differs_only_by_the_pointer_tie_break(seed):
    """Stock allocator: which of several equal-size octree nodes is expanded first depends on heap addresses, so a few keypoints per
    frame differ (and differ from call to call).  Everything the two runs have in common must be identical down to the descriptor."""
    img = synth(640, 480, seed)
    okp, odesc = orc.Extractor(1000, 1.2, 8, 20, 7).extract(img)
    rkp, rdesc = ref_py.RefExtractor(1000, 1.2, 8, 20, 7).extract(img)
    ref = {(float(a[0]), float(a[1]), int(a[5])): (a[2:5].tobytes(), d.tobytes()) for a, d in zip(rkp, rdesc)}
    ora = {(float(x), float(y), int(o)): (np.array([s, a, r], np.float32).tobytes(), d.tobytes())
           for x, y, o, s, a, r, d in zip(okp["x"], okp["y"], okp["octave"], okp["size"], okp["angle"], okp["response"], odesc)}
    common = set(ref) & set(ora)
    assert len(common) >= 0.97 * len(ora) and abs(len(ref) - len(ora)) <= 16
    assert all(ref[k] == ora[k] for k in common)


# ---- matcher: the reference's own src/ORBmatcher.cc (compiled unmodified over plain-data stand-ins for KeyFrame / Frame / MapPoint)
matcher = pytest.mark.skipif(not ref_py.matcher_available(), reason="oracle/_ref/libref_orbmatcher.so not built")


def _scene(*a, **k):
    import test_gpu_match as tgm            # the scene generator of the GPU parity tests (planted matches, TH_LOW edges, ties, claims)
    return tgm._scene(*a, **k)


@matcher
@pytest.mark.parametrize("n1,n2,seed", [(300, 280, 0), (2000, 2000, 1), (1000, 40, 2), (50, 900, 3), (0, 10, 4), (10, 0, 5)])
@pytest.mark.parametrize("ratio,ori", [(0.7, True), (0.75, True), (0.6, False), (1.0, True)])
def test_reference_search_by_bow_kf_frame_equals_oracle(n1, n2, seed, ratio, ori):
    s = _scene(n1, n2, seed)
    f1, f2 = orc.FeatVec(s["node1"]), orc.FeatVec(s["node2"])
    on, om = orc.search_bow_kf_f(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["ang2"], f2, ratio, ori)
    rn, rm = ref_py.ref_search_bow_kf_f(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["ang2"], f2, ratio, ori)
    assert rn == on and np.array_equal(rm, om)


@matcher
@pytest.mark.parametrize("ratio", [0.5, 0.75, 0.6, 1.0])
def test_reference_ratio_test_equality_edge(ratio):
    """Constructed equalities of `float(best) < nnratio * float(second)` and of `best <= TH_LOW` (src/ORBmatcher.cc:231-233): the
    reference's own code, the oracle and the definition agree case by case."""
    import test_gpu_match as tgm
    s = tgm._ratio_edge_scene()
    f1, f2 = orc.FeatVec(s["node1"]), orc.FeatVec(s["node2"])
    on, om = orc.search_bow_kf_f(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["ang2"], f2, ratio, False)
    rn, rm = ref_py.ref_search_bow_kf_f(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["ang2"], f2, ratio, False)
    assert rn == on and np.array_equal(rm, om)
    hits = 0
    for i, (b, s2) in enumerate(s["cases"]):
        lo, hi = min(b, s2), max(b, s2)
        want = lo <= 50 and np.float32(lo) < np.float32(ratio) * np.float32(hi)
        assert (rm[2 * i] == i or rm[2 * i + 1] == i) == bool(want), (i, b, s2, ratio)
        hits += bool(want)
    assert 0 < hits < len(s["cases"]) or ratio == 1.0


@matcher
@pytest.mark.parametrize("n1,n2,seed", [(300, 280, 10), (2000, 2000, 11), (700, 64, 12), (0, 0, 13)])
@pytest.mark.parametrize("ratio,ori", [(0.75, True), (0.6, False)])
def test_reference_search_by_bow_kf_kf_equals_oracle(n1, n2, seed, ratio, ori):
    s = _scene(n1, n2, seed)
    f1, f2 = orc.FeatVec(s["node1"]), orc.FeatVec(s["node2"])
    on, om = orc.search_bow_kf_kf(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["flag2"], s["ang2"], f2, ratio, ori)
    rn, rm = ref_py.ref_search_bow_kf_kf(s["d1"], s["flag1"], s["ang1"], f1, s["d2"], s["flag2"], s["ang2"], f2, ratio, ori)
    assert rn == on and np.array_equal(rm, om)


@matcher
@pytest.mark.parametrize("n1,n2,seed", [(300, 280, 20), (2000, 2000, 21), (64, 900, 22), (5, 0, 23)])
@pytest.mark.parametrize("only_stereo,ori", [(False, False), (True, False), (False, True)])
def test_reference_search_for_triangulation_equals_oracle(n1, n2, seed, only_stereo, ori):
    s = _scene(n1, n2, seed, tri=True)
    rng = np.random.default_rng(seed)
    F12 = (rng.normal(0, 1, (3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-3], [1e-3, 1e-3, 1e-1]])).astype(np.float32)
    sf2 = (1.2 ** np.arange(8)).astype(np.float32)
    sig2 = (sf2 * sf2 * 5000).astype(np.float32)
    args = (s["d1"], s["flag1"], s["ur1"], s["x1"], s["y1"], s["ang1"], orc.FeatVec(s["node1"]),
            s["d2"], s["flag2"], s["ur2"], s["x2"], s["y2"], s["ang2"], s["oct2"], orc.FeatVec(s["node2"]),
            F12, 320.0, 240.0, sf2, sig2, only_stereo, ori)
    on, op = orc.search_triangulation(*args)
    rn, rp = ref_py.ref_search_triangulation(*args)
    assert len(rp) == len(op) and np.array_equal(rp, op)
    if n1 >= 2000 and not only_stereo:
        assert len(op) >= 10


@matcher
def test_reference_descriptor_distance_and_three_maxima_equal_oracle():
    rng = np.random.default_rng(3)
    for _ in range(500):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert ref_py.ref_descriptor_distance(a, b) == orc.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
    cases = [rng.integers(0, 50, 30) for _ in range(100)] + [np.zeros(30, int), np.full(30, 4), np.eye(30, dtype=int)[7] * 9,
                                                             np.array([100, 9, 10, 11] + [0] * 26)]
    for h in cases:
        assert ref_py.ref_three_maxima(h) == tuple(orc.three_maxima(h))


# ---- the tracker's window searches through the reference's own code (Frame::GetFeaturesInArea is a stand-in restated from
# src/Frame.cc:445-498: that translation unit cannot be compiled here)
def _grid(fa, blocked):
    import proj_util as pu
    return orc.Grid(fa["desc"], fa["x"], fa["y"], fa["octave"], pu.SCALE, fa["bounds"], angle=fa["angle"], uright=fa["uright"], blocked=blocked)


@matcher
@pytest.mark.parametrize("n,npts,seed,cluster,stereo,th", [
    (2000, 3000, 10, False, True, 1.0), (2000, 3000, 11, True, True, 3.0), (1500, 800, 12, False, False, 5.0),
    (300, 4000, 13, True, True, 3.0), (1, 5, 14, False, True, 1.0), (50, 0, 15, False, True, 1.0), (8000, 9000, 16, False, True, 3.0)])
def test_reference_search_by_projection_map_points_equals_oracle(n, npts, seed, cluster, stereo, th):
    import proj_util as pu
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=stereo, cluster=cluster)
    og = _grid(fa, (rng.random(n) < 0.15).astype(np.uint8))
    mp = pu.map_points_for(fa, npts, rng)
    wn, want = orc.search_projection_map(og, th=th, nnratio=0.8, **mp)
    rn, got = ref_py.ref_search_projection_map(og, th=th, nnratio=0.8, **mp)
    assert rn == wn and np.array_equal(got, want)


@matcher
@pytest.mark.parametrize("n,nlast,seed,mono,tz,ori,cluster", [
    (2000, 2000, 30, True, 0.0, True, False), (2000, 2000, 31, False, 0.5, True, False), (2000, 2000, 32, False, -0.5, True, True),
    (1200, 3000, 33, False, 0.0, False, True), (2000, 0, 34, True, 0.0, True, False), (3, 10, 35, False, 0.0, True, False),
    (8000, 8000, 36, False, 0.0, True, False)])
def test_reference_search_by_projection_last_frame_equals_oracle(n, nlast, seed, mono, tz, ori, cluster):
    import proj_util as pu
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=not mono, cluster=cluster)
    og = _grid(fa, (rng.random(n) < 0.1).astype(np.uint8))
    lf = pu.last_frame_for(fa, nlast, rng, tz=tz)
    mbf, mb, th = 40.0, 40.0 / lf["fx"], 15.0 if mono else 7.0
    args = (lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], mbf, mb, lf["has_point"], lf["world"], lf["octave"], lf["angle"],
            lf["desc"], lf["claims"], th, mono)
    wn, want = orc.search_projection_frame(og, *args, ori)
    rn, got = ref_py.ref_search_projection_frame(og, *args, ori)
    # the reference sets culled entries back to NULL; the oracle marks them -2
    assert rn == wn and np.array_equal(got, np.where(want == -2, -1, want))


@matcher
@pytest.mark.parametrize("n2,n1,seed,window,ori", [(2000, 2000, 40, 100, True), (2000, 2000, 41, 10, True), (1000, 3000, 42, 50, False),
                                                   (2000, 0, 43, 100, True), (2, 9, 44, 100, True), (6000, 6000, 45, 100, True)])
def test_reference_search_for_initialization_equals_oracle(n2, n1, seed, window, ori):
    import proj_util as pu
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n2, rng, stereo=False)
    fa["octave"][rng.random(n2) < 0.5] = 0
    og = _grid(fa, None)
    f1 = pu.init_frame1_for(fa, n1, rng)
    prev_o, prev_r = f1["prev"].copy(), f1["prev"].copy()
    wn, want = orc.search_initialization(og, f1["desc1"], f1["octave1"], f1["angle1"], prev_o, window, 0.9, ori)
    rn, got = ref_py.ref_search_initialization(og, f1["desc1"], f1["octave1"], f1["angle1"], prev_r, window, 0.9, ori)
    assert rn == wn and np.array_equal(got, want) and np.array_equal(prev_r, prev_o)


def _drop_predict_scale_ub(kp, Tn):
    """This fork's MapPoint::PredictScale is unclamped and its result indexes mvScaleFactors (src/MapPoint.cc:633-642): out-of-range
    levels are undefined behaviour in the reference and clamped in the oracle / product.  Points that would get there are taken out
    of the scene for both sides (margin 1e-3 of a level)."""
    R, t = Tn[:, :3].astype(np.float64), Tn[:, 3].astype(np.float64)
    dist = np.linalg.norm(kp["world"].astype(np.float64) - (-R.T @ t), axis=1)
    lv = np.log(kp["mf_max"].astype(np.float64) / dist) / np.log(1.2)
    st = kp["state"].copy()
    st[~((lv > -0.999) & (lv < 6.999))] = 0
    return st


LOG_SF = np.float32(np.log(np.float32(1.2)))


@matcher
@pytest.mark.parametrize("n,npts,seed,ori,orbdist,cluster", [(2000, 2500, 70, True, 100, False), (2000, 2500, 71, False, 64, True),
                                                            (500, 3000, 72, True, 100, True), (4000, 4000, 73, True, 64, False)])
def test_reference_relocalisation_projection_equals_oracle(n, npts, seed, ori, orbdist, cluster):
    """SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1465-1602."""
    import proj_util as pu
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=False, cluster=cluster)
    og = _grid(fa, (rng.random(n) < 0.2).astype(np.uint8))
    kp = pu.kf_points_for(fa, npts, rng)
    st = _drop_predict_scale_ub(kp, kp["T"])
    a = (kp["T"], kp["fx"], kp["fy"], kp["cx"], kp["cy"], LOG_SF)
    b = (kp["world"], kp["mf_max"], kp["mf_min"], kp["angle"], kp["desc"], 10.0, orbdist, ori)
    on, oo = orc.search_projection_kf(og, *a, st == 1, *b)
    rn, ro = ref_py.ref_search_projection_kf(og, *a, st, *b)
    assert rn == on and on > 100 and np.array_equal(ro, np.where(oo == -2, -1, oo))


@matcher
@pytest.mark.parametrize("n,npts,seed,scale", [(2000, 2500, 80, 1.0), (2000, 2500, 81, 2.5), (600, 3000, 82, 0.4), (4000, 4000, 83, 1.7)])
def test_reference_loop_closing_projection_equals_oracle(n, npts, seed, scale):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:293-406, with a similarity of scale `scale`."""
    import proj_util as pu
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=False)
    fa["x"] = np.clip(fa["x"], 0, 639.5).astype(np.float32)
    og = _grid(fa, (rng.random(n) < 0.3).astype(np.uint8))
    kq = pu.kf_points_for(fa, npts, rng, sim_scale=scale)
    S = kq["S"].astype(np.float64)
    st = _drop_predict_scale_ub(kq, S / np.sqrt((S[0, :3] ** 2).sum()))
    a = (kq["S"], kq["fx"], kq["fy"], kq["cx"], kq["cy"], LOG_SF)
    b = (kq["world"], kq["mf_max"], kq["mf_min"], kq["normal"], kq["desc"], 10)
    on, oo = orc.search_projection_sim3(og, *a, st == 1, *b)
    rn, ro = ref_py.ref_search_projection_sim3(og, *a, st, *b)
    assert rn == on and on > 100 and np.array_equal(ro, oo)


# ---- DBoW2: the reference's own vocabulary class (Thirdparty/DBoW2 compiled unmodified) against the oracle's transform ------------
dbow2 = pytest.mark.skipif(not ref_py.dbow2_available(), reason="oracle/_ref/libref_dbow2.so not built")


@dbow2
@pytest.mark.parametrize("k,L,ragged,levelsup,interleave", [(10, 3, False, 1, False), (10, 4, False, 4, False), (10, 4, False, 2, True),
                                                            (7, 5, True, 3, False), (20, 2, False, 0, False), (3, 6, True, 4, True),
                                                            (18, 3, True, 1, True), (10, 3, False, 7, False)])
def test_reference_dbow2_transform_equals_oracle(k, L, ragged, levelsup, interleave, tmp_path):
    """TemplatedVocabulary::transform (per feature :1231-1272 and features -> BowVector + FeatureVector :1140-1207), FORB::distance
    (FORB.cpp:81-101), loaded by the reference's own loadFromTextFile: word ids, weights (raw doubles), node ids, the normalised
    BowVector and the FeatureVector equal the oracle's on the trees of tests/test_gpu_vocab.py."""
    from vocab_util import make_tree, write_text
    from orbslam_mapsave_b200.synth import synth_descriptors
    parent, desc, weight, is_leaf = make_tree(k, L, seed=k * 10 + L, ragged=ragged, interleave=interleave)
    write_text(tmp_path / "voc.txt", k, L, parent, desc, weight, is_leaf)
    voc = ref_py.RefVocabulary(tmp_path / "voc.txt")
    assert voc.info() == dict(k=k, L=L, n_nodes=len(parent), n_words=int(is_leaf.sum()))
    feats = synth_descriptors(3000, 5, dup_of=desc[1:], dup_rate=0.7, max_flip=30)
    rw, rwt, rnid = voc.transform_raw(feats, levelsup)
    ow, owt, onid = orc.voc_transform(parent, desc, weight, is_leaf, L, feats, levelsup)
    assert np.array_equal(rw, ow) and np.array_equal(rwt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(rnid, onid)
    # One undefined behaviour of the reference is kept out of the FeatureVector comparison: when the descent ends in a leaf ABOVE level
    # L - levelsup (ragged trees only; never in a complete tree like ORBvoc) the per-feature transform leaves *nid unwritten, and
    # transform(features, ...) then files the feature under an uninitialised NodeId (:1164-1175).  Oracle and product define it as 0.
    depth = np.zeros(len(parent), np.int32)
    for i in range(1, len(parent)):
        depth[i] = depth[parent[i]] + 1
    defined = (depth[np.nonzero(is_leaf)[0][ow]] >= L - levelsup) | (L - levelsup <= 0)
    feats, ow, owt, onid = feats[defined], ow[defined], owt[defined], onid[defined]
    assert defined.mean() > 0.5
    (bid, bval), (fnode, foff, ffeat) = voc.transform(feats, levelsup)
    obid, obval = orc.voc_bow(ow, owt, 0, 0)
    assert np.array_equal(bid, obid) and np.array_equal(bval.view(np.uint64), obval.view(np.uint64))
    keep = owt > 0
    ofv = orc.FeatVec(onid[keep])
    assert np.array_equal(fnode, ofv.ids) and np.array_equal(foff, ofv.off) and np.array_equal(ffeat, np.nonzero(keep)[0][ofv.feat])


@dbow2
def test_reference_dbow2_binary_format_round_trip(tmp_path):
    """The fork's binary vocabulary (saveToBinaryFile :1515-1536 / loadFromBinaryFile :1467-1512, weights stored as float32): a file
    written by the reference loads in the reference with the same transform as the oracle on float32-rounded weights.  (Its loader reads
    one record past the end — `while(!f.eof())` — which leaves a duplicate of the last node behind; a duplicate sibling never wins
    "first child with the smallest distance", so results are unaffected.)"""
    from vocab_util import make_tree, write_text
    from orbslam_mapsave_b200.synth import synth_descriptors
    k, L = 10, 3
    parent, desc, weight, is_leaf = make_tree(k, L, seed=99)
    write_text(tmp_path / "voc.txt", k, L, parent, desc, weight, is_leaf)
    voc = ref_py.RefVocabulary(tmp_path / "voc.txt")
    voc.save_binary(tmp_path / "voc.bin")
    raw = (tmp_path / "voc.bin").read_bytes()
    assert len(raw) == 24 + (len(parent) - 1) * 41 and np.frombuffer(raw[:24], np.int32).tolist() == [len(parent), 41, k, L, 0, 0]
    voc2 = ref_py.RefVocabulary(tmp_path / "voc.bin", binary=True)
    feats = synth_descriptors(1500, 6, dup_of=desc[1:], dup_rate=0.6)
    got = voc2.transform_raw(feats, 2)
    want = orc.voc_transform(parent, desc, weight.astype(np.float32).astype(np.float64), is_leaf, L, feats, 2)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


# ---- Fuse x2 and SearchBySim3 through the reference's own code: the oracle's candidate searches (+ a replay of the reference's map
# bookkeeping from the per-point result) against the object graph the reference itself leaves behind
def _level_ok(dist, mf_max):
    """Points whose PredictScale leaves [0, 8) index mvScaleFactors out of range in this fork (undefined; clamped in oracle / product)."""
    lv = np.log(mf_max.astype(np.float64) / np.maximum(dist, 1e-12)) / np.log(1.2)
    return (lv > -0.999) & (lv < 6.999)


@matcher
@pytest.mark.parametrize("variant,seed,n,npts,cluster", [(0, 300, 1200, 1500, False), (1, 301, 1200, 1500, True), (0, 302, 400, 3000, True),
                                                         (1, 303, 3000, 800, False), (0, 304, 2000, 40, False)])
def test_reference_fuse_equals_oracle_search_plus_bookkeeping(variant, seed, n, npts, cluster):
    import proj_util as pu
    from test_gpu_host_cpp import _Pt, _emulate_fuse
    rng = np.random.default_rng(seed)
    log_sf = float(np.log(np.float32(1.2)))
    K = np.array([520.0, 520.0, 320.0, 240.0, 40.0, log_sf], np.float32)
    is2 = (1.0 / (pu.SCALE * pu.SCALE)).astype(np.float32)
    fa = pu.frame_arrays(n, rng, stereo=True, cluster=cluster)
    fa["x"] = np.clip(fa["x"], 0, 639.5).astype(np.float32)
    fa["y"] = np.clip(fa["y"], 0, 479.5).astype(np.float32)
    R = pu.rot_small(rng, 3.0).astype(np.float32)
    t = rng.uniform(-0.1, 0.1, 3).astype(np.float32)
    T = np.concatenate([R, t[:, None]], 1).astype(np.float32)
    Ow = (-(R.astype(np.float64).T @ t.astype(np.float64))).astype(np.float32)
    scale = 1.0 if variant == 0 else 1.7
    A, b = (T[:, :3] * np.float32(scale)).astype(np.float32), (T[:, 3] * np.float32(scale)).astype(np.float32)
    S = np.concatenate([np.concatenate([A, b[:, None]], 1), np.array([[0, 0, 0, 1]], np.float32)], 0).astype(np.float32)
    kstate = rng.choice(np.array([0, 0, 1, 1, 2], np.uint8), n)
    knobs = rng.integers(1, 6, n)
    kP = pu.points_for_transform(fa, n, rng, A, b, False)
    P = pu.points_for_transform(fa, npts, rng, A, b, False)
    cstate = (rng.choice(np.array([0, 1, 1, 1, 1, 2, 3], np.uint8), npts) if variant == 0 else
              rng.choice(np.array([1, 1, 1, 1, 2, 3], np.uint8), npts))
    dist = np.linalg.norm(P["world"].astype(np.float64) - Ow.astype(np.float64), axis=1) if npts else np.zeros(0)
    cstate[(cstate == 1) & ~_level_ok(dist, P["mf_max"])] = 2
    cnobs = rng.integers(0, 6, npts)
    free = np.nonzero(kstate == 0)[0]
    in_at = np.full(npts, -1, np.int32)
    three = np.nonzero(cstate == 3)[0][:len(free)]
    cstate[np.setdiff1d(np.nonzero(cstate == 3)[0], three)] = 2
    in_at[three] = rng.choice(free, len(three), replace=False)
    th = 3.0 if variant == 0 else 4.0
    og = orc.Grid(fa["desc"], fa["x"], fa["y"], fa["octave"], pu.SCALE, fa["bounds"], angle=fa["angle"], uright=fa["uright"])
    skip = (cstate != 1).astype(np.uint8)
    best = orc.fuse_search(og, variant, T if variant == 0 else S[:3], Ow if variant == 0 else None, K[0], K[1], K[2], K[3], K[4],
                           np.float32(log_sf), skip, P["world"], P["mf_max"], P["mf_min"], P["normal"], P["desc"], th, is2)
    kfp = [_Pt(kstate[j] == 2, knobs[j]) for j in range(n)]
    cand = [_Pt(cstate[i] == 2, cnobs[i]) for i in range(npts)]
    kf_ptr = [None] * n
    for j in range(n):
        if kstate[j]:
            kf_ptr[j] = kfp[j]
            kfp[j].obs["kf"] = j
    for i in three:
        cand[i].obs["kf"] = int(in_at[i])
        kf_ptr[in_at[i]] = cand[i]
    nf, rep = _emulate_fuse(variant, fa["uright"], kf_ptr, kfp, cand, cstate, best)
    idx = {id(p): j for j, p in enumerate(kfp)}
    idx.update({id(p): 100000 + i for i, p in enumerate(cand)})
    enc = lambda p: -1 if p is None else idx[id(p)]
    rn, rkf, rpts, rrep = ref_py.ref_fuse(variant, og, K, is2, T, Ow, S, kstate, knobs, kP, cstate, cnobs, P, in_at, th)
    assert rn == nf and (npts < 100 or nf > 50)
    assert np.array_equal(rkf, np.array([enc(p) for p in kf_ptr], np.int32))
    assert np.array_equal(rpts, np.array([(int(p.bad), p.nobs) for p in kfp + cand], np.int32).reshape(-1, 2))
    if variant == 1:
        assert np.array_equal(rrep, np.array([enc(p) for p in rep], np.int32))


@matcher
@pytest.mark.parametrize("seed,n,s12,th", [(310, 1200, 1.15, 7.5), (311, 2500, 0.8, 7.5), (312, 300, 1.6, 10.0)])
def test_reference_search_by_sim3_equals_oracle(seed, n, s12, th):
    import proj_util as pu
    rng = np.random.default_rng(seed)
    log_sf = float(np.log(np.float32(1.2)))
    K = np.array([520.0, 520.0, 320.0, 240.0, 40.0, log_sf], np.float32)

    def frame():
        fa = pu.frame_arrays(n, rng, stereo=True)
        fa["x"] = np.clip(fa["x"], 0, 639.5).astype(np.float32)
        fa["y"] = np.clip(fa["y"], 0, 479.5).astype(np.float32)
        return fa
    fa1, fa2 = frame(), frame()
    Ra, ta = pu.rot_small(rng, 3.0).astype(np.float32), rng.uniform(-0.1, 0.1, 3).astype(np.float32)
    Rb, tb = pu.rot_small(rng, 3.0).astype(np.float32), rng.uniform(-0.1, 0.1, 3).astype(np.float32)
    s12 = np.float32(s12)
    R12 = (Ra.astype(np.float64) @ Rb.astype(np.float64).T).astype(np.float32)
    t12 = (ta.astype(np.float64) - float(s12) * (R12.astype(np.float64) @ tb.astype(np.float64))).astype(np.float32)
    # the reference's own arithmetic for sR12, sR21, t21 (:1121-1123): float scaling like cv::Mat::convertTo, `-sR21*t12` accumulated in double
    sR12 = (R12 * s12).astype(np.float32)
    sR21 = (R12.T * np.float32(1.0 / float(s12))).astype(np.float32)
    t21 = (-(sR21.astype(np.float64) @ t12.astype(np.float64))).astype(np.float32)
    perm = rng.permutation(n)
    inv = np.argsort(perm)
    tgt2 = np.where(rng.random(n) < 0.15, rng.integers(0, n, n), inv)
    P1 = pu.points_for_transform(fa2, n, rng, (Rb / s12).astype(np.float32), tb, True, tgt=perm)
    P2 = pu.points_for_transform(fa1, n, rng, (Ra * s12).astype(np.float32), ta, True, tgt=tgt2)
    st1 = rng.choice(np.array([0, 1, 1, 1, 2], np.uint8), n)
    st2 = rng.choice(np.array([0, 1, 1, 1, 2], np.uint8), n)
    c1 = (sR21.astype(np.float64) @ (Ra.astype(np.float64) @ P1["world"].astype(np.float64).T + ta.astype(np.float64)[:, None])).T + t21.astype(np.float64)
    c2 = (sR12.astype(np.float64) @ (Rb.astype(np.float64) @ P2["world"].astype(np.float64).T + tb.astype(np.float64)[:, None])).T + t12.astype(np.float64)
    st1[(st1 == 1) & ~_level_ok(np.linalg.norm(c1, axis=1), P1["mf_max"])] = 2
    st2[(st2 == 1) & ~_level_ok(np.linalg.norm(c2, axis=1), P2["mf_max"])] = 2
    pre = np.where((rng.random(n) < 0.1) & (st1 == 1), rng.integers(0, n, n), -1).astype(np.int32)
    pre[(pre >= 0) & (st2[np.maximum(pre, 0)] == 0)] = -1
    og1 = orc.Grid(fa1["desc"], fa1["x"], fa1["y"], fa1["octave"], pu.SCALE, fa1["bounds"], angle=fa1["angle"], uright=fa1["uright"])
    og2 = orc.Grid(fa2["desc"], fa2["x"], fa2["y"], fa2["octave"], pu.SCALE, fa2["bounds"], angle=fa2["angle"], uright=fa2["uright"])
    already1 = pre >= 0
    already2 = np.zeros(n, bool)
    already2[pre[pre >= 0]] = True
    m1 = orc.sim3_direction(og2, Ra, ta, sR21, t21, K[0], K[1], K[2], K[3], np.float32(log_sf), (st1 == 1) & ~already1, P1["world"], P1["mf_max"],
                            P1["mf_min"], P1["desc"], th)
    m2 = orc.sim3_direction(og1, Rb, tb, sR12, t12, K[0], K[1], K[2], K[3], np.float32(log_sf), (st2 == 1) & ~already2, P2["world"], P2["mf_max"],
                            P2["mf_min"], P2["desc"], th)
    want12, nfound = pre.copy(), 0
    for i1 in range(n):
        if m1[i1] >= 0 and m2[m1[i1]] == i1:
            want12[i1] = m1[i1]
            nfound += 1
    Tz = lambda R, t: np.concatenate([R, t[:, None]], 1).astype(np.float32)
    rn, r12 = ref_py.ref_search_by_sim3(og1, og2, K, Tz(Ra, ta), Tz(Rb, tb), st1, P1, st2, P2, s12, R12, t12, th, pre)
    assert rn == nfound and nfound > 20 and np.array_equal(r12, want12)


# ---- MapPoint: the reference's own src/MapPoint.cc + include/MapPoint.h compiled unmodified (stand-ins for KeyFrame / Frame / Map,
# recording stand-ins for Boost's archives; ORBmatcher::DescriptorDistance forwarded to the reference's own in libref_orbmatcher.so)
mappoint = pytest.mark.skipif(not ref_py.mappoint_available(), reason="oracle/_ref/libref_mappoint.so not built")


@mappoint
def test_reference_compute_distinctive_descriptors_equals_oracle():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:483-548): least median Hamming distance to the other observations,
    median = sorted row[int(0.5 * (N - 1))], first index wins ties, observations in bad keyframes left out — the oracle's choice on
    the ragged sets of tests/test_gpu_match.py::test_distinctive_descriptors_vs_oracle (sizes 1 / 2 / 32 / 33 / 257 / 1000, duplicates)."""
    from orbslam_mapsave_b200.synth import synth_descriptors
    rng = np.random.default_rng(17)
    sizes = [1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 64, 100, 257, 1000] + [int(v) for v in rng.integers(1, 40, 120)]
    for k, n in enumerate(sizes):
        base = synth_descriptors(1, 1000 + k)
        d = synth_descriptors(n, 2000 + k, dup_of=base, dup_rate=0.8, max_flip=60)
        if k % 5 == 0 and n > 3:
            d[n // 2] = d[0]                                   # exact duplicates: equal medians, first index wins
            d[n - 1] = d[1]
        bad = (rng.random(n) < (0.3 if k % 3 == 0 else 0.0)).astype(np.uint8)
        good = np.nonzero(bad == 0)[0]
        ri, rdesc = ref_py.ref_distinctive(d, bad)
        if len(good) == 0:
            assert ri == -1
            continue
        oi = orc.distinctive(d[good])
        assert np.array_equal(rdesc, d[good][oi]), (k, n, ri, oi)
    assert ref_py.ref_distinctive(synth_descriptors(5, 1), point_bad=True)[0] == -1        # a bad point keeps its descriptor (:491-492)


@mappoint
def test_reference_mappoint_save_field_sequence_matches_the_archive_layout():
    """The order and width of what MapPoint::save (src/MapPoint.cc:58-140) hands to the archive, recorded from the reference's own
    code, equals the field list the map-archive reader / writer (orbslam_mapsave_b200/csrc/orb_map.cpp) is built on — stated here
    as the tag sequence.  (This pins field order and widths, not Boost's class-info framing.)"""
    mat = lambda nbytes: ["Mat{", "i4", "i4", "u8", "u8", f"a{nbytes}", "}"]
    for n_obs, has_ref in [(0, True), (3, True), (5, False)]:
        raw, tags = ref_py.ref_mappoint_save_fields(n_obs, has_ref)
        want = (["u8", "u8", "i8", "i8", "i4", "f4", "f4", "f4", "b1", "i4", "f4"] + ["u8"] * 7 + mat(0) + ["u8"] + mat(12) + ["u4"] +
                ["b1", "u8", "u8"] * n_obs + mat(12) + mat(32 if n_obs else 0) + (["b1", "u8"] if has_ref else ["b1"]) +
                ["i4", "i4", "b1", "f4", "f4"])
        assert tags == want, (n_obs, has_ref, tags)
        # spot values: mnId = 41 (nNextId before construction), nNextId = 42 afterwards, mnFirstKFid = mnFirstFrame = 100
        assert np.frombuffer(raw[:32], np.int64).tolist() == [41, 42, 100, 100]
        rawt, tagst = ref_py.ref_mappoint_save_fields(n_obs, has_ref, track=True)
        assert tagst == tags and np.frombuffer(rawt[36:48], np.float32).tolist() == [11.0, 12.0, 13.0]
        if n_obs == 0:
            continue                       # (the builder always stores a 1 x 32 descriptor; a point without observations has none)
        # the same logical point through the product's map builder: its framing-free record equals the reference's bytes
        import ctypes as C
        from orbslam_mapsave_b200 import capi
        L = capi.lib()
        h = C.c_void_p()
        capi.check(L.orbmap_create(C.byref(h)))
        # MapPoint::nNextId is a static every record carries: 42 here, so a point with id 41 must be the newest in the map
        desc = np.frombuffer(raw, np.uint8)[-(32 + (9 if has_ref else 1) + 17):][:32].copy()
        kf_ids = np.array([100 + 7 * i for i in range(n_obs)], np.int64)
        feat = np.array([i % 4 for i in range(n_obs)], np.int64)
        pos, nrm = np.array([1.5, -2.25, 8.0], np.float32), np.zeros(3, np.float32)
        capi.check(L.orbmap_add_mappoint(h, 41, 100, capi._p(pos), capi._p(nrm), capi._p(desc), 100 if has_ref else -1, n_obs, capi._p(kf_ids),
                                         capi._p(feat), 1, 1, 0.0, 0.0))
        nb = C.c_int64()
        buf = np.zeros(4096, np.uint8)
        capi.check(L.orbmap_mappoint_record(h, 0, capi._p(buf), len(buf), C.byref(nb)))
        L.orbmap_destroy(h)
        assert buf[:nb.value].tobytes() == raw


# ---- Frame::ComputeStereoMatches: the reference's own src/Frame.cc compiled unmodified (the extractors it reads are stand-ins holding
# the pyramids; DescriptorDistance is the reference's own)
frame = pytest.mark.skipif(not ref_py.frame_available(), reason="oracle/_ref/libref_frame.so not built")


@frame
@pytest.mark.parametrize("W,H,nf,nl,seed,mbf", [(640, 480, 1000, 8, 0, 40.0), (752, 480, 1200, 8, 1, 47.9), (1280, 720, 2000, 8, 2, 386.0),
                                              (424, 240, 600, 6, 3, 20.0)])
def test_reference_compute_stereo_matches_equals_oracle(W, H, nf, nl, seed, mbf):
    """mvuRight / mvDepth of the reference's Frame::ComputeStereoMatches (src/Frame.cc:584-756) equal the oracle's bit for bit on the
    stereo pairs of tests/test_gpu_stereo.py (keypoints of the extractor stay >= 19 px inside every level, so none of the reference's
    unchecked accesses — row table, 11x11 windows — leaves the image)."""
    from test_gpu_stereo import _stereo_pair
    left, right = _stereo_pair(W, H, seed)
    oL, oR = orc.Extractor(nf, 1.2, nl, 20, 7), orc.Extractor(nf, 1.2, nl, 20, 7)
    okL, odL = oL.extract(left)
    okR, odR = oR.extract(right)
    t = oL.tables()
    mb = mbf / 500.0
    lv = ([oL.level(l) for l in range(nl)], [oR.level(l) for l in range(nl)])
    want_u, want_d = orc.stereo_matches(lv[0], lv[1], t["scale"], t["inv_scale"], okL, odL, okR, odR, mbf, mb)
    got_u, got_d = ref_py.ref_stereo_matches(lv[0], lv[1], t["scale"], t["inv_scale"], okL, odL, okR, odR, mbf, mb)
    assert (want_u >= 0).sum() > len(okL) // 4
    assert np.array_equal(got_u.view(np.uint32), want_u.view(np.uint32)) and np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))


@frame
@pytest.mark.parametrize("n,seed,cluster", [(2000, 50, False), (6000, 51, True), (3, 52, False)])
def test_reference_feature_grid_and_features_in_area_equal_oracle(n, seed, cluster):
    """Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:341-356, 500-510: `round`) and Frame::GetFeaturesInArea (:445-498) of the
    reference's own Frame.cc: the grid equals the CSR the oracle / product are handed, and the candidate lists (order included) equal the
    oracle's restatement that every window search starts from."""
    import proj_util as pu
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=False, cluster=cluster)
    og = orc.Grid(fa["desc"], fa["x"], fa["y"], fa["octave"], pu.SCALE, fa["bounds"], angle=fa["angle"])
    nq = 300
    q = np.stack([rng.uniform(-30, 670, nq), rng.uniform(-30, 510, nq), rng.choice([1.0, 3.0, 7.5, 15.0, 40.0, 100.0], nq)], 1).astype(np.float32)
    lev = np.stack([rng.integers(-1, 6, nq), rng.integers(-1, 8, nq)], 1).astype(np.int32)
    off, feat, lists = ref_py.ref_grid_and_areas(fa["x"], fa["y"], fa["octave"], fa["bounds"], q, lev)
    assert np.array_equal(off, og.off) and np.array_equal(feat, og.feat[:len(feat)]) and len(feat) == og.off[-1]
    total = 0
    for k in range(nq):
        want = orc.features_in_area(og, q[k, 0], q[k, 1], q[k, 2], lev[k, 0], lev[k, 1])
        assert np.array_equal(lists[k], want), k
        total += len(want)
    assert n < 100 or total > 1000
