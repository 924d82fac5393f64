"""GPU parity tests of the C++ drop-in classes (orbslam_mapsave_b200/host: ORB_SLAM2::ORBextractor / ORBmatcher with the
reference's own signatures) against the CPU oracle.  The driver tests/cpp/host_api_driver is built by
__graft_entry__.build() and calls the classes the way Frame::ExtractORB / LocalMapping do in the reference."""
import os
import struct
import subprocess

import numpy as np
import pytest

from orbslam_mapsave_b200 import KP_DTYPE
from orbslam_mapsave_b200.synth import synth
from oracle import orb_oracle_py as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "host_api_driver")


def _driver():
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "orbslam_mapsave_b200", "host")])
    return DRIVER


@pytest.mark.parametrize("W,H,seed,nf,nl,sf,ini,mn,masked", [(640, 480, 40, 1000, 8, 1.2, 20, 7, False),
                                                              (752, 480, 41, 1200, 8, 1.2, 20, 7, True)])
def test_cpp_orbextractor_matches_oracle(tmp_path, W, H, seed, nf, nl, sf, ini, mn, masked):
    img = synth(W, H, seed)
    (tmp_path / "in.raw").write_bytes(img.tobytes())
    args = [_driver(), "extract", str(tmp_path / "in.raw"), str(W), str(H), str(nf), str(sf), str(nl), str(ini), str(mn),
            str(tmp_path / "out")]
    mask = None
    if masked:
        mask = np.full((H, W), 255, np.uint8)
        mask[100:300, 300:500] = 0
        (tmp_path / "mask.raw").write_bytes(mask.tobytes())
        args.append(str(tmp_path / "mask.raw"))
    subprocess.check_call(args)
    kp = np.frombuffer((tmp_path / "out.kp").read_bytes(), KP_DTYPE)
    desc = np.frombuffer((tmp_path / "out.desc").read_bytes(), np.uint8).reshape(-1, 32)
    oex = orc.Extractor(nf, sf, nl, ini, mn)
    okp, odesc = oex.extract(img, mask)
    assert len(kp) == len(okp) == len(desc)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kp[f], okp[f]), f
    assert np.abs(kp["angle"].astype(np.float64) - okp["angle"]).max() <= 1e-3
    assert np.unpackbits(desc ^ odesc).sum() <= 1e-4 * desc.size * 8
    # getters and the public bordered pyramid
    meta = (tmp_path / "out.meta").read_bytes()
    nlev, = struct.unpack_from("<i", meta, 0)
    sfac, = struct.unpack_from("<f", meta, 4)
    assert nlev == nl and sfac == np.float32(sf)
    tabs = np.frombuffer(meta, np.float32, 4 * nl, 8).reshape(4, nl)
    t = oex.tables()
    for got, key in zip(tabs, ("scale", "inv_scale", "sigma2", "inv_sigma2")):
        assert np.array_equal(got, t[key]), key
    pos = 8 + 16 * nl
    for l in range(nl):
        w, h = struct.unpack_from("<ii", meta, pos)
        pos += 8
        got = np.frombuffer(meta, np.uint8, (w + 38) * (h + 38), pos).reshape(h + 38, w + 38)
        pos += (w + 38) * (h + 38)
        assert np.array_equal(got, oex.level(l, bordered=True)), f"bordered pyramid level {l}"


def _flip(d, nbits, rng):
    d = d.copy()
    for b in rng.choice(256, nbits, replace=False):
        d[b >> 3] ^= np.uint8(1 << (b & 7))
    return d


@pytest.mark.parametrize("seed,ratio,ori,only_stereo", [(0, 0.7, 1, 0), (1, 0.6, 0, 0), (2, 0.75, 1, 1)])
def test_cpp_orbmatcher_matches_oracle(tmp_path, seed, ratio, ori, only_stereo):
    rng = np.random.default_rng(seed)
    n = [900, 1100]
    desc = [rng.integers(0, 256, (n[0], 32), dtype=np.uint8), rng.integers(0, 256, (n[1], 32), dtype=np.uint8)]
    node = [rng.integers(0, 60, n[0]) * 2 + 1, rng.integers(0, 66, n[1]) * 2 + 1]
    k = 400
    src, dst = rng.choice(n[0], k, replace=False), rng.choice(n[1], k, replace=False)
    for s, t in zip(src, dst):
        desc[1][t] = _flip(desc[0][s], int(rng.choice([0, 2, 10, 30, 49, 50, 51])), rng)
        node[1][t] = node[0][s]
    state = [rng.choice([0, 1, 1, 2], n[i]).astype(np.uint8) for i in range(2)]
    ang = [rng.uniform(0, 360, n[i]).astype(np.float32) for i in range(2)]
    ang[1][dst] = np.mod(ang[0][src] + rng.choice([0, 1, 14, 44], k), 360).astype(np.float32)
    x = [rng.uniform(0, 640, n[i]).astype(np.float32) for i in range(2)]
    y = [rng.uniform(0, 480, n[i]).astype(np.float32) for i in range(2)]
    ur = [np.where(rng.random(n[i]) < 0.4, 50.0, -1.0).astype(np.float32) for i in range(2)]
    oc = [rng.integers(0, 8, n[i]).astype(np.int32) for i in range(2)]
    F12 = (rng.normal(0, 1, (3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-3], [1e-3, 1e-3, 1e-1]])).astype(np.float32)
    Ow = rng.normal(0, 1, 3).astype(np.float32)
    R, _ = np.linalg.qr(rng.normal(0, 1, (3, 3)))
    R = R.astype(np.float32)
    t = np.array([0.1, -0.2, 2.5], np.float32)
    K = np.array([500.0, 510.0, 320.0, 240.0], np.float32)
    sf2 = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    sig2 = (sf2 * sf2 * 4000).astype(np.float32)
    blob = b""
    for i in range(2):
        blob += struct.pack("<i", n[i]) + desc[i].tobytes() + node[i].astype(np.int32).tobytes() + state[i].tobytes()
        blob += ang[i].tobytes() + x[i].tobytes() + y[i].tobytes() + ur[i].tobytes() + oc[i].tobytes()
    blob += struct.pack("<fii", ratio, ori, only_stereo) + F12.tobytes() + Ow.tobytes() + R.tobytes() + t.tobytes() + K.tobytes()
    blob += struct.pack("<i", 8) + sf2.tobytes() + sig2.tobytes()
    (tmp_path / "in.bin").write_bytes(blob)
    subprocess.check_call([_driver(), "match", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    out = (tmp_path / "out.bin").read_bytes()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, out, pos)
        pos += struct.calcsize(fmt)
        return v

    fv = [orc.FeatVec(node[0]), orc.FeatVec(node[1])]
    good = [(state[i] == 1).astype(np.uint8) for i in range(2)]
    hasmp = [(state[i] != 0).astype(np.uint8) for i in range(2)]
    # SearchByBoW(KF, Frame)
    nm, sz = take("<ii")
    got = np.frombuffer(out, np.int32, sz, pos); pos += 4 * sz
    on, om = orc.search_bow_kf_f(desc[0], good[0], ang[0], fv[0], desc[1], ang[1], fv[1], np.float32(ratio), ori)
    assert nm == on and np.array_equal(got, om)
    # SearchByBoW(KF, KF)
    nm, sz = take("<ii")
    got = np.frombuffer(out, np.int32, sz, pos); pos += 4 * sz
    on, om = orc.search_bow_kf_kf(desc[0], good[0], ang[0], fv[0], desc[1], good[1], ang[1], fv[1], np.float32(ratio), ori)
    assert nm == on and np.array_equal(got, om)
    # SearchForTriangulation: epipole as the class computes it (cv::gemm on 3x3 floats: float32, left to right; prim_gemm3.npz)
    C2 = (((R[:, 0] * Ow[0] + R[:, 1] * Ow[1]) + R[:, 2] * Ow[2]) + t).astype(np.float32)
    invz = np.float32(1.0) / C2[2]
    ex = K[0] * C2[0] * invz + K[2]
    ey = K[1] * C2[1] * invz + K[3]
    nm, sz = take("<ii")
    got = np.frombuffer(out, np.int32, 2 * sz, pos).reshape(-1, 2); pos += 8 * sz
    on, op = orc.search_triangulation(desc[0], hasmp[0], ur[0], x[0], y[0], ang[0], fv[0], desc[1], hasmp[1], ur[1], x[1], y[1], ang[1],
                                      oc[1], fv[1], F12, ex, ey, sf2, sig2, only_stereo, ori)
    assert nm == on and np.array_equal(got, op)
    d, lo, hi, hl = take("<iiii")
    assert d == orc.descriptor_distance(desc[0][0], desc[1][0]) and (lo, hi, hl) == (50, 100, 30)
    # batched overloads: every element equals the single call; DescriptorDistances equals the bit hack row by row
    on, om = orc.search_bow_kf_f(desc[0], good[0], ang[0], fv[0], desc[1], ang[1], fv[1], np.float32(ratio), ori)
    for _ in range(2):
        nm, sz = take("<ii")
        got = np.frombuffer(out, np.int32, sz, pos); pos += 4 * sz
        assert nm == on and np.array_equal(got, om)
    on, om = orc.search_bow_kf_kf(desc[0], good[0], ang[0], fv[0], desc[1], good[1], ang[1], fv[1], np.float32(ratio), ori)
    for _ in range(2):
        nm, sz = take("<ii")
        got = np.frombuffer(out, np.int32, sz, pos); pos += 4 * sz
        assert nm == on and np.array_equal(got, om)
    nd, = take("<i")
    got = np.frombuffer(out, np.int32, nd, pos); pos += 4 * nd
    assert nd == min(n) and np.array_equal(got, np.unpackbits(desc[0][:nd] ^ desc[1][:nd], axis=1).sum(1))


def test_cpp_vocabulary_transform_matches_oracle(tmp_path):
    """ORBVocabularyB200::transform(features, BowVector, FeatureVector, 4) == DBoW2 semantics (oracle), doubles bit-exact."""
    import orbslam_mapsave_b200 as orb
    from vocab_util import make_tree
    from orbslam_mapsave_b200.synth import synth_descriptors
    k, L = 10, 4
    parent, desc, weight, is_leaf = make_tree(k, L, seed=7)
    orb.ORBVocabulary.from_arrays(k, L, parent, desc, weight, is_leaf).saveToBinaryFile(tmp_path / "voc.bin")
    feats = synth_descriptors(2000, 9, dup_of=desc[1:], dup_rate=0.7, max_flip=30)
    (tmp_path / "desc.raw").write_bytes(feats.tobytes())
    subprocess.check_call([_driver(), "bow", str(tmp_path / "voc.bin"), str(tmp_path / "desc.raw"), "4", str(tmp_path / "bow.bin")])
    out = (tmp_path / "bow.bin").read_bytes()
    nw, nb = struct.unpack_from("<ii", out, 0)
    pos = 8
    words, vals = [], []
    for _ in range(nb):
        w, v = struct.unpack_from("<id", out, pos)
        pos += 12
        words.append(w)
        vals.append(v)
    nn, = struct.unpack_from("<i", out, pos)
    pos += 4
    ids, lists = [], []
    for _ in range(nn):
        i, c = struct.unpack_from("<ii", out, pos)
        pos += 8
        lists.append(np.frombuffer(out, np.int32, c, pos))
        pos += 4 * c
        ids.append(i)
    w32 = weight.astype(np.float32).astype(np.float64)           # the binary format stores float weights
    ow, owt, onid = orc.voc_transform(parent, desc, w32, is_leaf, L, feats, 4)
    bw, bv = orc.voc_bow(ow, owt, 0, 0)
    assert nw == int(is_leaf.sum())
    assert np.array_equal(np.array(words, np.int32), bw)
    assert np.array_equal(np.array(vals, np.float64).view(np.uint64), bv.view(np.uint64))
    keep = owt > 0
    ofv = orc.FeatVec(onid[keep])
    assert ids == ofv.ids.tolist()
    idx = np.nonzero(keep)[0]
    for j, lst in enumerate(lists):
        assert np.array_equal(lst, idx[ofv.feat[ofv.off[j]:ofv.off[j + 1]]])


@pytest.mark.parametrize("seed,mono,ori", [(0, False, 1), (1, True, 1), (2, False, 0)])
def test_cpp_window_searches_match_oracle(tmp_path, seed, mono, ori):
    """ORBmatcher::SearchByProjection(F, MapPoints) / (CurrentFrame, LastFrame) / SearchForInitialization through the C++ class,
    on Frames assembled like Frame::Frame does (AssignFeaturesToGrid), against the oracle."""
    import proj_util as pu
    rng = np.random.default_rng(100 + seed)
    ratio = 0.8

    def frame_blob(fa):
        ur = fa["uright"] if fa["uright"] is not None else np.full(len(fa["x"]), -1, np.float32)
        return (struct.pack("<i", len(fa["x"])) + fa["desc"].tobytes() + fa["x"].tobytes() + fa["y"].tobytes() + fa["angle"].tobytes() +
                ur.astype(np.float32).tobytes() + fa["octave"].tobytes() + np.array(fa["bounds"], np.float32).tobytes() +
                struct.pack("<i", 8) + pu.SCALE.tobytes())

    blob = struct.pack("<fi", ratio, ori)
    # A: local map points
    n = 1500
    faA = pu.frame_arrays(n, rng, stereo=not mono, cluster=seed == 2)
    blockedA = (rng.random(n) < 0.15).astype(np.uint8)
    mp = pu.map_points_for(faA, 2000, rng)
    th = 3.0
    blob += frame_blob(faA) + blockedA.tobytes() + struct.pack("<if", 2000, th)
    blob += mp["in_view"].tobytes() + mp["claims"].tobytes() + mp["desc"].tobytes()
    blob += mp["proj_x"].tobytes() + mp["proj_y"].tobytes() + mp["proj_xr"].tobytes() + mp["view_cos"].tobytes() + mp["level"].tobytes()
    # B: last frame
    faB = pu.frame_arrays(n, rng, stereo=not mono)
    blockedB = (rng.random(n) < 0.1).astype(np.uint8)
    lf = pu.last_frame_for(faB, 1800, rng, tz=0.3 if seed == 2 else 0.0)
    mbf, mb, thB = 40.0, np.float32(40.0) / np.float32(lf["fx"]), 15.0 if mono else 7.0
    T4 = lambda T: np.concatenate([T, np.array([[0, 0, 0, 1]], np.float32)], 0).astype(np.float32)
    blob += frame_blob(faB) + blockedB.tobytes() + T4(lf["Tcw"]).tobytes() + T4(lf["Tlw"]).tobytes()
    blob += np.array([lf["fx"], lf["fy"], lf["cx"], lf["cy"], mbf, mb], np.float32).tobytes() + struct.pack("<fii", thB, int(mono), 1800)
    blob += lf["has_point"].tobytes() + lf["claims"].tobytes() + lf["desc"].tobytes() + lf["world"].tobytes() + lf["angle"].tobytes() + lf["octave"].tobytes()
    # C: initialisation
    faC = pu.frame_arrays(n, rng, stereo=False)
    faC["octave"][rng.random(n) < 0.5] = 0
    f1 = pu.init_frame1_for(faC, 1600, rng)
    blob += frame_blob(faC) + struct.pack("<ii", 1600, 100) + f1["desc1"].tobytes() + f1["octave1"].tobytes() + f1["angle1"].tobytes() + f1["prev"].tobytes()
    (tmp_path / "in.bin").write_bytes(blob)
    subprocess.check_call([_driver(), "project", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    out = (tmp_path / "out.bin").read_bytes()
    pos = 0
    # A
    (nm,) = struct.unpack_from("<i", out, pos); pos += 4
    got = np.frombuffer(out, np.int32, n, pos); pos += 4 * n
    ogA = orc.Grid(faA["desc"], faA["x"], faA["y"], faA["octave"], pu.SCALE, faA["bounds"], angle=faA["angle"], uright=faA["uright"], blocked=blockedA)
    on, oo = orc.search_projection_map(ogA, th=th, nnratio=ratio, **mp)
    assert nm == on and np.array_equal(got, oo) and on > 100
    # B
    (nm,) = struct.unpack_from("<i", out, pos); pos += 4
    got = np.frombuffer(out, np.int32, n, pos); pos += 4 * n
    ogB = orc.Grid(faB["desc"], faB["x"], faB["y"], faB["octave"], pu.SCALE, faB["bounds"], angle=faB["angle"], uright=faB["uright"], blocked=blockedB)
    on, oo = orc.search_projection_frame(ogB, lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], mbf, mb, lf["has_point"], lf["world"],
                                         lf["octave"], lf["angle"], lf["desc"], lf["claims"], thB, mono, ori)
    had_point = (np.arange(n) % 3 == 0) | (blockedB != 0)
    want = np.where((oo == -2) & ~had_point, -1, oo)        # NULLing an entry that held nothing is invisible to the caller
    assert nm == on and np.array_equal(got, want) and (oo >= 0).sum() > 100
    (batch_ok,) = struct.unpack_from("<i", out, pos); pos += 4      # the batched overload on three copies == the single call
    assert batch_ok == 1
    # C
    nm, sz = struct.unpack_from("<ii", out, pos); pos += 8
    got = np.frombuffer(out, np.int32, sz, pos); pos += 4 * sz
    gprev = np.frombuffer(out, np.float32, 2 * 1600, pos).reshape(-1, 2)
    ogC = orc.Grid(faC["desc"], faC["x"], faC["y"], faC["octave"], pu.SCALE, faC["bounds"], angle=faC["angle"])
    oprev = f1["prev"].copy()
    on, om = orc.search_initialization(ogC, f1["desc1"], f1["octave1"], f1["angle1"], oprev, 100, ratio, ori)
    assert nm == on and sz == 1600 and np.array_equal(got, om) and np.array_equal(gprev, oprev) and on > 50


@pytest.mark.parametrize("seed,ori,scale", [(0, 1, 1.0), (1, 0, 2.5)])
def test_cpp_relocalisation_and_loop_projection_match_oracle(tmp_path, seed, ori, scale):
    """SearchByProjection(CurrentFrame, KeyFrame, sAlreadyFound, th, ORBdist) and SearchByProjection(KeyFrame, Scw, vpPoints,
    vpMatched, th) through the C++ class (projection on the host, window search on the GPU) against the oracle."""
    import proj_util as pu
    rng = np.random.default_rng(200 + seed)
    n, npts = 1500, 1800
    log_sf = float(np.log(np.float32(1.2)))

    def frame_blob(fa):
        ur = np.full(len(fa["x"]), -1, np.float32)
        return (struct.pack("<i", len(fa["x"])) + fa["desc"].tobytes() + fa["x"].tobytes() + fa["y"].tobytes() + fa["angle"].tobytes() +
                ur.tobytes() + fa["octave"].tobytes() + np.array(fa["bounds"], np.float32).tobytes() + struct.pack("<i", 8) + pu.SCALE.tobytes())

    T4 = lambda T: np.concatenate([T, np.array([[0, 0, 0, 1]], np.float32)], 0).astype(np.float32)
    blob = struct.pack("<i", ori)
    # relocalisation
    faD = pu.frame_arrays(n, rng, stereo=False, cluster=seed == 1)
    blockedD = (rng.random(n) < 0.2).astype(np.uint8)
    kp = pu.kf_points_for(faD, npts, rng)
    thD, orbdist = 10.0, 100 if seed == 0 else 64
    K = np.array([kp["fx"], kp["fy"], kp["cx"], kp["cy"], log_sf], np.float32)
    blob += frame_blob(faD) + blockedD.tobytes() + T4(kp["T"]).tobytes() + K.tobytes() + struct.pack("<fii", thD, orbdist, npts)
    blob += kp["state"].tobytes() + kp["desc"].tobytes() + kp["world"].tobytes() + kp["normal"].tobytes() + kp["mf_max"].tobytes() + kp["mf_min"].tobytes() + kp["angle"].tobytes()
    # loop closing
    faE = pu.frame_arrays(n, rng, stereo=False)
    faE["x"] = np.clip(faE["x"], 0, 639.5).astype(np.float32)
    blockedE = (rng.random(n) < 0.3).astype(np.uint8)
    kq = pu.kf_points_for(faE, npts, rng, sim_scale=scale)
    thE = 10
    blob += frame_blob(faE) + blockedE.tobytes() + T4(kq["S"]).tobytes() + K.tobytes() + struct.pack("<ii", thE, npts)
    blob += kq["state"].tobytes() + kq["desc"].tobytes() + kq["world"].tobytes() + kq["normal"].tobytes() + kq["mf_max"].tobytes() + kq["mf_min"].tobytes()
    (tmp_path / "in.bin").write_bytes(blob)
    subprocess.check_call([_driver(), "project2", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    out = (tmp_path / "out.bin").read_bytes()
    pos = 0
    (nm,) = struct.unpack_from("<i", out, pos); pos += 4
    got = np.frombuffer(out, np.int32, n, pos); pos += 4 * n
    og = orc.Grid(faD["desc"], faD["x"], faD["y"], faD["octave"], pu.SCALE, faD["bounds"], angle=faD["angle"], blocked=blockedD)
    on, oo = orc.search_projection_kf(og, kp["T"], kp["fx"], kp["fy"], kp["cx"], kp["cy"], np.float32(log_sf), kp["state"] == 1, kp["world"],
                                      kp["mf_max"], kp["mf_min"], kp["angle"], kp["desc"], thD, orbdist, ori)
    want = np.where(oo == -2, -1, oo)            # a culled entry held nothing before: NULL again
    assert nm == on and np.array_equal(got, want) and (oo >= 0).sum() > 100
    (nm,) = struct.unpack_from("<i", out, pos); pos += 4
    got = np.frombuffer(out, np.int32, n, pos); pos += 4 * n
    og = orc.Grid(faE["desc"], faE["x"], faE["y"], faE["octave"], pu.SCALE, faE["bounds"], angle=faE["angle"], blocked=blockedE)
    on, oo = orc.search_projection_sim3(og, kq["S"], kq["fx"], kq["fy"], kq["cx"], kq["cy"], np.float32(log_sf), kq["state"] == 1, kq["world"],
                                        kq["mf_max"], kq["mf_min"], kq["normal"], kq["desc"], thE)
    assert nm == on and np.array_equal(got, oo) and on > 100


def test_cpp_stereo_matches_oracle(tmp_path):
    """Two C++ ORBextractors + ORBextractor::ComputeStereoMatches (the stereo Frame constructor's path, src/Frame.cc:78-95) with
    the pyramid download switched off, against the oracle's ComputeStereoMatches on the oracle's pyramids."""
    from test_gpu_stereo import _stereo_pair
    left, right = _stereo_pair(640, 480, 7)
    (tmp_path / "l.raw").write_bytes(left.tobytes())
    (tmp_path / "r.raw").write_bytes(right.tobytes())
    mbf, mb = 40.0, np.float32(40.0) / np.float32(500.0)
    subprocess.check_call([_driver(), "stereo", str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), "640", "480", "1000", "8", repr(mbf),
                           repr(float(mb)), str(tmp_path / "out.bin")])
    out = (tmp_path / "out.bin").read_bytes()
    (n,) = struct.unpack_from("<i", out, 0)
    got_u = np.frombuffer(out, np.float32, n, 4)
    got_d = np.frombuffer(out, np.float32, n, 4 + 4 * n)
    oL, oR = orc.Extractor(1000, 1.2, 8, 20, 7), orc.Extractor(1000, 1.2, 8, 20, 7)
    okL, odL = oL.extract(left)
    okR, odR = oR.extract(right)
    t = oL.tables()
    want_u, want_d = orc.stereo_matches([oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)], t["scale"], t["inv_scale"],
                                        okL, odL, okR, odR, mbf, float(mb))
    assert n == len(okL) and np.array_equal(got_u.view(np.uint32), want_u.view(np.uint32))
    assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32)) and (want_u >= 0).sum() > 200


class _Pt:
    def __init__(self, bad, nobs):
        self.bad, self.nobs, self.obs = bool(bad), int(nobs), {}


def _emulate_fuse(variant, uright, kf_ptr, kfp, cand, cand_state, best):
    """The bookkeeping of Fuse (src/ORBmatcher.cc:952-971 / 1083-1099) and MapPoint::Replace / AddObservation
    (src/MapPoint.cc:93-116, 226-283) replayed in Python from the oracle's per-point search result."""
    def add_obs(p, idx):
        if "kf" in p.obs:
            return
        p.obs["kf"] = idx
        p.nobs += 2 if uright[idx] >= 0 else 1

    def replace(this, other):
        if this is other:
            return
        obs, this.obs, this.bad = this.obs, {}, True
        for _, idx in obs.items():
            if "kf" not in other.obs:
                kf_ptr[idx] = other
                add_obs(other, idx)
            else:
                kf_ptr[idx] = None
    n, rep = 0, [None] * len(cand)
    for i, p in enumerate(cand):
        if variant == 0:
            if cand_state[i] == 0 or p.bad or "kf" in p.obs:
                continue
        if best[i] < 0:
            continue
        inkf = kf_ptr[best[i]]
        if inkf is not None:
            if not inkf.bad:
                if variant == 0:
                    if inkf.nobs > p.nobs:
                        replace(p, inkf)
                    else:
                        replace(inkf, p)
                else:
                    rep[i] = inkf
        else:
            add_obs(p, best[i])
            kf_ptr[best[i]] = p
        n += 1
    return n, rep


def test_cpp_fuse_and_sim3_match_oracle(tmp_path):
    """ORBmatcher::Fuse (both overloads) and SearchBySim3 through the C++ class: candidate loops on the GPU
    (orbm_search_windows_best), map updates on the host, against the oracle's searches + a Python replay of the bookkeeping."""
    import proj_util as pu
    rng = np.random.default_rng(300)
    n, npts = 1200, 1500
    log_sf = float(np.log(np.float32(1.2)))
    K = np.array([520.0, 520.0, 320.0, 240.0, 40.0, log_sf], np.float32)
    is2 = (1.0 / (pu.SCALE * pu.SCALE)).astype(np.float32)

    def frame_blob(fa):
        return (struct.pack("<i", len(fa["x"])) + fa["desc"].tobytes() + fa["x"].tobytes() + fa["y"].tobytes() + fa["angle"].tobytes() +
                fa["uright"].astype(np.float32).tobytes() + fa["octave"].tobytes() + np.array(fa["bounds"], np.float32).tobytes() + struct.pack("<i", 8) + pu.SCALE.tobytes())

    def kf_blob(fa, T, Ow):
        return frame_blob(fa) + K.tobytes() + is2.tobytes() + np.ascontiguousarray(T, np.float32).tobytes() + np.ascontiguousarray(Ow, np.float32).tobytes()

    def pts_blob(state, nobs, P):
        return (struct.pack("<i", len(state)) + state.tobytes() + nobs.astype(np.int32).tobytes() + P["desc"].tobytes() + P["world"].tobytes() +
                P["normal"].tobytes() + P["mf_max"].tobytes() + P["mf_min"].tobytes())

    def frame(seed_cluster=False):
        fa = pu.frame_arrays(n, rng, stereo=True, cluster=seed_cluster)
        fa["x"] = np.clip(fa["x"], 0, 639.5).astype(np.float32)
        fa["y"] = np.clip(fa["y"], 0, 479.5).astype(np.float32)
        return fa

    def pose(deg=3.0):
        R = pu.rot_small(rng, deg).astype(np.float32)
        t = rng.uniform(-0.1, 0.1, 3).astype(np.float32)
        return R, t

    blob, expect = b"", []
    for variant in (0, 1):
        fa = frame(variant == 1)
        R, t = pose()
        T = np.concatenate([R, t[:, None]], 1).astype(np.float32)
        Ow = (-(R.astype(np.float64).T @ t.astype(np.float64))).astype(np.float32)
        scale = 1.0 if variant == 0 else 1.7
        A, b = (T[:, :3] * np.float32(scale)).astype(np.float32), (T[:, 3] * np.float32(scale)).astype(np.float32)
        S = np.concatenate([A, b[:, None]], 1).astype(np.float32)
        kstate = rng.choice(np.array([0, 0, 1, 1, 2], np.uint8), n)
        knobs = rng.integers(1, 6, n)
        kP = pu.points_for_transform(fa, n, rng, A, b, False)
        P = pu.points_for_transform(fa, npts, rng, A, b, False)
        cstate = rng.choice(np.array([0, 1, 1, 1, 1, 2, 3], np.uint8), npts) if variant == 0 else rng.choice(np.array([1, 1, 1, 1, 2, 3], np.uint8), npts)
        cnobs = rng.integers(0, 6, npts)
        free = np.nonzero(kstate == 0)[0]
        in_at = np.full(npts, -1, np.int32)
        three = np.nonzero(cstate == 3)[0]
        in_at[three] = rng.choice(free, len(three), replace=False)
        th = 3.0 if variant == 0 else 4.0
        blob += kf_blob(fa, T, Ow) + pts_blob(kstate, knobs, kP) + pts_blob(cstate, cnobs, P) + struct.pack("<f", th) + in_at.tobytes()
        if variant == 1:
            blob += np.concatenate([S, np.array([[0, 0, 0, 1]], np.float32)], 0).astype(np.float32).tobytes()
        og = orc.Grid(fa["desc"], fa["x"], fa["y"], fa["octave"], pu.SCALE, fa["bounds"], angle=fa["angle"], uright=fa["uright"])
        skip = (cstate != 1).astype(np.uint8)
        best = orc.fuse_search(og, variant, T if variant == 0 else S, Ow if variant == 0 else None, K[0], K[1], K[2], K[3], K[4], np.float32(log_sf),
                               skip, P["world"], P["mf_max"], P["mf_min"], P["normal"], P["desc"], th, is2)
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
        enc = lambda p: -1 if p is None else (kfp.index(p) if p in kfp else 100000 + cand.index(p))
        expect.append((nf, [enc(p) for p in rep], [enc(p) for p in kf_ptr], [(int(p.bad), p.nobs) for p in kfp + cand], int((best >= 0).sum())))
    # SearchBySim3
    fa1, fa2 = frame(), frame()
    (Ra, ta), (Rb, tb) = pose(), pose()
    s12 = np.float32(1.15)
    R12 = (Ra.astype(np.float64) @ Rb.astype(np.float64).T).astype(np.float32)
    t12 = (ta.astype(np.float64) - float(s12) * (R12.astype(np.float64) @ tb.astype(np.float64))).astype(np.float32)
    # the class's own arithmetic for sR12, sR21, t21 (float scaling like cv::Mat::convertTo, float32 left-to-right products)
    sR12 = (R12 * s12).astype(np.float32)
    sR21 = (R12.T * np.float32(1.0 / float(s12))).astype(np.float32)
    t21 = (-((sR21[:, 0] * t12[0] + sR21[:, 1] * t12[1]) + sR21[:, 2] * t12[2])).astype(np.float32)
    perm = rng.permutation(n)                                  # KF1 feature i <-> KF2 feature perm[i], with a few broken pairs
    inv = np.argsort(perm)
    broken = rng.random(n) < 0.15
    tgt2 = np.where(broken, rng.integers(0, n, n), inv)
    P1 = pu.points_for_transform(fa2, n, rng, (Rb / s12).astype(np.float32), tb, True, tgt=perm)   # KF1's points land on KF2's features
    P2 = pu.points_for_transform(fa1, n, rng, (Ra * s12).astype(np.float32), ta, True, tgt=tgt2)
    st1 = rng.choice(np.array([0, 1, 1, 1, 2], np.uint8), n)
    st2 = rng.choice(np.array([0, 1, 1, 1, 2], np.uint8), n)
    pre = np.where((rng.random(n) < 0.1) & (st1 == 1), rng.integers(0, n, n), -1).astype(np.int32)
    pre[(pre >= 0) & (st2[np.maximum(pre, 0)] == 0)] = -1
    Tz = lambda R, t: np.concatenate([R, t[:, None]], 1).astype(np.float32)
    th3 = 7.5
    blob += kf_blob(fa1, Tz(Ra, ta), np.zeros(3)) + kf_blob(fa2, Tz(Rb, tb), np.zeros(3))
    blob += pts_blob(st1, np.ones(n), P1) + pts_blob(st2, np.ones(n), P2) + struct.pack("<ff", float(s12), th3) + R12.tobytes() + t12.tobytes() + pre.tobytes()
    og1 = orc.Grid(fa1["desc"], fa1["x"], fa1["y"], fa1["octave"], pu.SCALE, fa1["bounds"], angle=fa1["angle"], uright=fa1["uright"])
    og2 = orc.Grid(fa2["desc"], fa2["x"], fa2["y"], fa2["octave"], pu.SCALE, fa2["bounds"], angle=fa2["angle"], uright=fa2["uright"])
    already1 = pre >= 0
    already2 = np.zeros(n, bool)
    already2[pre[pre >= 0]] = True
    m1 = orc.sim3_direction(og2, Ra, ta, sR21, t21, K[0], K[1], K[2], K[3], np.float32(log_sf), (st1 == 1) & ~already1, P1["world"], P1["mf_max"],
                            P1["mf_min"], P1["desc"], th3)
    m2 = orc.sim3_direction(og1, Rb, tb, sR12, t12, K[0], K[1], K[2], K[3], np.float32(log_sf), (st2 == 1) & ~already2, P2["world"], P2["mf_max"],
                            P2["mf_min"], P2["desc"], th3)
    want12 = pre.copy()
    nfound = 0
    for i1 in range(n):
        if m1[i1] >= 0 and m2[m1[i1]] == i1:
            want12[i1] = m1[i1]
            nfound += 1
    (tmp_path / "in.bin").write_bytes(blob)
    subprocess.check_call([_driver(), "fuse", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    out = (tmp_path / "out.bin").read_bytes()
    pos = 0
    for variant in (0, 1):
        nf, rep, kfptr, pts, nbest = expect[variant]
        if variant == 1:
            got = np.frombuffer(out, np.int32, npts, pos); pos += 4 * npts
            assert np.array_equal(got, np.array(rep, np.int32))
        (gn,) = struct.unpack_from("<i", out, pos); pos += 4
        got = np.frombuffer(out, np.int32, n, pos); pos += 4 * n
        assert gn == nf and np.array_equal(got, np.array(kfptr, np.int32)) and nf > 100, (variant, gn, nf, nbest)
        got = np.frombuffer(out, np.int32, 2 * (n + npts), pos).reshape(-1, 2); pos += 8 * (n + npts)
        assert np.array_equal(got, np.array(pts, np.int32))
    (gn,) = struct.unpack_from("<i", out, pos); pos += 4
    got = np.frombuffer(out, np.int32, n, pos)
    assert gn == nfound and np.array_equal(got, want12) and nfound > 50, (gn, nfound)
