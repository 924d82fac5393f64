"""Pin the CPU oracle (oracle/orb_oracle.cpp) against the cv2-4.13 golden vectors in tests/golden/.

These are the "known-answer tests" of the hot path: the reference ships none, so the vectors were
generated from the OpenCV primitives the reference calls (tests/golden/make_golden.py).
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import orb_oracle_py as orc


def test_pattern_table_hash():
    # SHA-256 printed by tools/gen_pattern.py when it extracted the table from src/ORBextractor.cc:149-407
    assert hashlib.sha256(orc.pattern().tobytes()).hexdigest() == \
        "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"


def test_fast_matches_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "prim_fast.npz"))
    n = 0
    for i in range(24):
        roi = g[f"roi{i}"]
        for t in (20, 7):
            want = g[f"roi{i}_t{t}"]
            got = orc.fast(roi, t, True)
            got = np.stack([got["x"], got["y"], got["r"]], 1).reshape(-1, 3)
            assert np.array_equal(got, want), (i, t)
            n += len(want)
    assert n > 100


def test_fast_nonms_matches_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "prim_fast_nonms.npz"))
    for i in (1, 2):
        got = orc.fast(g[f"roi{i}"], 7, False)
        assert np.array_equal(np.stack([got["x"], got["y"]], 1).reshape(-1, 2), g[f"roi{i}_t7"])


def test_fast_strided_view():
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    view = big[5:40, 7:45]
    a = orc.fast(view, 20, True)
    b = orc.fast(np.ascontiguousarray(view), 20, True)
    assert np.array_equal(a, b)


def test_resize_matches_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "prim_resize.npz"))
    a = orc.resize(g["small_src"], 133, 100)
    assert np.array_equal(a, g["small_133x100"])
    assert np.array_equal(orc.resize(a, 111, 83), g["small_111x83"])
    assert np.array_equal(orc.resize(g["rnd_src"], 109, 81), g["rnd_109x81"])
    assert np.array_equal(orc.resize(g["rnd_src"], 200, 150), g["rnd_200x150"])


def test_blur_and_border_match_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "prim_blur.npz"))
    r = np.load(os.path.join(golden_dir, "prim_resize.npz"))
    assert np.array_equal(orc.blur(r["small_src"]), g["small_blur"])
    assert np.array_equal(orc.blur(r["rnd_src"]), g["rnd_blur"])
    assert np.array_equal(orc.border101(r["rnd_src"], 19), g["rnd_border19"])


def test_fast_atan2_matches_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "prim_atan2.npz"))
    got = np.array([orc.fast_atan2(y, x) for y, x in g["yx"]], np.float32)
    assert np.array_equal(got.view(np.uint32), g["angle"].view(np.uint32))


def _chain_image(g, W, H, seed):
    """The fixture's image, or (large cases) the synthetic frame regenerated from its seed and checked against the recorded CRC."""
    if "image" in g.files:
        return g["image"]
    import zlib
    from orbslam_mapsave_b200.synth import synth as _synth
    img = _synth(W, H, seed)
    assert zlib.crc32(img.tobytes()) == int(g["image_crc"]), "synth() no longer reproduces the frame this fixture was made from"
    return img


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "chain_*.npz"))),
                         ids=lambda p: os.path.basename(p)[6:-4])
def test_full_extractor_matches_cv2_chain(path):
    """Oracle == the reference's control flow chained over the real cv2 primitives (bit-exact everything)."""
    g = np.load(path)
    W, H, seed, nf, nl, ini, mn, use_mask = (int(v) for v in g["params"])
    img = _chain_image(g, W, H, seed)
    mask = None
    if use_mask:
        mask = np.full(img.shape, 255, np.uint8)
        mask[60:220, 150:260] = 0
    ex = orc.Extractor(nf, float(g["scaleFactor"]), nl, ini, mn)
    kp, desc = ex.extract(img, mask)
    assert np.array_equal(ex.tables()["quota"], g["quota"])
    assert [len(ex.candidates(l)) for l in range(nl)] == g["cand_counts"].tolist()
    assert [int(ex.level(l).astype(np.int64).sum()) for l in range(nl)] == g["level_sums"].tolist()
    assert np.array_equal(ex.level(nl - 1), g["last_level"])
    for l in range(nl):
        b = ex.blurred(l)
        assert (int(b.astype(np.int64).sum()) if b is not None else -1) == int(g["blur_sums"][l])
    want = g["kp"]
    assert len(kp) == len(want)
    for f in ("x", "y", "size", "response", "octave", "class_id", "angle"):
        assert np.array_equal(kp[f], want[f]), f
    assert np.array_equal(desc, g["desc"])


def test_small_gemm_matches_cv2(golden_dir):
    """`Rcw*x3Dw+tcw` (src/ORBmatcher.cc:1365) = cv::gemm(A, B, 1, C, 1): float32 left to right for 3x3 floats;
    `-Rcw.t()*tcw` (:1347) takes the general path with double accumulation (tests/golden/make_golden_gemm.py)."""
    import ctypes as C
    g = np.load(os.path.join(golden_dir, "prim_gemm3.npz"))
    L = orc.lib()
    L.orc_rt_apply.argtypes = [C.c_void_p] * 3
    L.orc_minus_rt_t.argtypes = [C.c_void_p] * 2
    R, x, t = g["R"], g["x"], g["t"]
    out = np.zeros(3, np.float32)
    for i in range(len(R)):
        T = np.ascontiguousarray(np.concatenate([R[i], t[i][:, None]], 1), np.float32)
        L.orc_rt_apply(T.ctypes.data, x[i].ctypes.data, out.ctypes.data)
        assert np.array_equal(out, g["out"][i]), i
        T2 = np.ascontiguousarray(np.concatenate([R[i], x[i][:, None]], 1), np.float32)
        L.orc_minus_rt_t(T2.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, g["outT"][i]), i


def test_descriptor_distance_and_top2_match_cv2(golden_dir):
    """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1650-1666, the bit-twiddling popcount) == cv2.norm(NORM_HAMMING), and the
    oracle's top-2 distances == cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (tests/golden/make_golden_hamming.py, cv2 4.13)."""
    g = np.load(os.path.join(golden_dir, "prim_hamming.npz"))
    got = np.array([orc.descriptor_distance(x, y) for x, y in zip(g["a"], g["b"])], np.int32)
    assert np.array_equal(got, g["dist"])
    bi, b1, b2 = orc.hamming_top2(g["q"], g["db"])
    assert np.array_equal(b1, g["best"]) and np.array_equal(b2, g["second"])
    uniq = g["best"] < g["second"]
    assert np.array_equal(bi[uniq], g["best_idx"][uniq])
