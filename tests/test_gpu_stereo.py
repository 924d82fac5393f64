"""GPU parity test of Frame::ComputeStereoMatches on the device-resident pyramids of two extractor handles,
against the oracle restatement run on the oracle's own pyramids.  mvuRight / mvDepth are float32 results of integer SADs and a
handful of IEEE operations: bit-exact."""
import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
from oracle import orb_oracle_py as orc

pytestmark = pytest.mark.gpu


def _stereo_pair(W, H, seed, max_disp=40):
    """A right image = the left scene shifted left by a disparity that grows toward the bottom of the image, plus fresh noise."""
    rng = np.random.default_rng(seed)
    left = synth(W + max_disp + 8, H, seed)
    rows = np.arange(H)
    disp = (3 + (max_disp - 3) * rows / (H - 1)).astype(np.int32)
    right = np.stack([left[y, disp[y]:disp[y] + W] for y in range(H)])
    left = left[:, :W].copy()
    right = np.clip(right.astype(np.int32) + rng.integers(-2, 3, right.shape), 0, 255).astype(np.uint8)
    return left, right


@pytest.mark.parametrize("W,H,nf,nl,seed,mbf", [(640, 480, 1000, 8, 0, 40.0), (752, 480, 1200, 8, 1, 47.9), (1280, 720, 2000, 8, 2, 386.0),
                                              (424, 240, 600, 6, 3, 20.0)])
def test_compute_stereo_matches_vs_oracle(W, H, nf, nl, seed, mbf):
    left, right = _stereo_pair(W, H, seed)
    exL, exR = orb.ORBextractor(nf, 1.2, nl, 20, 7), orb.ORBextractor(nf, 1.2, nl, 20, 7)
    kpL, dL = exL(left, download_pyramid=False)
    kpR, dR = exR(right, download_pyramid=False)
    oL, oR = orc.Extractor(nf, 1.2, nl, 20, 7), orc.Extractor(nf, 1.2, nl, 20, 7)
    okL, odL = oL.extract(left)
    okR, odR = oR.extract(right)
    assert np.array_equal(kpL["x"], okL["x"]) and np.array_equal(dL, odL) and np.array_equal(kpR["y"], okR["y"]) and np.array_equal(dR, odR)
    fx = 500.0
    mb = mbf / fx
    t = oL.tables()
    want_u, want_d = orc.stereo_matches([oL.level(l) for l in range(nl)], [oR.level(l) for l in range(nl)], t["scale"], t["inv_scale"],
                                        okL, odL, okR, odR, mbf, mb)
    got_u, got_d = exL.ComputeStereoMatches(exR, kpL, dL, kpR, dR, mbf, mb)
    assert np.array_equal(got_u.view(np.uint32), want_u.view(np.uint32))
    assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
    matched = (want_u >= 0).sum()
    assert matched > len(kpL) // 4, matched
    # the planted disparity is recovered
    rows = kpL["y"][want_u >= 0]
    true_disp = 3 + 37 * rows / (H - 1)
    assert np.median(np.abs((kpL["x"] - want_u)[want_u >= 0] - true_disp)) < 1.5


def test_stereo_no_right_keypoints_and_flat_image():
    left, _ = _stereo_pair(640, 480, 5)
    flat = np.full((480, 640), 90, np.uint8)
    exL, exR = orb.ORBextractor(1000, 1.2, 8, 20, 7), orb.ORBextractor(1000, 1.2, 8, 20, 7)
    kpL, dL = exL(left, download_pyramid=False)
    kpR, dR = exR(flat, download_pyramid=False)
    assert len(kpR) == 0
    u, d = exL.ComputeStereoMatches(exR, kpL, dL, kpR, dR, 40.0, 0.08)
    assert (u == -1).all() and (d == -1).all()
