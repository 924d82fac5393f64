"""Threading contract of the drop-in boundary: distinct ORBextractor instances run concurrently (the stereo Frame
constructor extracts left and right on two std::threads, src/Frame.cc:78-81), and the matcher entry points are callable from
several threads at once (Tracking, LocalMapping and LoopClosing each own an ORBmatcher).  Results must equal the serial ones."""
import threading

import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth, synth_descriptors

import proj_util as pu
from oracle import orb_oracle_py as orc

pytestmark = pytest.mark.gpu


def test_two_extractors_on_two_threads_match_serial():
    imgs = [synth(640, 480, 300 + i) for i in range(12)]
    serial = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    want = [serial(im, download_pyramid=False) for im in imgs]
    exs = [orb.ORBextractor(1000, 1.2, 8, 20, 7) for _ in range(2)]
    got = [None] * len(imgs)
    errs = []

    def work(t):
        try:
            for rep in range(3):
                for i in range(t, len(imgs), 2):
                    got[i] = exs[t](imgs[i], download_pyramid=(rep == 2))
        except Exception as e:          # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for (k, d), (wk, wd) in zip(got, want):
        assert np.array_equal(k.view(np.uint8), wk.view(np.uint8)) and np.array_equal(d, wd)


def test_matcher_calls_from_four_threads_match_serial():
    rng = np.random.default_rng(9)
    db = synth_descriptors(1500, 1)
    qs = [synth_descriptors(900 + 50 * i, 20 + i, dup_of=db) for i in range(4)]
    fa = pu.frame_arrays(1500, rng, stereo=True)
    g, _ = pu.make_grids(fa, None, orb, orc)
    mps = [pu.map_points_for(fa, 1200, np.random.default_rng(40 + i)) for i in range(4)]
    m = orb.ORBmatcher(0.8, True)
    want_top2 = [m.hamming_top2(q, db) for q in qs]
    want_proj = [m.SearchByProjectionMapPoints(g, th=3.0, **mp) for mp in mps]
    out, errs = [None] * 4, []

    def work(i):
        try:
            mm = orb.ORBmatcher(0.8, True)
            for _ in range(10):
                a = mm.hamming_top2(qs[i], db)
                b = mm.SearchByProjectionMapPoints(g, th=3.0, **mps[i])
            out[i] = (a, b)
        except Exception as e:          # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(4):
        assert all(np.array_equal(x, y) for x, y in zip(out[i][0], want_top2[i]))
        assert out[i][1][0] == want_proj[i][0] and np.array_equal(out[i][1][1], want_proj[i][1])
