"""GPU parity tests of the projection / window searches through the C ABI against the CPU oracle: feature
ownership, match tables and counts are index work and must be bit-exact, including the serial claim order of the reference."""
import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from oracle import orb_oracle_py as orc

import proj_util as pu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,npts,seed,cluster,stereo,th", [
    (2000, 3000, 10, False, True, 1.0), (2000, 3000, 11, True, True, 3.0), (1500, 800, 12, False, False, 5.0),
    (300, 4000, 13, True, True, 3.0), (1, 5, 14, False, True, 1.0), (50, 0, 15, False, True, 1.0), (8000, 9000, 16, False, True, 3.0)])
def test_search_by_projection_map_points(n, npts, seed, cluster, stereo, th):
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=stereo, cluster=cluster)
    blocked = (rng.random(n) < 0.15).astype(np.uint8)
    g, og = pu.make_grids(fa, blocked, orb, orc)
    mp = pu.map_points_for(fa, npts, rng)
    want_n, want = orc.search_projection_map(og, th=th, nnratio=0.8, **mp)
    got_n, got = orb.ORBmatcher(0.8, True).SearchByProjectionMapPoints(g, th=th, **mp)
    assert got_n == want_n and np.array_equal(got, want)
    if npts >= 800:
        assert want_n > min(npts, n) // 10


def test_projection_map_all_points_chase_one_feature():
    """Worst case for the claim replay: every map point wants the same feature; exactly the first acceptable claimer keeps it."""
    rng = np.random.default_rng(20)
    fa = pu.frame_arrays(400, rng, stereo=False)
    g, og = pu.make_grids(fa, None, orb, orc)
    k = 300
    t = int(np.argmin(np.abs(fa["x"] - 320) + np.abs(fa["y"] - 240)))
    mp = dict(in_view=np.ones(k, np.uint8), proj_x=np.full(k, fa["x"][t], np.float32), proj_y=np.full(k, fa["y"][t], np.float32),
              proj_xr=np.full(k, -1, np.float32), level=np.full(k, fa["octave"][t], np.int32), view_cos=np.full(k, 0.9, np.float32),
              desc=np.stack([pu.flip(fa["desc"][t], int(b), rng) for b in rng.integers(0, 60, k)]),
              claims=(np.arange(k) % 7 != 0).astype(np.uint8))
    want_n, want = orc.search_projection_map(og, th=3.0, nnratio=0.8, **mp)
    got_n, got = orb.ORBmatcher(0.8).SearchByProjectionMapPoints(g, th=3.0, **mp)
    assert got_n == want_n and np.array_equal(got, want)


@pytest.mark.parametrize("n,nlast,seed,mono,tz,ori,cluster", [
    (2000, 2000, 30, True, 0.0, True, False), (2000, 2000, 31, False, 0.5, True, False), (2000, 2000, 32, False, -0.5, True, True),
    (1200, 3000, 33, False, 0.0, False, True), (2000, 0, 34, True, 0.0, True, False), (3, 10, 35, False, 0.0, True, False),
    (8000, 8000, 36, False, 0.0, True, False)])
def test_search_by_projection_last_frame(n, nlast, seed, mono, tz, ori, cluster):
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=not mono, cluster=cluster)
    blocked = (rng.random(n) < 0.1).astype(np.uint8)
    g, og = pu.make_grids(fa, blocked, orb, orc)
    lf = pu.last_frame_for(fa, nlast, rng, tz=tz)
    mbf, mb, th = 40.0, 40.0 / lf["fx"], 15.0 if mono else 7.0
    args = (lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], mbf, mb, lf["has_point"], lf["world"], lf["octave"], lf["angle"],
            lf["desc"], lf["claims"], th, mono)
    want_n, want = orc.search_projection_frame(og, *args, ori)
    got_n, got = orb.ORBmatcher(0.9, ori).SearchByProjectionFrame(g, *args)
    assert got_n == want_n and np.array_equal(got, want)
    if nlast >= 2000:
        assert (want >= 0).sum() > nlast // 10
        if ori:
            assert (want == -2).sum() > 0            # the rotation cull fired


@pytest.mark.parametrize("ori", [True, False])
def test_search_by_projection_last_frame_batch(ori):
    """orbm_search_by_projection_frame_batch: ragged jobs (different feature counts, mono / stereo, forward / backward motion, an empty
    LastFrame, a three-feature frame) in one call == the oracle per job == the single call per job."""
    specs = [(2000, 2000, 130, True, 0.0, False), (2000, 2000, 131, False, 0.5, False), (1500, 2500, 132, False, -0.5, True),
             (1200, 3000, 133, False, 0.0, True), (2000, 0, 134, True, 0.0, False), (3, 10, 135, False, 0.0, False),
             (4000, 4000, 136, False, 0.0, False), (2000, 2000, 137, True, 0.0, True)]
    jobs, ojobs = [], []
    for n, nlast, seed, mono, tz, cluster in specs:
        rng = np.random.default_rng(seed)
        fa = pu.frame_arrays(n, rng, stereo=not mono, cluster=cluster)
        blocked = (rng.random(n) < 0.1).astype(np.uint8)
        g, og = pu.make_grids(fa, blocked, orb, orc)
        lf = pu.last_frame_for(fa, nlast, rng, tz=tz)
        mbf, mb, th = 40.0, 40.0 / lf["fx"], 15.0 if mono else 7.0
        args = (lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], mbf, mb, lf["has_point"], lf["world"], lf["octave"], lf["angle"],
                lf["desc"], lf["claims"], th, mono)
        jobs.append((g,) + args)
        ojobs.append((og,) + args)
    m = orb.ORBmatcher(0.9, ori)
    got = m.SearchByProjectionFrameBatch(jobs)
    assert len(got) == len(specs)
    total = 0
    for k, (gn, gown) in enumerate(got):
        wn, wown = orc.search_projection_frame(*ojobs[k], ori)
        sn, sown = m.SearchByProjectionFrame(*jobs[k])
        assert gn == wn == sn, k
        assert np.array_equal(gown, wown) and np.array_equal(gown, sown), k
        total += wn
    assert total > 3000
    assert m.SearchByProjectionFrameBatch([]) == []


@pytest.mark.parametrize("n2,n1,seed,window,ori", [(2000, 2000, 40, 100, True), (2000, 2000, 41, 10, True), (1000, 3000, 42, 50, False),
                                                   (2000, 0, 43, 100, True), (2, 9, 44, 100, True), (6000, 6000, 45, 100, True)])
def test_search_for_initialization(n2, n1, seed, window, ori):
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n2, rng, stereo=False)
    fa["octave"][rng.random(n2) < 0.5] = 0
    g, og = pu.make_grids(fa, None, orb, orc)
    f1 = pu.init_frame1_for(fa, n1, rng)
    prev_o, prev_g = f1["prev"].copy(), f1["prev"].copy()
    want_n, want = orc.search_initialization(og, f1["desc1"], f1["octave1"], f1["angle1"], prev_o, window, 0.9, ori)
    got_n, got = orb.ORBmatcher(0.9, ori).SearchForInitialization(g, f1["desc1"], f1["octave1"], f1["angle1"], prev_g, window)
    assert got_n == want_n and np.array_equal(got, want) and np.array_equal(prev_g, prev_o)
    if n1 >= 2000:
        assert want_n > 100


def test_window_search_rejects_bad_arguments():
    rng = np.random.default_rng(50)
    fa = pu.frame_arrays(20, rng)
    g, _ = pu.make_grids(fa, None, orb, orc)
    mp = pu.map_points_for(fa, 10, rng)
    mp["level"][:] = 9                       # predicted level outside mvScaleFactors
    mp["in_view"][:] = 1
    with pytest.raises(orb.OrbError):
        orb.ORBmatcher().SearchByProjectionMapPoints(g, **mp)


@pytest.mark.parametrize("n,nq,seed,cluster,ori,th", [(2000, 2500, 60, False, True, 100), (2000, 2500, 61, True, True, 100),
                                                      (3000, 1500, 62, True, False, 50), (400, 0, 63, False, True, 50)])
def test_search_windows_core(n, nq, seed, cluster, ori, th):
    """orbm_search_windows: explicit windows, best candidate, every accepted match blocks (the core of the relocalisation and
    loop-closing SearchByProjection overloads)."""
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=False, cluster=cluster)
    blocked = (rng.random(n) < 0.2).astype(np.uint8)
    g, og = pu.make_grids(fa, blocked, orb, orc)
    tgt = rng.integers(0, n, nq)
    dup = rng.random(nq) < 0.4
    if nq:
        tgt[dup] = tgt[rng.integers(0, nq, dup.sum())]
    desc = np.stack([pu.flip(fa["desc"][t], int(rng.choice([0, 5, 30, 49, 50, 51, 99, 100, 101])), rng) for t in tgt]) if nq else np.zeros((0, 32), np.uint8)
    u = (fa["x"][tgt] + rng.normal(0, 2, nq)).astype(np.float32)
    v = (fa["y"][tgt] + rng.normal(0, 2, nq)).astype(np.float32)
    lvl = fa["octave"][tgt] + rng.integers(-1, 2, nq)
    r = (rng.choice([7.0, 15.0, 40.0], nq) * np.float32(1.2) ** np.clip(lvl, 0, 7)).astype(np.float32)
    active = (rng.random(nq) < 0.9).astype(np.uint8)
    angle = ((fa["angle"][tgt] + rng.choice([0.0, 0.0, 120.0], nq)) % 360).astype(np.float32)
    args = (active, u, v, r, lvl - 1, lvl + rng.integers(0, 2, nq), desc)
    want_n, want = orc.search_windows(og, *args, angle, th, ori)
    got_n, got = orb.ORBmatcher(0.9, ori).SearchWindows(g, *args, angle=angle, th_dist=th)
    assert got_n == want_n and np.array_equal(got, want)
    if nq:
        assert (want >= 0).sum() > 200


@pytest.mark.parametrize("n,nq,seed,cluster,chi2,th", [(2000, 2500, 70, False, False, 100), (2000, 2500, 71, True, True, 50),
                                                       (3000, 1500, 72, False, True, 50), (10, 0, 73, False, False, 50)])
def test_search_windows_best(n, nq, seed, cluster, chi2, th):
    """orbm_search_windows_best: the independent candidate loops of SearchBySim3 / Fuse, with and without Fuse's chi-square gate."""
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(n, rng, stereo=True, cluster=cluster)
    g, og = pu.make_grids(fa, None, orb, orc)
    tgt = rng.integers(0, n, nq)
    desc = np.stack([pu.flip(fa["desc"][t], int(rng.choice([0, 5, 30, 49, 50, 51, 99, 100, 101])), rng) for t in tgt]) if nq else np.zeros((0, 32), np.uint8)
    u = (fa["x"][tgt] + rng.normal(0, 1.5, nq)).astype(np.float32)
    v = (fa["y"][tgt] + rng.normal(0, 1.5, nq)).astype(np.float32)
    lvl = fa["octave"][tgt] + rng.integers(0, 2, nq)
    r = (rng.choice([3.0, 4.0, 7.5], nq) * np.float32(1.2) ** np.clip(lvl, 0, 7)).astype(np.float32)
    active = (rng.random(nq) < 0.9).astype(np.uint8)
    ur = (np.where(fa["uright"][tgt] >= 0, fa["uright"][tgt], u - 10) + rng.normal(0, 1.5, nq)).astype(np.float32) if chi2 else None
    is2 = (1.0 / (pu.SCALE * pu.SCALE)).astype(np.float32) if chi2 else None
    args = (active, u, v, r, lvl - 1, lvl, desc, ur, is2)
    want = orc.search_windows_best(og, *args, th)
    got = orb.ORBmatcher().SearchWindowsBest(g, *args, th_dist=th)
    assert np.array_equal(got, want)
    if nq:
        assert (want >= 0).sum() > 200
