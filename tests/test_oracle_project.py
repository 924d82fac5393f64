"""CPU: the C++ oracle's window searches against an independent pure-Python restatement of the same reference
functions, and the product's AssignFeaturesToGrid mirror against the oracle's.  The reference ships no tests for these
functions, so two independent restatements agreeing is the pin available here."""
import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from oracle import orb_oracle_py as orc

import proj_util as pu
import pyref_project as py


@pytest.mark.parametrize("seed,cluster", [(0, False), (1, True), (2, False)])
def test_projection_map_oracle_vs_python(seed, cluster):
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(260, rng, stereo=seed != 2, cluster=cluster)
    blocked = (rng.random(260) < 0.15).astype(np.uint8)
    g, og = pu.make_grids(fa, blocked, orb, orc)
    assert np.array_equal(g.cell_offsets, og.off) and np.array_equal(g.cell_features, og.feat[:len(g.cell_features)])
    mp = pu.map_points_for(fa, 400, rng)
    th = 1.0 if seed == 0 else 3.0
    n, owner = orc.search_projection_map(og, th=th, nnratio=0.8, **mp)
    pn, powner = py.search_projection_map(py.PyFrame(fa, pu.SCALE, blocked), mp, th, 0.8)
    assert n == pn and np.array_equal(owner, powner)
    assert n > 50


@pytest.mark.parametrize("seed,mono,tz,ori", [(3, True, 0.0, True), (4, False, 0.5, True), (5, False, -0.5, False), (6, False, 0.0, True)])
def test_projection_frame_oracle_vs_python(seed, mono, tz, ori):
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(300, rng, stereo=not mono, cluster=seed == 6)
    blocked = (rng.random(300) < 0.1).astype(np.uint8)
    _, og = pu.make_grids(fa, blocked, orb, orc)
    lf = pu.last_frame_for(fa, 350, rng, tz=tz)
    mbf, mb, th = 40.0, 40.0 / lf["fx"], 15.0 if mono else 7.0
    n, owner = orc.search_projection_frame(og, lf["Tcw"], lf["Tlw"], lf["fx"], lf["fy"], lf["cx"], lf["cy"], mbf, mb, lf["has_point"],
                                           lf["world"], lf["octave"], lf["angle"], lf["desc"], lf["claims"], th, mono, ori)
    pn, powner = py.search_projection_frame(py.PyFrame(fa, pu.SCALE, blocked), lf, mbf, mb, th, mono, ori)
    assert n == pn and np.array_equal(owner, powner)
    assert (owner >= 0).sum() > 30


@pytest.mark.parametrize("seed,window,ori", [(7, 10, True), (8, 100, True), (9, 30, False)])
def test_initialization_oracle_vs_python(seed, window, ori):
    rng = np.random.default_rng(seed)
    fa = pu.frame_arrays(300, rng, stereo=False)
    fa["octave"][rng.random(300) < 0.5] = 0
    _, og = pu.make_grids(fa, None, orb, orc)
    f1 = pu.init_frame1_for(fa, 320, rng)
    prev = f1["prev"].copy()
    n, m12 = orc.search_initialization(og, f1["desc1"], f1["octave1"], f1["angle1"], prev, window, 0.9, ori)
    pn, pm12, pprev = py.search_initialization(py.PyFrame(fa, pu.SCALE), f1, window, 0.9, ori)
    assert n == pn and np.array_equal(m12, pm12) and np.array_equal(prev, pprev)
    assert n > 20
