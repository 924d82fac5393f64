"""Synthetic scenes for the projection / window searches.  The same arrays feed the oracle and the product; both
build the feature grid with their own restatement of Frame::AssignFeaturesToGrid (the test compares the two grids as well)."""
import numpy as np

SCALE = (1.2 ** np.arange(8)).astype(np.float32)


def flip(d, nbits, rng):
    d = d.copy()
    for b in rng.choice(256, nbits, replace=False):
        d[b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def frame_arrays(n, rng, w=640, h=480, stereo=True, cluster=False):
    """Random undistorted keypoints of a frame: positions (a few slightly outside the bounds, as undistortion can produce),
    octaves with the extractor's geometric distribution, angles, mvuRight (negative for monocular points)."""
    if cluster:                                   # dense blobs: many features per window, heavy claim contention
        cx = rng.uniform(40, w - 40, max(n // 60, 1))
        cy = rng.uniform(40, h - 40, len(cx))
        k = rng.integers(0, len(cx), n)
        x = (cx[k] + rng.normal(0, 9, n)).astype(np.float32)
        y = (cy[k] + rng.normal(0, 9, n)).astype(np.float32)
    else:
        x = rng.uniform(-3, w + 3, n).astype(np.float32)
        y = rng.uniform(-3, h + 3, n).astype(np.float32)
    p = 1.0 / SCALE.astype(np.float64)
    octave = rng.choice(8, n, p=p / p.sum()).astype(np.int32)
    angle = rng.uniform(0, 360, n).astype(np.float32)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if stereo:
        uright = np.where(rng.random(n) < 0.7, x - rng.uniform(2, 40, n).astype(np.float32), np.float32(-1)).astype(np.float32)
    else:
        uright = None
    return dict(desc=desc, x=x, y=y, octave=octave, angle=angle, uright=uright, bounds=(0.0, 0.0, float(w), float(h)))


def make_grids(fa, blocked, orb, orc):
    kw = dict(angle=fa["angle"], uright=fa["uright"], blocked=blocked)
    g = orb.matcher.GridView(fa["desc"], fa["x"], fa["y"], fa["octave"], SCALE, fa["bounds"], **kw)
    og = orc.Grid(fa["desc"], fa["x"], fa["y"], fa["octave"], SCALE, fa["bounds"], **kw)
    return g, og


def map_points_for(fa, n_points, rng, contention=0.3):
    """Map points as Frame::isInFrustum leaves them: most sit near a feature with a descriptor a few bits away from it (several
    points may chase the same feature), the rest are distractors."""
    n = len(fa["x"])
    tgt = rng.integers(0, n, n_points)
    dup = rng.random(n_points) < contention
    tgt[dup] = tgt[rng.integers(0, n_points, dup.sum())]          # contention: reuse another point's target
    desc = np.empty((n_points, 32), np.uint8)
    for i, t in enumerate(tgt):
        nb = int(rng.choice([0, 3, 10, 30, 60, 95, 100, 101, 128]))
        desc[i] = flip(fa["desc"][t], nb, rng)
    level = np.clip(fa["octave"][tgt] + rng.integers(0, 2, n_points), 0, 7).astype(np.int32)
    proj_x = (fa["x"][tgt] + rng.normal(0, 1.5, n_points)).astype(np.float32)
    proj_y = (fa["y"][tgt] + rng.normal(0, 1.5, n_points)).astype(np.float32)
    ur = fa["uright"][tgt] if fa["uright"] is not None else np.full(n_points, -1, np.float32)
    proj_xr = (np.where(ur > 0, ur, proj_x - 10) + rng.normal(0, 2.0, n_points)).astype(np.float32)
    view_cos = rng.choice(np.array([0.9, 0.997, 0.998, 0.9981, 0.9995], np.float32), n_points)
    in_view = (rng.random(n_points) < 0.9).astype(np.uint8)
    claims = (rng.random(n_points) < 0.8).astype(np.uint8)
    return dict(in_view=in_view, proj_x=proj_x, proj_y=proj_y, proj_xr=proj_xr, level=level, view_cos=view_cos, desc=desc, claims=claims)


def rot_small(rng, deg=1.5):
    a = np.deg2rad(rng.uniform(-deg, deg, 3))
    cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def last_frame_for(fa, n_last, rng, fx=520.0, fy=520.0, cx=320.0, cy=240.0, tz=0.0):
    """LastFrame features with map points that project (through Tcw_cur) near features of the current frame."""
    n = len(fa["x"])
    R = rot_small(rng).astype(np.float32)
    t = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), tz], np.float32)
    Tcw = np.concatenate([R, t[:, None]], 1).astype(np.float32)
    Rl = rot_small(rng).astype(np.float32)
    Tlw = np.concatenate([Rl, np.array([[0.0], [0.0], [0.0]], np.float32)], 1).astype(np.float32)
    tgt = rng.integers(0, n, n_last)
    dup = rng.random(n_last) < 0.3
    tgt[dup] = tgt[rng.integers(0, n_last, dup.sum())]
    z = rng.uniform(1.0, 8.0, n_last)
    u = fa["x"][tgt].astype(np.float64) + rng.normal(0, 2.0, n_last)
    v = fa["y"][tgt].astype(np.float64) + rng.normal(0, 2.0, n_last)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    behind = rng.random(n_last) < 0.03
    pc[behind, 2] *= -1                                               # some points behind the camera
    world = ((pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)   # R^T (pc - t)
    desc = np.empty((n_last, 32), np.uint8)
    for i, tt in enumerate(tgt):
        desc[i] = flip(fa["desc"][tt], int(rng.choice([0, 3, 10, 30, 60, 95, 100, 101, 128])), rng)
    octave = np.clip(fa["octave"][tgt] + rng.integers(-1, 2, n_last), 0, 7).astype(np.int32)
    angle = ((fa["angle"][tgt] + rng.choice([0.0, 0.0, 0.0, 45.0, 200.0], n_last) + rng.normal(0, 3, n_last)) % 360).astype(np.float32)
    has_point = (rng.random(n_last) < 0.85).astype(np.uint8)
    claims = (rng.random(n_last) < 0.8).astype(np.uint8)
    return dict(Tcw=Tcw, Tlw=Tlw, fx=fx, fy=fy, cx=cx, cy=cy, has_point=has_point, world=world, octave=octave, angle=angle, desc=desc,
                claims=claims)


def init_frame1_for(fa2, n1, rng, shift=(4.0, -3.0)):
    """F1 of SearchForInitialization: features displaced from F2's by a small motion, several F1 features per F2 feature."""
    n2 = len(fa2["x"])
    tgt = rng.integers(0, n2, n1)
    desc1 = np.empty((n1, 32), np.uint8)
    for i, t in enumerate(tgt):
        desc1[i] = flip(fa2["desc"][t], int(rng.choice([0, 2, 8, 20, 40, 49, 50, 51, 70, 128])), rng)
    x1 = (fa2["x"][tgt] - shift[0] + rng.normal(0, 2.0, n1)).astype(np.float32)
    y1 = (fa2["y"][tgt] - shift[1] + rng.normal(0, 2.0, n1)).astype(np.float32)
    octave1 = np.where(rng.random(n1) < 0.7, 0, fa2["octave"][tgt]).astype(np.int32)
    angle1 = ((fa2["angle"][tgt] + rng.choice([0.0, 0.0, 0.0, 90.0], n1) + rng.normal(0, 3, n1)) % 360).astype(np.float32)
    prev = np.ascontiguousarray(np.stack([x1, y1], 1), np.float32)
    return dict(desc1=desc1, octave1=octave1, angle1=angle1, prev=prev)


def kf_points_for(fa, n_pts, rng, fx=520.0, fy=520.0, cx=320.0, cy=240.0, sim_scale=1.0):
    """Map points seen from a pose (Tcw, optionally a similarity s*[R|t]) that project near features of `fa`, with the
    scale-invariance distances MapPoint::UpdateNormalAndDepth would give them (so PredictScale lands near the feature's octave)."""
    n = len(fa["x"])
    R = rot_small(rng, 3.0).astype(np.float32)
    t = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1)], np.float32)
    T = np.concatenate([R, t[:, None]], 1).astype(np.float32)
    S = (T * np.float32(sim_scale)).astype(np.float32)
    Tn = (S.astype(np.float64) / np.sqrt((S[0, :3].astype(np.float64) ** 2).sum()))           # what the search decomposes back
    tgt = rng.integers(0, n, n_pts)
    dup = rng.random(n_pts) < 0.3
    tgt[dup] = tgt[rng.integers(0, n_pts, dup.sum())]
    z = rng.uniform(1.0, 8.0, n_pts)
    u = fa["x"][tgt].astype(np.float64) + rng.normal(0, 2.0, n_pts)
    v = fa["y"][tgt].astype(np.float64) + rng.normal(0, 2.0, n_pts)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    pc[rng.random(n_pts) < 0.03, 2] *= -1
    world = ((pc - Tn[:, 3]) @ Tn[:, :3]).astype(np.float32)
    Ow = -Tn[:, :3].T @ Tn[:, 3]
    PO = world.astype(np.float64) - Ow
    dist = np.linalg.norm(PO, axis=1)
    lvl = np.clip(fa["octave"][tgt] + rng.integers(-1, 2, n_pts), 0, 7)
    mf_max = (dist * SCALE[lvl].astype(np.float64) * rng.uniform(0.85, 1.0, n_pts)).astype(np.float32)
    far = rng.random(n_pts) < 0.05
    mf_max[far] *= np.float32(0.3)                                     # outside the scale-invariance range
    mf_min = (mf_max / SCALE[7]).astype(np.float32)
    normal = PO / dist[:, None]
    side = rng.random(n_pts) < 0.1
    normal[side] = np.roll(normal[side], 1, axis=1) * np.array([1, -1, 1])            # oblique views (viewing-angle test)
    desc = np.stack([flip(fa["desc"][tt], int(rng.choice([0, 3, 10, 30, 49, 50, 51, 95, 100, 101, 128])), rng) for tt in tgt])
    angle = ((fa["angle"][tgt] + rng.choice([0.0, 0.0, 0.0, 45.0, 200.0], n_pts) + rng.normal(0, 3, n_pts)) % 360).astype(np.float32)
    state = rng.choice(np.array([0, 1, 1, 1, 1, 1, 2, 3], np.uint8), n_pts)       # 0 none, 1 good, 2 bad, 3 already found
    return dict(T=T, S=S, fx=fx, fy=fy, cx=cx, cy=cy, world=world, mf_max=mf_max, mf_min=mf_min, normal=normal.astype(np.float32),
                desc=desc, angle=angle, state=state)


def points_for_transform(fa, n_pts, rng, A, b, dist_from_camera, fx=520.0, fy=520.0, cx=320.0, cy=240.0, tgt=None):
    """World points P with A @ P + b landing (through K) near features of `fa`; scale-invariance distances consistent with the
    distance the search will measure (|A P + b| for SearchBySim3, |P - Ow| otherwise)."""
    n = len(fa["x"])
    if tgt is None:
        tgt = rng.integers(0, n, n_pts)
    z = rng.uniform(1.0, 8.0, n_pts)
    u = fa["x"][tgt].astype(np.float64) + rng.normal(0, 1.5, n_pts)
    v = fa["y"][tgt].astype(np.float64) + rng.normal(0, 1.5, n_pts)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    pc[rng.random(n_pts) < 0.03, 2] *= -1
    world = np.linalg.solve(A.astype(np.float64), (pc - b.astype(np.float64)).T).T.astype(np.float32)
    if dist_from_camera:
        dist = np.linalg.norm(pc, axis=1)
        PO = pc
    else:
        sc = np.sqrt((A[0].astype(np.float64) ** 2).sum())
        Rn, tn = A.astype(np.float64) / sc, b.astype(np.float64) / sc
        PO = world.astype(np.float64) - (-Rn.T @ tn)
        dist = np.linalg.norm(PO, axis=1)
    lvl = np.clip(fa["octave"][tgt] + rng.integers(0, 2, n_pts), 0, 7)
    mf_max = (dist * SCALE[lvl].astype(np.float64) * rng.uniform(0.88, 1.0, n_pts)).astype(np.float32)
    mf_max[rng.random(n_pts) < 0.05] *= np.float32(0.3)
    mf_min = (mf_max / SCALE[7]).astype(np.float32)
    normal = (PO / np.maximum(dist, 1e-9)[:, None])
    side = rng.random(n_pts) < 0.1
    normal[side] = np.roll(normal[side], 1, axis=1) * np.array([1, -1, 1])
    desc = np.stack([flip(fa["desc"][tt], int(rng.choice([0, 3, 10, 30, 49, 50, 51, 95, 100, 101, 128])), rng) for tt in tgt])
    return dict(world=world, mf_max=mf_max, mf_min=mf_min, normal=normal.astype(np.float32), desc=desc, tgt=tgt)
