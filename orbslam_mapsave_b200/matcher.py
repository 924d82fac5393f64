"""Python mirror of ORB_SLAM2::ORBmatcher's Hamming searches (include/ORBmatcher.h:37-101) over the C ABI.

KeyFrame / Frame are represented by `View`: exactly the fields the reference functions read
(mDescriptors, MapPoint flags, mvKeysUn/mvKeys, mvuRight, mFeatVec)."""
import ctypes as C

import numpy as np

from . import capi


class FeatureVector:
    """DBoW2::FeatureVector flattened (node ids ascending, CSR offsets, feature indices ascending inside a node)."""

    def __init__(self, node_of_feature=None, ids=None, off=None, feat=None):
        if node_of_feature is not None:
            node_of_feature = np.asarray(node_of_feature, np.int64)
            order = np.argsort(node_of_feature, kind="stable")
            ids, counts = np.unique(node_of_feature, return_counts=True)
            off = np.concatenate([[0], np.cumsum(counts)])
            feat = order
        self.ids = np.ascontiguousarray(ids, np.int32)
        self.off = np.ascontiguousarray(off, np.int32)
        self.feat = np.ascontiguousarray(feat, np.int32)

    def c(self):
        return capi.FeatVecC(len(self.ids), capi._p(self.ids), capi._p(self.off), capi._p(self.feat))


class View:
    def __init__(self, desc, fv, angle, flag=None, x=None, y=None, octave=None, uright=None):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.n = len(self.desc)
        self.fv = fv
        self.angle = f32(angle)
        self.flag = None if flag is None else np.ascontiguousarray(flag, np.uint8)
        self.x, self.y, self.uright = f32(x), f32(y), f32(uright)
        self.octave = None if octave is None else np.ascontiguousarray(octave, np.int32)

    def c(self):
        return capi.ViewC(self.n, capi._p(self.desc), capi._p(self.flag), capi._p(self.angle), capi._p(self.x), capi._p(self.y),
                          capi._p(self.octave), capi._p(self.uright), self.fv.c())


class GridView:
    """The Frame / KeyFrame fields the window searches read (include/Frame.h:98,141-197): undistorted keypoints, descriptors,
    mvuRight, the image bounds and the feature grid mGrid flattened to CSR (cell = ix * rows + iy, push order inside a cell)."""

    def __init__(self, desc, x, y, octave, scale_factors, bounds, angle=None, uright=None, blocked=None, grid_cols=64, grid_rows=48):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.n = len(self.desc)
        self.x, self.y, self.angle, self.uright = f32(x), f32(y), f32(angle), f32(uright)
        self.octave = np.ascontiguousarray(octave, np.int32)
        self.blocked = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
        self.scale_factors = f32(scale_factors)
        self.min_x, self.min_y, self.max_x, self.max_y = (np.float32(v) for v in bounds)
        self.grid_cols, self.grid_rows = int(grid_cols), int(grid_rows)
        # mfGridElementWidthInv / HeightInv (src/Frame.cc:101-102): float(cols) / float(maxX - minX)
        self.inv_w = np.float32(self.grid_cols) / np.float32(self.max_x - self.min_x)
        self.inv_h = np.float32(self.grid_rows) / np.float32(self.max_y - self.min_y)
        self._assign_features_to_grid()

    def _assign_features_to_grid(self):
        """Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:341-356, 500-510): cell = round((pt - min) * inv)."""
        cround = lambda v: np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), -np.floor(-v.astype(np.float64) + 0.5)).astype(np.int64)
        px = cround((self.x - self.min_x) * self.inv_w)
        py = cround((self.y - self.min_y) * self.inv_h)
        ok = (px >= 0) & (px < self.grid_cols) & (py >= 0) & (py < self.grid_rows)
        cell = px[ok] * self.grid_rows + py[ok]
        idx = np.nonzero(ok)[0]
        order = np.argsort(cell, kind="stable")
        counts = np.bincount(cell, minlength=self.grid_cols * self.grid_rows)
        self.cell_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.cell_features = np.ascontiguousarray(idx[order], np.int32)

    def c(self):
        p = capi._p
        return capi.GridViewC(self.n, p(self.desc), p(self.x), p(self.y), p(self.octave), p(self.angle), p(self.uright), p(self.blocked),
                              self.grid_cols, self.grid_rows, self.min_x, self.min_y, self.max_x, self.max_y, self.inv_w, self.inv_h,
                              p(self.cell_offsets), p(self.cell_features), p(self.scale_factors), len(self.scale_factors))


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = capi.TH_HIGH, capi.TH_LOW, capi.HISTO_LENGTH     # src/ORBmatcher.cc:37-39

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio, self.mbCheckOrientation, self.device = float(nnratio), bool(checkOri), int(device)

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        """Hamming distance of descriptor rows (src/ORBmatcher.cc:1650-1666); a, b: (32,) or (n, 32) uint8."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.zeros(len(a), np.int32)
        capi.check(capi.lib().orbm_descriptor_distance(capi._p(a), capi._p(b), len(a), capi._p(out), device))
        return int(out[0]) if len(out) == 1 else out

    def hamming_top2(self, q, db):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        db = np.ascontiguousarray(db, np.uint8).reshape(-1, 32)
        bi, bd, sd = (np.zeros(len(q), np.int32) for _ in range(3))
        capi.check(capi.lib().orbm_hamming_top2(capi._p(q), len(q), capi._p(db), len(db), capi._p(bi), capi._p(bd), capi._p(sd),
                                                self.device))
        return bi, bd, sd

    def SearchByBoW(self, kf, other, kf_kf=False):
        """kf_kf=False: SearchByBoW(KeyFrame*, Frame&) -> (nmatches, match21[F.N]);
        kf_kf=True: SearchByBoW(KeyFrame*, KeyFrame*) -> (nmatches, match12[KF1.N])."""
        v1, v2 = kf.c(), other.c()
        n = C.c_int()
        if kf_kf:
            m = np.zeros(max(kf.n, 1), np.int32)
            capi.check(capi.lib().orbm_search_by_bow_kf_kf(C.byref(v1), C.byref(v2), self.mfNNratio, int(self.mbCheckOrientation),
                                                           capi._p(m), C.byref(n), self.device))
            return n.value, m[:kf.n]
        m = np.zeros(max(other.n, 1), np.int32)
        capi.check(capi.lib().orbm_search_by_bow_kf_frame(C.byref(v1), C.byref(v2), self.mfNNratio, int(self.mbCheckOrientation),
                                                          capi._p(m), C.byref(n), self.device))
        return n.value, m[:other.n]

    def SearchByBoWBatch(self, anchor, others, kf_kf=False):
        """One view against many in one call (the candidate loops of Tracking::Relocalization / LoopClosing::ComputeSim3).
        kf_kf=False: SearchByBoW(KeyFrame* others[i], Frame& anchor); kf_kf=True: SearchByBoW(KeyFrame* anchor, KeyFrame* others[i]).
        Returns (nmatches[len(others)], matches[len(others), anchor.N])."""
        k = len(others)
        arr = (capi.ViewC * max(k, 1))(*[o.c() for o in others])
        va = anchor.c()
        m = np.zeros((max(k, 1), max(anchor.n, 1)), np.int32)
        nm = np.zeros(max(k, 1), np.int32)
        out = np.zeros(max(k * anchor.n, 1), np.int32)
        capi.check(capi.lib().orbm_search_by_bow_batch(C.byref(va), arr, k, int(kf_kf), self.mfNNratio, int(self.mbCheckOrientation),
                                                       capi._p(out), capi._p(nm), self.device))
        if anchor.n:
            m = out[:k * anchor.n].reshape(k, anchor.n)
        return nm[:k], m[:k, :anchor.n]

    def SearchForTriangulationBatch(self, kf1, neighbours, F12s, epipoles, scale_factors2, level_sigma2_2, bOnlyStereo):
        """SearchForTriangulation(kf1, neighbours[i], F12s[i], ...) for every neighbour in one call (LocalMapping::CreateNewMapPoints).
        scale_factors2 / level_sigma2_2: (len(neighbours), n_levels).  Returns (nmatches[k], [pairs_i])."""
        k = len(neighbours)
        arr = (capi.ViewC * max(k, 1))(*[o.c() for o in neighbours])
        va = kf1.c()
        F = np.ascontiguousarray(F12s, np.float32).reshape(max(k, 1), 9) if k else np.zeros((1, 9), np.float32)
        ep = np.ascontiguousarray(epipoles, np.float32).reshape(-1, 2) if k else np.zeros((1, 2), np.float32)
        sf2 = np.ascontiguousarray(scale_factors2, np.float32).reshape(max(k, 1), -1)
        s2 = np.ascontiguousarray(level_sigma2_2, np.float32).reshape(max(k, 1), -1)
        pairs = np.zeros((max(k, 1), max(kf1.n, 1), 2), np.int32)
        flat = np.zeros(max(2 * k * kf1.n, 1), np.int32)
        npairs, nm = np.zeros(max(k, 1), np.int32), np.zeros(max(k, 1), np.int32)
        capi.check(capi.lib().orbm_search_for_triangulation_batch(C.byref(va), arr, k, capi._p(F), capi._p(ep), capi._p(sf2), capi._p(s2),
                                                                  sf2.shape[1], int(bOnlyStereo), int(self.mbCheckOrientation),
                                                                  capi._p(flat), capi._p(npairs), capi._p(nm), self.device))
        if kf1.n:
            pairs = flat[:2 * k * kf1.n].reshape(k, kf1.n, 2)
        return nm[:k], [pairs[i, :npairs[i]].copy() for i in range(k)]

    def SearchForTriangulation(self, kf1, kf2, F12, ex, ey, scale_factors2, level_sigma2_2, bOnlyStereo):
        v1, v2 = kf1.c(), kf2.c()
        F12 = np.ascontiguousarray(F12, np.float32).reshape(9)
        sf2 = np.ascontiguousarray(scale_factors2, np.float32)
        s2 = np.ascontiguousarray(level_sigma2_2, np.float32)
        pairs = np.zeros((max(kf1.n, 1), 2), np.int32)
        npairs, nm = C.c_int(), C.c_int()
        capi.check(capi.lib().orbm_search_for_triangulation(C.byref(v1), C.byref(v2), capi._p(F12), float(ex), float(ey), capi._p(sf2),
                                                            capi._p(s2), len(sf2), int(bOnlyStereo), int(self.mbCheckOrientation),
                                                            capi._p(pairs), C.byref(npairs), C.byref(nm), self.device))
        return nm.value, pairs[:npairs.value].copy()

    def SearchByProjectionMapPoints(self, frame, in_view, proj_x, proj_y, proj_xr, level, view_cos, desc, claims, th=1.0):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129).  Arrays per map point as left by
        Frame::isInFrustum.  Returns (nmatches, owner[F.N]): the map point stored in F.mvpMapPoints[idx], -1 = untouched."""
        u8 = lambda a: np.ascontiguousarray(a, np.uint8)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        in_view, claims, desc = u8(in_view), u8(claims), u8(desc).reshape(-1, 32)
        proj_x, proj_y, proj_xr, view_cos = f32(proj_x), f32(proj_y), f32(proj_xr), f32(view_cos)
        level = np.ascontiguousarray(level, np.int32)
        g = frame.c()
        owner = np.zeros(max(frame.n, 1), np.int32)
        n = C.c_int()
        capi.check(capi.lib().orbm_search_by_projection_map(C.byref(g), len(in_view), capi._p(in_view), capi._p(proj_x), capi._p(proj_y),
                                                            capi._p(proj_xr), capi._p(level), capi._p(view_cos), capi._p(desc),
                                                            capi._p(claims), float(th), self.mfNNratio, capi._p(owner), C.byref(n),
                                                            self.device))
        return n.value, owner[:frame.n]

    def SearchByProjectionFrame(self, cur, Tcw_cur, Tcw_last, fx, fy, cx, cy, mbf, mb, has_point, world, octave, angle, desc, claims,
                                th, bMono):
        """SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (src/ORBmatcher.cc:1331-1463).
        Returns (nmatches, owner[cur.N]): LastFrame feature index, -1 = untouched, -2 = set to NULL by the rotation cull."""
        u8 = lambda a: np.ascontiguousarray(a, np.uint8)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        has_point, claims, desc = u8(has_point), u8(claims), u8(desc).reshape(-1, 32)
        world, angle = f32(world).reshape(-1, 3), f32(angle)
        octave = np.ascontiguousarray(octave, np.int32)
        Tc, Tl = f32(Tcw_cur).reshape(-1)[:12].copy(), f32(Tcw_last).reshape(-1)[:12].copy()
        g = cur.c()
        owner = np.zeros(max(cur.n, 1), np.int32)
        n = C.c_int()
        capi.check(capi.lib().orbm_search_by_projection_frame(C.byref(g), capi._p(Tc), capi._p(Tl), fx, fy, cx, cy, mbf, mb, len(has_point),
                                                              capi._p(has_point), capi._p(world), capi._p(octave), capi._p(angle),
                                                              capi._p(desc), capi._p(claims), float(th), int(bMono),
                                                              int(self.mbCheckOrientation), capi._p(owner), C.byref(n), self.device))
        return n.value, owner[:cur.n]

    def SearchByProjectionFrameBatch(self, jobs):
        """Many-frame form (orbm_search_by_projection_frame_batch): `jobs` is a list of argument tuples of SearchByProjectionFrame, one per
        independent (CurrentFrame, LastFrame) pair.  Returns [(nmatches, owner), ...], identical to calling the single form per job."""
        u8 = lambda a: np.ascontiguousarray(a, np.uint8)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        J = (capi.FrameSearchJobC * max(len(jobs), 1))()
        keep, owners = [], []
        for k, (cur, Tcw_cur, Tcw_last, fx, fy, cx, cy, mbf, mb, has_point, world, octave, angle, desc, claims, th, bMono) in enumerate(jobs):
            has_point, claims, desc = u8(has_point), u8(claims), u8(desc).reshape(-1, 32)
            world, angle = f32(world).reshape(-1, 3), f32(angle)
            octave = np.ascontiguousarray(octave, np.int32)
            Tc, Tl = f32(Tcw_cur).reshape(-1)[:12].copy(), f32(Tcw_last).reshape(-1)[:12].copy()
            g = cur.c()
            owner = np.zeros(max(cur.n, 1), np.int32)
            keep.append((g, has_point, claims, desc, world, angle, octave, Tc, Tl))
            owners.append((owner, cur.n))
            j = J[k]
            j.cur = C.pointer(g)
            j.Tcw_cur, j.Tcw_last = capi._p(Tc), capi._p(Tl)
            j.fx, j.fy, j.cx, j.cy, j.mbf, j.mb = fx, fy, cx, cy, mbf, mb
            j.n_last = len(has_point)
            j.has_point, j.world, j.octave, j.angle = capi._p(has_point), capi._p(world), capi._p(octave), capi._p(angle)
            j.desc, j.claims, j.th, j.mono, j.owner = capi._p(desc), capi._p(claims), float(th), int(bMono), capi._p(owner)
        capi.check(capi.lib().orbm_search_by_projection_frame_batch(C.cast(J, C.c_void_p), len(jobs), int(self.mbCheckOrientation), self.device))
        return [(J[k].n_matches, owners[k][0][:owners[k][1]]) for k in range(len(jobs))]

    def SearchForInitialization(self, f2, desc1, octave1, angle1, prev_matched, windowSize=10):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:408-523).
        prev_matched: (n1, 2) float32, updated in place.  Returns (nmatches, vnMatches12)."""
        desc1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32)
        octave1 = np.ascontiguousarray(octave1, np.int32)
        angle1 = np.ascontiguousarray(angle1, np.float32)
        assert prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous and prev_matched.shape == (len(desc1), 2)
        g = f2.c()
        m = np.zeros(max(len(desc1), 1), np.int32)
        n = C.c_int()
        capi.check(capi.lib().orbm_search_for_initialization(C.byref(g), len(desc1), capi._p(desc1), capi._p(octave1), capi._p(angle1),
                                                             capi._p(prev_matched), int(windowSize), self.mfNNratio,
                                                             int(self.mbCheckOrientation), capi._p(m), C.byref(n), self.device))
        return n.value, m[:len(desc1)]

    def SearchWindows(self, target, active, u, v, r, min_level, max_level, desc, angle=None, th_dist=capi.TH_HIGH):
        """Core of SearchByProjection(Frame, KeyFrame, ...) / (KeyFrame, Scw, ...) (src/ORBmatcher.cc:1465-1602, 293-406) on
        explicit windows.  Returns (nmatches, owner[target.N]) with -1 = untouched, -2 = culled to NULL."""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        active, desc = np.ascontiguousarray(active, np.uint8), np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        u, v, r, min_level, max_level = f32(u), f32(v), f32(r), i32(min_level), i32(max_level)
        ori = self.mbCheckOrientation and angle is not None
        angle = f32(angle) if ori else None
        g = target.c()
        owner = np.zeros(max(target.n, 1), np.int32)
        n = C.c_int()
        capi.check(capi.lib().orbm_search_windows(C.byref(g), len(active), capi._p(active), capi._p(u), capi._p(v), capi._p(r),
                                                  capi._p(min_level), capi._p(max_level), capi._p(desc), capi._p(angle), int(th_dist),
                                                  int(ori), capi._p(owner), C.byref(n), self.device))
        return n.value, owner[:target.n]

    def SearchWindowsBest(self, target, active, u, v, r, min_level, max_level, desc, ur=None, inv_level_sigma2=None, th_dist=capi.TH_LOW):
        """Independent window search (SearchBySim3 / the candidate loops of Fuse, src/ORBmatcher.cc:828-1329): per query the
        nearest feature in its window, -1 if none within th_dist; with ur + inv_level_sigma2 Fuse's chi-square gate applies."""
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        active, desc = np.ascontiguousarray(active, np.uint8), np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        u, v, r, ur, is2, min_level, max_level = f32(u), f32(v), f32(r), f32(ur), f32(inv_level_sigma2), i32(min_level), i32(max_level)
        g = target.c()
        best = np.zeros(max(len(active), 1), np.int32)
        capi.check(capi.lib().orbm_search_windows_best(C.byref(g), len(active), capi._p(active), capi._p(u), capi._p(v), capi._p(r),
                                                       capi._p(min_level), capi._p(max_level), capi._p(desc), capi._p(ur), capi._p(is2),
                                                       int(th_dist), capi._p(best), self.device))
        return best[:len(active)]

    @staticmethod
    def ComputeThreeMaxima(histo, device=0):
        histo = np.ascontiguousarray(histo, np.int32)
        ind = np.zeros(3, np.int32)
        capi.check(capi.lib().orbm_three_maxima(capi._p(histo), len(histo), capi._p(ind), device))
        return tuple(int(v) for v in ind)


def allpairs_multi(desc, th_low=capi.TH_LOW, ratio=0.75, devices=(0,), want_best=False):
    """All-pairs keyframe matching (BASELINE config 4) on the listed GPUs from this one process (orbm_allpairs_multi).
    desc: (n_kf, per_kf, 32) uint8 host array.  Returns the (n_kf, n_kf) uint16 table of ratio-test matches per (query keyframe, keyframe)
    and, with want_best, per query descriptor the keyframe holding its nearest descriptor and that distance."""
    desc = np.ascontiguousarray(desc, np.uint8)
    n_kf, per_kf = desc.shape[0], desc.shape[1]
    dev = np.ascontiguousarray(devices, np.int32)
    counts = np.zeros((n_kf, n_kf), np.uint16)
    bk = np.zeros((n_kf, per_kf), np.int32) if want_best else None
    bd = np.zeros((n_kf, per_kf), np.int32) if want_best else None
    capi.check(capi.lib().orbm_allpairs_multi(capi._p(desc), n_kf, per_kf, int(th_low), float(ratio), capi._p(dev), len(dev), capi._p(counts),
                                              capi._p(bk), capi._p(bd)))
    return (counts, bk, bd) if want_best else counts


def distinctive_descriptors(desc, offsets, device=0):
    """MapPoint::ComputeDistinctiveDescriptors for a batch of map points (src/MapPoint.cc:483-548): index, inside each set
    desc[offsets[s]:offsets[s+1]], of the descriptor with the least median distance to the set."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros(max(len(offsets) - 1, 1), np.int32)
    capi.check(capi.lib().orbm_distinctive_descriptors(capi._p(desc), capi._p(offsets), len(offsets) - 1, capi._p(out), device))
    return out[:len(offsets) - 1]


def popc_peak(device=0):
    v, clk = C.c_double(), C.c_double()
    capi.check(capi.lib().orbm_popc_peak(device, C.byref(v), C.byref(clk)))
    return v.value, clk.value
