"""Test infrastructure: an independent (pure Python, struct.pack) statement of the byte stream System::SaveMap produces
(src/System.cc:565-574: boost::archive::binary_oarchive with no_header, `oa << mpMap`).  Written procedurally in the order of
Map::save (src/Map.cc:31-74), MapPoint::save (src/MapPoint.cc:59-140) and KeyFrame::save (src/KeyFrame.cc:86-306) so that it
shares no code with the product's reader / writer (orbslam_mapsave_b200/csrc/orb_map.cpp).  Boost is not available in this
image: the class-info framing (tracking byte + 32-bit version the first time a class appears, 64-bit collection sizes, 32-bit
item versions for non-arithmetic vectors, array-optimised arithmetic vectors) is restated from Boost's archive format.

A logical map is a dict: {"mappoints": [...], "keyframes": [...], "origins": [...], "max_kf_id": int}; see make_random_map().
"""
import struct

import numpy as np


class _Stream:
    def __init__(self):
        self.b = bytearray()
        self.seen = set()

    def cls(self, name):
        # first appearance of a class saved by value: tracking_type(bool) = 0, version_type(uint32) = 0
        if name not in self.seen:
            self.seen.add(name)
            self.b += struct.pack("<BI", 0, 0)

    def i32(self, v):
        self.b += struct.pack("<i", int(v))

    def u32(self, v):
        self.b += struct.pack("<I", int(v))

    def u64(self, v):
        self.b += struct.pack("<Q", int(v))

    def i64(self, v):
        self.b += struct.pack("<q", int(v))

    def f32(self, v):
        self.b += np.float32(v).tobytes()

    def f64(self, v):
        self.b += struct.pack("<d", float(v))

    def boolean(self, v):
        self.b += b"\x01" if v else b"\x00"

    def mat(self, m):
        """m: None (empty Mat) or a 2-D numpy array of uint8 / float32."""
        self.cls("Mat")
        if m is None:
            self.i32(0), self.i32(0), self.u64(0), self.u64(0)      # cv::Mat(): dims == 0, so elemSize() is 0 and type() is CV_8UC1
            return
        m = np.ascontiguousarray(m)
        assert m.ndim == 2 and m.dtype in (np.uint8, np.float32)
        self.i32(m.shape[1]), self.i32(m.shape[0])
        self.u64(m.dtype.itemsize), self.u64(0 if m.dtype == np.uint8 else 5)
        self.b += m.tobytes()

    def keypoints(self, kps):
        self.cls("vector<KeyPoint>")
        self.u64(len(kps)), self.u32(0)
        for k in kps:
            self.cls("KeyPoint")
            self.f32(k["angle"]), self.i32(k["class_id"]), self.i32(k["octave"]), self.f32(k["response"]), self.f32(k["response"])
            self.f32(k["x"]), self.f32(k["y"])

    def vec_f32(self, v):
        # no class preamble: vectors of arithmetic types are `object_serializable` (BOOST_SERIALIZATION_COLLECTION_TRAITS)
        v = np.ascontiguousarray(v, np.float32)
        self.u64(len(v))
        self.b += v.tobytes()

    def vec_i32(self, v):
        v = np.ascontiguousarray(v, np.int32)
        self.u64(len(v))
        self.b += v.tobytes()

    def grid(self, g):
        self.cls("vector<vector<vector<size_t>>>")
        self.u64(len(g)), self.u32(0)
        for col in g:
            self.cls("vector<vector<size_t>>")
            self.u64(len(col)), self.u32(0)
            for cell in col:
                self.u64(len(cell))
                self.b += np.ascontiguousarray(cell, np.uint64).tobytes()

    def id_list(self, ids):
        self.i32(len(ids))
        for v in ids:
            if v is None or v < 0:
                self.boolean(False)
            else:
                self.boolean(True), self.u64(v)


def _mappoint(s, p):
    s.cls("MapPoint")
    s.u64(p["id"]), s.u64(p["next_id"]), s.i64(p["first_kf"]), s.i64(p["first_frame"]), s.i32(p["n_obs"])
    s.f32(p.get("proj_x", 0)), s.f32(p.get("proj_y", 0)), s.f32(p.get("proj_xr", 0)), s.boolean(p.get("track_in_view", False))
    s.i32(p.get("track_scale_level", 0)), s.f32(p.get("track_view_cos", 0))
    for f in ("track_ref_frame", "last_frame_seen", "ba_local_kf", "fuse_candidate_kf", "loop_point_kf", "corrected_by_kf",
              "corrected_ref"):
        s.u64(p.get(f, 0))
    s.mat(p.get("pos_gba")), s.u64(p.get("ba_global_kf", 0)), s.mat(p["world_pos"])
    s.u32(len(p["obs"]))
    for kf, idx in p["obs"]:
        if kf is None or kf < 0:
            s.boolean(False)
        else:
            s.boolean(True), s.u64(kf), s.u64(idx)
    s.mat(p["normal"]), s.mat(p["desc"])
    if p["ref_kf"] is None or p["ref_kf"] < 0:
        s.boolean(False)
    else:
        s.boolean(True), s.u64(p["ref_kf"])
    s.i32(p["visible"]), s.i32(p["found"]), s.boolean(p["bad"]), s.f32(p["min_dist"]), s.f32(p["max_dist"])


def _keyframe(s, k):
    s.cls("KeyFrame")
    s.u64(k["next_id"]), s.u64(k["id"]), s.u64(k["frame_id"]), s.f64(k["timestamp"])
    s.i32(k["grid_cols"]), s.i32(k["grid_rows"]), s.f32(k["grid_inv_w"]), s.f32(k["grid_inv_h"])
    for f in ("track_ref_frame", "fuse_target_kf", "ba_local_kf", "ba_fixed_kf", "loop_query"):
        s.u64(k.get(f, 0))
    s.i32(k.get("loop_words", 0)), s.f32(k.get("loop_score", 0)), s.u64(k.get("reloc_query", 0)), s.i32(k.get("reloc_words", 0))
    s.f32(k.get("reloc_score", 0))
    s.mat(k.get("tcw_gba")), s.mat(k.get("tcw_bef_gba")), s.u64(k.get("ba_global_kf", 0))
    for f in ("fx", "fy", "cx", "cy", "invfx", "invfy", "bf", "b", "th_depth"):
        s.f32(k[f])
    s.i32(k["n"])
    s.keypoints(k["keys"]), s.keypoints(k["keys_un"]), s.vec_f32(k["uright"]), s.vec_f32(k["depth"])
    s.mat(k["desc"]), s.mat(k.get("tcp"))
    s.i32(k["n_levels"]), s.f32(k["scale_factor"]), s.f32(k["log_scale_factor"])
    s.vec_f32(k["scale_factors"]), s.vec_f32(k["level_sigma2"]), s.vec_f32(k["inv_level_sigma2"])
    s.i32(k["min_x"]), s.i32(k["min_y"]), s.i32(k["max_x"]), s.i32(k["max_y"])
    s.mat(k["K"]), s.mat(k["Tcw"]), s.mat(k["Twc"]), s.mat(k["Ow"]), s.mat(k["Cw"])
    s.id_list(k["mappoint_ids"])
    s.grid(k["grid"])
    s.i32(len(k["connected"]))
    for kf, w in k["connected"]:
        if kf is None or kf < 0:
            s.boolean(False)
        else:
            s.boolean(True), s.u64(kf), s.i32(w)
    s.id_list(k["ordered_ids"])
    s.vec_i32(k["ordered_weights"])
    s.boolean(k["first_connection"])
    if k["parent"] is None or k["parent"] < 0:
        s.boolean(False)
    else:
        s.boolean(True), s.u64(k["parent"])
    s.id_list(k["children"]), s.id_list(k["loop_edges"])
    s.boolean(k["not_erase"]), s.boolean(k["to_be_erased"]), s.boolean(k["bad"]), s.f32(k["half_baseline"])


def serialize_map(m, second_copy=True):
    """Bytes of `oa << mpMap`.  second_copy: Map::save appends the map points once more after the 0xdeadbeef marker
    (src/Map.cc:68-73); Map::load never reads them."""
    s = _Stream()
    if m is None:
        return struct.pack("<h", -1)                                  # NULL pointer tag
    # pointer preamble: class_id 0, then (new class) tracking = 1 and version = 0, then object_id 0
    s.b += struct.pack("<hBII", 0, 1, 0, 0)
    s.i32(len(m["mappoints"]))
    for p in m["mappoints"]:
        _mappoint(s, p)
    s.i32(len(m["keyframes"]))
    for k in m["keyframes"]:
        _keyframe(s, k)
    s.i32(len(m["origins"]))
    for k in m["origins"]:
        _keyframe(s, k)
    s.u64(m["max_kf_id"])
    s.u32(0xdeadbeef)
    if second_copy:
        s.i32(len(m["mappoints"]))
        for p in m["mappoints"]:
            _mappoint(s, p)
    return bytes(s.b)


# ---- a random, self-consistent logical map -----------------------------------------------------------------------------------

def _pose(rng):
    a = rng.normal(size=(3, 3))
    q, _ = np.linalg.qr(a)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = q.astype(np.float32)
    T[:3, 3] = rng.normal(size=3).astype(np.float32)
    return T


def derived_pose(Tcw, half_baseline):
    """KeyFrame::SetPose (src/KeyFrame.cc:792-806) in float32, products accumulated left to right."""
    Tcw = np.asarray(Tcw, np.float32)
    Twc = np.zeros((4, 4), np.float32)
    Ow = np.zeros((3, 1), np.float32)
    for r in range(3):
        acc = np.float32(0)
        for c in range(3):
            Twc[r, c] = Tcw[c, r]
            acc = np.float32(acc + np.float32(Tcw[c, r] * Tcw[c, 3]))
        Ow[r, 0] = -acc
        Twc[r, 3] = -acc
    Twc[3, 3] = 1
    center = np.array([half_baseline, 0, 0, 1], np.float32)
    Cw = np.zeros((4, 1), np.float32)
    for r in range(4):
        acc = np.float32(0)
        for c in range(4):
            acc = np.float32(acc + np.float32(Twc[r, c] * center[c]))
        Cw[r, 0] = acc
    return Twc, Ow, Cw


def assign_grid(keys_un, min_x, min_y, inv_w, inv_h, cols=64, rows=48):
    """Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:341-356, 500-510)."""
    g = [[[] for _ in range(rows)] for _ in range(cols)]
    for i, k in enumerate(keys_un):
        fx = np.float32(np.float32(np.float32(k["x"]) - np.float32(min_x)) * np.float32(inv_w))
        fy = np.float32(np.float32(np.float32(k["y"]) - np.float32(min_y)) * np.float32(inv_h))
        gx = int(np.floor(abs(float(fx)) + 0.5) * (1 if fx >= 0 else -1))     # C round(): half away from zero
        gy = int(np.floor(abs(float(fy)) + 0.5) * (1 if fy >= 0 else -1))
        if 0 <= gx < cols and 0 <= gy < rows:
            g[gx][gy].append(i)
    return g


def make_random_map(seed, n_kf=4, n_feat=120, n_mp=60, n_levels=8, w=640, h=480, with_origin=True, empty_kf=False, gba=True):
    rng = np.random.default_rng(seed)
    sf = np.float32(1.2) ** np.arange(n_levels, dtype=np.float32)
    kfs = []
    for i in range(n_kf):
        n = 0 if (empty_kf and i == 1) else n_feat + int(rng.integers(0, 17))
        keys = [dict(x=np.float32(rng.uniform(0, w)), y=np.float32(rng.uniform(0, h)), angle=np.float32(rng.uniform(0, 360)),
                     response=np.float32(rng.integers(7, 200)), octave=int(rng.integers(0, n_levels)), class_id=-1) for _ in range(n)]
        keys_un = [dict(k, x=np.float32(k["x"] + rng.uniform(-3, 3)), y=np.float32(k["y"] + rng.uniform(-3, 3))) for k in keys]
        Tcw = _pose(rng)
        hb = np.float32(0.04)
        Twc, Ow, Cw = derived_pose(Tcw, hb)
        inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
        K = np.array([[500, 0, w / 2], [0, 500, h / 2], [0, 0, 1]], np.float32)
        kfs.append(dict(
            next_id=(n_kf - 1) * 3 + 1, id=i * 3, frame_id=i * 7 + 1, timestamp=1403636580.0 + 0.05 * i, grid_cols=64, grid_rows=48,
            grid_inv_w=inv_w, grid_inv_h=inv_h, fx=500.0, fy=500.0, cx=w / 2, cy=h / 2, invfx=np.float32(1) / np.float32(500),
            invfy=np.float32(1) / np.float32(500), bf=40.0, b=0.08, th_depth=3.2, n=n, keys=keys, keys_un=keys_un,
            uright=rng.uniform(-1, w, n).astype(np.float32), depth=rng.uniform(-1, 10, n).astype(np.float32),
            desc=rng.integers(0, 256, (n, 32), dtype=np.uint8) if n else np.zeros((0, 32), np.uint8),
            n_levels=n_levels, scale_factor=np.float32(1.2), log_scale_factor=np.float32(np.log(np.float32(1.2))),
            scale_factors=sf, level_sigma2=sf * sf, inv_level_sigma2=np.float32(1) / (sf * sf), min_x=0, min_y=0, max_x=w, max_y=h,
            K=K, Tcw=Tcw, Twc=Twc, Ow=Ow, Cw=Cw, mappoint_ids=[-1] * n,
            grid=assign_grid(keys_un, 0, 0, inv_w, inv_h), connected=[], ordered_ids=[], ordered_weights=[], first_connection=(i == 0),
            parent=(None if i == 0 else (i - 1) * 3), children=([] if i == n_kf - 1 else [(i + 1) * 3]),
            loop_edges=([(n_kf - 1) * 3] if i == 0 and n_kf > 2 else []), not_erase=bool(i % 2), to_be_erased=False, bad=False,
            half_baseline=hb))
    for i, k in enumerate(kfs):
        others = [o for j, o in enumerate(kfs) if j != i]
        k["connected"] = [(o["id"], int(rng.integers(15, 200))) for o in others]
        if len(others) > 1:
            k["connected"][1] = (-1, 0)                                 # an entry stored without an id
        k["ordered_ids"] = [o["id"] for o in others][::-1]
        k["ordered_weights"] = [int(rng.integers(15, 200)) for _ in others]
    mps = []
    for j in range(n_mp):
        obs = []
        for k in kfs:
            free = [f for f in range(k["n"]) if k["mappoint_ids"][f] < 0]
            if free and rng.random() < 0.7:
                f = int(free[int(rng.integers(0, len(free)))])
                k["mappoint_ids"][f] = j * 2 + 5
                obs.append((k["id"], f))
        if j == 3:
            obs.insert(0, (-1, 0))                                      # "Empty observation" branch (src/MapPoint.cc:99-106)
        desc = rng.integers(0, 256, (1, 32), dtype=np.uint8)
        mps.append(dict(id=j * 2 + 5, next_id=(n_mp - 1) * 2 + 6, first_kf=0, first_frame=0, n_obs=len(obs),
                        world_pos=rng.normal(size=(3, 1)).astype(np.float32), obs=obs, normal=rng.normal(size=(3, 1)).astype(np.float32),
                        desc=desc, ref_kf=(obs[-1][0] if obs and obs[-1][0] >= 0 else None), visible=int(rng.integers(1, 50)),
                        found=int(rng.integers(1, 50)), bad=False, min_dist=np.float32(rng.uniform(0.1, 1)),
                        max_dist=np.float32(rng.uniform(5, 20)), pos_gba=(rng.normal(size=(3, 1)).astype(np.float32) if gba and j % 5 == 0 else None)))
    return dict(mappoints=mps, keyframes=kfs, origins=([kfs[0]] if with_origin and kfs else []), max_kf_id=max([k["id"] for k in kfs], default=0))
