"""ctypes binding of the CPU oracle (oracle/orb_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never by
the product package (orbslam_mapsave_b200), which must fail loudly without its CUDA library instead.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
XYR_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("r", "<i4")])


def build(force=False):
    src = os.path.join(_HERE, "orb_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.orc_fast.restype = i32
        L.orc_fast.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32]
        L.orc_resize.restype = None
        L.orc_resize.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32]
        L.orc_blur.restype = None
        L.orc_blur.argtypes = [vp, i32, i32, i32, vp, i32]
        L.orc_border101.restype = None
        L.orc_border101.argtypes = [vp, i32, i32, i32, vp, i32, i32]
        L.orc_fast_atan2.restype = f32
        L.orc_fast_atan2.argtypes = [f32, f32]
        L.orc_cv_round.restype = i32
        L.orc_cv_round.argtypes = [C.c_double]
        L.orc_pattern.restype = vp
        L.orc_extractor_create.restype = vp
        L.orc_extractor_create.argtypes = [i32, f32, i32, i32, i32]
        L.orc_extractor_destroy.restype = None
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extractor_tables.restype = None
        L.orc_extractor_tables.argtypes = [vp] * 7
        L.orc_extractor_extract.restype = i32
        L.orc_extractor_extract.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp, i32]
        L.orc_extractor_level_dims.restype = None
        L.orc_extractor_level_dims.argtypes = [vp, i32, vp, vp]
        L.orc_extractor_level_copy.restype = None
        L.orc_extractor_level_copy.argtypes = [vp, i32, i32, vp, i32]
        L.orc_extractor_blurred_copy.restype = i32
        L.orc_extractor_blurred_copy.argtypes = [vp, i32, vp, i32]
        L.orc_extractor_candidates.restype = i32
        L.orc_extractor_candidates.argtypes = [vp, i32, vp, i32]
        L.orc_extractor_level_keypoints.restype = i32
        L.orc_extractor_level_keypoints.argtypes = [vp, i32, vp, i32]
        L.orc_octree.restype = i32
        L.orc_octree.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32]
        L.orc_descriptor.restype = None
        L.orc_descriptor.argtypes = [vp, i32, f32, f32, f32, vp]
        L.orc_extract_batch_mt.restype = C.c_double
        L.orc_extract_batch_mt.argtypes = [vp, i32, i32, i32, i32, f32, i32, i32, i32, i32, vp]
        L.orc_descriptor_distance.restype = i32
        L.orc_descriptor_distance.argtypes = [vp, vp]
        L.orc_three_maxima.restype = None
        L.orc_three_maxima.argtypes = [vp, i32, vp]
        L.orc_hamming_top2.restype = None
        L.orc_hamming_top2.argtypes = [vp, i32, vp, i32, vp, vp, vp]
        L.orc_hamming_top2_mt.restype = C.c_double
        L.orc_hamming_top2_mt.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32]
        L.orc_search_bow_kf_f.restype = i32
        L.orc_search_bow_kf_f.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, f32, i32, vp]
        L.orc_search_bow_kf_kf.restype = i32
        L.orc_search_bow_kf_kf.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, f32, i32, vp]
        L.orc_search_triangulation.restype = i32
        L.orc_search_triangulation.argtypes = ([vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp] +
                                               [vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp] +
                                               [vp, f32, f32, vp, vp, i32, i32, vp, vp])
        L.orc_voc_transform.restype = None
        L.orc_voc_transform.argtypes = [i32, vp, vp, vp, vp, i32, vp, i32, i32, vp, vp, vp]
        L.orc_distinctive.restype = i32
        L.orc_distinctive.argtypes = [vp, i32]
        L.orc_voc_bow.restype = i32
        L.orc_voc_bow.argtypes = [i32, vp, vp, i32, i32, vp, vp, i32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


# ------------------------------------------------------------------ primitives
def fast(img, threshold, nms=True):
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    cap = img.shape[0] * img.shape[1]
    out = np.zeros(max(cap, 1), XYR_DTYPE)
    n = lib().orc_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], int(threshold), int(nms), _p(out), cap)
    return out[:n]


def resize(src, dw, dh):
    src = np.asarray(src)
    assert src.dtype == np.uint8 and src.strides[1] == 1
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def blur(src):
    src = np.asarray(src)
    assert src.dtype == np.uint8 and src.strides[1] == 1
    dst = np.zeros(src.shape, np.uint8)
    lib().orc_blur(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), src.shape[1])
    return dst


def border101(src, b=19):
    src = _u8(src)
    h, w = src.shape
    dst = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    lib().orc_border101(_p(src), w, h, w, _p(dst), w + 2 * b, b)
    return dst


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(float(y), float(x)))


def pattern():
    buf = (C.c_byte * 1024).from_address(lib().orc_pattern())
    return np.frombuffer(buf, dtype=np.int8).copy()


# ------------------------------------------------------------------ extractor
class Extractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = lib().orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        quota = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        lib().orc_extractor_tables(self.h, _p(sf), _p(isf), _p(s2), _p(is2), _p(quota), _p(umax))
        return dict(scale=sf, inv_scale=isf, sigma2=s2, inv_sigma2=is2, quota=quota, umax=umax)

    def extract(self, img, mask=None):
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        cap = self.nfeatures + 4 * self.nlevels + 64
        while True:
            kp = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            m = None
            if mask is not None:
                m = np.ascontiguousarray(mask, dtype=np.uint8)
            n = lib().orc_extractor_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                            _p(m), (m.shape[1] if m is not None else 0), _p(kp), _p(desc), cap)
            if n >= 0:
                return kp[:n].copy(), desc[:n].copy()
            cap *= 2

    def level(self, l, bordered=False):
        w, h = C.c_int(), C.c_int()
        lib().orc_extractor_level_dims(self.h, l, C.byref(w), C.byref(h))
        w, h = w.value, h.value
        if bordered:
            out = np.zeros((h + 38, w + 38), np.uint8)
        else:
            out = np.zeros((h, w), np.uint8)
        lib().orc_extractor_level_copy(self.h, l, int(bordered), _p(out), out.shape[1])
        return out

    def blurred(self, l):
        out = np.zeros_like(self.level(l))
        ok = lib().orc_extractor_blurred_copy(self.h, l, _p(out), out.shape[1])
        return out if ok else None

    def candidates(self, l):
        cap = 1 << 16
        while True:
            out = np.zeros(cap, XYR_DTYPE)
            n = lib().orc_extractor_candidates(self.h, l, _p(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = n

    def level_keypoints(self, l):
        cap = self.nfeatures + 64
        while True:
            out = np.zeros(cap, KP_DTYPE)
            n = lib().orc_extractor_level_keypoints(self.h, l, _p(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = n


def octree(x, y, resp, min_x, max_x, min_y, max_y, N):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    r = np.ascontiguousarray(resp, np.float32)
    cap = len(x) + 8
    ox, oy, orr = (np.zeros(cap, np.float32) for _ in range(3))
    n = lib().orc_octree(_p(x), _p(y), _p(r), len(x), min_x, max_x, min_y, max_y, N, _p(ox), _p(oy), _p(orr), cap)
    return ox[:n], oy[:n], orr[:n]


def descriptor(blurred, x, y, angle):
    blurred = np.asarray(blurred)
    assert blurred.dtype == np.uint8 and blurred.strides[1] == 1
    d = np.zeros(32, np.uint8)
    lib().orc_descriptor(_p(blurred), blurred.strides[0], float(x), float(y), float(angle), _p(d))
    return d


def extract_batch_mt(frames, nfeatures, scale_factor, nlevels, ini_th, min_th, nthreads):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    tot = C.c_long()
    secs = lib().orc_extract_batch_mt(_p(frames), n, w, h, nfeatures, scale_factor, nlevels, ini_th, min_th, nthreads,
                                      C.byref(tot))
    return secs, tot.value


# ------------------------------------------------------------------ matcher
def descriptor_distance(a, b):
    a, b = _u8(a), _u8(b)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def three_maxima(hist):
    hist = np.ascontiguousarray(hist, np.int32)
    ind = np.zeros(3, np.int32)
    lib().orc_three_maxima(_p(hist), len(hist), _p(ind))
    return tuple(int(v) for v in ind)


def hamming_top2(q, db, nthreads=1):
    q, db = _u8(q), _u8(db)
    nq, ndb = len(q), len(db)
    bi, b1, b2 = (np.zeros(nq, np.int32) for _ in range(3))
    if nthreads <= 1:
        lib().orc_hamming_top2(_p(q), nq, _p(db), ndb, _p(bi), _p(b1), _p(b2))
        return bi, b1, b2
    secs = lib().orc_hamming_top2_mt(_p(q), nq, _p(db), ndb, _p(bi), _p(b1), _p(b2), nthreads)
    return bi, b1, b2, secs


class FeatVec:
    """Flattened DBoW2::FeatureVector: node ids ascending, CSR offsets, feature indices."""

    def __init__(self, node_of_feature):
        node_of_feature = np.asarray(node_of_feature, np.int64)
        order = np.argsort(node_of_feature, kind="stable")
        ids, counts = np.unique(node_of_feature, return_counts=True)
        self.ids = ids.astype(np.int32)
        self.off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.feat = order.astype(np.int32)

    @property
    def n(self):
        return len(self.ids)


def search_bow_kf_f(desc1, valid1, angle1, fv1, desc2, angle2, fv2, nnratio, check_ori):
    desc1, desc2 = _u8(desc1), _u8(desc2)
    valid1 = _u8(valid1)
    angle1 = np.ascontiguousarray(angle1, np.float32)
    angle2 = np.ascontiguousarray(angle2, np.float32)
    m = np.zeros(len(desc2), np.int32)
    n = lib().orc_search_bow_kf_f(_p(desc1), len(desc1), _p(valid1), _p(angle1), fv1.n, _p(fv1.ids), _p(fv1.off), _p(fv1.feat),
                                  _p(desc2), len(desc2), _p(angle2), fv2.n, _p(fv2.ids), _p(fv2.off), _p(fv2.feat),
                                  float(nnratio), int(check_ori), _p(m))
    return n, m


def search_bow_kf_kf(desc1, valid1, angle1, fv1, desc2, valid2, angle2, fv2, nnratio, check_ori):
    desc1, desc2 = _u8(desc1), _u8(desc2)
    valid1, valid2 = _u8(valid1), _u8(valid2)
    angle1 = np.ascontiguousarray(angle1, np.float32)
    angle2 = np.ascontiguousarray(angle2, np.float32)
    m = np.zeros(len(desc1), np.int32)
    n = lib().orc_search_bow_kf_kf(_p(desc1), len(desc1), _p(valid1), _p(angle1), fv1.n, _p(fv1.ids), _p(fv1.off), _p(fv1.feat),
                                   _p(desc2), len(desc2), _p(valid2), _p(angle2), fv2.n, _p(fv2.ids), _p(fv2.off), _p(fv2.feat),
                                   float(nnratio), int(check_ori), _p(m))
    return n, m


def search_triangulation(desc1, hasmp1, uright1, kx1, ky1, ang1, fv1,
                         desc2, hasmp2, uright2, kx2, ky2, ang2, oct2, fv2,
                         F12, ex, ey, sf2, sigma2_2, only_stereo, check_ori):
    f = lambda a: np.ascontiguousarray(a, np.float32)
    desc1, desc2, hasmp1, hasmp2 = _u8(desc1), _u8(desc2), _u8(hasmp1), _u8(hasmp2)
    uright1, kx1, ky1, ang1 = f(uright1), f(kx1), f(ky1), f(ang1)
    uright2, kx2, ky2, ang2 = f(uright2), f(kx2), f(ky2), f(ang2)
    oct2 = np.ascontiguousarray(oct2, np.int32)
    F12, sf2, sigma2_2 = f(F12).reshape(9), f(sf2), f(sigma2_2)
    pairs = np.zeros((len(desc1), 2), np.int32)
    npairs = C.c_int()
    n = lib().orc_search_triangulation(
        _p(desc1), len(desc1), _p(hasmp1), _p(uright1), _p(kx1), _p(ky1), _p(ang1), fv1.n, _p(fv1.ids), _p(fv1.off), _p(fv1.feat),
        _p(desc2), len(desc2), _p(hasmp2), _p(uright2), _p(kx2), _p(ky2), _p(ang2), _p(oct2), fv2.n, _p(fv2.ids), _p(fv2.off), _p(fv2.feat),
        _p(F12), float(ex), float(ey), _p(sf2), _p(sigma2_2), int(only_stereo), int(check_ori), _p(pairs), C.byref(npairs))
    return n, pairs[:npairs.value].copy()


# ------------------------------------------------------------------ vocabulary (DBoW2 transform)
def voc_transform(parent, ndesc, nweight, is_leaf, L, feat, levelsup):
    parent = np.ascontiguousarray(parent, np.int32)
    ndesc, is_leaf, feat = _u8(ndesc), _u8(is_leaf), _u8(feat)
    nweight = np.ascontiguousarray(nweight, np.float64)
    n = len(feat)
    w, nid = np.zeros(n, np.int32), np.zeros(n, np.int32)
    wt = np.zeros(n, np.float64)
    lib().orc_voc_transform(len(parent), _p(parent), _p(ndesc), _p(nweight), _p(is_leaf), int(L), _p(feat), n, int(levelsup),
                            _p(w), _p(wt), _p(nid))
    return w, wt, nid


def voc_bow(word_id, weight, weighting=0, scoring=0):
    word_id = np.ascontiguousarray(word_id, np.int32)
    weight = np.ascontiguousarray(weight, np.float64)
    cap = len(word_id) + 1
    ow, ov = np.zeros(cap, np.int32), np.zeros(cap, np.float64)
    k = lib().orc_voc_bow(len(word_id), _p(word_id), _p(weight), int(weighting), int(scoring), _p(ow), _p(ov), cap)
    return ow[:k], ov[:k]


def distinctive(desc):
    desc = _u8(desc).reshape(-1, 32)
    return lib().orc_distinctive(_p(desc), len(desc))


# ------------------------------------------------------------------ projection / window searches
class GridViewC(C.Structure):
    """orc_grid_view of orb_oracle.cpp (same layout as the product's orbm_grid_view)"""
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p), ("octave", C.c_void_p),
                ("angle", C.c_void_p), ("uright", C.c_void_p), ("blocked", C.c_void_p), ("grid_cols", C.c_int), ("grid_rows", C.c_int),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float), ("inv_w", C.c_float),
                ("inv_h", C.c_float), ("cell_offsets", C.c_void_p), ("cell_features", C.c_void_p), ("scale_factors", C.c_void_p),
                ("n_levels", C.c_int)]


class Grid:
    """A Frame as the window searches see it; the grid is built by the oracle's AssignFeaturesToGrid restatement."""

    def __init__(self, desc, x, y, octave, scale_factors, bounds, angle=None, uright=None, blocked=None, grid_cols=64, grid_rows=48):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        self.desc = _u8(desc).reshape(-1, 32)
        self.n = len(self.desc)
        self.x, self.y, self.angle, self.uright = f32(x), f32(y), f32(angle), f32(uright)
        self.octave = np.ascontiguousarray(octave, np.int32)
        self.blocked = None if blocked is None else _u8(blocked)
        self.sf = f32(scale_factors)
        self.bounds = tuple(np.float32(v) for v in bounds)
        self.cols, self.rows = grid_cols, grid_rows
        self.inv_w = np.float32(grid_cols) / np.float32(self.bounds[2] - self.bounds[0])
        self.inv_h = np.float32(grid_rows) / np.float32(self.bounds[3] - self.bounds[1])
        self.off = np.zeros(grid_cols * grid_rows + 1, np.int32)
        self.feat = np.zeros(max(self.n, 1), np.int32)
        f = lib().orc_assign_features_to_grid
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        f(self.n, _p(self.x), _p(self.y), grid_cols, grid_rows, self.bounds[0], self.bounds[1], self.inv_w, self.inv_h, _p(self.off), _p(self.feat))

    def c(self):
        return GridViewC(self.n, _p(self.desc), _p(self.x), _p(self.y), _p(self.octave), _p(self.angle), _p(self.uright), _p(self.blocked),
                         self.cols, self.rows, self.bounds[0], self.bounds[1], self.bounds[2], self.bounds[3], self.inv_w, self.inv_h,
                         _p(self.off), _p(self.feat), _p(self.sf), len(self.sf))


def search_projection_map(grid, in_view, proj_x, proj_y, proj_xr, level, view_cos, desc, claims, th, nnratio):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    in_view, claims, desc = _u8(in_view), _u8(claims), _u8(desc).reshape(-1, 32)
    proj_x, proj_y, proj_xr, view_cos = f32(proj_x), f32(proj_y), f32(proj_xr), f32(view_cos)
    level = np.ascontiguousarray(level, np.int32)
    owner = np.zeros(max(grid.n, 1), np.int32)
    f = lib().orc_search_projection_map
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_int] + [C.c_void_p] * 8 + [C.c_float, C.c_float, C.c_void_p]
    g = grid.c()
    n = f(C.byref(g), len(in_view), _p(in_view), _p(proj_x), _p(proj_y), _p(proj_xr), _p(level), _p(view_cos), _p(desc), _p(claims),
          float(th), float(nnratio), _p(owner))
    return n, owner[:grid.n]


def search_projection_frame(grid, Tcw, Tlw, fx, fy, cx, cy, mbf, mb, has_point, world, octave, angle, desc, claims, th, mono, check_ori):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    has_point, claims, desc = _u8(has_point), _u8(claims), _u8(desc).reshape(-1, 32)
    world, angle = f32(world).reshape(-1, 3), f32(angle)
    octave = np.ascontiguousarray(octave, np.int32)
    Tc, Tl = f32(Tcw).reshape(-1)[:12].copy(), f32(Tlw).reshape(-1)[:12].copy()
    owner = np.zeros(max(grid.n, 1), np.int32)
    f = lib().orc_search_projection_frame
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_void_p, C.c_void_p] + [C.c_float] * 6 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_int, C.c_void_p]
    g = grid.c()
    n = f(C.byref(g), _p(Tc), _p(Tl), fx, fy, cx, cy, mbf, mb, len(has_point), _p(has_point), _p(world), _p(octave), _p(angle), _p(desc),
          _p(claims), float(th), int(mono), int(check_ori), _p(owner))
    return n, owner[:grid.n]


def search_initialization(grid2, desc1, octave1, angle1, prev_matched, window_size, nnratio, check_ori):
    desc1 = _u8(desc1).reshape(-1, 32)
    octave1 = np.ascontiguousarray(octave1, np.int32)
    angle1 = np.ascontiguousarray(angle1, np.float32)
    assert prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous
    m = np.zeros(max(len(desc1), 1), np.int32)
    f = lib().orc_search_initialization
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    g = grid2.c()
    n = f(C.byref(g), len(desc1), _p(desc1), _p(octave1), _p(angle1), _p(prev_matched), int(window_size), float(nnratio), int(check_ori), _p(m))
    return n, m[:len(desc1)]


def search_windows(grid, active, u, v, r, min_level, max_level, desc, angle, th_dist, check_ori):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    active, desc = _u8(active), _u8(desc).reshape(-1, 32)
    u, v, r, min_level, max_level = f32(u), f32(v), f32(r), i32(min_level), i32(max_level)
    angle = f32(angle) if angle is not None else None
    owner = np.zeros(max(grid.n, 1), np.int32)
    f = lib().orc_search_windows
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_void_p]
    g = grid.c()
    n = f(C.byref(g), len(active), _p(active), _p(u), _p(v), _p(r), _p(min_level), _p(max_level), _p(desc), _p(angle), int(th_dist),
          int(check_ori), _p(owner))
    return n, owner[:grid.n]


def search_projection_kf(grid, Tcw, fx, fy, cx, cy, log_sf, valid, world, mf_max, mf_min, angle, desc, th, orb_dist, check_ori):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    valid, desc = _u8(valid), _u8(desc).reshape(-1, 32)
    world, mf_max, mf_min, angle = f32(world).reshape(-1, 3), f32(mf_max), f32(mf_min), f32(angle)
    T = f32(Tcw).reshape(-1)[:12].copy()
    owner = np.zeros(max(grid.n, 1), np.int32)
    f = lib().orc_search_projection_kf
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_void_p] + [C.c_float] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_int, C.c_void_p]
    g = grid.c()
    n = f(C.byref(g), _p(T), fx, fy, cx, cy, log_sf, len(valid), _p(valid), _p(world), _p(mf_max), _p(mf_min), _p(angle), _p(desc),
          float(th), int(orb_dist), int(check_ori), _p(owner))
    return n, owner[:grid.n]


def search_projection_sim3(grid, Scw, fx, fy, cx, cy, log_sf, valid, world, mf_max, mf_min, normal, desc, th):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    valid, desc = _u8(valid), _u8(desc).reshape(-1, 32)
    world, normal, mf_max, mf_min = f32(world).reshape(-1, 3), f32(normal).reshape(-1, 3), f32(mf_max), f32(mf_min)
    S = f32(Scw).reshape(-1)[:12].copy()
    owner = np.zeros(max(grid.n, 1), np.int32)
    f = lib().orc_search_projection_sim3
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_void_p] + [C.c_float] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    g = grid.c()
    n = f(C.byref(g), _p(S), fx, fy, cx, cy, log_sf, len(valid), _p(valid), _p(world), _p(mf_max), _p(mf_min), _p(normal), _p(desc),
          int(th), _p(owner))
    return n, owner[:grid.n]


def stereo_matches(levels_left, levels_right, scale_factors, inv_scale_factors, kp_left, desc_left, kp_right, desc_right, mbf, mb):
    """Frame::ComputeStereoMatches restatement; levels_*: lists of border-less pyramid levels (uint8 arrays)."""
    nl = len(levels_left)
    L = [np.ascontiguousarray(a, np.uint8) for a in levels_left]
    R = [np.ascontiguousarray(a, np.uint8) for a in levels_right]
    assert all(a.shape == b.shape for a, b in zip(L, R))
    pl = (C.c_void_p * nl)(*[a.ctypes.data for a in L])
    pr = (C.c_void_p * nl)(*[a.ctypes.data for a in R])
    w = np.array([a.shape[1] for a in L], np.int32)
    h = np.array([a.shape[0] for a in L], np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    isf = np.ascontiguousarray(inv_scale_factors, np.float32)
    kl, kr = np.ascontiguousarray(kp_left, KP_DTYPE), np.ascontiguousarray(kp_right, KP_DTYPE)
    dl, dr = _u8(desc_left).reshape(-1, 32), _u8(desc_right).reshape(-1, 32)
    u = np.zeros(max(len(kl), 1), np.float32)
    d = np.zeros(max(len(kl), 1), np.float32)
    f = lib().orc_stereo_matches
    f.restype = None
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    f(nl, pl, pr, _p(w), _p(h), _p(sf), _p(isf), _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), float(mbf), float(mb), _p(u), _p(d))
    return u[:len(kl)], d[:len(kl)]


def search_windows_best(grid, active, u, v, r, min_level, max_level, desc, ur, inv_sigma2, th_dist):
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    active, desc = _u8(active), _u8(desc).reshape(-1, 32)
    u, v, r, ur, inv_sigma2, min_level, max_level = f32(u), f32(v), f32(r), f32(ur), f32(inv_sigma2), i32(min_level), i32(max_level)
    best = np.zeros(max(len(active), 1), np.int32)
    f = lib().orc_search_windows_best
    f.restype = None
    f.argtypes = [C.POINTER(GridViewC), C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_void_p]
    g = grid.c()
    f(C.byref(g), len(active), _p(active), _p(u), _p(v), _p(r), _p(min_level), _p(max_level), _p(desc), _p(ur), _p(inv_sigma2), int(th_dist), _p(best))
    return best[:len(active)]


def fuse_search(grid, variant, T, Ow, fx, fy, cx, cy, bf, log_sf, skip, world, mf_max, mf_min, normal, desc, th, inv_sigma2):
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    skip, desc = _u8(skip), _u8(desc).reshape(-1, 32)
    world, normal, mf_max, mf_min, inv_sigma2 = f32(world).reshape(-1, 3), f32(normal).reshape(-1, 3), f32(mf_max), f32(mf_min), f32(inv_sigma2)
    Tf, Owf = f32(T).reshape(-1)[:12].copy(), f32(Ow if Ow is not None else np.zeros(3))
    best = np.zeros(max(len(skip), 1), np.int32)
    f = lib().orc_fuse_search
    f.restype = None
    f.argtypes = [C.POINTER(GridViewC), C.c_int, C.c_void_p, C.c_void_p] + [C.c_float] * 6 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_void_p, C.c_void_p]
    g = grid.c()
    f(C.byref(g), int(variant), _p(Tf), _p(Owf), fx, fy, cx, cy, bf, log_sf, len(skip), _p(skip), _p(world), _p(mf_max), _p(mf_min),
      _p(normal), _p(desc), float(th), _p(inv_sigma2), _p(best))
    return best[:len(skip)]


def sim3_direction(grid_to, Rfw, tfw, sR, t, fx, fy, cx, cy, log_sf, valid, world, mf_max, mf_min, desc, th):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    valid, desc = _u8(valid), _u8(desc).reshape(-1, 32)
    world, mf_max, mf_min = f32(world).reshape(-1, 3), f32(mf_max), f32(mf_min)
    Rfw, tfw, sR, t = f32(Rfw).reshape(9), f32(tfw).reshape(3), f32(sR).reshape(9), f32(t).reshape(3)
    out = np.zeros(max(len(valid), 1), np.int32)
    f = lib().orc_sim3_direction
    f.restype = None
    f.argtypes = [C.POINTER(GridViewC)] + [C.c_void_p] * 4 + [C.c_float] * 5 + [C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_void_p]
    g = grid_to.c()
    f(C.byref(g), _p(Rfw), _p(tfw), _p(sR), _p(t), fx, fy, cx, cy, log_sf, len(valid), _p(valid), _p(world), _p(mf_max), _p(mf_min), _p(desc),
      float(th), _p(out))
    return out[:len(valid)]


def features_in_area(grid, x, y, r, min_level=-1, max_level=-1):
    """Frame::GetFeaturesInArea restatement on its own (src/Frame.cc:445-498)."""
    out = np.zeros(max(grid.n, 1), np.int32)
    f = lib().orc_features_in_area
    f.restype = C.c_int
    f.argtypes = [C.POINTER(GridViewC), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    g = grid.c()
    n = f(C.byref(g), float(x), float(y), float(r), int(min_level), int(max_level), _p(out), len(out))
    return out[:n].copy()
