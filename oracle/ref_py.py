"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref — the reference's own src/ORBextractor.cc compiled unmodified against
oracle/ref_shim/cvshim.hpp (oracle/Makefile).  Only tests/ and bench.py's CPU arm may use this.  Nothing here reads
/root/reference at run time: the binaries are built where the checkout exists and travel with the repo snapshot."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libref_orbextractor.so")
CLI = os.path.join(_HERE, "_ref", "ref_extract_cli")
KP_FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")


def available():
    return os.path.exists(LIB) and os.path.exists(CLI)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.refx_create.restype = C.c_void_p
        L.refx_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.refx_destroy.argtypes = [C.c_void_p]
        L.refx_extract.restype = C.c_int
        L.refx_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


class RefExtractor:
    """ORB_SLAM2::ORBextractor of the reference, stock heap: its octree tie-break compares node pointers (src/ORBextractor.cc:683),
    so a few keypoints per frame depend on the allocator and change from call to call."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self._h = lib().refx_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self._cap = nfeatures + 8 * nlevels + 64

    def __del__(self):
        try:
            lib().refx_destroy(self._h)
        except Exception:
            pass

    def extract(self, img, mask=None):
        img = np.ascontiguousarray(img, np.uint8)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        kp = np.zeros((self._cap, 7), np.float32)
        desc = np.zeros((self._cap, 32), np.uint8)
        n = lib().refx_extract(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                               mask.ctypes.data if mask is not None else None, mask.strides[0] if mask is not None else 0,
                               kp.ctypes.data, desc.ctypes.data, self._cap)
        assert n <= self._cap
        return kp[:n], desc[:n]


def extract_monotonic_heap(img, nfeatures, scale_factor, nlevels, ini_th, min_th, mask=None):
    """The same translation unit in a fresh process whose operator new never reuses an address (ref_shim/ref_cli.cpp): pointer
    order == creation order, the canonical tie-break.  Deterministic; returns (kp[n, 7] float32, desc[n, 32])."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    with tempfile.TemporaryDirectory() as d:
        img.tofile(os.path.join(d, "i.raw"))
        mpath = "-"
        if mask is not None:
            mpath = os.path.join(d, "m.raw")
            np.ascontiguousarray(mask, np.uint8).tofile(mpath)
        out = os.path.join(d, "o.bin")
        subprocess.check_call([CLI, str(w), str(h), str(nfeatures), repr(float(scale_factor)), str(nlevels), str(ini_th), str(min_th),
                               os.path.join(d, "i.raw"), mpath, out])
        raw = open(out, "rb").read()
    n = int(np.frombuffer(raw[:4], np.int32)[0])
    kp = np.frombuffer(raw[4:4 + 28 * n], np.float32).reshape(n, 7).copy()
    desc = np.frombuffer(raw[4 + 28 * n:4 + 60 * n], np.uint8).reshape(n, 32).copy()
    return kp, desc


# ---- matcher: the reference's own src/ORBmatcher.cc (oracle/_ref/libref_orbmatcher.so) --------------------------------------
MLIB = os.path.join(_HERE, "_ref", "libref_orbmatcher.so")


def matcher_available():
    return os.path.exists(MLIB)


class _Side(C.Structure):
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("flag", C.c_void_p), ("angle", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p),
                ("octave", C.c_void_p), ("uright", C.c_void_p), ("n_nodes", C.c_int), ("node_ids", C.c_void_p), ("off", C.c_void_p),
                ("feat", C.c_void_p)]


_mlib = None


def mlib():
    global _mlib
    if _mlib is None:
        L = C.CDLL(MLIB)
        L.refm_descriptor_distance.restype = C.c_int
        L.refm_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.refm_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.refm_search_bow_kf_frame.restype = C.c_int
        L.refm_search_bow_kf_frame.argtypes = [C.POINTER(_Side), C.POINTER(_Side), C.c_float, C.c_int, C.c_void_p]
        L.refm_search_bow_kf_kf.restype = C.c_int
        L.refm_search_bow_kf_kf.argtypes = [C.POINTER(_Side), C.POINTER(_Side), C.c_float, C.c_int, C.c_void_p]
        L.refm_search_triangulation.restype = C.c_int
        L.refm_search_triangulation.argtypes = [C.POINTER(_Side), C.POINTER(_Side), C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int]
        _mlib = L
    return _mlib


def _side(desc, fv, angle, flag=None, x=None, y=None, octave=None, uright=None):
    """fv: an object with .ids / .off / .feat int32 arrays (oracle.orb_oracle_py.FeatVec).  Returns (struct, keep-alive list)."""
    keep = [np.ascontiguousarray(desc, np.uint8).reshape(-1, 32), np.ascontiguousarray(angle, np.float32)]
    p = lambda a: a.ctypes.data if a is not None else None
    opt = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
    flag, x, y, octave, uright = opt(flag, np.uint8), opt(x, np.float32), opt(y, np.float32), opt(octave, np.int32), opt(uright, np.float32)
    ids, off, feat = (np.ascontiguousarray(a, np.int32) for a in (fv.ids, fv.off, fv.feat))
    keep += [flag, x, y, octave, uright, ids, off, feat]
    s = _Side(len(keep[0]), p(keep[0]), p(flag), p(keep[1]), p(x), p(y), p(octave), p(uright), len(ids), p(ids), p(off), p(feat))
    return s, keep


def ref_descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return mlib().refm_descriptor_distance(a.ctypes.data, b.ctypes.data)


def ref_three_maxima(counts):
    c = np.ascontiguousarray(counts, np.int32)
    out = np.zeros(3, np.int32)
    mlib().refm_three_maxima(c.ctypes.data, len(c), out.ctypes.data)
    return tuple(int(v) for v in out)


def ref_search_bow_kf_f(desc1, valid1, angle1, fv1, desc2, angle2, fv2, nnratio, check_ori):
    s1, k1 = _side(desc1, fv1, angle1, flag=valid1)
    s2, k2 = _side(desc2, fv2, angle2)
    m = np.zeros(max(len(k2[0]), 1), np.int32)
    n = mlib().refm_search_bow_kf_frame(C.byref(s1), C.byref(s2), float(nnratio), int(check_ori), m.ctypes.data)
    return n, m[:len(k2[0])]


def ref_search_bow_kf_kf(desc1, valid1, angle1, fv1, desc2, valid2, angle2, fv2, nnratio, check_ori):
    s1, k1 = _side(desc1, fv1, angle1, flag=valid1)
    s2, k2 = _side(desc2, fv2, angle2, flag=valid2)
    m = np.zeros(max(len(k1[0]), 1), np.int32)
    n = mlib().refm_search_bow_kf_kf(C.byref(s1), C.byref(s2), float(nnratio), int(check_ori), m.ctypes.data)
    return n, m[:len(k1[0])]


def ref_search_triangulation(desc1, hasmp1, uright1, kx1, ky1, ang1, fv1, desc2, hasmp2, uright2, kx2, ky2, ang2, oct2, fv2,
                             F12, ex, ey, sf2, sigma2_2, only_stereo, check_ori, nnratio=0.6):
    s1, k1 = _side(desc1, fv1, ang1, flag=hasmp1, x=kx1, y=ky1, octave=np.zeros(len(desc1), np.int32), uright=uright1)
    s2, k2 = _side(desc2, fv2, ang2, flag=hasmp2, x=kx2, y=ky2, octave=oct2, uright=uright2)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sf2, sg = np.ascontiguousarray(sf2, np.float32), np.ascontiguousarray(sigma2_2, np.float32)
    cap = max(len(k1[0]), 1)
    pairs = np.zeros((cap, 2), np.int32)
    n = mlib().refm_search_triangulation(C.byref(s1), C.byref(s2), F.ctypes.data, float(ex), float(ey), sf2.ctypes.data, sg.ctypes.data,
                                         len(sf2), int(only_stereo), float(nnratio), int(check_ori), pairs.ctypes.data, cap)
    return n, pairs[:n].copy()


# ---- window searches of the tracker through the reference's own code (grid: an oracle.orb_oracle_py.Grid; same struct layout) ----
def _wlib():
    L = mlib()
    if not getattr(L, "_win_ready", False):
        vp, f32, i32 = C.c_void_p, C.c_float, C.c_int
        L.refm_search_projection_map.restype = i32
        L.refm_search_projection_map.argtypes = [vp, i32] + [vp] * 8 + [f32, f32, vp]
        L.refm_search_projection_frame.restype = i32
        L.refm_search_projection_frame.argtypes = [vp, vp, vp] + [f32] * 6 + [i32] + [vp] * 6 + [f32, i32, i32, f32, vp]
        L.refm_search_initialization.restype = i32
        L.refm_search_initialization.argtypes = [vp, i32, vp, vp, vp, vp, i32, f32, i32, vp]
        L._win_ready = True
    return L


def _a(x, t):
    return np.ascontiguousarray(x, t)


def ref_search_projection_map(grid, in_view, proj_x, proj_y, proj_xr, level, view_cos, desc, claims, th, nnratio):
    g = grid.c()
    iv, cl, d = _a(in_view, np.uint8), _a(claims, np.uint8), _a(desc, np.uint8).reshape(-1, 32)
    px, py, pr, vc, lv = _a(proj_x, np.float32), _a(proj_y, np.float32), _a(proj_xr, np.float32), _a(view_cos, np.float32), _a(level, np.int32)
    owner = np.zeros(max(grid.n, 1), np.int32)
    n = _wlib().refm_search_projection_map(C.addressof(g), len(iv), iv.ctypes.data, px.ctypes.data, py.ctypes.data, pr.ctypes.data,
                                           lv.ctypes.data, vc.ctypes.data, d.ctypes.data, cl.ctypes.data, float(th), float(nnratio),
                                           owner.ctypes.data)
    return n, owner[:grid.n]


def ref_search_projection_frame(grid, Tcw, Tlw, fx, fy, cx, cy, mbf, mb, has_point, world, octave, angle, desc, claims, th, mono, check_ori,
                                nnratio=0.9):
    g = grid.c()
    hp, cl, d = _a(has_point, np.uint8), _a(claims, np.uint8), _a(desc, np.uint8).reshape(-1, 32)
    w, an, oc = _a(world, np.float32).reshape(-1, 3), _a(angle, np.float32), _a(octave, np.int32)
    Tc, Tl = _a(Tcw, np.float32).reshape(-1)[:12].copy(), _a(Tlw, np.float32).reshape(-1)[:12].copy()
    owner = np.zeros(max(grid.n, 1), np.int32)
    n = _wlib().refm_search_projection_frame(C.addressof(g), Tc.ctypes.data, Tl.ctypes.data, fx, fy, cx, cy, mbf, mb, len(hp), hp.ctypes.data,
                                             w.ctypes.data, oc.ctypes.data, an.ctypes.data, d.ctypes.data, cl.ctypes.data, float(th),
                                             int(mono), int(check_ori), float(nnratio), owner.ctypes.data)
    return n, owner[:grid.n]


def ref_search_initialization(grid2, desc1, octave1, angle1, prev_matched, window_size, nnratio, check_ori):
    g = grid2.c()
    d, oc, an = _a(desc1, np.uint8).reshape(-1, 32), _a(octave1, np.int32), _a(angle1, np.float32)
    assert prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous
    m = np.zeros(max(len(d), 1), np.int32)
    n = _wlib().refm_search_initialization(C.addressof(g), len(d), d.ctypes.data, oc.ctypes.data, an.ctypes.data, prev_matched.ctypes.data,
                                           int(window_size), float(nnratio), int(check_ori), m.ctypes.data)
    return n, m[:len(d)]


def ref_search_projection_kf(grid, Tcw, fx, fy, cx, cy, log_sf, state, world, mf_max, mf_min, angle, desc, th, orb_dist, check_ori):
    """state: 0 none / 1 good / 2 bad / 3 already found (uint8), per keyframe feature."""
    L = mlib()
    L.refm_search_projection_kf.restype = C.c_int
    L.refm_search_projection_kf.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_int, C.c_void_p]
    g = grid.c()
    st, d = _a(state, np.uint8), _a(desc, np.uint8).reshape(-1, 32)
    w, mx, mn, an = _a(world, np.float32).reshape(-1, 3), _a(mf_max, np.float32), _a(mf_min, np.float32), _a(angle, np.float32)
    T = _a(Tcw, np.float32).reshape(-1)[:12].copy()
    owner = np.zeros(max(grid.n, 1), np.int32)
    n = L.refm_search_projection_kf(C.addressof(g), T.ctypes.data, fx, fy, cx, cy, log_sf, len(st), st.ctypes.data, w.ctypes.data, mx.ctypes.data,
                                    mn.ctypes.data, an.ctypes.data, d.ctypes.data, float(th), int(orb_dist), int(check_ori), owner.ctypes.data)
    return n, owner[:grid.n]


def ref_search_projection_sim3(grid, Scw, fx, fy, cx, cy, log_sf, state, world, mf_max, mf_min, normal, desc, th):
    L = mlib()
    L.refm_search_projection_sim3.restype = C.c_int
    L.refm_search_projection_sim3.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    g = grid.c()
    st, d = _a(state, np.uint8), _a(desc, np.uint8).reshape(-1, 32)
    w, nr, mx, mn = _a(world, np.float32).reshape(-1, 3), _a(normal, np.float32).reshape(-1, 3), _a(mf_max, np.float32), _a(mf_min, np.float32)
    S = _a(Scw, np.float32).reshape(-1)[:12].copy()
    owner = np.zeros(max(grid.n, 1), np.int32)
    n = L.refm_search_projection_sim3(C.addressof(g), S.ctypes.data, fx, fy, cx, cy, log_sf, len(st), st.ctypes.data, w.ctypes.data, mx.ctypes.data,
                                      mn.ctypes.data, nr.ctypes.data, d.ctypes.data, int(th), owner.ctypes.data)
    return n, owner[:grid.n]


# ---- DBoW2: the reference's own Thirdparty/DBoW2 vocabulary (oracle/_ref/libref_dbow2.so) --------------------------------------
VLIB = os.path.join(_HERE, "_ref", "libref_dbow2.so")


def dbow2_available():
    return os.path.exists(VLIB)


_vlib = None


def vlib():
    global _vlib
    if _vlib is None:
        L = C.CDLL(VLIB)
        vp, i32 = C.c_void_p, C.c_int
        L.refv_load.restype = vp
        L.refv_load.argtypes = [C.c_char_p, i32]
        L.refv_destroy.argtypes = [vp]
        L.refv_save_binary.argtypes = [vp, C.c_char_p]
        L.refv_save_text.argtypes = [vp, C.c_char_p]
        L.refv_info.argtypes = [vp] * 5
        L.refv_transform_raw.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.refv_transform.restype = i32
        L.refv_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        _vlib = L
    return _vlib


class RefVocabulary:
    """ORB_SLAM2::ORBVocabulary (TemplatedVocabulary<FORB::TDescriptor, FORB>) of the reference, loaded from one of its two on-disk
    formats by its own loaders."""

    def __init__(self, path, binary=False):
        self._h = vlib().refv_load(str(path).encode(), int(binary))
        if not self._h:
            raise ValueError(f"the reference's loader rejected {path}")

    def __del__(self):
        try:
            vlib().refv_destroy(self._h)
        except Exception:
            pass

    def info(self):
        v = [C.c_int() for _ in range(4)]
        vlib().refv_info(self._h, *[C.byref(x) for x in v])
        return dict(zip(("k", "L", "n_nodes", "n_words"), (x.value for x in v)))

    def save_binary(self, path):
        vlib().refv_save_binary(self._h, str(path).encode())

    def save_text(self, path):
        vlib().refv_save_text(self._h, str(path).encode())

    def transform_raw(self, desc, levelsup):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        w, nid, wt = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        vlib().refv_transform_raw(self._h, d.ctypes.data, n, int(levelsup), w.ctypes.data, wt.ctypes.data, nid.ctypes.data)
        return w, wt, nid

    def transform(self, desc, levelsup):
        """(BowVector word ids, values), (FeatureVector node ids, CSR offsets, feature indices)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bid, bval = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
        fnode, foff, ffeat = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32)
        nn = C.c_int()
        nb = vlib().refv_transform(self._h, d.ctypes.data, n, int(levelsup), bid.ctypes.data, bval.ctypes.data, fnode.ctypes.data,
                                   foff.ctypes.data, ffeat.ctypes.data, C.byref(nn))
        k = nn.value
        return (bid[:nb], bval[:nb]), (fnode[:k], foff[:k + 1], ffeat[:foff[k]])


# ---- Fuse x2 / SearchBySim3 through the reference's own src/ORBmatcher.cc (object graph built from flat arrays) ------------------
class _Points(C.Structure):
    _fields_ = [("n", C.c_int), ("state", C.c_void_p), ("nobs", C.c_void_p), ("desc", C.c_void_p), ("world", C.c_void_p),
                ("normal", C.c_void_p), ("mf_max", C.c_void_p), ("mf_min", C.c_void_p)]


def _points(state, nobs, P):
    keep = [_a(state, np.uint8), _a(nobs, np.int32), _a(P["desc"], np.uint8).reshape(-1, 32), _a(P["world"], np.float32).reshape(-1, 3),
            _a(P["normal"], np.float32).reshape(-1, 3), _a(P["mf_max"], np.float32), _a(P["mf_min"], np.float32)]
    return _Points(len(keep[0]), *[k.ctypes.data for k in keep]), keep


def ref_fuse(variant, grid, K6, inv_sigma2, T, Ow, Scw, kf_state, kf_nobs, kfP, cand_state, cand_nobs, candP, in_at, th):
    """ORBmatcher::Fuse (variant 0: src/ORBmatcher.cc:828-972; variant 1, with a similarity: :974-1103).  Returns
    (nFused, keyframe feature -> point, (isBad, Observations) per point [keyframe's points, then the candidates], vpReplacePoint)."""
    L = mlib()
    L.refm_fuse.restype = C.c_int
    L.refm_fuse.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.POINTER(_Points), C.POINTER(_Points), C.c_void_p, C.c_float] + [C.c_void_p] * 3
    g = grid.c()
    kp, k1 = _points(kf_state, kf_nobs, kfP)
    cp, k2 = _points(cand_state, cand_nobs, candP)
    K6, is2 = _a(K6, np.float32), _a(inv_sigma2, np.float32)
    Tm = _a(T, np.float32).reshape(-1)[:12].copy()
    Owm = _a(Ow if Ow is not None else np.zeros(3), np.float32)
    Sm = _a(Scw if Scw is not None else np.eye(4), np.float32).reshape(-1).copy()
    ia = _a(in_at, np.int32)
    n, npts = grid.n, len(ia)
    kf_ptr, pts, rep = np.zeros(max(n, 1), np.int32), np.zeros((n + npts, 2), np.int32), np.full(max(npts, 1), -1, np.int32)
    nf = L.refm_fuse(int(variant), C.addressof(g), K6.ctypes.data, is2.ctypes.data, Tm.ctypes.data, Owm.ctypes.data, Sm.ctypes.data,
                     C.byref(kp), C.byref(cp), ia.ctypes.data, float(th), kf_ptr.ctypes.data, pts.ctypes.data, rep.ctypes.data)
    return nf, kf_ptr[:n], pts, rep[:npts]


def ref_search_by_sim3(grid1, grid2, K6, T1, T2, st1, P1, st2, P2, s12, R12, t12, th, pre12):
    """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1105-1329): (nFound, vpMatches12 as KF2 feature indices)."""
    L = mlib()
    L.refm_search_by_sim3.restype = C.c_int
    L.refm_search_by_sim3.argtypes = [C.c_void_p] * 5 + [C.POINTER(_Points), C.POINTER(_Points), C.c_float, C.c_void_p, C.c_void_p, C.c_float,
                                                          C.c_void_p, C.c_void_p]
    g1, g2 = grid1.c(), grid2.c()
    p1, k1 = _points(st1, np.ones(len(st1)), P1)
    p2, k2 = _points(st2, np.ones(len(st2)), P2)
    K6 = _a(K6, np.float32)
    T1m, T2m = _a(T1, np.float32).reshape(-1)[:12].copy(), _a(T2, np.float32).reshape(-1)[:12].copy()
    R, t, pre = _a(R12, np.float32).reshape(9).copy(), _a(t12, np.float32).reshape(3).copy(), _a(pre12, np.int32)
    out = np.zeros(max(grid1.n, 1), np.int32)
    n = L.refm_search_by_sim3(C.addressof(g1), C.addressof(g2), K6.ctypes.data, T1m.ctypes.data, T2m.ctypes.data, C.byref(p1), C.byref(p2),
                              float(s12), R.ctypes.data, t.ctypes.data, float(th), pre.ctypes.data, out.ctypes.data)
    return n, out[:grid1.n]


# ---- MapPoint: the reference's own src/MapPoint.cc (oracle/_ref/libref_mappoint.so) ------------------------------------------
PLIB = os.path.join(_HERE, "_ref", "libref_mappoint.so")


def mappoint_available():
    return os.path.exists(PLIB) and os.path.exists(MLIB)


_plib = None


def plib():
    global _plib
    if _plib is None:
        L = C.CDLL(PLIB)
        L.refp_distinctive.restype = C.c_int
        L.refp_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.refp_save_fields.restype = C.c_int
        L.refp_save_fields.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        _plib = L
    return _plib


def ref_distinctive(desc, kf_bad=None, point_bad=False):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:483-548) on a point observed once by each of len(desc) keyframes.
    Returns (index of the chosen observation or -1 when mDescriptor is left alone, mDescriptor)."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    bad = None if kf_bad is None else np.ascontiguousarray(kf_bad, np.uint8)
    out = np.zeros(32, np.uint8)
    i = plib().refp_distinctive(d.ctypes.data, bad.ctypes.data if bad is not None else None, len(d), int(point_bad), out.ctypes.data)
    return i, out


def ref_mappoint_save_fields(n_obs, has_ref=True, track=False):
    """(raw bytes, field tags) MapPoint::save (src/MapPoint.cc:58-140) hands to the archive, in call order."""
    buf = np.zeros(1 << 16, np.uint8)
    f = C.create_string_buffer(1 << 14)
    n = plib().refp_save_fields(int(n_obs), int(has_ref), int(track), buf.ctypes.data, len(buf), f, len(f))
    assert n >= 0
    return buf[:n].tobytes(), f.value.decode().strip(";").split(";")


# ---- Frame: the reference's own src/Frame.cc (oracle/_ref/libref_frame.so) --------------------------------------------------
FLIB = os.path.join(_HERE, "_ref", "libref_frame.so")


def frame_available():
    return os.path.exists(FLIB) and os.path.exists(MLIB)


def ref_stereo_matches(levels_left, levels_right, scale_factors, inv_scale_factors, kp_left, desc_left, kp_right, desc_right, mbf, mb):
    """Frame::ComputeStereoMatches (src/Frame.cc:584-756) of the reference on border-less pyramid levels; same arguments as the oracle's
    stereo_matches.  Returns (mvuRight, mvDepth)."""
    from .orb_oracle_py import KP_DTYPE
    L = C.CDLL(FLIB)
    nl = len(levels_left)
    Ls = [np.ascontiguousarray(a, np.uint8) for a in levels_left]
    Rs = [np.ascontiguousarray(a, np.uint8) for a in levels_right]
    pl = (C.c_void_p * nl)(*[a.ctypes.data for a in Ls])
    pr = (C.c_void_p * nl)(*[a.ctypes.data for a in Rs])
    w = np.array([a.shape[1] for a in Ls], np.int32)
    h = np.array([a.shape[0] for a in Ls], np.int32)
    sf, isf = _a(scale_factors, np.float32), _a(inv_scale_factors, np.float32)
    kl, kr = np.ascontiguousarray(kp_left, KP_DTYPE), np.ascontiguousarray(kp_right, KP_DTYPE)
    dl, dr = _a(desc_left, np.uint8).reshape(-1, 32), _a(desc_right, np.uint8).reshape(-1, 32)
    u, d = np.zeros(max(len(kl), 1), np.float32), np.zeros(max(len(kl), 1), np.float32)
    f = L.reff_stereo_matches
    f.restype = None
    f.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    f(nl, pl, pr, w.ctypes.data, h.ctypes.data, sf.ctypes.data, isf.ctypes.data, kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data,
      dr.ctypes.data, len(kr), float(mbf), float(mb), u.ctypes.data, d.ctypes.data)
    return u[:len(kl)], d[:len(kl)]


def ref_grid_and_areas(x, y, octave, bounds, queries, levels):
    """Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea of the reference: (cell_offsets, cell_features, per-query index lists)."""
    L = C.CDLL(FLIB)
    x, y, oc, b = _a(x, np.float32), _a(y, np.float32), _a(octave, np.int32), _a(bounds, np.float32)
    q, ql = _a(queries, np.float32).reshape(-1, 3), _a(levels, np.int32).reshape(-1, 2)
    n, nq = len(x), len(q)
    off, feat = np.zeros(64 * 48 + 1, np.int32), np.zeros(max(n, 1), np.int32)
    cap = max(n, 1) * max(nq, 1)
    qoff, qidx = np.zeros(nq + 1, np.int32), np.zeros(cap, np.int32)
    f = L.reff_grid_and_areas
    f.restype = C.c_int
    f.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int]
    e = f(n, x.ctypes.data, y.ctypes.data, oc.ctypes.data, b.ctypes.data, off.ctypes.data, feat.ctypes.data, nq, q.ctypes.data, ql.ctypes.data,
          qoff.ctypes.data, qidx.ctypes.data, cap)
    return off, feat[:e], [qidx[qoff[k]:qoff[k + 1]].copy() for k in range(nq)]
