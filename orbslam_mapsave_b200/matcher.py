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

    @staticmethod
    def ComputeThreeMaxima(histo, device=0):
        histo = np.ascontiguousarray(histo, np.int32)
        ind = np.zeros(3, np.int32)
        capi.check(capi.lib().orbm_three_maxima(capi._p(histo), len(histo), capi._p(ind), device))
        return tuple(int(v) for v in ind)


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
