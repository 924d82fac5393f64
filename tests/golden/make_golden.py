#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the REAL OpenCV (cv2 4.13.0) primitives.

The reference (skaegy/ORBSLAM_MapSave) has no tests or golden vectors for the ORB path and cannot be compiled
here (OpenCV/Boost C++ headers absent).  Its arithmetic lives in OpenCV calls (src/ORBextractor.cc:808,813,
1089,1123,1125,102), so the oracle is pinned against those primitives as exposed by Python cv2:

  prim_*.npz      inputs + cv2 outputs for cv::FAST(t, nms=true), cv::resize(INTER_LINEAR), cv::GaussianBlur(7x7,2),
                  cv::copyMakeBorder(REFLECT_101), cv::fastAtan2
  chain_*.npz     full extractor outputs from `cv_chain_extract` below: the control flow of
                  ORBextractor::operator() restated in Python but calling the real cv2 primitives at every
                  OpenCV call site (an implementation independent of oracle/orb_oracle.cpp)

Run in the build container:  python tests/golden/make_golden.py     (needs cv2; not run on the GPU box)
"""
import ctypes
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from orbslam_mapsave_b200.synth import synth  # noqa: E402

cv2.setNumThreads(1)
_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
f32 = np.float32

PATTERN = None


def load_pattern():
    global PATTERN
    if PATTERN is None:
        txt = open(os.path.join(HERE, "..", "..", "orbslam_mapsave_b200", "csrc", "orb_pattern_31.inc")).read()
        txt = txt[txt.index("*/") + 2:]
        PATTERN = np.array([int(v) for v in txt.replace("\n", "").split(",") if v.strip()], np.int32).reshape(512, 2)
    return PATTERN


def cv_round(v):
    return int(np.rint(v))


# ---------------------------------------------------------------- octree (reference ORBextractor.cc:538-762)
class _Node:
    __slots__ = ("keys", "UL", "UR", "BL", "BR", "nomore", "seq", "alive")

    def __init__(self):
        self.keys = []
        self.nomore = False
        self.alive = True


def _divide(p):
    halfX = int(math.ceil(float(f32(p.UR[0] - p.UL[0]) / f32(2))))
    halfY = int(math.ceil(float(f32(p.BR[1] - p.UL[1]) / f32(2))))
    n1, n2, n3, n4 = _Node(), _Node(), _Node(), _Node()
    n1.UL = p.UL; n1.UR = (p.UL[0] + halfX, p.UL[1]); n1.BL = (p.UL[0], p.UL[1] + halfY); n1.BR = (p.UL[0] + halfX, p.UL[1] + halfY)
    n2.UL = n1.UR; n2.UR = p.UR; n2.BL = n1.BR; n2.BR = (p.UR[0], p.UL[1] + halfY)
    n3.UL = n1.BL; n3.UR = n1.BR; n3.BL = p.BL; n3.BR = (n1.BR[0], p.BL[1])
    n4.UL = n3.UR; n4.UR = n2.BR; n4.BL = n3.BR; n4.BR = p.BR
    for k in p.keys:
        if k[0] < n1.UR[0]:
            (n1 if k[1] < n1.BR[1] else n3).keys.append(k)
        elif k[1] < n1.BR[1]:
            n2.keys.append(k)
        else:
            n4.keys.append(k)
    for n in (n1, n2, n3, n4):
        if len(n.keys) == 1:
            n.nomore = True
    return n1, n2, n3, n4


def py_octree(keys, minX, maxX, minY, maxY, N):
    """keys: list of (x, y, response) floats relative to minBorder.  Python list used as the std::list
    (index 0 = front).  Pointer tie-break -> creation sequence number (canonical rule)."""
    seq = [0]
    nIni = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))   # C round() for positive values
    hX = f32(maxX - minX) / f32(nIni)
    L = []
    for i in range(nIni):
        n = _Node()
        n.UL = (int(hX * f32(i)), 0); n.UR = (int(hX * f32(i + 1)), 0)
        n.BL = (n.UL[0], maxY - minY); n.BR = (n.UR[0], maxY - minY)
        n.seq = seq[0]; seq[0] += 1
        L.append(n)
    ini = list(L)
    for k in keys:
        ini[int(f32(k[0]) / hX)].keys.append(k)
    L = [n for n in L if len(n.keys) > 0]
    for n in L:
        if len(n.keys) == 1:
            n.nomore = True
    finish = False
    while not finish:
        prev = len(L)
        vsize = []
        nToExpand = 0
        newfront = []       # children in creation order; final list = reversed(newfront) + survivors
        survivors = []
        for n in L:
            if n.nomore:
                survivors.append(n)
                continue
            for c in _divide(n):
                if len(c.keys) > 0:
                    c.seq = seq[0]; seq[0] += 1
                    newfront.append(c)
                    if len(c.keys) > 1:
                        nToExpand += 1
                        vsize.append(c)
        L = newfront[::-1] + survivors
        if len(L) >= N or len(L) == prev:
            finish = True
        elif len(L) + nToExpand * 3 > N:
            while not finish:
                prev = len(L)
                vprev = sorted(vsize, key=lambda n: (len(n.keys), n.seq))
                vsize = []
                front = []
                size = len(L)
                for n in reversed(vprev):
                    for c in _divide(n):
                        if len(c.keys) > 0:
                            c.seq = seq[0]; seq[0] += 1
                            front.append(c)
                            size += 1
                            if len(c.keys) > 1:
                                vsize.append(c)
                    n.alive = False
                    size -= 1
                    if size >= N:
                        break
                L = front[::-1] + [n for n in L if n.alive]
                assert len(L) == size
                if len(L) >= N or len(L) == prev:
                    finish = True
    out = []
    for n in L:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        out.append(best)
    return out


# ---------------------------------------------------------------- the extractor chained from cv2 primitives
def cv_chain_extract(img, nfeatures, scaleFactor, nlevels, iniTh, minTh, mask=None, return_stages=False):
    EDGE = 19
    pat = load_pattern()
    sfd = float(f32(scaleFactor))                      # double member initialised from a float (ORBextractor.h:103)
    sf = [f32(1.0)]
    for i in range(1, nlevels):
        sf.append(f32(float(sf[-1]) * sfd))
    isf = [f32(1.0) / s for s in sf]
    factor = f32(1.0 / sfd)
    nDes = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    quota, tot = [], 0
    for l in range(nlevels - 1):
        quota.append(cv_round(nDes)); tot += quota[-1]
        nDes = nDes * factor
    quota.append(max(nfeatures - tot, 0))
    umax = [0] * 16
    vmax = int(math.floor(float(f32(15) * f32(math.sqrt(2.0)) / f32(2) + f32(1))))
    vmin = int(math.ceil(float(f32(15) * f32(math.sqrt(2.0)) / f32(2))))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(225.0 - v * v))
    v0 = 0
    for v in range(15, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1

    image = img.copy()
    if mask is not None:
        image = np.where(mask != 0, img, 0).astype(np.uint8)
    rows, cols = image.shape
    pyr = []
    for l in range(nlevels):
        w, h = cv_round(f32(cols) * isf[l]), cv_round(f32(rows) * isf[l])
        if l == 0:
            pyr.append(image)
        else:
            pyr.append(cv2.resize(pyr[l - 1], (w, h), interpolation=cv2.INTER_LINEAR))
    bordered = [cv2.copyMakeBorder(p, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101) for p in pyr]

    det_ini = cv2.FastFeatureDetector_create(iniTh, True)
    det_min = cv2.FastFeatureDetector_create(minTh, True)
    cands, lvl_kps = [], []
    for l in range(nlevels):
        im = pyr[l]
        minBX = minBY = EDGE - 3
        maxBX, maxBY = im.shape[1] - EDGE + 3, im.shape[0] - EDGE + 3
        width, height = f32(maxBX - minBX), f32(maxBY - minBY)
        nCols, nRows = int(width / f32(30)), int(height / f32(30))
        wCell, hCell = int(math.ceil(float(width / f32(nCols)))), int(math.ceil(float(height / f32(nRows))))
        todist = []
        for i in range(nRows):
            iniY = minBY + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                roi = im[iniY:maxY, iniX:maxX]
                kps = det_ini.detect(roi)
                if len(kps) == 0:
                    kps = det_min.detect(roi)
                for k in kps:
                    todist.append((float(k.pt[0]) + j * wCell, float(k.pt[1]) + i * hCell, float(k.response)))
        cands.append(todist)
        kept = py_octree(todist, minBX, maxBX, minBY, maxBY, quota[l])
        size = float(int(f32(31) * sf[l]))
        kps = []
        for (x, y, r) in kept:
            kps.append([x + minBX, y + minBY, size, -1.0, r, l, -1])
        lvl_kps.append(kps)
    # orientation (IC_Angle on the un-blurred level)
    for l in range(nlevels):
        im = pyr[l].astype(np.int64)
        for kp in lvl_kps[l]:
            cx, cy = cv_round(kp[0]), cv_round(kp[1])
            m10 = m01 = 0
            for v in range(-15, 16):
                d = umax[abs(v)]
                row = im[cy + v, cx - d:cx + d + 1]
                us = np.arange(-d, d + 1)
                m10 += int((us * row).sum())
                m01 += v * int(row.sum())
            kp[3] = float(cv2.fastAtan2(float(m01), float(m10)))
    # descriptors on the blurred clone
    all_kp, all_desc, blurred = [], [], []
    factorPI = f32(math.pi / float(f32(180.0)))
    for l in range(nlevels):
        if not lvl_kps[l]:
            blurred.append(None)
            continue
        work = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        blurred.append(work)
        px, py = pat[:, 0].astype(np.float32), pat[:, 1].astype(np.float32)
        for kp in lvl_kps[l]:
            ang = f32(kp[3]) * factorPI
            a, b = f32(_libm.cosf(float(ang))), f32(_libm.sinf(float(ang)))
            cx, cy = cv_round(kp[0]), cv_round(kp[1])
            yy = np.rint(px * b + py * a).astype(np.int64) + cy
            xx = np.rint(px * a - py * b).astype(np.int64) + cx
            vals = work[yy, xx].astype(np.int32)
            bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
            all_desc.append(np.packbits(bits, bitorder="little"))
            x, y = f32(kp[0]), f32(kp[1])
            if l != 0:
                x, y = x * sf[l], y * sf[l]
            all_kp.append((x, y, kp[2], kp[3], kp[4], kp[5], kp[6]))
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    kp_arr = np.array(all_kp, dtype=kp_dtype) if all_kp else np.zeros(0, kp_dtype)
    desc_arr = np.stack(all_desc) if all_desc else np.zeros((0, 32), np.uint8)
    if return_stages:
        return kp_arr, desc_arr, dict(pyr=pyr, bordered=bordered, cands=cands, blurred=blurred, quota=quota,
                                      sf=np.array(sf, np.float32), umax=umax)
    return kp_arr, desc_arr


# ---------------------------------------------------------------- fixture writers
def make_primitives():
    rng = np.random.default_rng(12345)
    base = synth(640, 480, 7)
    # FAST: random ROIs (random-noise, synthetic-scene, tiny, and non-contiguous views), thresholds 20 and 7
    rois, outs = [], {}
    for i in range(24):
        if i % 3 == 0:
            h, w = int(rng.integers(7, 41)), int(rng.integers(7, 41))
            roi = rng.integers(0, 256, (h, w), dtype=np.uint8)
        else:
            h, w = int(rng.integers(20, 45)), int(rng.integers(20, 45))
            y, x = int(rng.integers(0, 480 - h)), int(rng.integers(0, 640 - w))
            roi = base[y:y + h, x:x + w]
        rois.append(np.ascontiguousarray(roi))
    fast = {}
    for i, roi in enumerate(rois):
        fast[f"roi{i}"] = roi
        for t in (20, 7):
            view = base.copy()[:roi.shape[0], :roi.shape[1]]          # exercise a strided (non-contiguous) view
            view[:] = roi
            kps = cv2.FastFeatureDetector_create(t, True).detect(view)
            fast[f"roi{i}_t{t}"] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps], np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, "prim_fast.npz"), **fast)
    # FAST without NMS on two ROIs (score definition)
    nn = {}
    for i in (1, 2):
        kps = cv2.FastFeatureDetector_create(7, False).detect(rois[i])
        nn[f"roi{i}"] = rois[i]
        nn[f"roi{i}_t7"] = np.array([(int(k.pt[0]), int(k.pt[1])) for k in kps], np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "prim_fast_nonms.npz"), **nn)
    # resize: checksums for big cases, full arrays for small ones
    rs = {}
    small = synth(160, 120, 3)
    rs["small_src"] = small
    rs["small_133x100"] = cv2.resize(small, (133, 100), interpolation=cv2.INTER_LINEAR)
    rs["small_111x83"] = cv2.resize(rs["small_133x100"], (111, 83), interpolation=cv2.INTER_LINEAR)
    rnd = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    rs["rnd_src"] = rnd
    rs["rnd_109x81"] = cv2.resize(rnd, (109, 81), interpolation=cv2.INTER_LINEAR)
    rs["rnd_200x150"] = cv2.resize(rnd, (200, 150), interpolation=cv2.INTER_LINEAR)      # upscale: exercises clamps
    np.savez_compressed(os.path.join(HERE, "prim_resize.npz"), **rs)
    # blur + border
    bl = {"small_blur": cv2.GaussianBlur(small, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101),
          "rnd_blur": cv2.GaussianBlur(rnd, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101),
          "rnd_border19": cv2.copyMakeBorder(rnd, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)}
    np.savez_compressed(os.path.join(HERE, "prim_blur.npz"), **bl)
    # fastAtan2
    yx = rng.integers(-200000, 200000, (4000, 2)).astype(np.float32)
    yx[:50] = rng.integers(-3, 4, (50, 2)).astype(np.float32)
    yx[50] = (0, 0)
    at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "prim_atan2.npz"), yx=yx, angle=at)


def make_chain():
    cases = [
        ("c1_640x480_seed0", dict(W=640, H=480, seed=0, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniTh=20, minTh=7, mask=False)),
        ("small_320x240_seed5", dict(W=320, H=240, seed=5, nfeatures=500, scaleFactor=1.2, nlevels=6, iniTh=20, minTh=7, mask=False)),
        ("mask_400x300_seed9", dict(W=400, H=300, seed=9, nfeatures=600, scaleFactor=1.2, nlevels=7, iniTh=20, minTh=7, mask=True)),
        ("rgbd_424x240_seed11", dict(W=424, H=240, seed=11, nfeatures=300, scaleFactor=1.5, nlevels=3, iniTh=15, minTh=3, mask=False)),
        # BASELINE config 3 geometry: 16:9, so DistributeOctTree starts from nIni = round(1280/720) = 2 root nodes (:542-562)
        ("c3_1280x720_seed3", dict(W=1280, H=720, seed=3, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniTh=20, minTh=7, mask=False)),
        # the values the fork ships in Examples/ORB_RGB640x480.yaml (2000 features, iniThFAST 32); image regenerated from the seed
        ("fork_yaml_640x480_seed2", dict(W=640, H=480, seed=2, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniTh=32, minTh=7, mask=False,
                                         store_image=False)),
        # BASELINE config 5 (4K, 8000 features, 12 levels).  The 8 MB image is not stored: the tests regenerate it with
        # synth(W, H, seed) and check its CRC against the one recorded here.
        ("c5_3840x2160_seed1", dict(W=3840, H=2160, seed=1, nfeatures=8000, scaleFactor=1.2, nlevels=12, iniTh=20, minTh=7, mask=False,
                                    store_image=False)),
    ]
    only = set(sys.argv[2:]) if len(sys.argv) > 2 and sys.argv[1] == "chain" else None
    for name, c in cases:
        if only is not None and name not in only:
            continue
        img = synth(c["W"], c["H"], c["seed"])
        mask = None
        if c["mask"]:
            mask = np.full(img.shape, 255, np.uint8)
            mask[60:220, 150:260] = 0                      # an OpDetector-style "person" hole (DetectHumanPose.cpp:291-301)
        kp, desc, st = cv_chain_extract(img, c["nfeatures"], c["scaleFactor"], c["nlevels"], c["iniTh"], c["minTh"], mask, True)
        cand_counts = np.array([len(x) for x in st["cands"]], np.int32)
        lvl_sum = np.array([int(p.astype(np.int64).sum()) for p in st["pyr"]], np.int64)
        blur_sum = np.array([int(b.astype(np.int64).sum()) if b is not None else -1 for b in st["blurred"]], np.int64)
        params = np.array([c["W"], c["H"], c["seed"], c["nfeatures"], c["nlevels"], c["iniTh"], c["minTh"], int(c["mask"])], np.int32)
        import zlib
        extra = dict(image=img) if c.get("store_image", True) else dict(image_crc=np.uint32(zlib.crc32(img.tobytes())))
        np.savez_compressed(os.path.join(HERE, f"chain_{name}.npz"), params=params, scaleFactor=np.float32(c["scaleFactor"]),
                            kp=kp, desc=desc, cand_counts=cand_counts, level_sums=lvl_sum, blur_sums=blur_sum,
                            quota=np.array(st["quota"], np.int32), last_level=st["pyr"][-1], **extra)
        print(name, len(kp), cand_counts.tolist())


if __name__ == "__main__":
    # `python make_golden.py` regenerates everything; `python make_golden.py chain <name>...` only the named chain cases
    if not (len(sys.argv) > 1 and sys.argv[1] == "chain"):
        make_primitives()
    make_chain()
