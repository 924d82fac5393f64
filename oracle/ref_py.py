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
