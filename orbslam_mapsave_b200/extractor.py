"""Python mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:50-116) over the C ABI — used by tests and bench.
The C++ drop-in class lives in orbslam_mapsave_b200/host/ORBextractor.{h,cc}; both call the same entry points."""
import ctypes as C

import numpy as np

from . import capi


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1          # include/ORBextractor.h:54

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width=None, height=None, max_batch=1, device=0):
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(scaleFactor), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        self.device, self.max_batch = int(device), int(max_batch)
        self._h = None
        self._size = None
        self.mvImagePyramid = []
        self._tables = None
        if width is not None:
            self._plan(int(width), int(height))

    # -- handle management (the C++ class re-plans the same way when the image size changes)
    def _plan(self, w, h):
        if self._size == (w, h):
            return
        self.close()
        h_ = C.c_void_p()
        capi.check(capi.lib().orbx_create(C.byref(h_), self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST,
                                          self.minThFAST, w, h, self.max_batch, self.device))
        self._h, self._size = h_, (w, h)
        n = self.nlevels
        t = [np.zeros(n, np.float32) for _ in range(4)] + [np.zeros(n, np.int32)]
        capi.check(capi.lib().orbx_tables(self._h, *[capi._p(a) for a in t]))
        self._tables = t

    def close(self):
        if self._h is not None:
            capi.lib().orbx_destroy(self._h)
            self._h, self._size = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def max_keypoints(self):
        return capi.lib().orbx_max_keypoints(self._h)

    # -- reference getters (include/ORBextractor.h:68-88)
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(np.float32(self.scaleFactor))

    def _need_tables(self):
        if self._tables is None:       # the constructor's tables depend on the parameters only: no handle, no device
            n = self.nlevels
            t = [np.zeros(n, np.float32) for _ in range(4)] + [np.zeros(n, np.int32)]
            capi.check(capi.lib().orbx_compute_tables(self.nfeatures, self.scaleFactor, n, *[capi._p(a) for a in t]))
            self._tables = t
        return self._tables

    def GetScaleFactors(self):
        return self._need_tables()[0].copy()

    def GetInverseScaleFactors(self):
        return self._need_tables()[1].copy()

    def GetScaleSigmaSquares(self):
        return self._need_tables()[2].copy()

    def GetInverseScaleSigmaSquares(self):
        return self._need_tables()[3].copy()

    def features_per_level(self):
        return self._need_tables()[4].copy()

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        capi.check(capi.lib().orbx_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    # -- operator()(image, mask, keypoints, descriptors)  (src/ORBextractor.cc:1042-1108)
    def __call__(self, image, mask=None, download_pyramid=True):
        image = np.asarray(image)
        if image.size == 0:
            return None                 # reference: silent return, outputs untouched (:1045-1046)
        assert image.dtype == np.uint8 and image.ndim == 2, "image must be 8UC1 (ORBextractor.cc:1051)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        self._plan(w, h)
        m = None
        if mask is not None and np.asarray(mask).size:
            m = np.asarray(mask)
            assert m.dtype == np.uint8 and m.shape == image.shape
            if m.strides[1] != 1:
                m = np.ascontiguousarray(m)
        cap = self.max_keypoints()
        kp = np.zeros(cap, capi.KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int()
        capi.check(capi.lib().orbx_extract(self._h, capi._p(image), w, h, image.strides[0], capi._p(m),
                                           m.strides[0] if m is not None else 0, capi._p(kp), capi._p(desc), cap, C.byref(n)))
        if download_pyramid:
            self.mvImagePyramid = self.pyramid(0)
        return kp[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, masks=None):
        """images: (B, H, W) uint8 host array.  Returns (kp[B, cap], desc[B, cap, 32], n[B])."""
        images = np.ascontiguousarray(images, np.uint8)
        B, h, w = images.shape
        self._plan(w, h)
        if masks is not None:
            masks = np.ascontiguousarray(masks, np.uint8)
        cap = self.max_keypoints()
        kp = np.zeros((B, cap), capi.KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        capi.check(capi.lib().orbx_extract_batch(self._h, capi._p(images), B, w, h, w, w * h, capi._p(masks), w, w * h,
                                                 capi._p(kp), capi._p(desc), cap, capi._p(n)))
        return kp, desc, n

    def extract_batch_list(self, images):
        """images: a list of (H, W) uint8 host arrays, each its own allocation (row stride = the arrays' common strides[0]).
        Returns (kp[B, cap], desc[B, cap, 32], n[B]) like extract_batch (orbx_extract_batch_ptrs)."""
        B = len(images)
        h, w = images[0].shape
        stride = images[0].strides[0]
        assert all(im.dtype == np.uint8 and im.shape == (h, w) and im.strides == (stride, 1) for im in images)
        self._plan(w, h)
        cap = self.max_keypoints()
        kp = np.zeros((B, cap), capi.KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in images])
        capi.check(capi.lib().orbx_extract_batch_ptrs(self._h, C.cast(ptrs, C.c_void_p), B, w, h, stride, capi._p(kp), capi._p(desc), cap,
                                                      capi._p(n)))
        return kp, desc, n

    def extract_batch_device(self, d_images, d_kp, d_desc, d_n, cap, d_masks=None, stream=None, stages=capi.STAGE_ALL):
        """All arguments are torch CUDA tensors (or objects with data_ptr()); asynchronous on `stream`."""
        B, h, w = d_images.shape
        self._plan(w, h)
        capi.check(capi.lib().orbx_run_stages_device(self._h, capi._p(d_images), B, capi._p(d_masks), capi._p(d_kp),
                                                     capi._p(d_desc), cap, capi._p(d_n), stages,
                                                     C.c_void_p(stream) if stream else None))

    def ComputeStereoMatches(self, right, kp_left, desc_left, kp_right, desc_right, mbf, mb, frame_left=0, frame_right=0):
        """Frame::ComputeStereoMatches (src/Frame.cc:584-756) on the device-resident pyramids of this (left) extractor and
        `right`, after both have extracted their image.  Returns (mvuRight, mvDepth)."""
        kp_left = np.ascontiguousarray(kp_left, capi.KP_DTYPE)
        kp_right = np.ascontiguousarray(kp_right, capi.KP_DTYPE)
        desc_left = np.ascontiguousarray(desc_left, np.uint8).reshape(-1, 32)
        desc_right = np.ascontiguousarray(desc_right, np.uint8).reshape(-1, 32)
        u = np.zeros(max(len(kp_left), 1), np.float32)
        d = np.zeros(max(len(kp_left), 1), np.float32)
        capi.check(capi.lib().orbx_stereo_matches(self._h, right._h, frame_left, frame_right, capi._p(kp_left), capi._p(desc_left),
                                                  len(kp_left), capi._p(kp_right), capi._p(desc_right), len(kp_right), float(mbf), float(mb),
                                                  capi._p(u), capi._p(d)))
        return u[:len(kp_left)], d[:len(kp_left)]

    def check_status(self):
        capi.check(capi.lib().orbx_check_status(self._h))

    def launch_count(self):
        return capi.lib().orbx_launch_count(self._h)

    # -- views of the last pass
    def pyramid_level(self, frame, level, bordered=False):
        w, h = self.level_size(level)
        out = np.zeros((h + 38, w + 38) if bordered else (h, w), np.uint8)
        capi.check(capi.lib().orbx_get_pyramid_level(self._h, frame, level, int(bordered), capi._p(out), out.shape[1]))
        return out

    def pyramid(self, frame, bordered=False):
        """All levels of one frame with a single call (what the C++ class uses for mvImagePyramid)."""
        b = 38 if bordered else 0
        outs = [np.zeros((h + b, w + b), np.uint8) for (w, h) in (self.level_size(l) for l in range(self.nlevels))]
        ptrs = (C.c_void_p * self.nlevels)(*[o.ctypes.data for o in outs])
        strides = (C.c_int * self.nlevels)(*[o.shape[1] for o in outs])
        capi.check(capi.lib().orbx_get_pyramid(self._h, frame, int(bordered), ptrs, strides))
        return outs

    def blurred_level(self, frame, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        capi.check(capi.lib().orbx_get_blurred_level(self._h, frame, level, capi._p(out), w))
        return out

    def candidates(self, frame, level):
        cap = 1 << 14
        while True:
            out = np.zeros(cap, capi.CAND_DTYPE)
            n = C.c_int()
            capi.check(capi.lib().orbx_get_candidates(self._h, frame, level, capi._p(out), cap, C.byref(n)))
            if n.value <= cap:
                return out[:n.value].copy()
            cap = n.value
