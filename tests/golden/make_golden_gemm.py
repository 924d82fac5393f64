#!/usr/bin/env python3
"""Golden vectors for the small float matrix products of the projection searches (src/ORBmatcher.cc:1344-1365, 667-673):
cv::Mat expressions `Rcw*x3Dw+tcw` lower to cv::gemm(A, B, 1, C, 1); `cv::norm(PO)` is the L2 norm.  Generated with the real
cv2 4.13.0 (the only OpenCV in this image); tests/test_oracle_golden.py checks the oracle's restatement against them.
Run in the build container:  python tests/golden/make_golden_gemm.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(1234)
k = 4000
R = rng.normal(size=(k, 3, 3)).astype(np.float32)
x = (rng.normal(size=(k, 3)) * rng.choice([0.01, 1.0, 10.0, 300.0], (k, 1))).astype(np.float32)
t = rng.normal(size=(k, 3)).astype(np.float32)
out = np.stack([cv2.gemm(R[i], x[i].reshape(3, 1), 1.0, t[i].reshape(3, 1), 1.0).reshape(3) for i in range(k)])
out0 = np.stack([cv2.gemm(R[i], x[i].reshape(3, 1), 1.0, None, 0.0).reshape(3) for i in range(k)])
outT = np.stack([cv2.gemm(R[i], x[i].reshape(3, 1), -1.0, None, 0.0, flags=cv2.GEMM_1_T).reshape(3) for i in range(k)])   # -R^T x
norm = np.array([cv2.norm(x[i].reshape(3, 1)) for i in range(k)], np.float64)
np.savez_compressed(os.path.join(HERE, "prim_gemm3.npz"), R=R, x=x, t=t, out=out.astype(np.float32), out0=out0.astype(np.float32),
                    outT=outT.astype(np.float32), norm=norm)
print("wrote prim_gemm3.npz", out.dtype, out0.dtype)
