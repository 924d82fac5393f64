"""Per-chunk device timeline of orbx_extract_batch (ORBX_E2E_TRACE=1): where the host-buffer path spends its step.
usage: ORBX_E2E_TRACE=1 python tools/e2e_trace.py [frames] [chunk]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orbslam_mapsave_b200 as orb  # noqa: E402
from orbslam_mapsave_b200 import capi  # noqa: E402
from orbslam_mapsave_b200.synth import synth  # noqa: E402

nF = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 256
uniq = np.stack([synth(640, 480, s) for s in range(64)])
h_frames = torch.from_numpy(np.concatenate([uniq] * (nF // 64))).pin_memory()
ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=chunk)
cap = ex.max_keypoints()
h_kp = torch.zeros((nF, cap, 7), dtype=torch.float32).pin_memory()
h_desc = torch.zeros((nF, cap, 32), dtype=torch.uint8).pin_memory()
h_n = np.zeros(nF, np.int32)
for i in range(3):
    if i < 2:
        sys.stderr.write(f"--- call {i} (warm-up)\n")
    t0 = time.perf_counter()
    capi.check(capi.lib().orbx_extract_batch(ex.handle, capi._p(h_frames), nF, 640, 480, 640, 640 * 480, None, 0, 0, capi._p(h_kp),
                                             capi._p(h_desc), cap, capi._p(h_n)))
    sys.stderr.write(f"call {i}: {1e3 * (time.perf_counter() - t0):.2f} ms wall\n")
