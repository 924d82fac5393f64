"""Times the octree stage alone for several image sizes (experiment helper). usage: ORBX_OCT_T=128 python tools/octree_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200 import capi
from orbslam_mapsave_b200.synth import synth
out = []
for (W, H, nf, nl, B) in [(640, 480, 1000, 8, 256), (1280, 720, 2000, 8, 64), (3840, 2160, 8000, 12, 8)]:
    frames = torch.from_numpy(np.stack([synth(W, H, s) for s in range(min(B, 4))] * (B // min(B, 4)))).cuda()
    ex = orb.ORBextractor(nf, 1.2, nl, 20, 7, W, H, max_batch=B)
    cap = ex.max_keypoints()
    kp = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    st = ts.cuda_stream
    ex.extract_batch_device(frames, kp, desc, n, cap, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): ex.extract_batch_device(frames, kp, desc, n, cap, stream=st, stages=capi.STAGE_OCTREE)
    e0.record(ts)
    for _ in range(10): ex.extract_batch_device(frames, kp, desc, n, cap, stream=st, stages=capi.STAGE_OCTREE)
    e1.record(ts); torch.cuda.synchronize()
    out.append(f"{W}x{H}: {e0.elapsed_time(e1) / 10 / B * 1e3:.3f} us/frame")
print("ORBX_OCT_T=" + os.environ.get("ORBX_OCT_T", "default"), " | ".join(out), flush=True)
