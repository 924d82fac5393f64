"""Pass-size efficiency of the device-resident extractor and the effect of concurrent H2D / D2H traffic on it (experiment helper).
Prints us per frame for passes of 64..1024 frames: alone, and with pinned copies of the e2e path's sizes running on two other streams."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
W, H, TOTAL = 640, 480, 4096
uniq = np.stack([synth(W, H, s) for s in range(64)])
for B in (64, 128, 256, 512, 1024):
    NP = TOTAL // B
    frames = torch.from_numpy(np.concatenate([uniq] * (B // 64))).cuda()
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, W, H, max_batch=B)
    cap = ex.max_keypoints()
    kp = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    st = ts.cuda_stream
    h_in = torch.zeros((B, H, W), dtype=torch.uint8).pin_memory()
    d_in = torch.zeros((B, H, W), dtype=torch.uint8, device="cuda")
    h_out = torch.zeros((B, cap, 60), dtype=torch.uint8).pin_memory()
    d_out = torch.zeros((B, cap, 60), dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = []
    for copies in (0, 1, 3):
        for _ in range(3):
            ex.extract_batch_device(frames, kp, desc, n, cap, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for i in range(NP):
            if copies & 1:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if copies & 2:
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
            ex.extract_batch_device(frames, kp, desc, n, cap, stream=st)
        e1.record(ts)
        torch.cuda.synchronize()
        res.append(1e3 * e0.elapsed_time(e1) / TOTAL)
    print(f"pass of {B:5d} frames: {res[0]:.3f} us/frame alone, {res[1]:.3f} with H2D, {res[2]:.3f} with H2D + D2H", flush=True)
    del ex
