"""Does concurrent H2D/D2H traffic slow the extraction kernels down? (experiment helper)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth
W, H, B, NP = 640, 480, 256, 16
frames = torch.from_numpy(np.stack([synth(W, H, s) for s in range(32)] * (B // 32))).cuda()
ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, W, H, max_batch=B)
cap = ex.max_keypoints()
kp = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
n = torch.zeros(B, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream(); st = ts.cuda_stream
h_in = torch.zeros((B, H, W), dtype=torch.uint8).pin_memory(); d_in = torch.zeros((B, H, W), dtype=torch.uint8, device="cuda")
h_out = torch.zeros((B, cap, 60), dtype=torch.uint8).pin_memory(); d_out = torch.zeros((B, cap, 60), dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

def run(copies):
    for _ in range(3): ex.extract_batch_device(frames, kp, desc, n, cap, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for i in range(NP):
        if copies & 1:
            with torch.cuda.stream(s_in): d_in.copy_(h_in, non_blocking=True)
        if copies & 2:
            with torch.cuda.stream(s_out): h_out.copy_(d_out, non_blocking=True)
        ex.extract_batch_device(frames, kp, desc, n, cap, stream=st)
    e1.record(ts)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for c, name in [(0, "no copies"), (1, "H2D"), (2, "D2H"), (3, "H2D + D2H")]:
    print(f"{name:10s}: {run(c):.2f} ms per {NP} passes of {B} frames", flush=True)
