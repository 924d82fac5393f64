"""Times orbm_allpairs_device / orbm_hamming_top2 for one or more builds of liborb_b200.so (experiment helper).
usage: python tools/match_probe.py lib1.so [lib2.so ...]"""
import ctypes as C
import hashlib
import sys

import numpy as np
import torch


def run(path, nq=32, ndb=512, per=2000, steps=10):
    lib = C.CDLL(path)
    rng = np.random.default_rng(11)
    db = rng.integers(0, 256, (ndb, per, 32), dtype=np.uint8)
    db[1:, : per // 4] = db[0, : per // 4] ^ (rng.integers(0, 256, (ndb - 1, per // 4, 32), dtype=np.uint8)
                                               * (rng.random((ndb - 1, per // 4, 32)) < 0.05)).astype(np.uint8)
    d_db = torch.from_numpy(db).cuda()
    cnt = torch.zeros(nq * ndb, dtype=torch.int16, device="cuda")
    bk = torch.zeros(nq * per, dtype=torch.int32, device="cuda")
    bd = torch.zeros(nq * per, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f = lib.orbm_allpairs_device
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def step(full=False):
        rc = f(d_db.data_ptr(), ndb, per, 0, nq, 50, 0.75, cnt.data_ptr(), bk.data_ptr() if full else None, bd.data_ptr() if full else None, st)
        assert rc == 0, rc
    step(True)
    torch.cuda.synchronize()
    h = hashlib.sha1(cnt.cpu().numpy().tobytes() + bk.cpu().numpy().tobytes() + bd.cpu().numpy().tobytes()).hexdigest()[:12]
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pairs = nq * (ndb - 1) * per * per
    return pairs / ms * 1e3, h


if __name__ == "__main__":
    for p in sys.argv[1:]:
        v, h = run(p)
        print(f"{p}: {v / 1e9:.1f} G pairs/s  result-hash {h}", flush=True)
