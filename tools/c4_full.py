"""BASELINE.json config 4 at FULL size: all-pairs keyframe descriptor matching over K keyframes x P descriptors (default
10 000 x 2 000 = 640 MB of descriptors, replicated on every GPU), sharded by query keyframe across the ranks, with the one real
exchange of the path: an NCCL all-gather of the per-rank uint16 match-count tables (K x K x 2 B = 200 MB in total).

  python tools/c4_full.py --kf 1000                                    (1 GPU smoke run)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/c4_full.py

Prints one JSON line (rank 0): seconds (CUDA events, max over ranks), pairs/s, table checksum and the planted-match sanity check."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbslam_mapsave_b200 import capi                     # noqa: E402
from orbslam_mapsave_b200.sharding import shard_range, gather_match_tables   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kf", type=int, default=10000)
    ap.add_argument("--per", type=int, default=2000)
    ap.add_argument("--chunk", type=int, default=50, help="query keyframes per kernel launch")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, P = a.kf, a.per
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)                                   # every rank builds the same database
    db = torch.randint(0, 256, (K, P, 32), dtype=torch.uint8, device="cuda", generator=g)
    # keyframe k re-observes the first quarter of keyframe k-1's descriptors with ~2 % of the bits flipped
    q = P // 4
    fb = torch.zeros((K - 1, q, 32), dtype=torch.uint8, device="cuda")
    for bit in range(8):
        fb |= (torch.rand((K - 1, q, 32), device="cuda", generator=g) < 0.02).to(torch.uint8) << bit
    for k in range(1, K):
        db[k, :q] = db[k - 1, :q] ^ fb[k - 1]
    del fb
    q0, q1 = shard_range(K, rank, world)
    nq = q1 - q0
    cnt = torch.zeros((nq * K + 1) // 2 * 2, dtype=torch.uint16, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    lib = capi.lib()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for c0 in range(q0, q1, a.chunk):
        c1 = min(q1, c0 + a.chunk)
        off = (c0 - q0) * K
        assert off % 2 == 0 or a.chunk % 2 == 0
        capi.check(lib.orbm_allpairs_device(capi._p(db), K, P, c0, c1, 50, 0.75, cnt[off:].data_ptr(), None, None, stream))
    e1.record()
    table = cnt[: nq * K].view(nq, K)
    full = gather_match_tables(table, K, rank, world) if world > 1 else table
    e2.record()
    torch.cuda.synchronize()
    secs = torch.tensor([e0.elapsed_time(e2) / 1e3, e0.elapsed_time(e1) / 1e3, e1.elapsed_time(e2) / 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = full.to(torch.int64)
        band = torch.diagonal(t, offset=-1)                # count[k, k-1]: the planted re-observations
        far = t[torch.triu_indices(K, K, 60, device="cuda").unbind()]     # 60 keyframes apart: ~45 % of the bits re-flipped
        pairs = K * (K - 1) * P * P
        print(json.dumps({
            "workload": f"C4 all-pairs: {K} keyframes x {P} descriptors, top-2 + ratio 0.75 + TH_LOW, query-keyframe sharded x{world}",
            "n_gpus": world, "seconds_total": secs[0].item(), "seconds_matching": secs[1].item(), "seconds_all_gather": secs[2].item(),
            "pairs": pairs, "pairs_per_s": pairs / secs[0].item(), "match_table_bytes": int(full.numel() * 2),
            "table_checksum": int(t.sum().item()), "planted_band_min": int(band.min().item()), "planted_band_mean": float(band.float().mean().item()),
            "unrelated_max": int(far.max().item()), "diag_zero": bool((torch.diagonal(t) == 0).all().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
