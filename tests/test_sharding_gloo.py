"""CPU (gloo, world_size 2) test of the N>1 host logic: frame / query-keyframe sharding and the all-gather that assembles
the per-rank match tables.  The per-rank table is filled by the CPU oracle here — on GPUs the same code path
is fed by orbm_allpairs_device (bench.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbslam_mapsave_b200.sharding import gather_match_tables, shard_range
from orbslam_mapsave_b200.synth import synth_descriptors


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 10000):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def _table_rows(desc, q0, q1):
    from oracle import orb_oracle_py as orc
    n_kf = len(desc)
    out = np.zeros((q1 - q0, n_kf), np.int16)
    for qi, q in enumerate(range(q0, q1)):
        for k in range(n_kf):
            if k == q:
                continue
            _, b1, b2 = orc.hamming_top2(desc[q], desc[k])
            out[qi, k] = int(((b1 <= 50) & (b1.astype(np.float32) < np.float32(0.75) * b2.astype(np.float32))).sum())
    return out


def _worker(rank, world, port, n_kf, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = synth_descriptors(48, 5)
    desc = np.stack([synth_descriptors(48, 100 + k, dup_of=base, dup_rate=0.5) for k in range(n_kf)])
    q0, q1 = shard_range(n_kf, rank, world)
    local = torch.from_numpy(_table_rows(desc, q0, q1))
    full = gather_match_tables(local, n_kf)
    # every rank must hold the same full table
    raw = full.contiguous().view(torch.uint8)
    chk = [torch.empty_like(raw) for _ in range(world)]
    dist.all_gather(chk, raw)
    assert all(torch.equal(chk[0], c) for c in chk)
    if rank == 0:
        ret.put(full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_kf", [5, 6])          # uneven and even query shards
def test_match_table_allgather_world2(n_kf):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_kf
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_kf, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = ret.get(timeout=10)
    base = synth_descriptors(48, 5)
    desc = np.stack([synth_descriptors(48, 100 + k, dup_of=base, dup_rate=0.5) for k in range(n_kf)])
    assert np.array_equal(full, _table_rows(desc, 0, n_kf))
    assert full.sum() > 0
