"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed for the plumbing.

* Extraction shards by frame: frames are independent, there is NO data-path collective.
* All-pairs keyframe matching shards by query keyframe against a descriptor database replicated on every rank; the only
  exchange is one all-gather of the compact per-rank match tables (uint16 count per (query keyframe, database keyframe)).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [begin, end) of `n` units for `rank` of `world` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_match_tables(local_counts, n_queries, rank=None, world=None):
    """local_counts: uint16 / int16 / int32 tensor [q_local, n_db] for this rank's query shard = shard_range(n_queries, rank, world)
    (the kernel's counts are UNSIGNED 16 bit: view them as torch.uint16, an int16 view turns counts above 32767 negative).
    Returns the full [n_queries, n_db] table on every rank.  Uses the default process group (NCCL on GPUs, gloo on CPU)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return local_counts.clone()
    qmax = (n_queries + world - 1) // world
    pad = torch.zeros((qmax, local_counts.shape[1]), dtype=local_counts.dtype, device=local_counts.device)
    pad[: local_counts.shape[0]] = local_counts
    # gathered as raw bytes: neither NCCL nor gloo has a 16-bit integer type
    raw = pad.view(torch.uint8)
    parts = [torch.empty_like(raw) for _ in range(world)]
    dist.all_gather(parts, raw)
    rows = []
    for r in range(world):
        b, e = shard_range(n_queries, r, world)
        rows.append(parts[r].view(local_counts.dtype)[: e - b])
    return torch.cat(rows, 0)
