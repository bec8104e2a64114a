"""Multi-GPU plumbing: the path shards by read batch — the index is replicated per GPU, reads are
split into contiguous shards (mates of a pair stay together), there is NO collective on the alignment path; only the
timing / counters of a benchmark are reduced (max / sum) over torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import os


def env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_pairs(n_pairs: int, rank: int, world: int):
    """Contiguous shard of read PAIRS for `rank`: returns (first_read, last_read_exclusive) in read indices with
    reads 2i, 2i+1 = mates of pair i. Shards differ in size by at most one pair and cover every pair exactly once."""
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return 2 * lo, 2 * hi


def shard_srand_base(srand_base: int, first_read: int) -> int:
    """RANSAC stream of read i is srand(srand_base + i) with i the GLOBAL read index: a shard that
    starts at first_read passes this as its local base, so results do not depend on the sharding."""
    return (srand_base + first_read) & 0xFFFFFFFF


def reduce_max_sum(values_max, values_sum, device=None):
    """max over ranks of values_max, sum over ranks of values_sum (lists of floats)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values_max), list(values_sum)
    mx = torch.tensor(values_max, dtype=torch.float64, device=device)
    sm = torch.tensor(values_sum, dtype=torch.float64, device=device)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx.tolist(), sm.tolist()
