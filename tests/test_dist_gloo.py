"""world_size-2 gloo test of the multi-GPU host logic (no compute): shards cover every pair once, mates stay together,
the per-shard srand base reproduces the global RANSAC streams, and the benchmark reduction is max / sum over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ma_b200 import dist as madist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = madist.shard_pairs(n_pairs, rank, world)
    ms = 100.0 + 7 * rank
    mx, sm = madist.reduce_max_sum([ms], [float(hi - lo)])
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, madist.shard_srand_base(1000, lo)))
    if rank == 0:
        out.put((mx, sm, gathered))
    dist.destroy_process_group()


def test_sharding_and_reduction_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    n_pairs = 1001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, out)) for r in range(2)]
    for p in procs:
        p.start()
    mx, sm, gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mx == [107.0] and sm == [2.0 * n_pairs]
    (lo0, hi0, s0), (lo1, hi1, s1) = gathered
    assert lo0 == 0 and hi0 == lo1 and hi1 == 2 * n_pairs and lo1 % 2 == 0
    assert s0 == 1000 and s1 == 1000 + lo1


def test_shard_pairs_properties():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = madist.shard_pairs(n, r, world)
                assert lo == prev and lo % 2 == 0 and hi >= lo
                prev = hi
            assert prev == 2 * n
