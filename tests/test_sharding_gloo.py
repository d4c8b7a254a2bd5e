"""The N > 1 path on CPU: two gloo ranks shard one list of streams, plan their shards with the
host layer (no device needed for planning) and agree that the shards are disjoint, complete and
balanced.  The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bo_lz4_ada_b200 as lz
    from bo_lz4_ada_b200.sharding import shard_streams

    names = sorted(n for n in os.listdir(golden) if n.endswith(".lz4"))
    streams = [open(os.path.join(golden, n), "rb").read() for n in names]
    costs = [len(s) * 3 for s in streams]
    mine = shard_streams(costs, world, rank)
    # plan my shard on the host (block-table builder); no GPU involved
    src = b"".join(streams[i] for i in mine)
    offs, pos = [], 0
    for i in mine:
        offs.append((pos, len(streams[i])))
        pos += len(streams[i])
    batch = lz.Batch(None, src, offs)
    blocks = batch.block_count
    bad = sum(1 for k in range(len(mine)) if batch.host_outcome(k)["exception"] != "OK")
    # exchange shard membership and block counts
    flags = torch.zeros(len(streams), dtype=torch.int64)
    flags[mine] = 1
    dist.all_reduce(flags)
    totals = torch.tensor([blocks, bad, sum(costs[i] for i in mine)], dtype=torch.int64)
    gathered = [torch.zeros_like(totals) for _ in range(world)]
    dist.all_gather(gathered, totals)
    assert bool((flags == 1).all()), "shards must be disjoint and complete"
    assert sum(int(g[1]) for g in gathered) == 0
    # every rank sees the same global block count as a single-rank plan of everything
    if rank == 0:
        all_src = b"".join(streams)
        offs, pos = [], 0
        for s in streams:
            offs.append((pos, len(s)))
            pos += len(s)
        assert lz.Batch(None, all_src, offs).block_count == sum(int(g[0]) for g in gathered)
        loads = [int(g[2]) for g in gathered]
        assert max(loads) <= 1.35 * (sum(loads) / world) + max(costs)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, golden), nprocs=2, join=True)


def test_shard_streams_properties():
    from bo_lz4_ada_b200.sharding import shard_loads, shard_streams
    costs = [5, 1, 9, 3, 3, 7, 2, 8, 4, 6] * 13
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += shard_streams(costs, world, r)
        assert sorted(seen) == list(range(len(costs)))
        loads = shard_loads(costs, world)
        assert max(loads) - min(loads) <= max(costs)
