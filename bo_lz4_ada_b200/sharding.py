"""Host-side sharding of independent streams across the GPUs of one box.

The path has no exchange step: frames are independent, so every rank decodes its own subset and
nothing but the final timing is reduced.  A frame with a content checksum is never split (its
XXH32 is one serial chain), so the unit of sharding is the whole stream.
"""


def shard_streams(costs, world_size, rank):
    """Greedy longest-processing-time partition.  costs[i] = compressed + decompressed bytes of
    stream i (what the device stage moves).  Returns the sorted stream indices of `rank`.
    Deterministic: every rank computes the same partition without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return [i for i in range(len(costs)) if owner[i] == rank]


def shard_loads(costs, world_size):
    return [sum(costs[i] for i in shard_streams(costs, world_size, r)) for r in range(world_size)]
