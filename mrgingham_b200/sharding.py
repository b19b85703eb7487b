"""How a batch of frames is split over the GPUs of one node.

Images are independent (the reference parallelises only over images: image i goes to worker
i mod N, mrgingham-from-image.cc:50), so the batch is cut into contiguous shards, one per rank,
and nothing is exchanged on the data path: each rank runs its own detector on its own frames and
the per-frame results (a few KB) are concatenated in frame order by the caller.
"""


def shard_range(nframes, rank, world):
    """[lo, hi) of the frames rank `rank` of `world` processes owns: frame i -> rank floor(i*world/nframes)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (nframes * rank) // world, (nframes * (rank + 1)) // world


def owner_of(frame, nframes, world):
    """inverse of shard_range()"""
    for r in range(world):
        lo, hi = shard_range(nframes, r, world)
        if lo <= frame < hi:
            return r
    raise ValueError("frame out of range")
