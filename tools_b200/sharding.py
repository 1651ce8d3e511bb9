"""Target sharding across GPUs: targets are independent given the key, so each rank takes a
contiguous slice of the global target index range and the key is replicated (no collective on the
data path).  Passing the slice start as `first_index` keeps the Philox streams aligned with the
single-GPU run."""


def shard(total: int, rank: int, world: int):
    """Contiguous slice [start, stop) of range(total) for `rank` of `world`; sizes differ by <= 1."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
