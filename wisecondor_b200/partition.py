"""Who computes what (no device, no PyTorch): the partitions of the path over the GPUs of one node.

newref shards target-bin rows with the reference's own getPart (wisetools.py:358-361); test shards samples."""


def row_shard(rank, world, bincount):
    """Rows [start, end) of 0-based part `rank` of `world` (reference wisetools.py:358-361)."""
    return int(bincount / float(world) * rank), int(bincount / float(world) * (rank + 1))


def sample_shard(rank, world, nsamples):
    """Contiguous block of samples for `rank`: sizes differ by at most one."""
    base, extra = divmod(int(nsamples), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_shard_rows(world, bincount):
    return max(row_shard(r, world, bincount)[1] - row_shard(r, world, bincount)[0] for r in range(world))
