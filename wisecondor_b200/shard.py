"""Partitioning of the hot path over the GPUs of one node (one process per GPU, torch.distributed).

newref: target-bin rows are sharded with the reference's own getPart (wisetools.py:358-361) - rank r of W computes
exactly what `newrefpart r+1 W` would - every rank holds the full corrected matrix, and one all-gather of the
(rows x refsize) int32 + float64 blocks replaces toolNewrefPost's file concatenation (wisecondor.py:146-158).
test: samples are sharded; no communication.  Works with the nccl backend (CUDA tensors) and gloo (CPU tensors; used
by the CPU tests).
"""
import torch
import torch.distributed as dist


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


def allgather_rows(idx_local, dist_local, bincount, group=None, out_idx=None, out_dist=None, scratch=None):
    """Concatenate every rank's (rows x k) result block in rank order -> (bincount x k) on every rank.

    Parts differ by at most one row, so blocks are padded to the largest part for all_gather_into_tensor and the
    padding rows are dropped afterwards.  `scratch` may carry preallocated (pad_idx, pad_dist, all_idx, all_dist)."""
    world = dist.get_world_size(group)
    k = idx_local.shape[1]
    dev = idx_local.device
    rmax = max_shard_rows(world, bincount)
    if scratch is None:
        scratch = (torch.empty((rmax, k), dtype=idx_local.dtype, device=dev),
                   torch.empty((rmax, k), dtype=dist_local.dtype, device=dev),
                   torch.empty((world * rmax, k), dtype=idx_local.dtype, device=dev),
                   torch.empty((world * rmax, k), dtype=dist_local.dtype, device=dev))
    pad_idx, pad_dist, all_idx, all_dist = scratch
    rows = idx_local.shape[0]
    if pad_idx.data_ptr() != idx_local.data_ptr():      # the caller may have let the search write into the scratch
        pad_idx[:rows].copy_(idx_local)
    if pad_dist.data_ptr() != dist_local.data_ptr():
        pad_dist[:rows].copy_(dist_local)
    dist.all_gather_into_tensor(all_idx, pad_idx, group=group)
    dist.all_gather_into_tensor(all_dist, pad_dist, group=group)
    if out_idx is None:
        out_idx = torch.empty((bincount, k), dtype=idx_local.dtype, device=dev)
        out_dist = torch.empty((bincount, k), dtype=dist_local.dtype, device=dev)
    lens = [row_shard(r, world, bincount)[1] - row_shard(r, world, bincount)[0] for r in range(world)]
    torch.cat([all_idx[r * rmax:r * rmax + lens[r]] for r in range(world)], out=out_idx)      # drop the padding rows
    torch.cat([all_dist[r * rmax:r * rmax + lens[r]] for r in range(world)], out=out_dist)
    return out_idx, out_dist
