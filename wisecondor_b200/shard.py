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


class ShardedSearch(object):
    """newref search of one matrix on all ranks of the group, from HOST buffers (the multi-GPU form of
    device.newref_topk_host).  Every rank uploads only its slice of the corrected matrix over PCIe; one NCCL
    all-gather over NVLink assembles the full matrix on every GPU (277 MB at 600 x 50 kb: < 1 ms, against ~9 ms
    for eight simultaneous full uploads); each rank then searches its getPart rows and copies them back to (pinned)
    host memory - the counterpart of the reference's per-part npz files (wisecondor.py:128-132)."""

    def __init__(self, n, s, refsize, rank, world, device, group=None):
        from . import device as _dev
        self._dev = _dev
        self.n, self.s, self.k, self.rank, self.world, self.group = int(n), int(s), int(refsize), rank, world, group
        self.device = device
        self.rows_per = (self.n + world - 1) // world                 # matrix slices: equal, padded
        self.local = torch.empty((self.rows_per, self.s), dtype=torch.float64, device=device)
        self.full = torch.empty((world * self.rows_per, self.s), dtype=torch.float64, device=device)
        self.r0, self.r1 = row_shard(rank, world, self.n)            # search rows: the reference's getPart
        rows = self.r1 - self.r0
        self.idx = torch.empty((rows, self.k), dtype=torch.int32, device=device)
        self.dist = torch.empty((rows, self.k), dtype=torch.float64, device=device)
        self.h_idx = torch.empty((rows, self.k), dtype=torch.int32).pin_memory()
        self.h_dist = torch.empty((rows, self.k), dtype=torch.float64).pin_memory()

    def run(self, x_host, chrom_bins):
        """x_host: pinned CPU tensor [N][S] float64 (every rank holds it, reads only its slice).  Returns this rank's
        rows (indexes, distances) as numpy views of pinned memory."""
        a = self.rank * self.rows_per
        b = min(self.n, a + self.rows_per)
        if b > a:
            self.local[:b - a].copy_(x_host[a:b], non_blocking=True)
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, self.local, group=self.group)
            x = self.full[:self.n]
        else:
            x = self.local[:self.n]
        self._dev.newref_topk(x, chrom_bins, self.r0, self.r1, self.k, self.idx, self.dist)
        self.h_idx.copy_(self.idx, non_blocking=True)
        self.h_dist.copy_(self.dist, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.h_idx.numpy(), self.h_dist.numpy()
