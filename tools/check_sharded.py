"""Under torchrun: shard.ShardedSearch (1/N upload + NCCL all-gather of the matrix + getPart rows per rank) against the C
oracle on every rank's rows, and the all-gathered table against the whole.  TEST INFRASTRUCTURE use of oracle/."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import c_oracle  # noqa: E402
from wisecondor_b200 import shard, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bins = [int(b) for b in np.maximum(1, np.array(synth.chrom_bins(250000)) // 3)]
X = synth.corrected_like(bins, 61, seed=9)
n = X.shape[0]
job = shard.ShardedSearch(n, 61, 100, rank, world, dev)
idx, dst = job.run(torch.from_numpy(X).pin_memory(), bins)
oidx, odst = c_oracle.get_reference_rows(X, bins, job.r0, job.r1, 100)
ok = bool(np.array_equal(idx, oidx) and np.array_equal(dst, odst))
full_i, full_d = shard.allgather_rows(job.idx, job.dist, n)
if rank == 0:
    wi, wd = c_oracle.get_reference_rows(X, bins, 0, n, 100)
    ok = ok and bool(np.array_equal(full_i.cpu().numpy(), wi) and np.array_equal(full_d.cpu().numpy(), wd))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded search on %d ranks: %s" % (world, "identical to the oracle" if flag.item() == 1 else "MISMATCH"))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
