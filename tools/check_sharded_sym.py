"""Under torchrun: shard.SymmetricShardedSearch (block pairs of the symmetric search divided over the ranks; NCCL
all-reduce(MIN) of the thresholds, all-to-all of the column-side candidates, all-gather of the rows) against the C oracle.
TEST INFRASTRUCTURE use of oracle/.   torchrun --nproc-per-node N tools/check_sharded_sym.py [bins_divisor] [S] [k]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import c_oracle  # noqa: E402
from wisecondor_b200 import device, shard, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 61
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
bins = [int(b) for b in np.maximum(1, np.array(synth.chrom_bins(250000)) // div)]
X = synth.corrected_like(bins, S, seed=9)
n = X.shape[0]
job = shard.SymmetricShardedSearch(n, k, rank, world, dev)
x = torch.as_tensor(X, device=dev)
idx, dst = job.run(x, bins)
oidx, odst = c_oracle.get_reference_rows(X, bins, job.row0, max(job.row0, job.row1), k)
ok = bool(np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(dst.cpu().numpy(), odst))
full_i, full_d = job.gather()
if rank == 0:
    wi, wd = c_oracle.get_reference_rows(X, bins, 0, n, k)
    ok = ok and bool(np.array_equal(full_i.cpu().numpy(), wi) and np.array_equal(full_d.cpu().numpy(), wd))
st = device.last_search_stats(local)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("symmetric sharded search on %d ranks, %d bins: %s; rank 0: K5 %.3f ms (first pass %.3f), K6 %.3f ms, tiles %d of %d plain, "
          "fallback rows %d" % (world, n, "identical to the oracle" if flag.item() == 1 else "MISMATCH", st["dist_topk_ms"],
                                st["dist_topk_first_pass_ms"], st["finalize_ms"], st["tiles"], st["tiles_plain"], st["exhaustive_rows"]))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
