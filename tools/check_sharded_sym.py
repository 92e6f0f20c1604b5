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
binsize = int(sys.argv[4]) if len(sys.argv) > 4 else 250000
bins = [int(b) for b in np.maximum(1, np.array(synth.chrom_bins(binsize)) // div)]
X = synth.corrected_like(bins, S, seed=9)
n = X.shape[0]
job = shard.SymmetricShardedSearch(n, k, rank, world, dev)
x = torch.as_tensor(X, device=dev)
idx, dst = job.run(x, bins)
r0, r1 = job.row0, max(job.row0, job.row1)
sampled = n * S > 4e6            # large shapes: every rank checks sampled ranges of its own rows against the oracle
if sampled:
    ok, rows_checked = True, 0
    hi, hd = idx.cpu().numpy(), dst.cpu().numpy()
    for a in sorted({r0, max(r0, (r0 + r1) // 2 - 16), max(r0, r1 - 32)}):
        b = min(r1, a + 32)
        if b <= a:
            continue
        oi, od = c_oracle.get_reference_rows(X, bins, a, b, k)
        ok = ok and bool(np.array_equal(hi[a - r0:b - r0], oi) and np.array_equal(hd[a - r0:b - r0], od))
        rows_checked += b - a
else:
    oidx, odst = c_oracle.get_reference_rows(X, bins, r0, r1, k)
    ok = bool(np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(dst.cpu().numpy(), odst))
    rows_checked = r1 - r0
full_i, full_d = job.gather()
if rank == 0:
    if sampled:          # the gathered table against the single-GPU search (itself oracle-checked in the GPU tests)
        wi, wd = device.newref_topk(x, bins, 0, n, k)
        ok = ok and bool(torch.equal(full_i, wi) and torch.equal(full_d, wd))
    else:
        wi, wd = c_oracle.get_reference_rows(X, bins, 0, n, k)
        ok = ok and bool(np.array_equal(full_i.cpu().numpy(), wi) and np.array_equal(full_d.cpu().numpy(), wd))
st = device.last_search_stats(local)
flag = torch.tensor([1 if ok else 0], device=dev)
rows_t = torch.tensor([rows_checked], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.all_reduce(rows_t, op=dist.ReduceOp.SUM)
if rank == 0:
    print("symmetric sharded search on %d ranks, %d bins x %d samples, refsize %d: %s (%d rows against the C oracle%s); rank 0: K5 %.3f ms "
          "(first pass %.3f), K6 %.3f ms, tiles %d of %d plain, fallback rows %d" %
          (world, n, S, k, "identical to the oracle" if flag.item() == 1 else "MISMATCH", int(rows_t.item()),
           ", whole gathered table against the single-GPU search" if sampled else ", whole gathered table too",
           st["dist_topk_ms"], st["dist_topk_first_pass_ms"], st["finalize_ms"], st["tiles"], st["tiles_plain"], st["exhaustive_rows"]))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
