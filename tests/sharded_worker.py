"""torchrun worker: row-sharded search over NCCL equals the oracle (and therefore a single-shard search).
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/sharded_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import knn, synth  # noqa: E402
from rassengine_b200.sharded import ShardedIndex, shard_bounds  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
N, D, B = 60000, 1024, 70
X = synth.embeddings(N, D, 41)
synth.plant_duplicates(X, 8)
Q = synth.clustered_queries(X, B, seed=42)
for k in (10, 100):
    want_rows, _, want_scores = knn.knn_exact(X, Q, k)
    lo, hi = shard_bounds(N, world, rank)
    idx = ShardedIndex(dim=D, capacity_rows=hi - lo)
    idx.set_row_base(lo)
    idx.append_dev(torch.from_numpy(X[lo:hi]).cuda())
    rows, scores = idx.search_dev(torch.from_numpy(Q).cuda(), k)
    torch.cuda.synchronize()
    assert np.array_equal(rows.cpu().numpy(), want_rows), f"rank {rank}: merged ids differ from the oracle (k={k})"
    np.testing.assert_allclose(scores.cpu().numpy(), want_scores, rtol=1e-5)
    r2, s2 = idx.search(Q, k)                       # host-buffer flavour
    assert np.array_equal(r2, want_rows)
    idx.close()
dist.barrier()
if rank == 0:
    print(f"sharded x{world}: merged top-k identical to the oracle")
dist.destroy_process_group()
