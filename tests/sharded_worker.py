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
    # pipelined flavour: two batches in flight, exchange on a side stream, same ids
    Qd = torch.from_numpy(Q).cuda()
    Qd2 = torch.from_numpy(Q[::-1].copy()).cuda()
    tickets = [idx.search_dev_async(Qd, k), idx.search_dev_async(Qd2, k)]
    got = []
    for it in range(6):
        r, sc = idx.wait(tickets.pop(0))
        got.append((r.clone(), sc.clone()))
        tickets.append(idx.search_dev_async(Qd if it % 2 == 0 else Qd2, k))
    for t in tickets:
        r, sc = idx.wait(t)
        got.append((r.clone(), sc.clone()))
    torch.cuda.synchronize()
    for i, (r, sc) in enumerate(got):
        w = want_rows if i % 2 == 0 else want_rows[::-1]
        assert np.array_equal(r.cpu().numpy(), w), f"rank {rank}: pipelined ids differ (k={k}, batch {i})"
    # ... and with host buffers on both ends
    ta, tb = idx.search_async(Q, k), idx.search_async(Q[::-1].copy(), k)
    ra, sa_ = idx.wait_host(ta)
    rb, _ = idx.wait_host(tb)
    assert np.array_equal(ra, want_rows) and np.array_equal(rb, want_rows[::-1])
    np.testing.assert_allclose(sa_, want_scores, rtol=1e-5)
    idx.close()

# ---- hybrid over row-sharded postings: global statistics, global k nearest, two all-gathers ----
from oracle import bm25, fusion  # noqa: E402
Nh, V, Dh, Bh, k = 9000, 900, 256, 12, 10
indptr, doc, tf, doclen = synth.text_corpus(Nh, vocab=V, seed=17, median_len=50, max_len=200)
full = bm25.BM25Index(indptr, doc, tf, doclen)
Xh = synth.embeddings(Nh, Dh, 43)
Qh = synth.embeddings(Bh, Dh, 44)
qterms = synth.text_queries(Bh, vocab=V, seed=18)
lo, hi = shard_bounds(Nh, world, rank)
term_of = np.repeat(np.arange(V), np.diff(indptr))
mine = (doc >= lo) & (doc < hi)
l_indptr = np.zeros(V + 1, dtype=np.int64)
np.add.at(l_indptr, term_of[mine] + 1, 1)
l_indptr = np.cumsum(l_indptr)
idx = ShardedIndex(dim=Dh, capacity_rows=hi - lo)
idx.set_row_base(lo)
idx.append_dev(torch.from_numpy(Xh[lo:hi]).cuda())
idx.bm25_build(l_indptr, (doc[mine] - lo).astype(np.int32), tf[mine], doclen[lo:hi])
rows, scores = idx.search_hybrid_dev(torch.from_numpy(Qh).cuda(), qterms, 4.5, 2.0, k)
torch.cuda.synchronize()
rows, scores = rows.cpu().numpy(), scores.cpu().numpy()
knn_rows, _, knn_scores = knn.knn_exact(Xh, Qh, k)
for b in range(Bh):
    wr, ws = fusion.hybrid(full, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k)
    assert rows[b, :len(wr)].tolist() == wr.tolist(), f"rank {rank}: sharded hybrid ids differ (query {b})"
    np.testing.assert_allclose(scores[b, :len(wr)], ws, rtol=2e-6, atol=0)
rows_t, scores_t = idx.search_hybrid_dev(None, qterms, 4.5, 0.0, k)          # text only
torch.cuda.synchronize()
for b in range(Bh):
    tr, ts = bm25.topk(full.score(qterms[b], boost=4.5), k)
    assert rows_t[b].cpu().tolist() == tr.tolist() and scores_t[b].cpu().tolist() == ts.tolist()
idx.close()

# ---- the same hybrid on an L2 engine: the fused lists are raw scores (larger is better), whatever the vector metric --
import rassengine_b200 as rb  # noqa: E402
idx = ShardedIndex(dim=Dh, metric=rb.METRIC_L2, capacity_rows=hi - lo)
idx.set_row_base(lo)
idx.append_dev(torch.from_numpy(Xh[lo:hi]).cuda())
idx.bm25_build(l_indptr, (doc[mine] - lo).astype(np.int32), tf[mine], doclen[lo:hi])
rows, scores = idx.search_hybrid_dev(torch.from_numpy(Qh).cuda(), qterms, 4.5, 2.0, k)
torch.cuda.synchronize()
rows, scores = rows.cpu().numpy(), scores.cpu().numpy()
knn_rows, _, knn_scores = knn.knn_exact(Xh, Qh, k, metric=knn.L2)
for b in range(Bh):
    wr, ws = fusion.hybrid(full, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k)
    assert rows[b, :len(wr)].tolist() == wr.tolist(), f"rank {rank}: sharded L2 hybrid ids differ (query {b})"
    np.testing.assert_allclose(scores[b, :len(wr)], ws, rtol=2e-6, atol=0)
idx.close()
dist.barrier()
if rank == 0:
    print(f"sharded x{world}: merged top-k identical to the oracle (knn and hybrid)")
dist.destroy_process_group()
