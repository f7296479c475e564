"""BASELINE.json configs[3]: hybrid BM25 multi_match + vector fusion over N synthetic FHIR-narrative-shaped chunks
(~30k-term Zipf vocabulary), top-10.  Builds the corpus on the device, times rass_search_hybrid per query (the
reference issues one hybrid query per /ask request) and checks ids/scores against the oracle at full size -- which is
why it lives under tests/: only tests may execute oracle/ (it is a script, not collected by pytest).

    python tests/cfg4_hybrid_bench.py [N_DOCS] [N_QUERIES]      -> one JSON line
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rassengine_b200 as rb  # noqa: E402
from oracle import bm25, fusion, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 64
V, D, K = 30000, 1024, 10
dev = torch.device("cuda", 0)
t_start = time.time()

# ---- text corpus on the device: Zipf(1.07) term ids, clipped-lognormal lengths (median 120, max 512) ----------
g = torch.Generator(device=dev).manual_seed(4242)
p = 1.0 / torch.arange(1, V + 1, device=dev, dtype=torch.float64) ** 1.07
p = (p / p.sum()).float()
doclen = torch.empty(N, dtype=torch.int32, device=dev)
keys = []
CH = 250_000
for c0 in range(0, N, CH):
    m = min(CH, N - c0)
    ln = torch.exp(torch.randn(m, generator=g, device=dev) * 0.6 + np.log(120.0)).round().clamp_(1, 512).to(torch.int64)
    doclen[c0:c0 + m] = ln.to(torch.int32)
    tot = int(ln.sum())
    terms = torch.multinomial(p, tot, replacement=True, generator=g)
    docs = torch.repeat_interleave(torch.arange(c0, c0 + m, device=dev), ln)
    k_, cnt = torch.unique(terms * N + docs, return_counts=True)
    keys.append((k_, cnt.clamp_(max=65535).to(torch.int16)))
    del terms, docs
key = torch.cat([k for k, _ in keys])
tfv = torch.cat([c for _, c in keys])
del keys
order = torch.argsort(key)           # term-major, doc ascending within a term
key, tfv = key[order], tfv[order]
del order
term_of = key // N
doc_of = (key - term_of * N).to(torch.int32)
indptr = torch.zeros(V + 1, dtype=torch.int64, device=dev)
indptr[1:] = torch.cumsum(torch.bincount(term_of, minlength=V), 0)
nnz = int(key.numel())
del key, term_of
h_indptr, h_doc = indptr.cpu().numpy(), doc_of.cpu().numpy()
h_tf, h_len = tfv.cpu().numpy().view(np.uint16), doclen.cpu().numpy().astype(np.uint32)
del doc_of, tfv, indptr, doclen
torch.cuda.empty_cache()
t_text = time.time() - t_start

# ---- embeddings ----
e = rb.Engine(dim=D, capacity_rows=N)
for c0 in range(0, N, 500_000):
    m = min(500_000, N - c0)
    gg = torch.Generator(device=dev).manual_seed(1234 + c0 // 500_000)
    x = torch.randn((m, D), generator=gg, device=dev)
    x /= x.norm(dim=1, keepdim=True) + 1e-9
    torch.cuda.synchronize()
    e.append_dev(x.data_ptr(), m)
    del x
t0 = time.time()
e.bm25_build(h_indptr, h_doc, h_tf, h_len)
t_build = time.time() - t0

qterms = synth.text_queries(NQ, vocab=V, seed=4243)
Q = synth.embeddings(NQ, D, 5678)

# ---- timing: one hybrid query per call (what /ask does), then one batched call ----
for b in range(3):
    e.search_hybrid(Q[b:b + 1], [qterms[b]], 4.5, 2.0, K)
t0 = time.perf_counter()
res = []
posting_bytes = 0
for b in range(NQ):
    res.append(e.search_hybrid(Q[b:b + 1], [qterms[b]], 4.5, 2.0, K))
    posting_bytes += e.last_stats["bytes_streamed"] - N * D * 2
t_single = time.perf_counter() - t0
e.search_hybrid(Q, qterms, 4.5, 2.0, K)          # sizes the batch workspaces
t0 = time.perf_counter()
for _ in range(3):
    rows_b, scores_b = e.search_hybrid(Q, qterms, 4.5, 2.0, K)
t_batch = (time.perf_counter() - t0) / 3
st = e.last_stats
t0 = time.perf_counter()
for b in range(NQ):
    e.search_hybrid(None, [qterms[b]], 4.5, 2.0, K)
t_text_only = time.perf_counter() - t0

# ---- parity at full size: oracle BM25 + oracle fusion on the host over the same CSR; the kNN clause of the oracle
# fusion takes the device's fp64 exact scan (PATH_EXACT) as its k nearest, so every arithmetic step of the text and
# fusion kernels is checked against numpy at 5M docs ----
idx = bm25.BM25Index(h_indptr, h_doc, h_tf, h_len)
e.set_path(rb.PATH_EXACT)
n_chk = min(NQ, 16)
knn_rows, knn_scores = e.search_knn(Q[:n_chk], K)
e.set_path(rb.PATH_AUTO)
ok_ids = ok_scores = ok_text = 0
for b in range(n_chk):
    wr, ws = fusion.hybrid(idx, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, K)
    ok_ids += int(res[b][0][0, :len(wr)].tolist() == wr.tolist() and rows_b[b, :len(wr)].tolist() == wr.tolist())
    ok_scores += int(np.allclose(res[b][1][0, :len(wr)], ws, rtol=2e-6, atol=0))
    tr, ts = bm25.topk(idx.score(qterms[b], boost=4.5), K)
    r_t, s_t = e.search_hybrid(None, [qterms[b]], 4.5, 2.0, K)
    ok_text += int(r_t[0].tolist() == tr.tolist() and s_t[0].tolist() == ts.tolist())
sum_df = float(np.mean([sum(int(h_indptr[t + 1] - h_indptr[t]) for t in q) for q in qterms]))
out = {"workload": f"cfg4: hybrid BM25 multi_match + knn fusion, {N} chunks, vocab {V}, top-{K}", "n_docs": N, "nnz": nnz,
       "queries": NQ, "mean_postings_per_query": sum_df,
       "qps_one_query_per_call": NQ / t_single, "ms_per_query": 1e3 * t_single / NQ,
       "qps_batched_call": NQ / t_batch, "qps_text_only": NQ / t_text_only,
       "text_postings_per_s": sum_df * NQ / t_text_only,
       "parity": {"checked_queries": n_chk, "fused_ids_equal_oracle": ok_ids, "fused_scores_equal_oracle": ok_scores,
                  "text_only_bit_identical": ok_text},
       "last_stats": st, "setup_s": {"text_corpus": round(t_text, 1), "bm25_build": round(t_build, 1)}}
print(json.dumps(out), flush=True)
