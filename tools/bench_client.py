"""Latency of the drop-in surface itself (B200Indexer over B200Client) on a mid-size index: what one /ask request pays
on top of the device time.    python tools/bench_client.py [N_DOCS]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rassengine_b200 import indexer as ix  # noqa: E402
from rassengine_b200.client import B200Client  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
D = 1024
rng = np.random.default_rng(3)
words = [f"w{i:04d}" for i in range(3000)]
client = B200Client()
name = ix.get_index_name("lat")
ix.ensure_index_exists(client, name, ix.index_body(D))
t0 = time.perf_counter()
CH = 50_000
for c0 in range(0, N, CH):
    m = min(CH, N - c0)
    emb = rng.standard_normal((m, D), dtype=np.float32)
    tok = rng.integers(0, len(words), size=(m, 24))
    docs = [{"doc_id": f"d{c0 + i}", "doc_type": "unstructured", "patientId": f"p{(c0 + i) % 500}",
             "unstructuredText": " ".join(words[j] for j in tok[i])} for i in range(m)]
    ix.store_chunks(client, name, docs, emb, as_lists=False, flush=CH)
t_ingest = time.perf_counter() - t0
idxr = ix.B200Indexer(client, name)
q = rng.standard_normal((1, D)).astype(np.float32)
out = {"docs": N, "ingest_docs_per_s": N / t_ingest}
# refresh: the first hybrid query after the last bulk makes the new rows' text searchable (device-side segments + one
# commit, rassengine_b200/csrc/postings.cu); then one more bulk and the query after it
t0 = time.perf_counter()
idxr.hybrid_search("w0001 w0500 w2999 w1234", q, k=3)
out["first_hybrid_after_ingest_ms"] = (time.perf_counter() - t0) * 1e3
m = 10_000
emb = rng.standard_normal((m, D), dtype=np.float32)
tok = rng.integers(0, len(words), size=(m, 24))
docs = [{"doc_id": f"x{i}", "doc_type": "unstructured", "patientId": f"p{i % 500}",
         "unstructuredText": " ".join(words[j] for j in tok[i])} for i in range(m)]
ix.store_chunks(client, name, docs, emb, as_lists=False, flush=m)
t0 = time.perf_counter()
idxr.hybrid_search("w0001 w0500 w2999 w1234", q, k=3)
out["first_hybrid_after_a_10k_bulk_ms"] = (time.perf_counter() - t0) * 1e3
ti = client._indices[name].text
out["text_syncs"] = {"device_commits": ti.device_commits, "host_rebuilds": ti.host_rebuilds}


def lat(fn, n=100):
    for _ in range(5):
        fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e3


eng = client._indices[name].engine
out["engine_search_knn_k3_ms"] = lat(lambda: eng.search_knn(q, 3))
out["semantic_search_k3_ms"] = lat(lambda: idxr.semantic_search(q, k=3))
out["semantic_search_k10_ms"] = lat(lambda: idxr.semantic_search(q, k=10))
out["semantic_search_k10_patient_ms"] = lat(lambda: idxr.semantic_search(q, k=10, patient_id="p7"))
out["hybrid_search_k3_ms"] = lat(lambda: idxr.hybrid_search("w0001 w0500 w2999 w1234", q, k=3), 50)
out["hybrid_search_k10_patient_ms"] = lat(lambda: idxr.hybrid_search("w0001 w0500 w2999 w1234", q, k=10, patient_id="p7"), 50)
out["explanatory_search_k3_ms"] = lat(lambda: idxr.explanatory_search("w0001 w0500", k=3), 50)
print(json.dumps(out), flush=True)
client.close()
