"""Ingest throughput of the vector store: rows/s through rass_append from pageable and pinned host
memory and from device memory, and through the two client paths (reference-shaped python lists vs numpy rows).

    python tools/bench_ingest.py [N_ROWS]      -> one JSON line
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
from rassengine_b200 import indexer as ix  # noqa: E402
from rassengine_b200.client import B200Client  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
D = 1024
rng = np.random.default_rng(1)
X = rng.standard_normal((N, D), dtype=np.float32)
out = {"rows": N, "dim": D, "bytes_per_row": D * 4}


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for name, src in (("pageable", X), ("pinned", torch.from_numpy(X).pin_memory().numpy())):
    with rb.Engine(dim=D, capacity_rows=N) as e:
        e.append(src[:1000])                                   # warm the staging buffers
        t = timed(lambda: e.append(src))
        out[f"append_{name}_rows_per_s"] = N / t
        out[f"append_{name}_GBps"] = N * D * 4 / t / 1e9
xd = torch.from_numpy(X[: N // 2]).cuda()
with rb.Engine(dim=D, capacity_rows=N) as e:
    e.append_dev(xd.data_ptr(), 1000)
    t = timed(lambda: e.append_dev(xd.data_ptr(), xd.shape[0]))
    out["append_dev_rows_per_s"] = xd.shape[0] / t
    out["append_dev_GBps_read"] = xd.shape[0] * D * 4 / t / 1e9
del xd

# client paths on a smaller slice (the python-list path is slow by construction: 1024 float objects per row)
M = min(N, 20000)
docs = [{"doc_id": f"d{i}", "doc_type": "unstructured", "patientId": f"p{i % 100}", "unstructuredText": "note text"}
        for i in range(M)]
for name, kw in (("lists_flush64", dict(as_lists=True)), ("numpy_flush4096", dict(as_lists=False, flush=4096))):
    c = B200Client()
    idx = ix.get_index_name("ingest")
    ix.ensure_index_exists(c, idx, ix.index_body(D))
    t = timed(lambda: ix.store_chunks(c, idx, docs, X[:M], **kw))
    out[f"store_chunks_{name}_rows_per_s"] = M / t
    c.close()
print(json.dumps(out), flush=True)
