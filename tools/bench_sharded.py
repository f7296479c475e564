#!/usr/bin/env python
"""Single-process multi-GPU throughput: ONE sharded engine handle (rass_create_sharded) over N GPUs of the box, the way
the reference's single uvicorn process would hold it (app/main.py:337-343, :357) -- no torchrun, no NCCL: every shard's
finish kernel stores its top-k list into the coordinator GPU's gather buffer through peer-mapped memory, one merge kernel
per batch.  Same corpus, queries and two-batches-in-flight loop as bench.py's cfg2 (whose driver contract launches one
process per GPU instead).

    python tools/bench_sharded.py [--gpus N] [--rows R] [--batch B] [--k K] [--steps S] [--warmup W]   -> one JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import rassengine_b200 as rb
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--overlap", type=int, default=1, help="1: every shard runs its two async slots on their own streams")
    ap.add_argument("--check", type=int, default=8, help="queries compared with the fp64 exact scan of every shard")
    a = ap.parse_args()
    D, CH = 1024, 500_000
    devices = list(range(a.gpus))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    e = rb.Engine(dim=D, devices=devices, capacity_rows=a.rows)
    t0 = time.time()
    for c0 in range(0, a.rows, CH):
        m = min(CH, a.rows - c0)
        g = torch.Generator(device=dev).manual_seed(1234 + c0 // CH)
        x = torch.randn((CH, D), generator=g, device=dev)[:m]
        x = (x / (x.norm(dim=1, keepdim=True) + 1e-9)).contiguous()
        torch.cuda.synchronize()
        e.append_dev(x.data_ptr(), m)
        del x
    fill_s = time.time() - t0
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.set_async_overlap(bool(a.overlap))
    gq = torch.Generator(device=dev).manual_seed(5678)
    B, k = a.batch, a.k
    qs = [torch.randn((B, D), generator=gq, device=dev) for _ in range(8)]
    rows = [torch.empty((B, k), dtype=torch.int64, device=dev) for _ in range(2)]
    scores = [torch.empty((B, k), dtype=torch.float32, device=dev) for _ in range(2)]
    acc = {"scan_ms": 0.0, "n": 0, "again": 0}

    def run(n):
        pending = []
        for i in range(n):
            s = i & 1
            if len(pending) == 2:
                sl = pending.pop(0)
                final, st = e.search_knn_dev_wait(sl)
                acc["scan_ms"] += st["scan_ms"]
                acc["n"] += 1
                acc["again"] += 0 if final else 1
            e.search_knn_dev_async(qs[i % 8].data_ptr(), B, k, rows[s].data_ptr(), scores[s].data_ptr(), 0, s, 0)
            pending.append(s)
        for sl in pending:
            final, st = e.search_knn_dev_wait(sl)
            acc["scan_ms"] += st["scan_ms"]
            acc["n"] += 1

    run(a.warmup)
    acc.update(scan_ms=0.0, n=0, again=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(a.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # parity of the merged list: the fast path against the fp64 exact scan (RASS_PATH_EXACT) through the same handle
    n_chk = min(a.check, B)
    r_fast, s_fast = e.search_knn(qs[0][:n_chk].cpu().numpy(), k)
    e.set_path(rb.PATH_EXACT)
    r_exact, _ = e.search_knn(qs[0][:n_chk].cpu().numpy(), k)
    e.set_path(rb.PATH_AUTO)
    # ... and against the fp64 CPU oracle over all rows read back through the sharded handle (test infrastructure)
    import bench
    from oracle import knn as oknn
    qc = qs[0][:n_chk].cpu().numpy()
    o_rows, o_cos = bench.cpu_oracle_topk(e, 0, a.rows, qc, k, max(1, min(6, (os.cpu_count() or 4) - 1)))
    # the async (peer-store) path's merged list for the same queries
    e.search_knn_dev_async(qs[0].data_ptr(), B, k, rows[0].data_ptr(), scores[0].data_ptr(), 0, 0, 0)
    e.search_knn_dev_wait(0)
    r_async = rows[0][:n_chk].cpu().numpy()
    # host-buffer call (what the drop-in client issues), blocking
    qh = qs[0].cpu().numpy()
    for _ in range(3):
        e.search_knn(qh, k)
    t0 = time.perf_counter()
    n_host = max(10, a.steps // 4)
    for _ in range(n_host):
        e.search_knn(qh, k)
    t_host = (time.perf_counter() - t0) / n_host
    out = {"workload": f"single-process sharded handle: exact cosine top-{k}, {a.rows} x {D}, batch {B}, {a.gpus} GPUs",
           "n_gpus": a.gpus, "qps": a.steps * B / (ms * 1e-3), "ms_per_step": ms / a.steps,
           "scan_ms_max_over_shards": acc["scan_ms"] / max(1, acc["n"]),
           "scan_gbs_per_gpu": (a.rows / a.gpus) * D * 2 / (acc["scan_ms"] / max(1, acc["n"]) * 1e-3) / 1e9,
           "blocking_host_call_ms": t_host * 1e3, "blocking_host_call_qps": B / t_host,
           "batches_repeated": acc["again"], "ids_equal_fp64_scan": bool(np.array_equal(r_fast, r_exact)),
           "ids_equal_cpu_oracle": bool(np.array_equal(r_fast, o_rows) and np.array_equal(r_async, o_rows)),
           "queries_checked_cpu": int(n_chk),
           "max_score_rel_err": float(np.max(np.abs(s_fast - oknn.score_from_cos(o_cos)) / oknn.score_from_cos(o_cos))),
           "fill_s": round(fill_s, 1), "async_overlap": bool(a.overlap), "exchange": "peer stores into the coordinator's gather buffer + one merge kernel"}
    print(json.dumps(out), flush=True)
    e.close()


if __name__ == "__main__":
    main()
