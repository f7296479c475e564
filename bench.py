#!/usr/bin/env python
"""Headline benchmark: exact top-10 QPS over a 10M x 1024 corpus (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5]
                    [--rows R] [--batch B] [--k K]
    torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, NCCL)

--workload selects the BASELINE.json configuration: cfg2 (default, the one the metric is quoted on: 10M x 1024,
batch 64, top-10, HBM roofline), cfg3 (10M x 1024, batch 8192, top-100, tensor roofline) or cfg5 (100M x 1024
bf16 corpus row-sharded over >= 2 GPUs, batch 1024, top-10, tensor roofline).  The default run also reports the
batch-1, batch-1024 and batch-8192/top-100 regimes of the same corpus under `batch1` / `batched`.

One step = one query batch (default 64 queries) answered exactly against the whole corpus.  The corpus is
synthetic (seeded N(0,1) rows, L2-normalised as app/main.py:1250-1251 does), generated on the device and
ingested through the engine's append path before the timed region.  With N GPUs the SAME 10M-row corpus is
row-sharded (strong scaling): local exact top-k, one NCCL all-gather of k candidates per query, device merge.

Prints ONE JSON line (rank 0).  `value` = queries/s with queries resident in HBM; `e2e` = the same through the
host-buffer API (pinned host queries in, host results out, copies inside the timed region); `roofline` = the
scan kernel's algorithmic bytes / its CUDA-event time against the measured HBM copy peak; `cpu_baseline` = the
numpy port of the reference's exact CPU scan on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "exact top-10 QPS @10M x 1024 (cosine, ids identical to the fp64 CPU oracle)"
DIM = 1024
CHUNK = 500_000
SEED_CORPUS, SEED_QUERIES = 1234, 5678


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg5"])
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--cpu-sample-rows", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 / batched side measurements")
    a = ap.parse_args()
    rows, batch, k = {"cfg2": (10_000_000, 64, 10), "cfg3": (10_000_000, 8192, 100),
                      "cfg5": (100_000_000, 1024, 10)}[a.workload]
    a.rows = a.rows if a.rows is not None else rows
    a.batch = a.batch if a.batch is not None else batch
    a.k = a.k if a.k is not None else k
    a.bf16_only = a.workload == "cfg5"
    if a.cpu_sample_rows is None:
        a.cpu_sample_rows = 2_000_000 if a.batch <= 64 else 200_000
    return a


def workload_name(a):
    store = ("bf16 corpus (the bf16 values are the data)" if a.bf16_only
             else "synthetic fp32 unit rows (bf16 scan shadow + fp64 rerank)")
    return f"{a.workload}: exact cosine top-{a.k}, {a.rows} x {DIM} {store}, query batch {a.batch}"


def scan_kernel_name(B):
    if B <= 1:
        return "scan_stream_kernel (128-bit streaming GEMV + warp select)"
    if B <= 64:
        return "scan_umma_kernel (TMA + tcgen05, 64 queries/pass)"
    return "scan_gemm_kernel (TMA + tcgen05 cta_group::2, 256 queries/pass)"


def peaks():
    """(HBM GB/s, bf16 TFLOP/s sustained, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), \
            "measured (MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        return [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]

    def stop(self):
        if self.proc:
            self.proc.terminate()

    @staticmethod
    def summarise(lines):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the numpy port of the reference's exact CPU scan (oracle.knn.knn_fp32_baseline)
# ---------------------------------------------------------------------------------------------------
def cpu_scan_qps(a, sample_rows, batch, repeats=1, X=None):
    """Times fp32 sgemm + argpartition over `sample_rows` rows x `batch` queries and scales to a.rows."""
    from oracle import knn, synth
    if X is None:
        X = synth.embeddings(sample_rows, DIM, SEED_CORPUS)
    Q = synth.embeddings(batch, DIM, SEED_QUERIES)
    knn.knn_fp32_baseline(X[: min(20000, sample_rows)], Q, a.k)      # warm BLAS threads
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        knn.knn_fp32_baseline(X, Q, a.k)
        best = min(best, time.perf_counter() - t0)
    scale = a.rows / float(sample_rows)
    return batch / (best * scale), best


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = min(a.cpu_sample_rows, a.rows)
    from oracle import synth
    X = synth.embeddings(sample, DIM, SEED_CORPUS)
    for _ in range(max(0, min(a.warmup, 2))):
        cpu_scan_qps(a, sample, a.batch, X=X)
    times = []
    steps = max(1, min(a.steps, 10))
    for _ in range(steps):
        qps, t = cpu_scan_qps(a, sample, a.batch, X=X)
        times.append(t)
    t_step = float(np.mean(times)) * (a.rows / float(sample))
    qps = a.batch / t_step
    sample_desc = (f"numpy fp32 sgemm + argpartition over {sample} of {a.rows} rows x {a.batch} queries per step, "
                   f"time scaled x{a.rows / float(sample):g}; oracle port of the exact CPU scan "
                   "(OpenSearch/HNSW and FAISS cannot be installed offline)")
    out = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
           "steps": steps, "warmup": a.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(a)},
           "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample_desc},
           "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def fill_shard(index, lo, hi):
    """Device-side synthetic rows [lo, hi): chunk c of the global corpus is seeded SEED_CORPUS + c, so the data is
    the same for every GPU count."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    c0 = lo // CHUNK
    while c0 * CHUNK < hi:
        g = torch.Generator(device=dev).manual_seed(SEED_CORPUS + c0)
        x = torch.randn((CHUNK, DIM), generator=g, device=dev, dtype=torch.float32)
        x = x / (x.norm(dim=1, keepdim=True) + 1e-9)
        s = max(lo, c0 * CHUNK) - c0 * CHUNK
        e = min(hi, (c0 + 1) * CHUNK) - c0 * CHUNK
        part = x[s:e].contiguous()
        torch.cuda.synchronize()
        index.append_dev(part)
        del x, part
        c0 += 1


def run_ours(a):
    import torch
    import torch.distributed as dist
    import rassengine_b200 as rb
    from rassengine_b200.sharded import ShardedIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("RASS_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    lo, hi = shard_bounds(a.rows, world, rank)
    index = ShardedIndex(dim=DIM, capacity_rows=hi - lo, flags=rb.BF16_ONLY if a.bf16_only else 0)
    index.set_row_base(lo)
    fill_shard(index, lo, hi)
    eng = index.engine
    B, k = a.batch, a.k

    gq = torch.Generator(device=dev).manual_seed(SEED_QUERIES)
    n_batches = 8 if B <= 1024 else 2
    q_dev = [torch.randn((B, DIM), generator=gq, device=dev) for _ in range(n_batches)]
    q_host = [q.cpu().pin_memory() for q in q_dev]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity spot-check outside the timed region: fast path ids == fp64 full scan ids on this shard ----
    nb = min(B, 256)          # every query of a 64-query batch; a 256-query sample (one pass of the pair kernel) above
    rows_f = torch.empty((nb, k), dtype=torch.int64, device=dev)
    sc_f = torch.empty((nb, k), dtype=torch.float32, device=dev)
    eng.search_knn_dev(q_dev[0].data_ptr(), nb, k, rows_f.data_ptr(), sc_f.data_ptr())
    eng.set_path(rb.PATH_EXACT)
    rows_x = torch.empty_like(rows_f)
    sc_x = torch.empty_like(sc_f)
    eng.search_knn_dev(q_dev[0].data_ptr(), nb, k, rows_x.data_ptr(), sc_x.data_ptr())
    eng.set_path(rb.PATH_AUTO)
    parity_ok = bool((rows_f == rows_x).all())

    def timed(fn, steps, warmup, flush=None):
        for i in range(warmup):
            fn(i)
        if flush:
            flush()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        if flush:
            flush()          # the last batches in flight complete inside the timed region
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- device-resident arm ----
    stats = {"scan_ms": 0.0, "launches": 0, "fallback": 0, "certified": 0, "bytes": 0, "n": 0}
    index.merge_launches = 0

    def account():
        st = eng.last_stats
        stats["scan_ms"] += st["scan_ms"]
        stats["launches"] += st["launches"]
        stats["fallback"] += st["n_fallback"]
        stats["certified"] += st["n_certified"]
        stats["bytes"] += st["bytes_streamed"]
        stats["n"] += 1

    # the serving loop keeps two batches in flight: batch i+1 is enqueued before batch i is collected, so the
    # exchange (all-gather + merge, on a side stream) and the host work of a batch hide behind the next scan
    def pipelined(pick, kx):
        """(step, flush) closures of a loop that keeps two batches of pick(i) in flight."""
        tickets = []

        def step(i):
            tickets.append(index.search_dev_async(pick(i), kx))
            if len(tickets) == 2:
                index.wait(tickets.pop(0))
                account()

        def flush():
            while tickets:
                index.wait(tickets.pop(0))
                account()
        return step, flush

    step_dev, flush_dev = pipelined(lambda i: q_dev[i % n_batches], k)

    def reset():
        for kk in stats:
            stats[kk] = 0.0 if kk == "scan_ms" else 0
        index.merge_launches = 0

    for i in range(a.warmup):
        step_dev(i)
    flush_dev()
    reset()
    ms_dev, t0, t1 = timed(step_dev, a.steps, 0, flush=flush_dev)
    scan_ms_avg = stats["scan_ms"] / max(1, stats["n"])
    bytes_per_step = stats["bytes"] / max(1, stats["n"])
    launches = stats["launches"] + index.merge_launches
    fallback = stats["fallback"]
    clocks = ClockSampler.summarise(sampler.window(t0, t1)) if sampler else None

    # ---- end-to-end arm: pinned host queries in, host results out, every step (same two-in-flight loop) ----
    e2e_tickets = []

    def step_e2e(i):
        e2e_tickets.append(index.search_async(q_host[i % n_batches], k))
        if len(e2e_tickets) == 2:
            index.wait_host(e2e_tickets.pop(0))

    def flush_e2e():
        while e2e_tickets:
            index.wait_host(e2e_tickets.pop(0))

    ms_e2e, _, _ = timed(step_e2e, a.steps, a.warmup, flush=flush_e2e)

    # ---- batch-1 latency regime (same corpus), for the record ----
    b1 = None
    if B != 1 and a.workload == "cfg2" and not a.no_extras:
        q1 = [q[:1].contiguous() for q in q_dev]
        reset()

        # one blocking call per query: the latency regime (the streaming scan fills every SM, so a second batch in
        # flight would only delay the exchange of the first)
        def step_b1(i):
            index.search_dev(q1[i % n_batches], k)
            account()

        for i in range(a.warmup):
            step_b1(i)
        reset()
        ms_b1, _, _ = timed(step_b1, a.steps, 0)
        b1_scan = stats["scan_ms"] / max(1, stats["n"])
        b1 = {"qps": a.steps / (ms_b1 * 1e-3), "ms_per_query": ms_b1 / a.steps,
              "scan_gbs": (stats["bytes"] / max(1, stats["n"])) / (b1_scan * 1e-3) / 1e9 if b1_scan else None}
    # ---- large-batch regimes of the same corpus (tensor-core contraction), for the record ----
    batched = None
    if a.workload == "cfg2" and not a.no_extras:
        batched = {}
        for name, Bx, kx, nsteps in (("batch1024_top10", 1024, 10, 10), ("batch8192_top100", 8192, 100, 3)):
            qx = torch.randn((Bx, DIM), generator=gq, device=dev)
            reset()

            step_x, flush_x = pipelined(lambda i, qx=qx: qx, kx)
            for i in range(3):
                step_x(i)
            flush_x()
            reset()
            ms_x, _, _ = timed(step_x, nsteps, 0, flush=flush_x)
            scan_x = stats["scan_ms"] / max(1, stats["n"])
            flops = 2.0 * Bx * (hi - lo) * DIM
            batched[name] = {"qps": nsteps * Bx / (ms_x * 1e-3), "ms_per_batch": ms_x / nsteps, "k": kx,
                             "scan_kernel": scan_kernel_name(Bx), "scan_ms": scan_x,
                             "scan_tflops_per_gpu": flops / (scan_x * 1e-3) / 1e12 if scan_x else None,
                             "certificate_fallbacks": int(stats["fallback"])}
            del qx
    if sampler:
        sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_hbm, peak_tf, peak_src = peaks()
    tensor_bound = B > 64
    if tensor_bound:
        flops_per_step = 2.0 * B * (hi - lo) * DIM
        achieved = flops_per_step / (scan_ms_avg * 1e-3) / 1e12 if scan_ms_avg else 0.0
        peak = peak_tf
    else:
        achieved = bytes_per_step / (scan_ms_avg * 1e-3) / 1e9 if scan_ms_avg else 0.0
        peak = peak_hbm
    if batched:
        for v in batched.values():
            v["frac_of_tensor_peak"] = v["scan_tflops_per_gpu"] / peak_tf if v["scan_tflops_per_gpu"] else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            t = json.load(open(tp))
            traffic = (t.get("dram_bytes_per_row", 0) * (hi - lo) or None) if not tensor_bound else None
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": a.steps * B / (ms_dev * 1e-3), "unit": "queries/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev / a.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16 scan (fp32 accumulate) + fp64 rerank",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "rows": a.rows, "dim": DIM, "batch": B, "k": k,
                   "sharding": f"row-sharded x{world}, one NCCL all-gather of k candidates per query" if world > 1
                   else "single shard",
                   "l2": "inputs larger than L2: every step streams the whole bf16 shard "
                         f"({(hi - lo) * DIM * 2 / 1e9:.2f} GB per GPU)",
                   "scan_kernel": scan_kernel_name(B),
                   "loop": "two batches in flight (search_dev_async / wait): the exchange and host work of batch i "
                           "overlap the scan of batch i+1; the e2e arm is the same loop with pinned host queries in "
                           "and host results out"},
        "e2e": {"value": a.steps * B / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 4,
                "d2h_bytes_per_step": B * k * 12, "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(launches),
        "roofline": ({"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                      "frac": achieved / peak if peak else None, "traffic": None, "peak_source": peak_src,
                      "kernel": "scan_gemm_kernel", "algorithmic_flops_per_launch": 2.0 * B * (hi - lo) * DIM,
                      "kernel_ms": scan_ms_avg, "frac_of_nominal_2250TF": achieved / 2250.0} if tensor_bound else
                     {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                      "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                      "kernel": "scan_umma_kernel" if B > 1 else "scan_stream_kernel",
                      "algorithmic_bytes_per_launch": bytes_per_step, "kernel_ms": scan_ms_avg,
                      "frac_of_nominal_8TBs": achieved / 8000.0}),
        "parity": {"fast_path_ids_equal_fp64_scan": parity_ok, "queries_checked": int(nb),
                   "certificate_fallbacks_in_timed_region": int(fallback)},
        "batch1": b1,
        "batched": batched,
        "clocks": clocks,
    }
    if not a.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        sample = min(a.cpu_sample_rows, a.rows)
        Xs = eng.read_rows(0, sample)            # the same rows the GPU scanned
        qps, t = cpu_scan_qps(a, sample, B, repeats=5, X=Xs)
        out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"numpy fp32 sgemm + argpartition (oracle.knn.knn_fp32_baseline) over the "
                                         f"first {sample} of {a.rows} rows x {B} queries ({t:.2f} s), scaled "
                                         f"x{a.rows / float(sample):g}"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
