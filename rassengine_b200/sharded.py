"""Row-sharded corpus across the GPUs of one NVSwitch box: one process per GPU, local exact top-k, ONE NCCL
all-gather of k candidates per query, on-device merge.

This is the B200 form of OpenSearch's ``number_of_shards`` + coordinator merge (reference app/main.py:357;
SURVEY.md 8e): rank g of G owns rows [g * ceil(N / G), ...), searches them exactly, and contributes its [B, k]
(fp64 key, int64 global row) lists.  The exact rerank happens before the gather, so only exact keys travel and
the merged top-k is the top-k of the union, bit-identical to a single-shard search.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist

from . import _capi as capi


def shard_bounds(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


class _DeviceOps:
    """Tensor-level view of the raw-pointer engine calls (the seam the gloo tests replace with a CPU double)."""

    def __init__(self, engine, main_stream: int = 0):
        self.engine = engine
        self.main_stream = main_stream     # the stream the engine normally runs on (torch's current stream)

    def search(self, q, k, packed, scores):
        """packed: int64 [2, B, k] -- plane 0 receives the fp64 key bits, plane 1 the global rows."""
        self.engine.search_knn_dev(q.data_ptr(), q.shape[0], k, packed[1].data_ptr(), scores.data_ptr(),
                                   packed[0].data_ptr())
        return self.engine.last_stats

    def merge(self, gathered, world, B, k, out_rows, out_scores, out_keys=None, stride=None, stream=None, raw=False):
        """gathered: every rank's packed buffer in rank order ([2, B, k] int64 each, `stride` elements apart).
        raw: the keys are fused hybrid scores (larger is better whatever the vector metric), not cos / d^2."""
        plane = B * k * 8
        if stream is not None:
            self.engine.set_stream(stream)
        try:
            fn = self.engine.merge_scores_dev if raw else self.engine.merge_topk_dev
            fn(gathered.data_ptr(), gathered.data_ptr() + plane, world, B, k,
                                       out_rows.data_ptr(), out_scores.data_ptr(),
                                       out_keys.data_ptr() if out_keys is not None else 0,
                                       shard_stride=stride or 2 * B * k)
        finally:
            if stream is not None:
                self.engine.set_stream(self.main_stream)

    def search_async(self, q, k, packed, scores, slot, flag):
        self.engine.search_knn_dev_async(q.data_ptr(), q.shape[0], k, packed[1].data_ptr(), scores.data_ptr(),
                                         packed[0].data_ptr(), slot=slot, flag_ptr=flag.data_ptr())

    def search_wait(self, slot):
        return self.engine.search_knn_dev_wait(slot)

    def join(self, slot, stream):
        """`stream` waits (on the device) for the search enqueued in `slot` (RASS_OPT_ASYNC_OVERLAP)."""
        self.engine.async_join(slot, stream)

    def bm25_build(self, indptr, doc, tf, doclen, doc_count, sum_ttf, df):
        self.engine.bm25_build(indptr, doc, tf, doclen, global_doc_count=doc_count, global_sum_ttf=sum_ttf,
                               global_df=df)

    def fuse(self, B, qterms, w_text, knn_rows, knn_scores, w_knn, k, packed, scores, qweights=None, qflags=None):
        """packed: int64 [2, B, k] -- plane 0 receives the fused scores as fp64 bits, plane 1 the global rows."""
        self.engine.fuse_hybrid_dev(B, qterms, w_text, knn_rows.data_ptr() if knn_rows is not None else 0,
                                    knn_scores.data_ptr() if knn_scores is not None else 0, w_knn, k,
                                    packed[1].data_ptr(), scores.data_ptr(), packed[0].data_ptr(),
                                    qweights=qweights, qflags=qflags)


class ShardedIndex:
    def __init__(self, dim: int = 1024, metric: int = capi.METRIC_COSINE, flags: int = 0, capacity_rows: int = 0,
                 device: int | None = None, group=None, engine=None, ops=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dim = dim
        if engine is None and ops is None:
            from .engine import Engine
            if device is None:
                device = torch.cuda.current_device()
            engine = Engine(dim=dim, metric=metric, device=device, capacity_rows=capacity_rows, flags=flags)
            # run on torch's current stream so NCCL and the engine's kernels are ordered by the stream
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
            # pipelined searches: each of the two batches in flight on its own stream and workspace
            # (RASS_B200_NO_OVERLAP=1: both on the main stream, back to back -- the A/B switch of the measurement)
            self._overlap = os.environ.get("RASS_B200_NO_OVERLAP") != "1"
            engine.set_async_overlap(self._overlap)
            # the all-gather and the merge of batch i need an SM of their own while batch i+1 is being scanned (measured
            # at 2 x 1.25M rows, batch 64: 0.461 ms per step with none, 0.441 with 1 or 2, 0.448 with 8 reserved)
            reserve = os.environ.get("RASS_B200_SCAN_RESERVE")
            engine.set_scan_reserve_sms(int(reserve) if reserve is not None else (2 if self.world > 1 else 0))
        self.engine = engine
        self.ops = ops if ops is not None else _DeviceOps(
            engine, torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0)
        self.merge_launches = 0
        self._bufs: dict = {}          # (B, k, device) -> work tensors, reused across calls (no allocator traffic)
        self._pending = [None, None]   # search_dev_async tickets
        self._next_slot = 0
        self._side = None              # side stream of the pipelined exchange
        self._overlap = getattr(self, "_overlap", False)

    def _buffers(self, B: int, k: int, dev):
        key = (B, k, str(dev))
        b = self._bufs.get(key)
        if b is None:
            if len(self._bufs) > 8:
                self._bufs.clear()
            b = {"packed": torch.empty((2, B, k), dtype=torch.int64, device=dev),
                 "scores": torch.empty((B, k), dtype=torch.float32, device=dev)}
            if self.world > 1:
                b["gathered"] = torch.empty((self.world * 2, B, k), dtype=torch.int64, device=dev)
                b["out_rows"] = torch.empty((B, k), dtype=torch.int64, device=dev)
                b["out_scores"] = torch.empty((B, k), dtype=torch.float32, device=dev)
            self._bufs[key] = b
        return b

    def set_row_base(self, base: int):
        self.engine.set_row_base(base)

    # -- pipelined search: two batches in flight, the exchange of batch i overlaps the scan of batch i+1 ------------
    def search_dev_async(self, q: torch.Tensor, k: int, to_host: bool = False) -> int:
        """Enqueues the whole sharded search of one batch and returns a ticket; `wait(ticket)` hands out the result.
        At most two tickets may be outstanding.  The local search runs on the current stream; the all-gather and the
        merge run on a side stream behind an event, so the next batch's scan does not wait for them.  Every rank
        appends the number of its uncertified queries to its candidates, so after the gather all ranks know -- without
        another collective -- whether the batch has to be repeated through the blocking path."""
        slot = self._next_slot
        if self._pending[slot] is not None:
            raise RuntimeError("two batches are already in flight: wait() for the older ticket first")
        self._next_slot ^= 1
        B, dev = q.shape[0], q.device
        key = ("async", slot, B, k, str(dev))
        b = self._bufs.get(key)
        if b is None:
            n = 2 * B * k + 1                                     # candidates + the uncertified-queries count
            b = {"flat": torch.zeros(n, dtype=torch.int64, device=dev),
                 "scores": torch.empty((B, k), dtype=torch.float32, device=dev),
                 "gathered": torch.zeros(self.world * n, dtype=torch.int64, device=dev),
                 "out_rows": torch.empty((B, k), dtype=torch.int64, device=dev),
                 "out_scores": torch.empty((B, k), dtype=torch.float32, device=dev),
                 "flags_host": (torch.zeros(self.world, dtype=torch.int64).pin_memory() if dev.type == "cuda"
                                else torch.zeros(self.world, dtype=torch.int64)),
                 "ev_main": torch.cuda.Event() if dev.type == "cuda" else None,
                 "ev_side": torch.cuda.Event() if dev.type == "cuda" else None}
            b["packed"] = b["flat"][: 2 * B * k].view(2, B, k)
            self._bufs[key] = b
        if to_host and "rows_host" not in b:
            pin = dev.type == "cuda"
            b["rows_host"] = torch.empty((B, k), dtype=torch.int64, pin_memory=pin)
            b["scores_host"] = torch.empty((B, k), dtype=torch.float32, pin_memory=pin)
            b["ev_out"] = torch.cuda.Event() if pin else None
        self.ops.search_async(q, k, b["packed"], b["scores"], slot, b["flat"][-1:])
        cuda = dev.type == "cuda"
        if cuda and self._side is None:
            self._side = torch.cuda.Stream(device=dev)

        def side_after_search():
            # the side stream picks up behind this batch's search: behind the slot's own stream when the slots overlap
            # (the main stream must not wait for it: the next batch forks from there), else behind the main stream
            if self._overlap:
                self.ops.join(slot, self._side.cuda_stream)
            else:
                b["ev_main"].record(torch.cuda.current_stream())
                self._side.wait_event(b["ev_main"])

        if self.world == 1 and to_host:
            if cuda:
                side_after_search()
            with (torch.cuda.stream(self._side) if cuda else contextlib.nullcontext()):
                b["rows_host"].copy_(b["packed"][1], non_blocking=True)    # behind the search, beside the next batch
                b["scores_host"].copy_(b["scores"], non_blocking=True)
                if b["ev_out"] is not None:
                    b["ev_out"].record(self._side)
        if self.world > 1:
            if cuda:
                side_after_search()
                ctx = torch.cuda.stream(self._side)
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                dist.all_gather_into_tensor(b["gathered"], b["flat"], group=self.group)
                self.ops.merge(b["gathered"], self.world, B, k, b["out_rows"], b["out_scores"], stride=2 * B * k + 1,
                               stream=self._side.cuda_stream if cuda else None)
                # every rank's uncertified-queries count to pinned host memory, still on the side stream: reading it
                # in wait() must not touch the main stream, where the next batch's scan is already queued
                n = b["flat"].numel()
                b["flags_host"].copy_(b["gathered"][n - 1::n], non_blocking=True)
                if to_host:
                    b["rows_host"].copy_(b["out_rows"], non_blocking=True)
                    b["scores_host"].copy_(b["out_scores"], non_blocking=True)
                if cuda:
                    b["ev_side"].record(self._side)
            self.merge_launches += 1
        self._pending[slot] = (q, k, b)
        return slot

    def search_async(self, q_host, k: int) -> int:
        """Host-buffer flavour of search_dev_async: pinned (or pageable) [B, dim] fp32 queries in, results staged to
        pinned host memory behind the search; collect with wait_host(ticket)."""
        q = torch.as_tensor(q_host, dtype=torch.float32)
        qd = q.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
        return self.search_dev_async(qd, k, to_host=True)

    def wait_host(self, ticket: int):
        """-> (rows, scores) of that batch as host numpy arrays."""
        rows, scores, final, b = self._wait(ticket)
        if final and "rows_host" in b:
            if self.world == 1 and b.get("ev_out") is not None:
                b["ev_out"].synchronize()
            return b["rows_host"].numpy().copy(), b["scores_host"].numpy().copy()
        return rows.cpu().numpy(), scores.cpu().numpy()          # repeated through the blocking path (or not staged)

    def wait(self, ticket: int):
        """-> (rows int64 [B, k], scores fp32 [B, k]) of that batch, final on every rank."""
        return self._wait(ticket)[:2]

    def _wait(self, ticket: int):
        q, k, b = self._pending[ticket]
        self._pending[ticket] = None
        final, _ = self.ops.search_wait(ticket)
        if self.world > 1:
            if b["ev_side"] is not None:
                b["ev_side"].synchronize()
            final = int(b["flags_host"].sum()) == 0            # every rank reads the same gathered counts
        if final:
            return ((b["out_rows"], b["out_scores"]) if self.world > 1 else (b["packed"][1], b["scores"])) + (True, b)
        rows, scores = self.search_dev(q, k)           # rare: repeat the batch through the blocking path (all ranks)
        return rows.clone(), scores.clone(), False, b

    def append_dev(self, rows: torch.Tensor) -> int:
        assert rows.dtype == torch.float32 and rows.is_contiguous() and rows.shape[1] == self.dim
        return self.engine.append_dev(rows.data_ptr(), rows.shape[0])

    def search_dev(self, q: torch.Tensor, k: int):
        """q: [B, dim] fp32 on this rank's device (the same batch on every rank).  Returns (rows int64 [B, k] global
        row ids, scores fp32 [B, k]) on the device, identical on every rank.  The returned tensors are work buffers
        owned by the index: they are overwritten by the next call with the same (B, k) -- clone to keep them."""
        B = q.shape[0]
        dev = q.device
        b = self._buffers(B, k, dev)
        packed, scores = b["packed"], b["scores"]                            # plane 0: fp64 key bits, plane 1: rows
        self.ops.search(q, k, packed, scores)
        if self.world == 1:
            return packed[1], scores
        gathered, out_rows, out_scores = b["gathered"], b["out_rows"], b["out_scores"]
        dist.all_gather_into_tensor(gathered, packed, group=self.group)      # the one collective of the path
        self.ops.merge(gathered, self.world, B, k, out_rows, out_scores)
        self.merge_launches += 1
        return out_rows, out_scores

    # -- hybrid over row-sharded postings ----------------------------------------------------------
    def bm25_build(self, indptr, doc, tf, doclen):
        """Local CSR postings of this rank's rows (doc = local row).  docCount, sumTotalTermFreq and df are summed
        over the ranks first, so idf / avgdl -- and with them every score -- are the ones a single shard would
        compute (deliberately not OpenSearch's per-shard statistics)."""
        import numpy as np
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        doclen = np.ascontiguousarray(doclen, dtype=np.uint32)
        stats = torch.from_numpy(np.concatenate([[np.count_nonzero(doclen), int(doclen.astype(np.int64).sum())],
                                                 np.diff(indptr)]).astype(np.int64))
        if self.world > 1:
            dev = stats.device if dist.get_backend(self.group) != "nccl" else torch.device("cuda",
                                                                                         torch.cuda.current_device())
            stats = stats.to(dev)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
            stats = stats.cpu()
        stats = stats.numpy()
        self.ops.bm25_build(indptr, doc, tf, doclen, int(stats[0]), int(stats[1]), stats[2:].copy())

    def search_hybrid_dev(self, q: torch.Tensor | None, qterms, w_text: float, w_knn: float, k: int, qweights=None,
                          qflags=None):
        """bool.should of text clauses + knn over the row-sharded index.  q: [B, dim] or None; qterms: B term-id lists
        (global vocabulary) or None.  The knn clause matches the GLOBAL k nearest: they are found first (local top-k,
        all-gather, merge), every rank fuses its own rows against that list, and the fused local top-k are merged by
        a second all-gather.  Returns (global rows int64 [B, k], fused scores fp32 [B, k]) on the device."""
        B = q.shape[0] if q is not None else len(qterms)
        dev = q.device if q is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else torch.device("cpu")
        knn_rows = knn_scores = None
        if q is not None:
            r, s_ = self.search_dev(q, k)
            knn_rows, knn_scores = r.clone(), s_.clone()       # search_dev hands out its work buffers
        key = ("hyb", B, k, str(dev))
        b = self._bufs.get(key)
        if b is None:
            b = {"packed": torch.empty((2, B, k), dtype=torch.int64, device=dev),
                 "scores": torch.empty((B, k), dtype=torch.float32, device=dev),
                 "gathered": torch.empty((self.world * 2, B, k), dtype=torch.int64, device=dev),
                 "out_rows": torch.empty((B, k), dtype=torch.int64, device=dev),
                 "out_scores": torch.empty((B, k), dtype=torch.float32, device=dev),
                 "out_keys": torch.empty((B, k), dtype=torch.float64, device=dev)}
            self._bufs[key] = b
        self.ops.fuse(B, qterms, w_text, knn_rows, knn_scores, w_knn, k, b["packed"], b["scores"], qweights=qweights,
                      qflags=qflags)
        if self.world == 1:
            return b["packed"][1], b["scores"]
        dist.all_gather_into_tensor(b["gathered"], b["packed"], group=self.group)
        self.ops.merge(b["gathered"], self.world, B, k, b["out_rows"], b["out_scores"], b["out_keys"], raw=True)
        self.merge_launches += 1
        return b["out_rows"], b["out_scores"]

    def search(self, q_host, k: int):
        """Host-buffer flavour: q_host is a pinned or pageable [B, dim] fp32 tensor/array; results come back as
        host numpy arrays (the end-to-end path bench.py times)."""
        q = torch.as_tensor(q_host, dtype=torch.float32)
        qd = q.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
        rows, scores = self.search_dev(qd, k)
        return rows.cpu().numpy(), scores.cpu().numpy()

    def close(self):
        if self.engine is not None:
            self.engine.close()
