"""Request coalescing for the kNN path.

The reference answers one `/ask` at a time: every request issues its own `client.search` (app/main.py:1552, 1607),
so every request would stream the whole corpus for a single query (batch 1: ~330 QPS at 10M rows).  One corpus pass
answers 64 queries in the same time, so concurrent requests are worth coalescing: `MicroBatcher.submit` queues a
query and returns a future; a worker thread drains the queue into ONE engine call per batch (at most `max_batch`
queries, waiting at most `max_wait_s` for company) and hands every caller its own rows.  Results are identical to
unbatched calls: the engine's top-k is exact per query, and the top-k' of a query is the prefix of its top-k.

    batcher = MicroBatcher(engine.search_knn, max_batch=64, max_wait_s=0.002)
    rows, scores = batcher.search(q, k)                  # blocking (thread pool / sync code)
    rows, scores = await batcher.asearch(q, k)           # asyncio (the reference's FastAPI handlers)
"""
from __future__ import annotations

import asyncio
import queue
import threading
import time
from concurrent.futures import Future
from typing import Callable

import numpy as np


class MicroBatcher:
    def __init__(self, search_fn: Callable[[np.ndarray, int], tuple], max_batch: int = 64, max_wait_s: float = 0.002):
        """search_fn(Q [B, dim] fp32, k) -> (rows [B, k], scores [B, k]); called from the worker thread only."""
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self._fn = search_fn
        self.max_batch = max_batch
        self.max_wait_s = max_wait_s
        self._q: queue.Queue = queue.Queue()
        self._closed = False
        self._gate = threading.Lock()        # submit() and close() are serialised: nothing is queued behind the stop marker
        self.batches = 0            # engine calls issued
        self.requests = 0           # queries answered
        self._worker = threading.Thread(target=self._run, name="rass-microbatcher", daemon=True)
        self._worker.start()

    # -- callers -----------------------------------------------------------------------------------------
    def submit(self, q: np.ndarray, k: int) -> Future:
        """q: one query [dim] or [1, dim].  The future resolves to (rows [k], scores [k])."""
        if k < 1:
            raise ValueError("k must be >= 1")
        fut: Future = Future()
        item = (np.asarray(q, dtype=np.float32).reshape(-1), int(k), fut)
        with self._gate:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._q.put(item)
        return fut

    def search(self, q: np.ndarray, k: int):
        return self.submit(q, k).result()

    async def asearch(self, q: np.ndarray, k: int):
        return await asyncio.wrap_future(self.submit(q, k))

    def close(self):
        with self._gate:
            if self._closed:
                return
            self._closed = True
            self._q.put(None)                # the last item the queue will ever hold
        self._worker.join(timeout=10)
        # whatever the worker did not get to (it died, or the join timed out) must not leave callers blocked
        while True:
            try:
                item = self._q.get_nowait()
            except queue.Empty:
                break
            if item is not None and not item[2].done():
                item[2].set_exception(RuntimeError("MicroBatcher is closed"))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- worker ------------------------------------------------------------------------------------------
    def _collect(self):
        """Blocks for the first request, then gathers company until the batch is full or the window closes."""
        first = self._q.get()
        if first is None:
            return None
        batch = [first]
        deadline = time.monotonic() + self.max_wait_s
        while len(batch) < self.max_batch:
            left = deadline - time.monotonic()
            try:
                item = self._q.get(timeout=left) if left > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if item is None:
                self._q.put(None)          # leave the stop marker for the outer loop
                break
            batch.append(item)
        return batch

    def _run(self):
        while True:
            batch = self._collect()
            if batch is None:
                return
            # queries of different dimension cannot share a call; group by dim (normally one group)
            groups: dict[int, list] = {}
            for item in batch:
                groups.setdefault(item[0].size, []).append(item)
            for items in groups.values():
                try:
                    Q = np.stack([it[0] for it in items])
                    k = max(it[1] for it in items)
                    rows, scores = self._fn(Q, k)[:2]
                    self.batches += 1
                    self.requests += len(items)
                    for i, (_, ki, fut) in enumerate(items):
                        fut.set_result((rows[i, :ki].copy(), scores[i, :ki].copy()))
                except BaseException as e:  # every waiter learns about the failure
                    for _, _, fut in items:
                        if not fut.done():
                            fut.set_exception(e)
