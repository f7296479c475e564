"""Host logic of the request coalescer against a CPU double of the engine call."""
import asyncio
import threading
import time

import numpy as np
import pytest

from rassengine_b200.batcher import MicroBatcher


class FakeEngine:
    """search(Q, k): row j of query i is 1000 * tag(i) + j, where tag = first element of the query."""

    def __init__(self, delay=0.0):
        self.calls = []
        self.delay = delay

    def search(self, Q, k):
        self.calls.append((Q.shape[0], k))
        time.sleep(self.delay)
        tags = Q[:, 0].astype(np.int64)
        rows = tags[:, None] * 1000 + np.arange(k)[None, :]
        return rows, rows.astype(np.float32) / 7.0


def test_concurrent_requests_share_calls_and_get_their_own_rows():
    eng = FakeEngine(delay=0.01)
    with MicroBatcher(eng.search, max_batch=16, max_wait_s=0.05) as mb:
        out = {}

        def worker(tag):
            q = np.full(8, tag, dtype=np.float32)
            out[tag] = mb.search(q, 3 + tag % 4)

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(40)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert mb.requests == 40 and mb.batches < 40
    assert max(b for b, _ in eng.calls) <= 16 and sum(b for b, _ in eng.calls) == 40
    for tag, (rows, scores) in out.items():
        k = 3 + tag % 4
        assert rows.tolist() == [tag * 1000 + j for j in range(k)]          # the prefix of the batch's top-k_max
        np.testing.assert_allclose(scores, rows.astype(np.float32) / 7.0)


def test_single_request_is_not_held_longer_than_the_window():
    eng = FakeEngine()
    with MicroBatcher(eng.search, max_batch=64, max_wait_s=0.01) as mb:
        t0 = time.perf_counter()
        rows, _ = mb.search(np.ones(4, dtype=np.float32), 2)
        assert time.perf_counter() - t0 < 0.5
        assert rows.tolist() == [1000, 1001] and eng.calls == [(1, 2)]


def test_asyncio_front_end_and_error_propagation():
    eng = FakeEngine()

    async def main(mb):
        res = await asyncio.gather(*[mb.asearch(np.full(4, t, dtype=np.float32), 2) for t in range(10)])
        return [r[0].tolist() for r in res]

    with MicroBatcher(eng.search, max_batch=8, max_wait_s=0.02) as mb:
        got = asyncio.run(main(mb))
    assert got == [[t * 1000, t * 1000 + 1] for t in range(10)]

    def boom(Q, k):
        raise RuntimeError("device lost")

    with MicroBatcher(boom, max_wait_s=0.001) as mb:
        with pytest.raises(RuntimeError, match="device lost"):
            mb.search(np.ones(4, dtype=np.float32), 1)
        with pytest.raises(ValueError):
            mb.submit(np.ones(4, dtype=np.float32), 0)
    with pytest.raises(RuntimeError):
        mb.submit(np.ones(4, dtype=np.float32), 1)


def test_close_races_with_submit_without_stranding_a_caller():
    """submit() and close() are serialised: a request is either answered or refused, never queued behind the stop
    marker with a future nobody resolves (indices.delete / client.close run while other threads search)."""
    for _ in range(20):
        eng = FakeEngine(delay=0.001)
        mb = MicroBatcher(eng.search, max_batch=8, max_wait_s=0.001)
        results, refused = [], []

        def worker(tag):
            for j in range(50):
                try:
                    fut = mb.submit(np.array([tag, 0.0], dtype=np.float32), 2)
                except RuntimeError:
                    refused.append(tag)
                    return
                try:
                    results.append(fut.result(timeout=5))
                except RuntimeError:                      # refused after the fact: closed before the worker got to it
                    refused.append(tag)
                    return

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
        for t in ts:
            t.start()
        time.sleep(0.005)
        mb.close()
        for t in ts:
            t.join(timeout=10)
            assert not t.is_alive(), "a caller is blocked on a future that will never resolve"
        with pytest.raises(RuntimeError):
            mb.submit(np.zeros(2, dtype=np.float32), 1)
