"""GPU tests of the rows SURVEY.md 8f marks "next" around the hot path: snapshot/restore, the fast ingest path and
request coalescing.  Each is held to the same bar as the path itself: results identical to the plain calls."""
import threading

import numpy as np
import pytest

from oracle import knn, synth

pytestmark = pytest.mark.gpu


def _chunk_docs(n_docs=1200, vocab=300, dim=64):
    indptr, doc, tf, doclen = synth.text_corpus(n_docs, vocab=vocab, seed=6, median_len=30, max_len=90)
    texts = synth.docs_as_text(indptr, doc, tf, n_docs)
    raw = np.random.default_rng(10).standard_normal((n_docs, dim)).astype(np.float32) * 2.0
    docs = [{"doc_id": f"note-{i}", "doc_type": "unstructured", "patientId": f"pat-{i % 5}",
             "unstructuredText": texts[i]} for i in range(n_docs)]
    return docs, raw


@pytest.mark.parametrize("bf16", [False, True])
def test_engine_snapshot_roundtrip(tmp_path, bf16):
    import rassengine_b200 as rb
    flags = rb.BF16_ONLY if bf16 else 0
    X = synth.embeddings(9000, 384, 41)
    Q = synth.embeddings(7, 384, 42)
    path = str(tmp_path / "store.vec")
    with rb.Engine(dim=384, flags=flags) as e:
        e.append(X[:5000])
        e.append(X[5000:])
        e.tombstone(3)
        e.tombstone(8123)
        want = e.search_knn(Q, 10, want_keys=True)
        stored = e.read_rows(0, 9000)
        e.save(path)
    with rb.Engine(dim=384, flags=flags) as e2:
        e2.load(path)
        assert e2.rows() == 9000 and e2.count() == 8998
        np.testing.assert_array_equal(e2.read_rows(0, 9000), stored)
        got = e2.search_knn(Q, 10, want_keys=True)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
        with pytest.raises(rb.RassError):
            e2.load(path)                        # only an empty engine can be restored into
    with rb.Engine(dim=128) as e3:
        with pytest.raises(rb.RassError):
            e3.load(path)                        # dimension mismatch
        with pytest.raises(rb.RassError):
            e3.load(str(tmp_path / "missing.vec"))


def test_client_snapshot_restore_and_fast_ingest(tmp_path):
    """A restored client answers kNN and hybrid queries exactly like the one that was snapshotted; the numpy-row
    ingest path stores the same values as the reference-shaped list ingest."""
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    docs, raw = _chunk_docs()
    dim = raw.shape[1]
    name = ix.get_index_name("u")
    a = B200Client()
    ix.ensure_index_exists(a, name, ix.index_body(dim))
    assert ix.store_chunks(a, name, docs, raw) == (len(docs), [])                      # python lists, like the reference
    b = B200Client()
    ix.ensure_index_exists(b, name, ix.index_body(dim))
    assert ix.store_chunks(b, name, docs, raw, as_lists=False, flush=500) == (len(docs), [])   # numpy rows
    ea, eb = a._indices[name].engine, b._indices[name].engine
    np.testing.assert_array_equal(ea.read_rows(0, len(docs)), eb.read_rows(0, len(docs)))

    q = np.random.default_rng(2).standard_normal((1, dim)).astype(np.float32)
    text = docs[17]["unstructuredText"].split()[:4]
    qtext = " ".join(text)
    ia = ix.B200Indexer(a, name)
    want_sem = ia.semantic_search(q, k=7)
    want_hyb = ia.hybrid_search(qtext, q, k=7, patient_id="pat-2")
    assert len(want_sem) == 7 and want_hyb and len(want_sem[0][0]["embedding"]) == dim
    assert [h[0]["doc_id"] for h in ix.B200Indexer(b, name).semantic_search(q, k=7)] == [h[0]["doc_id"] for h in want_sem]

    assert a.snapshot(str(tmp_path)) == [name]
    a.close()
    c = B200Client()
    assert c.restore(str(tmp_path)) == [name]
    ic = ix.B200Indexer(c, name)
    assert c.count(index=name)["count"] == len(docs)
    assert ic.semantic_search(q, k=7) == want_sem                                       # ids, scores and whole _source
    assert ic.hybrid_search(qtext, q, k=7, patient_id="pat-2") == want_hyb
    b.close()
    c.close()


def test_microbatcher_over_the_engine_is_exact():
    """64 threads each ask one query; the coalesced passes return what 64 separate calls return."""
    import rassengine_b200 as rb
    from rassengine_b200.batcher import MicroBatcher
    X = synth.embeddings(40000, 1024, 51)
    Q = synth.embeddings(64, 1024, 52)
    want_rows, _, want_scores = knn.knn_exact(X, Q, 10)
    with rb.Engine(dim=1024) as e:
        e.append(X)
        got = {}
        with MicroBatcher(e.search_knn, max_batch=64, max_wait_s=0.05) as mb:
            def ask(i):
                got[i] = mb.search(Q[i], 10 if i % 2 else 5)
            ts = [threading.Thread(target=ask, args=(i,)) for i in range(64)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            assert mb.requests == 64 and mb.batches < 64
    for i in range(64):
        k = 10 if i % 2 else 5
        assert got[i][0].tolist() == want_rows[i, :k].tolist()
        np.testing.assert_allclose(got[i][1], want_scores[i, :k], rtol=1e-5)


def test_client_batch_window_coalesces_threads():
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    docs, raw = _chunk_docs(n_docs=600)
    dim = raw.shape[1]
    name = ix.get_index_name("w")
    plain, windowed = B200Client(), B200Client(batch_window_ms=30.0)
    for c in (plain, windowed):
        ix.ensure_index_exists(c, name, ix.index_body(dim))
        ix.store_chunks(c, name, docs, raw, as_lists=False, flush=600)
    qs = np.random.default_rng(4).standard_normal((12, 1, dim)).astype(np.float32)
    want = [ix.B200Indexer(plain, name).semantic_search(q, k=4) for q in qs]
    got = [None] * len(qs)
    iw = ix.B200Indexer(windowed, name)

    def ask(i):
        got[i] = iw.semantic_search(qs[i], k=4, patient_id="pat-1" if i % 3 == 0 else None)

    ts = [threading.Thread(target=ask, args=(i,)) for i in range(len(qs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(len(qs)):
        w = want[i] if i % 3 else [h for h in want[i] if h[0]["patientId"] == "pat-1"]
        assert got[i] == w
    mb = windowed._indices[name].batcher
    assert mb is not None and mb.requests == len(qs) and mb.batches < len(qs)
    plain.close()
    windowed.close()


def test_hybrid_requests_share_the_knn_pass_under_a_batch_window():
    """Concurrent hybrid requests with DIFFERENT patient filters: their knn clauses share corpus passes through the
    MicroBatcher, text clauses and filters are fused per request (rass_fuse_hybrid); answers equal the plain client's."""
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    docs, raw = _chunk_docs(n_docs=900)
    dim = raw.shape[1]
    name = ix.get_index_name("hw")
    plain, windowed = B200Client(), B200Client(batch_window_ms=40.0)
    for c in (plain, windowed):
        ix.ensure_index_exists(c, name, ix.index_body(dim))
        ix.store_chunks(c, name, docs, raw, as_lists=False, flush=900)
    rng = np.random.default_rng(14)
    qs = rng.standard_normal((10, 1, dim)).astype(np.float32)
    texts = [" ".join(docs[int(i)]["unstructuredText"].split()[:3]) for i in rng.integers(0, len(docs), size=10)]
    pats = [None if i % 4 == 0 else f"pat-{i % 5}" for i in range(10)]
    ip = ix.B200Indexer(plain, name)
    want = [ip.hybrid_search(texts[i], qs[i], k=5, patient_id=pats[i]) for i in range(10)]
    assert all(want)
    iw = ix.B200Indexer(windowed, name)
    got = [None] * 10

    def ask(i):
        got[i] = iw.hybrid_search(texts[i], qs[i], k=5, patient_id=pats[i])

    ts = [threading.Thread(target=ask, args=(i,)) for i in range(10)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert got == want
    mb = windowed._indices[name].batcher
    assert mb is not None and mb.requests == 10 and mb.batches < 10
    plain.close()
    windowed.close()


@pytest.mark.parametrize("bf16", [False, True])
def test_store_grows_in_place_across_many_appends(bf16):
    """The store's arrays grow by mapping more physical chunks behind the same pointers (csrc/vmm.cu; cudaMalloc + copy
    where the driver lacks the API): rows appended in uneven batches across several chunk boundaries read back unchanged,
    tombstones and overwrites made before a growth survive it, and searches in between and at the end equal the oracle."""
    import rassengine_b200 as rb
    flags = rb.BF16_ONLY if bf16 else 0
    n, dim = 150_000, 1024                       # x32 chunks hold 32k rows, so the store grows several times
    X = synth.embeddings(n, dim, 51)
    Q = synth.embeddings(5, dim, 52)
    with rb.Engine(dim=dim, flags=flags) as e:
        cuts = [0, 1, 700, 33_000, 33_001, 70_000, 131_073, n]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            assert e.append(X[lo:hi]) == lo
            if hi == 700:
                e.tombstone(5)
                e.overwrite(9, X[100])
            if hi in (33_001, n):
                stored = e.read_rows(0, hi)
                Xs = stored.copy()
                rows, scores = e.search_knn(Q, 10)
                alive = np.ones(hi, dtype=bool)
                alive[5] = False
                want_rows, _, want_scores = knn.knn_exact(Xs, Q, 10, alive=alive)
                assert np.array_equal(rows, want_rows)
                np.testing.assert_allclose(scores, want_scores, rtol=1e-5)
        stored = e.read_rows(0, n)
        if not bf16:
            keep = np.ones(n, dtype=bool)
            keep[9] = False
            assert np.array_equal(stored[keep], X[keep]) and np.array_equal(stored[9], X[100])
        assert e.rows() == n and e.count() == n - 1
        info = e.store_info()
        assert info["capacity_rows"] >= n
        import os
        assert info["grows_in_place"] == ("RASS_DEBUG_NO_VMM" not in os.environ)      # the B200 driver has the API
