"""Single-process multi-GPU handle (rass_create_sharded): ONE engine handle whose rows are spread
block-cyclically over several devices, every C-ABI entry point working on it, results bit-identical to the oracle and
to a single-device handle.  On a box with one GPU the shards share device 0 (the dispatcher, the row map, the gather
buffer and the merge are exercised all the same); with >= 2 GPUs the lists travel as peer stores over NVLink."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import bm25, fusion, knn, synth

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 4))))
    return sets


def _engine(**kw):
    import rassengine_b200 as rb
    return rb.Engine(**kw)


@pytest.mark.parametrize("devices", _device_sets())
def test_sharded_knn_store_and_filters_match_oracle(devices, tmp_path):
    import rassengine_b200 as rb
    N, D, B = 23000, 1024, 70                      # 23 blocks of 1000-ish rows: the last block is partial
    X = synth.embeddings(N, D, 41)
    synth.plant_duplicates(X, 8)
    Q = synth.clustered_queries(X, B, seed=42)
    with _engine(dim=D, devices=devices) as e:
        assert e.append(X[:7]) == 0 and e.append(X[7:3000]) == 7 and e.append(X[3000:]) == 3000
        assert e.rows() == N and e.count() == N
        np.testing.assert_array_equal(e.read_rows(1020, 10), X[1020:1030])          # across a block boundary
        pick = np.array([5, 1024, 22999, 2048, 1023], dtype=np.int64)
        np.testing.assert_array_equal(e.read_rows_list(pick), X[pick])
        for k in (10, 100):
            want_rows, _, want_scores = knn.knn_exact(X, Q, k)
            for path in (rb.PATH_STREAM, rb.PATH_UMMA, rb.PATH_GEMM, rb.PATH_AUTO):
                e.set_path(path)
                nq = 4 if path == rb.PATH_STREAM else B
                rows, scores, keys = e.search_knn(Q[:nq], k, want_keys=True)
                assert np.array_equal(rows, want_rows[:nq]), (k, path)
                np.testing.assert_allclose(scores, want_scores[:nq], rtol=1e-5)
                st = e.last_stats
                assert st["n_queries"] == nq and st["rows_scanned"] == N
        e.set_path(rb.PATH_AUTO)
        # tombstone / overwrite route to the owning shard
        best = int(want_rows[0, 0])
        e.tombstone(best)
        assert e.count() == N - 1
        alive = np.ones(N, dtype=bool)
        alive[best] = False
        wr, _, ws = knn.knn_exact(X, Q[:5], 10, alive=alive)
        rows, scores = e.search_knn(Q[:5], 10)
        assert np.array_equal(rows, wr)
        e.overwrite(best, X[best])
        assert e.count() == N
        # exact pre-filter: mask and row-list forms
        rng = np.random.default_rng(33)
        patient = rng.integers(0, 40, size=N)
        e.set_knn_prefilter(True)
        for mask in (patient == 3, np.arange(N) < 4):
            wr, _, ws = knn.knn_exact(X, Q[:6], 10, alive=mask)
            e.set_row_filter(mask)
            rows, scores = e.search_knn(Q[:6], 10)
            kk = wr.shape[1]
            assert np.array_equal(rows[:, :kk], wr) and (rows[:, kk:] == -1).all()
            e.set_row_filter_rows(np.flatnonzero(mask), N)
            rows2, _ = e.search_knn(Q[:6], 10)
            assert np.array_equal(rows2, rows)
        e.set_row_filter(None)
        e.set_knn_prefilter(False)
        # snapshot: the single-engine format, restorable onto any device count
        e.tombstone(17)
        snap = str(tmp_path / "sharded.vec")
        e.save(snap)
        alive = np.ones(N, dtype=bool)
        alive[17] = False
        wr, _, ws = knn.knn_exact(X, Q[:9], 10, alive=alive)
    for devs in (None, devices):
        with _engine(dim=D, devices=devs) as e2:
            e2.load(snap)
            assert e2.rows() == N and e2.count() == N - 1
            rows, scores = e2.search_knn(Q[:9], 10)
            assert np.array_equal(rows, wr)
            np.testing.assert_allclose(scores, ws, rtol=1e-5)


@pytest.mark.parametrize("devices", _device_sets())
def test_sharded_async_device_path(devices):
    """rass_search_knn_dev_async / _wait on a sharded handle: queries and outputs on the coordinator device, two batches
    in flight, the shards' finish kernels store into the coordinator's gather buffer, one merge kernel per batch."""
    import torch
    N, D, B, k = 40000, 1024, 64, 10
    X = synth.embeddings(N, D, 51)
    Q = synth.embeddings(2 * B, D, 52)
    want, _, want_scores = knn.knn_exact(X, Q, k)
    dev = torch.device("cuda", devices[0])
    with _engine(dim=D, devices=devices) as e:
        Xd = torch.from_numpy(X).to(dev)
        torch.cuda.synchronize(dev)              # the rows are complete before another stream / device reads them
        e.append_dev(Xd.data_ptr(), N)           # device rows of the coordinator, spread by peer copies
        del Xd
        qd = [torch.from_numpy(Q[:B]).to(dev), torch.from_numpy(Q[B:]).to(dev)]
        rows = [torch.empty((B, k), dtype=torch.int64, device=dev) for _ in range(2)]
        scores = [torch.empty((B, k), dtype=torch.float32, device=dev) for _ in range(2)]
        flag = [torch.full((1,), -5, dtype=torch.int64, device=dev) for _ in range(2)]
        torch.cuda.synchronize(dev)
        for it in range(3):
            for s in range(2):
                e.search_knn_dev_async(qd[s].data_ptr(), B, k, rows[s].data_ptr(), scores[s].data_ptr(), 0, s,
                                       flag[s].data_ptr())
            for s in range(2):
                final, st = e.search_knn_dev_wait(s)
                assert final and int(flag[s].item()) == 0, st
                assert np.array_equal(rows[s].cpu().numpy(), want[s * B:(s + 1) * B]), (it, s)
                np.testing.assert_allclose(scores[s].cpu().numpy(), want_scores[s * B:(s + 1) * B], rtol=1e-5)
        st = e.search_knn_dev(qd[0].data_ptr(), B, k, rows[0].data_ptr(), scores[0].data_ptr())
        assert np.array_equal(rows[0].cpu().numpy(), want[:B]) and st["rows_scanned"] == N


@pytest.mark.parametrize("devices", _device_sets())
def test_sharded_hybrid_matches_oracle(devices):
    """Postings split by the row map with corpus-wide statistics; the merged k nearest fused on every shard; raw-score
    merge.  Fused float32 scores bit-identical to a single-device handle's, ids identical to the oracle's -- cosine and
    L2 (the fused lists are larger-is-better whatever the vector metric)."""
    import rassengine_b200 as rb
    Nh, V, Dh, Bh, k = 9000, 900, 256, 12, 10
    indptr, doc, tf, doclen = synth.text_corpus(Nh, vocab=V, seed=17, median_len=50, max_len=200)
    full = bm25.BM25Index(indptr, doc, tf, doclen)
    Xh = synth.embeddings(Nh, Dh, 43)
    Qh = synth.embeddings(Bh, Dh, 44)
    qterms = synth.text_queries(Bh, vocab=V, seed=18)
    for metric, om in ((rb.METRIC_COSINE, knn.COSINE), (rb.METRIC_L2, knn.L2)):
        knn_rows, _, knn_scores = knn.knn_exact(Xh, Qh, k, metric=om)
        with _engine(dim=Dh, metric=metric, devices=devices) as e, _engine(dim=Dh, metric=metric) as one:
            for eng in (e, one):
                eng.append(Xh)
                eng.bm25_build(indptr, doc, tf, doclen)
            rows, scores = e.search_hybrid(Qh, qterms, 4.5, 2.0, k)
            rows1, scores1 = one.search_hybrid(Qh, qterms, 4.5, 2.0, k)
            assert np.array_equal(rows, rows1) and np.array_equal(scores.view(np.uint32), scores1.view(np.uint32))
            for b in range(Bh):
                wr, ws = fusion.hybrid(full, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k)
                assert rows[b, :len(wr)].tolist() == wr.tolist(), b
                np.testing.assert_allclose(scores[b, :len(wr)], ws, rtol=2e-6, atol=0)
            rows_t, scores_t = e.search_hybrid(None, qterms, 4.5, 0.0, k)                 # text only
            for b in range(Bh):
                tr, ts = bm25.topk(full.score(qterms[b], boost=4.5), k)
                assert rows_t[b].tolist() == tr.tolist() and scores_t[b].tolist() == ts.tolist()
            # a knn list obtained earlier (request coalescing) fused with this request's text clauses and filter
            alive = (np.arange(Nh) % 3 != 0)
            e.set_row_filter(alive)
            r_f, s_f = e.fuse_hybrid(qterms, 4.5, knn_rows, knn_scores, 2.0, k)
            e.set_row_filter(None)
            for b in range(Bh):
                wr, ws = fusion.hybrid(full, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k, alive=alive)
                assert r_f[b, :len(wr)].tolist() == wr.tolist()
                np.testing.assert_allclose(s_f[b, :len(wr)], ws, rtol=2e-6, atol=0)


def test_client_spreads_an_index_over_devices_and_answers_identically():
    """B200Client(devices=...) / settings.index.number_of_shards / RASS_B200_DEVICES: the drop-in client builds ONE sharded
    handle per index; hits and scores equal the single-device client's for knn, hybrid (fuzzy, filters) and host-side
    shapes."""
    from rassengine_b200 import indexer as ix
    from rassengine_b200.client import B200Client, resolve_devices
    rng = np.random.default_rng(5)
    docs = [{"doc_id": f"c{i}", "doc_type": "unstructured", "patientId": f"pat-{i % 7}",
             "unstructuredText": " ".join(rng.choice(["chest", "pain", "fever", "cough", "diabetes", "metformin", "normal",
                                                      "sinus", "rhythm", "denies", "nausea"], size=12))}
            for i in range(5000)]
    emb = rng.standard_normal((len(docs), 32)).astype(np.float32)
    out = []
    for devs in (None, [0, 0]):
        client = B200Client(devices=devs)
        name = ix.get_index_name("sh")
        ix.ensure_index_exists(client, name, ix.index_body(32))
        assert ix.store_chunks(client, name, docs, emb) == (len(docs), [])
        idxr = ix.B200Indexer(client, name)
        q = emb[11:12] + 0.01
        res = [idxr.semantic_search(q, k=5), idxr.semantic_search(q, k=5, patient_id="pat-4"),
               idxr.hybrid_search("chest pian fever", q, k=5), idxr.hybrid_search("diabetes metformin", q, k=5,
                                                                                  patient_id="pat-2")]
        out.append([[(d["doc_id"], s) for d, s in r] for r in res])
        assert client._get(name).engine.devices == (devs or [0])
        client.close()
    assert out[0] == out[1]
    assert resolve_devices(None, 0, {"settings": {"index": {"number_of_shards": 1}}}) == [0]
    os.environ["RASS_B200_DEVICES"] = "0,0"
    try:
        assert resolve_devices(None, 0, None) == [0, 0]
    finally:
        del os.environ["RASS_B200_DEVICES"]


def test_staged_copy_exchange_without_peer_access():
    """RASS_DEBUG_NO_PEER: the lists travel by cudaMemcpyPeer instead of peer stores (what a box without P2P does)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (shards on one device always store directly)")
    code = r'''
import numpy as np, sys
sys.path.insert(0, %r)
from oracle import knn, synth
import rassengine_b200 as rb
X = synth.embeddings(20000, 256, 1); Q = synth.embeddings(70, 256, 2)
want, _, _ = knn.knn_exact(X, Q, 10)
with rb.Engine(dim=256, devices=[0, 1]) as e:
    e.append(X)
    rows, _ = e.search_knn(Q, 10)
    assert np.array_equal(rows, want)
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RASS_DEBUG_NO_PEER="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("devices", _device_sets())
def test_sharded_text_ingest_equals_single_device(devices):
    """rass_text_add_rows / rass_text_commit on a handle over several GPUs: the token streams are split by the row map,
    every shard inverts and merges its own rows, the corpus-wide statistics are summed between the merge and the
    finalisation.  Same bulks (fresh rows, rewrites, a cleared row, a growing vocabulary) into a single-device handle:
    the statistics agree and hybrid / text-only searches return the same rows with bit-identical scores."""
    rng = np.random.default_rng(71)
    n_docs, dim = 9000, 256

    def zipf(vocab, n):
        p = 1.0 / np.arange(1, vocab + 1) ** 1.07
        return rng.choice(vocab, size=n, p=p / p.sum()).astype(np.int32)

    X = synth.embeddings(n_docs, dim, 72)
    Q = synth.embeddings(12, dim, 73)
    bulks = []                                         # (field, {row: ids}), vocab sizes after the bulk
    vocab = [800, 50]
    cuts = [0, 1, 1023, 1025, 4100, n_docs]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        vocab = [vocab[0] + 100, vocab[1] + 3]
        for f in range(2):
            ch = {r: zipf(vocab[f], int(rng.integers(1, 40))) for r in range(lo, hi) if not (f == 1 and r % 3)}
            bulks.append((f, ch, list(vocab), hi))
    # rewrites of rows on different shards, one of them cleared
    bulks.append((0, {5: zipf(vocab[0], 9), 1024: np.zeros(0, np.int32), 2047: zipf(vocab[0], 30), 8999: zipf(vocab[0], 3)},
                  list(vocab), n_docs))

    def feed(e):
        out = []
        for i, (f, ch, vs, n) in enumerate(bulks):
            rows = sorted(ch)
            indptr = np.zeros(len(rows) + 1, dtype=np.int64)
            indptr[1:] = np.cumsum([ch[r].size for r in rows])
            cat = np.concatenate([ch[r] for r in rows]) if indptr[-1] else np.zeros(0, np.int32)
            e.text_add_rows(f, rows, indptr, cat)
            if f == 1 or i == len(bulks) - 1:          # a commit per bulk of rows, and one after the rewrites
                e.text_commit(vs, n)
                out.append(e.text_stats())
        return out

    qterms = [[int(t) for t in zipf(vocab[0], int(rng.integers(2, 8)))] for _ in range(12)]
    with _engine(dim=dim) as a, _engine(dim=dim, devices=devices) as b:
        a.append(X)
        b.append(X)
        sa, sb = feed(a), feed(b)
        for (ia, da), (ib, db) in zip(sa, sb):
            assert np.array_equal(ia, ib) and np.array_equal(da, db)
        assert a.text_size() == b.text_size()
        for w_text, w_knn in ((4.5, 2.0), (1.0, 0.0)):
            ra, xa = a.search_hybrid(Q if w_knn else None, qterms, w_text, w_knn, 10)
            rb_, xb = b.search_hybrid(Q if w_knn else None, qterms, w_text, w_knn, 10)
            assert np.array_equal(ra, rb_)
            assert np.array_equal(xa.view(np.uint32), xb.view(np.uint32))
