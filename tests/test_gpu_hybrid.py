"""GPU parity of the BM25 + fusion kernels and of the OpenSearch-shaped boundary (client shim + B200Indexer)
against the numpy oracle.  Bar: ranked ids identical, fused float32 scores bit-identical (integer/byte work and
individually-rounded float ops), kNN scores within 1e-5 relative."""
import json
import os

import numpy as np
import pytest

from oracle import bm25, fusion, fuzzy, knn, multifield, synth

pytestmark = pytest.mark.gpu


def _engine(**kw):
    import rassengine_b200 as rb
    return rb.Engine(**kw)


def test_bm25_micro_golden_scores(golden_dir):
    """tests/golden/bm25_micro.json (hand-checkable, incl. a 45-token doc quantised to 44) through the kernel."""
    g = json.load(open(os.path.join(golden_dir, "bm25_micro.json")))
    idx = bm25.BM25Index.from_token_ids(g["docs"], g["vocab"])
    n = len(g["docs"])
    with _engine(dim=4) as e:
        e.append(np.eye(4, dtype=np.float32)[np.arange(n) % 4])
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        for c in g["cases"]:
            rows, scores = e.search_hybrid(None, [c["qterms"]], c["boost"], 0.0, n)
            want = {d: s for d, s in enumerate(c["scores"]) if s > 0}
            got = {int(r): float(s) for r, s in zip(rows[0], scores[0]) if r >= 0}
            assert got == want
            order = sorted(want, key=lambda d: (-want[d], d))
            assert [int(r) for r in rows[0] if r >= 0] == order


def _text_case(n_docs=6000, vocab=1500, dim=256, nq=24):
    indptr, doc, tf, doclen = synth.text_corpus(n_docs, vocab=vocab, seed=11, median_len=60, max_len=300)
    idx = bm25.BM25Index(indptr, doc, tf, doclen)
    X = synth.embeddings(n_docs, dim, 31)
    Q = synth.embeddings(nq, dim, 32)
    qterms = synth.text_queries(nq, vocab=vocab, seed=12)
    qterms[0] = qterms[0] + [qterms[0][0]]          # a duplicated query token counts twice
    qterms[1] = qterms[1] + [vocab + 5, -1]         # unknown tokens are ignored
    return idx, X, Q, qterms


@pytest.mark.parametrize("k", [10, 100])
def test_hybrid_matches_oracle(k):
    idx, X, Q, qterms = _text_case()
    knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
    with _engine(dim=X.shape[1]) as e:
        e.append(X)
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        rows, scores = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
        for b in range(Q.shape[0]):
            qt = [t for t in qterms[b] if 0 <= t < idx.vocab]
            want_rows, want_scores = fusion.hybrid(idx, qt, knn_rows[b], knn_scores[b], 4.5, 2.0, k)
            assert rows[b, :len(want_rows)].tolist() == want_rows.tolist(), b
            np.testing.assert_allclose(scores[b, :len(want_rows)], want_scores, rtol=2e-6)
        # text only and vector only degenerate forms
        rows_t, scores_t = e.search_hybrid(None, qterms[:4], 4.5, 2.0, k)
        for b in range(4):
            qt = [t for t in qterms[b] if 0 <= t < idx.vocab]
            wr, ws = bm25.topk(idx.score(qt, boost=4.5), k)
            assert rows_t[b, :len(wr)].tolist() == wr.tolist()
            assert scores_t[b, :len(wr)].tolist() == ws.tolist()       # bit-identical float32
        rows_v, scores_v = e.search_hybrid(Q[:4], None, 4.5, 2.0, k)
        assert np.array_equal(rows_v, knn_rows[:4])
        np.testing.assert_allclose(scores_v, np.float32(2.0) * knn_scores[:4], rtol=1e-5)


def test_hybrid_row_filter_matches_oracle_alive_mask():
    idx, X, Q, qterms = _text_case(n_docs=3000, nq=6)
    k = 10
    knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
    alive = (np.arange(3000) % 3 != 0)
    with _engine(dim=X.shape[1]) as e:
        e.append(X)
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        e.set_row_filter(alive)
        rows, scores = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
        e.set_row_filter(None)
        rows_all, _ = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
    for b in range(Q.shape[0]):
        qt = [t for t in qterms[b] if 0 <= t < idx.vocab]
        wr, ws = fusion.hybrid(idx, qt, knn_rows[b], knn_scores[b], 4.5, 2.0, k, alive=alive)
        assert rows[b, :len(wr)].tolist() == wr.tolist()
        np.testing.assert_allclose(scores[b, :len(wr)], ws, rtol=2e-6)
        wr2, _ = fusion.hybrid(idx, qt, knn_rows[b], knn_scores[b], 4.5, 2.0, k)
        assert rows_all[b, :len(wr2)].tolist() == wr2.tolist()


# ---------------------------------------------------------------------------------------------------------
# the drop-in boundary: client shim + B200Indexer, driven the way main.py / embedding_gen.py drive OpenSearch
# ---------------------------------------------------------------------------------------------------------
WORDS = ("patient reports chest pain radiating to the left arm since yesterday evening denies fever cough nausea "
         "history of hypertension diabetes mellitus type two metformin lisinopril aspirin daily allergic penicillin "
         "blood pressure pulse oxygen saturation normal sinus rhythm troponin negative discharged follow up cardiology "
         "paitent repotrs chets pian hypertention diabetis metforming asprin cardiolgy presure oxigen").split()


def test_fuzzy_expand_matches_oracle_on_the_term_dictionary():
    """rass_fuzzy_expand (device scan of the dictionary) against oracle/fuzzy.py: same terms, same edit counts,
    for exact tokens, misspellings, transpositions, short tokens and tokens absent from the dictionary."""
    vocab = sorted(set(WORDS)) + [synth.token(t) for t in range(0, 3000, 7)]
    with _engine(dim=4) as e:
        e.append(np.eye(4, dtype=np.float32))
        e.set_vocab(vocab)
        for tok in ["patient", "paitent", "chest", "chets", "pian", "to", "up", "aspirin", "metformin", "t00014",
                    "t00700", "zzzzzzzz", "cardiology", "hypertension", "a" * 40]:
            me = fuzzy.auto_max_edits(len(tok))
            want = {}
            for tid, term in enumerate(vocab):
                if abs(len(term) - len(tok)) <= me:
                    d = fuzzy.osa_distance(tok, term)
                    if d <= me:
                        want[tid] = d
            tid, ed = e.fuzzy_expand(tok, me)
            assert dict(zip(tid.tolist(), ed.tolist())) == want, tok


def test_fuzzy_expand_counts_edits_in_code_points():
    """Lucene's FuzzyQuery works on code points: "jose" is ONE edit from "jos\u00e9" (two UTF-8 bytes), a token of
    accented letters is as long as it looks.  Names (`patientName`) and note text carry such letters."""
    vocab = ["jos\u00e9", "jose", "josef", "m\u00fcller", "muller", "mueller", "\u00e9\u00e8\u00ea", "na\u00efve",
             "naive", "\u60a3\u8005", "o'neil", "oneil", "3.5", "35"]
    with _engine(dim=4) as e:
        e.append(np.eye(4, dtype=np.float32))
        e.set_vocab(vocab)
        for tok in ["jose", "jos\u00e9", "m\u00fcller", "muller", "\u00e9\u00e8\u00eb", "naive", "\u60a3\u8005", "oneil",
                    "3.5"]:
            me = fuzzy.auto_max_edits(len(tok))
            want = {}
            for tid, term in enumerate(vocab):
                if abs(len(term) - len(tok)) <= me:
                    d = fuzzy.osa_distance(tok, term)
                    if d <= me:
                        want[tid] = d
            tid, ed = e.fuzzy_expand(tok, me)
            assert dict(zip(tid.tolist(), ed.tolist())) == want, tok


def test_weighted_hybrid_matches_fuzzy_oracle():
    """rass_search_hybrid_weighted with the fuzzy rewrite's (term, weight) lists: fused float32 scores bit-identical to
    the oracle's, with and without the knn clause."""
    idx, X, Q, _ = _text_case(n_docs=5000, vocab=1200, dim=256, nq=8)
    vocab_terms = [synth.token(t) for t in range(idx.vocab)]
    rng = np.random.default_rng(77)
    k = 10
    knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
    with _engine(dim=256) as e:
        e.append(X)
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        for b in range(Q.shape[0]):
            tokens = [synth.token(int(t)) for t in rng.integers(0, idx.vocab, size=4)]
            ids, ws = fuzzy.weighted_terms(idx, vocab_terms, tokens, 4.5)
            text32 = fuzzy.score(idx, ids, ws)
            wr, wsc = fusion.hybrid(idx, None, knn_rows[b], knn_scores[b], 4.5, 2.0, k, text32=text32)
            rows, scores = e.search_hybrid(Q[b:b + 1], [ids], 0.0, 2.0, k, qweights=[ws])
            assert rows[0, :len(wr)].tolist() == wr.tolist()
            np.testing.assert_allclose(scores[0, :len(wr)], wsc, rtol=2e-6, atol=0)
            tr, ts = bm25.topk(text32, k)
            rows, scores = e.search_hybrid(None, [ids], 0.0, 0.0, k, qweights=[ws])
            assert rows[0].tolist() == tr.tolist() and scores[0].tolist() == ts.tolist()


def _chunk_docs(n_docs=1500, vocab=400, dim=64):
    indptr, doc, tf, doclen = synth.text_corpus(n_docs, vocab=vocab, seed=5, median_len=40, max_len=120)
    texts = synth.docs_as_text(indptr, doc, tf, n_docs)
    raw = np.random.default_rng(9).standard_normal((n_docs, dim)).astype(np.float32) * 3.0   # NOT unit: store normalises
    docs = [{"doc_id": f"txt-note-{i}", "doc_type": "unstructured", "resourceType": "DocumentReference",
             "file_path": f"/data/p{i % 7}.txt", "file_type": "txt", "patientId": f"pat-{i % 7}",
             "unstructuredText": texts[i]} for i in range(n_docs)]
    return docs, raw, (indptr, doc, tf, doclen)


def test_indexer_surface_end_to_end():
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    docs, raw, (indptr, doc, tf, doclen) = _chunk_docs()
    dim = raw.shape[1]
    client = B200Client(hosts=[{"host": "localhost", "port": 9200}], http_compress=True, use_ssl=False)
    name = ix.get_index_name("user-1")
    ix.ensure_index_exists(client, name, ix.index_body(dim))
    ix.ensure_index_exists(client, name, ix.index_body(dim))            # idempotent
    assert client.indices.exists(name)
    idxr = ix.B200Indexer(client, name)
    assert not idxr.has_any_data()
    ok, errors = ix.store_chunks(client, name, docs, raw)
    assert ok == len(docs) and not errors
    assert idxr.has_any_data() and client.count(index=name)["count"] == len(docs)
    assert "version" in client.info()

    X = knn.normalize_rows(raw).astype(np.float32)                      # what the store normalises to
    # the JSON round trip of the reference sends python floats: the stored rows are float32(list(X))
    rng = np.random.default_rng(3)
    q_emb = rng.standard_normal((1, dim)).astype(np.float32)
    q_unit = (q_emb / (np.linalg.norm(q_emb, axis=1, keepdims=True) + 1e-9)).astype(np.float32)
    k = 5
    want_rows, _, want_scores = knn.knn_exact(X, q_unit, k)

    hits = idxr.semantic_search(q_emb, k=k, query="ignored like ask() passes it")
    assert [h[0]["doc_id"] for h in hits] == [docs[r]["doc_id"] for r in want_rows[0]]
    np.testing.assert_allclose([h[1] for h in hits], want_scores[0], rtol=1e-5)
    assert len(hits[0][0]["embedding"]) == dim                          # _source comes back whole

    # hybrid: 4.5 * BM25(unstructuredText, fuzziness AUTO) + 2.0 * knn, boosted sum, no normalisation
    # (app/main.py:1574-1598); the text clause is the fuzzy rewrite restated in oracle/fuzzy.py
    bm = bm25.BM25Index(indptr, doc, tf, doclen)
    vocab_terms = [synth.token(t) for t in range(bm.vocab)]
    qtext = " ".join(synth.token(t) for t in (3, 17, 3, 250)) + " unknownword zz"
    qtokens = [synth.token(t) for t in (3, 17, 3, 250)] + ["unknownword", "zz"]

    def fused(w_text, w_knn, alive=None):
        ids, ws_ = fuzzy.weighted_terms(bm, vocab_terms, qtokens, w_text)
        assert len(ids) > len(qtokens)               # the synthetic tokens do have neighbours within 2 edits
        return fusion.hybrid(bm, None, want_rows[0], want_scores[0], w_text, w_knn, k, alive=alive,
                             text32=fuzzy.score(bm, ids, ws_))

    wr, ws = fused(4.5, 2.0)
    hits = idxr.hybrid_search(qtext, q_emb, k=k)
    assert [h[0]["doc_id"] for h in hits] == [docs[r]["doc_id"] for r in wr]
    np.testing.assert_allclose([h[1] for h in hits], ws, rtol=2e-6)
    # north-star core
    core = idxr.search(q_emb, qtext, top_k=k)
    assert [h["_id"] for h in core] == [docs[r]["doc_id"] for r in wr] and all("_score" in h for h in core)
    # multi-intent: same shape, boosts 1.0*3 / 1.5
    wr3, ws3 = fused(3.0, 1.5)
    hits = idxr.multi_intent_search(qtext, q_emb, k=k)
    assert [h[0]["doc_id"] for h in hits] == [docs[r]["doc_id"] for r in wr3]

    # patient filter: post-filter on the k nearest for knn (nmslib semantics), pre-filter mask for hybrid
    pid = "pat-3"
    hits = idxr.semantic_search(q_emb, k=k, patient_id=pid)
    assert [h[0]["doc_id"] for h in hits] == [docs[r]["doc_id"] for r in want_rows[0] if docs[r]["patientId"] == pid]
    alive = np.array([d["patientId"] == pid for d in docs])
    wrf, wsf = fused(4.5, 2.0, alive=alive)
    hits = idxr.hybrid_search(qtext, q_emb, k=k, patient_id=pid)
    assert [h[0]["doc_id"] for h in hits] == [docs[r]["doc_id"] for r in wrf]
    assert all(h[0]["patientId"] == pid for h in hits)

    # knn_filter="pre": the exact top-k AMONG the patient's rows (device pass mask inside the scan)
    pre_client = B200Client(knn_filter="pre")
    ix.ensure_index_exists(pre_client, name, ix.index_body(dim))
    ix.store_chunks(pre_client, name, docs, raw)
    wr_pre, _, ws_pre = knn.knn_exact(X, q_unit, k, alive=alive)
    hits = ix.B200Indexer(pre_client, name).semantic_search(q_emb, k=k, patient_id=pid)
    assert [h[0]["doc_id"] for h in hits] == [docs[r]["doc_id"] for r in wr_pre[0]]
    np.testing.assert_allclose([h[1] for h in hits], ws_pre[0], rtol=1e-5)
    pre_client.close()

    # bulk "index" on an existing _id overwrites in place (same row, new vector and text)
    from rassengine_b200.client import bulk
    new = dict(docs[int(want_rows[0][0])])
    new["unstructuredText"] = "replaced text"
    new["embedding"] = (-X[int(want_rows[0][0])]).tolist()
    s, e = bulk(client, [{"_op_type": "index", "_index": name, "_id": new["doc_id"], "_source": new}])
    assert (s, e) == (1, []) and client.count(index=name)["count"] == len(docs)
    hits = idxr.semantic_search(q_emb, k=k)
    assert new["doc_id"] not in [h[0]["doc_id"] for h in hits]

    # error conventions: empty inputs short-circuit, unsupported DSL / unknown index never raise out of the indexer
    assert idxr.semantic_search(np.zeros((0, dim), dtype=np.float32)) == []
    assert idxr.hybrid_search("   ", q_emb) == []
    assert ix.B200Indexer(client, "no-such-index").semantic_search(q_emb) == []
    assert idxr.semantic_search(q_emb, k=k, filter_clause=[{"text": "x", "label": "Y"}]) == []   # ask()'s NER list
    assert idxr.exact_match_search("zzz not in any document") == []
    assert sum(b["doc_count"] for b in idxr.aggregate_search("x")["by_patient"]["buckets"]) <= len(docs)
    assert client.search(index=name, body={"size": 3, "query": {"match_phrase": {"unstructuredText": "zzz"}}}
                         )["hits"]["hits"] == []                    # keyword-side shapes are answered host-side (N4)
    with pytest.raises(NotImplementedError):
        client.search(index=name, body={"size": 3, "query": {"span_near": {"clauses": []}}})
    client.close()


def test_opensearchpy_shim_import(monkeypatch):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "rassengine_b200", "shim"))
    sys.modules.pop("opensearchpy", None)
    from opensearchpy import OpenSearch, RequestsHttpConnection
    from opensearchpy.helpers import bulk
    c = OpenSearch(hosts=[{"host": "localhost", "port": 9200}], http_compress=True, use_ssl=False,
                   verify_certs=False, connection_class=RequestsHttpConnection)
    c.indices.create(index="i", body={"settings": {"index": {"knn": True}}, "mappings": {"properties": {
        "embedding": {"type": "knn_vector", "dimension": 8, "method": {"name": "hnsw", "engine": "nmslib",
                                                                         "space_type": "cosinesimil"}}}}})
    vecs = np.eye(8, dtype=np.float32)
    ok, err = bulk(c, [{"_op_type": "index", "_index": "i", "_id": f"d{i}", "_source": {"doc_id": f"d{i}",
                        "embedding": vecs[i].tolist()}, "_routing": "p"} for i in range(8)])
    assert ok == 8 and not err
    resp = c.search(index="i", body={"size": 2, "query": {"knn": {"embedding": {"vector": vecs[5].tolist(), "k": 2}}},
                                     "terminate_after": 2}, routing="p")
    assert [h["_id"] for h in resp["hits"]["hits"]] == ["d5", "d0"]
    assert resp["hits"]["hits"][0]["_score"] == pytest.approx(1.0) and resp["hits"]["hits"][1]["_score"] == pytest.approx(0.5)
    c.close()
    sys.modules.pop("opensearchpy", None)


def test_mixed_index_multi_field_dismax_matches_oracle():
    """structured FHIR documents share the index with the chunks (app/main.py:1223-1240); the hybrid
    query's two multi_match clauses then score max-over-fields per clause and sum over clauses
    (oracle/multifield.py).  Ranked ids identical, fused float32 scores bit-identical."""
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    rng = np.random.default_rng(21)
    names = ["john smith", "jon smyth", "maria garcia", "john garcia", "wei chen", "smith johnson"]
    conds = ["chest pain", "type two diabetes mellitus", "essential hypertension", "chronic chest pain syndrome",
             "acute bronchitis", "hypertensive heart disease"]
    genders = ["male", "female", "other"]
    n_struct, n_chunk, dim = 240, 600, 64
    docs = []
    for i in range(n_struct):
        kind = i % 3
        d = {"doc_id": f"res-{i}", "doc_type": "structured", "patientId": f"pat-{i % 9}"}
        if kind == 0:
            d.update(resourceType="Patient", patientName=names[int(rng.integers(len(names)))],
                     patientGender=genders[int(rng.integers(3))], patientAddress=f"{int(rng.integers(1, 99))} main street")
        elif kind == 1:
            d.update(resourceType="Condition", conditionCodeText=conds[int(rng.integers(len(conds)))],
                     conditionNote="patient reports " + conds[int(rng.integers(len(conds)))],
                     conditionClinicalStatus=["active", "resolved"][int(rng.integers(2))])
        else:
            d.update(resourceType="Practitioner", practitionerName="dr " + names[int(rng.integers(len(names)))],
                     practitionerGender=genders[int(rng.integers(3))])
        docs.append(d)
    words = ("patient john reports chest pain and hypertension denies diabetes smith male female follow up with dr "
             "garcia regarding chronic pain medication metformin daily active resolved street").split()
    chunk_raw = rng.standard_normal((n_chunk, dim)).astype(np.float32)
    chunks = [{"doc_id": f"note-{i}", "doc_type": "unstructured", "patientId": f"pat-{i % 9}",
               "unstructuredText": " ".join(words[int(j)] for j in rng.integers(0, len(words), size=int(rng.integers(5, 40))))}
              for i in range(n_chunk)]
    client = B200Client()
    name = ix.get_index_name("mixed")
    ix.ensure_index_exists(client, name, ix.index_body(dim))
    assert ix.store_structured(client, name, docs[:100]) == (100, [])
    assert ix.store_chunks(client, name, chunks, chunk_raw, as_lists=False, flush=256) == (n_chunk, [])
    assert ix.store_structured(client, name, docs[100:]) == (n_struct - 100, [])
    all_docs = docs[:100] + chunks + docs[100:]                      # row order = ingest order
    types = {f.split("^")[0]: "text" for f in ix.TEXT_FIELDS}
    types.update({f.split("^")[0]: "keyword" for f in ix.KEYWORD_FIELDS})
    fields = multifield.build(all_docs, types)
    assert {"unstructuredText", "patientName", "conditionCodeText", "conditionNote", "patientGender"} <= set(fields)
    n = len(all_docs)
    X = np.zeros((n, dim), dtype=np.float32)
    alive = np.zeros(n, dtype=bool)
    X[100:100 + n_chunk] = knn.normalize_rows(chunk_raw)
    alive[100:100 + n_chunk] = True                                  # structured documents carry no vector
    spec = lambda lst: [(f.split("^")[0], float(f.split("^")[1]) if "^" in f else 1.0) for f in lst]
    idxr = ix.B200Indexer(client, name)
    k = 8
    for qtext in ("john smith chest pain", "male", "diabetes hypertention garcia", "active", "smyth"):
        q_emb = rng.standard_normal((1, dim)).astype(np.float32)
        q_unit = (q_emb / (np.linalg.norm(q_emb, axis=1, keepdims=True) + 1e-9)).astype(np.float32)
        knn_rows, _, knn_scores = knn.knn_exact(X, q_unit, k, alive=alive)
        for (w_t, w_kw, w_knn), call in (((1.5, 1.0, 2.0), idxr.hybrid_search), ((1.0, 0.5, 1.5), idxr.multi_intent_search)):
            total = multifield.text_total(fields, [(qtext, spec(ix.TEXT_FIELDS), w_t, True),
                                                   (qtext, spec(ix.KEYWORD_FIELDS), w_kw, False)], n)
            wr, ws = fusion.hybrid(None, None, knn_rows[0], knn_scores[0], 0.0, w_knn, k, text64=total)
            hits = call(qtext, q_emb, k=k)
            assert [h[0]["doc_id"] for h in hits] == [all_docs[r]["doc_id"] for r in wr], qtext
            np.testing.assert_allclose([h[1] for h in hits], ws, rtol=2e-6, atol=0)
    # a structured document and a chunk both surface for a name query; a keyword value matches only as a whole string
    hits = idxr.hybrid_search("john smith", rng.standard_normal((1, dim)).astype(np.float32), k=k)
    assert {h[0]["doc_type"] for h in hits} == {"structured", "unstructured"} or len(hits) == k
    client.close()


def test_multifield_micro_golden_through_the_client(golden_dir):
    """tests/golden/multifield_micro.json (hand-sized corpus: per-field statistics, a transposition typo, a 2-character
    token, keyword values with a space, a 47-token field quantised to 46) through the OpenSearch-shaped client:
    ranked ids equal the golden totals' order (score desc, row asc), scores their float32 image."""
    from rassengine_b200.client import B200Client
    g = json.load(open(os.path.join(golden_dir, "multifield_micro.json")))
    props = {f: {"type": t} for f, t in g["types"].items()}
    props["embedding"] = {"type": "knn_vector", "dimension": 8, "method": {"space_type": "cosinesimil"}}
    client = B200Client()
    client.indices.create(index="micro", body={"mappings": {"properties": props}})
    ok, errs = client.bulk_actions([{"_op_type": "index", "_index": "micro", "_id": f"d{i}", "_source": d}
                                    for i, d in enumerate(g["docs"])])
    assert (ok, errs) == (len(g["docs"]), [])
    spec = lambda lst: [f"{f}^{b:g}" for f, b in lst]
    for c in g["cases"]:
        body = {"size": len(g["docs"]), "query": {"bool": {"should": [
            {"multi_match": {"query": c["query"], "fields": spec(g["text_fields"]), "type": "best_fields",
                             "operator": "or", "fuzziness": "AUTO", "boost": c["w_text"]}},
            {"multi_match": {"query": c["query"], "fields": spec(g["keyword_fields"]), "type": "best_fields",
                             "operator": "or", "boost": c["w_keyword"]}}], "minimum_should_match": 1}}}
        hits = client.search(index="micro", body=body)["hits"]["hits"]
        tot32 = np.asarray(c["totals"], dtype=np.float64).astype(np.float32)
        want = sorted((i for i in range(len(tot32)) if c["totals"][i] > 0), key=lambda i: (-float(tot32[i]), i))
        assert [h["_id"] for h in hits] == [f"d{i}" for i in want], c["query"]
        assert [np.float32(h["_score"]) for h in hits] == [tot32[i] for i in want], c["query"]
    client.close()


def test_hybrid_many_tiles_batched_matches_oracle():
    """74 tiles of 4096 docs and a batched launch: tiles prune each other through the per-query bound, the frequent
    terms go through the per-tile posting offsets; ids identical, fused float32 scores bit-identical."""
    n_docs, vocab, dim, nq, k = 300_000, 4000, 256, 24, 10
    indptr, doc, tf, doclen = synth.text_corpus(n_docs, vocab=vocab, seed=71, median_len=40, max_len=160)
    idx = bm25.BM25Index(indptr, doc, tf, doclen)
    assert int(np.diff(indptr).max()) > 512 * 20          # some terms take the offset table, most of the rest do not
    X = synth.embeddings(n_docs, dim, 72)
    Q = synth.embeddings(nq, dim, 73)
    qterms = synth.text_queries(nq, vocab=vocab, seed=74)
    knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
    with _engine(dim=dim, capacity_rows=n_docs) as e:
        e.append(X)
        e.bm25_build(indptr, doc, tf, doclen)
        rows_b, scores_b = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
        rows_t, scores_t = e.search_hybrid(None, qterms, 4.5, 0.0, k)
        for b in range(nq):
            wr, ws = fusion.hybrid(idx, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k)
            assert rows_b[b].tolist() == wr.tolist(), b
            np.testing.assert_allclose(scores_b[b], ws, rtol=2e-6, atol=0)
            tr, ts = bm25.topk(idx.score(qterms[b], boost=4.5), k)
            assert rows_t[b].tolist() == tr.tolist() and scores_t[b].tolist() == ts.tolist(), b
        one_r, one_s = e.search_hybrid(Q[5:6], [qterms[5]], 4.5, 2.0, k)          # the same query on its own
        assert one_r[0].tolist() == rows_b[5].tolist() and one_s[0].tolist() == scores_b[5].tolist()


def test_multi_field_many_tiles_matches_oracle():
    """rass_bm25_build_fields + group / clause marks over 13 tiles: clause = max over its field groups of the float
    field sums, text score = double sum of the clauses, then the knn clause; batched, with cross-tile pruning."""
    n_docs, dim, nq, k = 50_000, 128, 10, 10
    V0, V1 = 800, 600
    ip0, d0, tf0, dl0 = synth.text_corpus(n_docs, vocab=V0, seed=81, median_len=30, max_len=100)
    ip1, d1, tf1, dl1 = synth.text_corpus(n_docs, vocab=V1, seed=82, median_len=8, max_len=30)
    has1 = (np.arange(n_docs) % 3) != 0                      # a third of the documents lack the second field
    keep = has1[d1]
    t_of = np.repeat(np.arange(V1), np.diff(ip1))[keep]
    ip1 = np.zeros(V1 + 1, dtype=np.int64)
    np.add.at(ip1, t_of + 1, 1)
    ip1 = np.cumsum(ip1)
    d1, tf1 = d1[keep], tf1[keep]
    dl1 = np.where(has1, dl1, 0).astype(np.uint32)
    f0, f1 = bm25.BM25Index(ip0, d0, tf0, dl0), bm25.BM25Index(ip1, d1, tf1, dl1)
    indptr = np.concatenate([ip0[:-1], ip1 + ip0[-1]])
    term_field = np.concatenate([np.zeros(V0, np.int32), np.ones(V1, np.int32)])
    X = synth.embeddings(n_docs, dim, 83)
    Q = synth.embeddings(nq, dim, 84)
    knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
    rng = np.random.default_rng(85)
    qterms, qweights, qflags, want = [], [], [], []
    for b in range(nq):
        a0 = [int(t) for t in rng.integers(0, 60, size=3)]          # frequent terms of field 0
        a1 = [int(t) for t in rng.integers(0, V1, size=3)]
        c1 = [int(t) for t in rng.integers(0, 40, size=2)]          # second clause: field 1 only
        w0, w1, w2 = f0.term_weights(a0, 4.5), f1.term_weights(a1, 3.0), f1.term_weights(c1, 1.5)
        qterms.append(a0 + [V0 + t for t in a1] + [V0 + t for t in c1])
        qweights.append(np.concatenate([w0, w1, w2]))
        qflags.append([0, 0, 1, 0, 0, 1 | 2, 0, 1 | 2])
        clause1 = np.maximum(f0.score(a0, boost=4.5), f1.score(a1, boost=3.0))
        total = clause1.astype(np.float64) + f1.score(c1, boost=1.5).astype(np.float64)
        want.append(fusion.hybrid(None, None, knn_rows[b], knn_scores[b], 0.0, 2.0, k, text64=total))
    with _engine(dim=dim, capacity_rows=n_docs) as e:
        e.append(X)
        e.bm25_build_fields(indptr, np.concatenate([d0, d1]), np.concatenate([tf0, tf1]), term_field,
                            np.stack([dl0, dl1]))
        rows, scores = e.search_hybrid(Q, qterms, 0.0, 2.0, k, qweights=qweights, qflags=qflags)
        for b in range(nq):
            wr, ws = want[b]
            assert rows[b].tolist() == wr.tolist(), b
            np.testing.assert_allclose(scores[b], ws, rtol=2e-6, atol=0)
        r1, s1 = e.search_hybrid(Q[3:4], [qterms[3]], 0.0, 2.0, k, qweights=[qweights[3]], qflags=[qflags[3]])
        assert r1[0].tolist() == rows[3].tolist() and s1[0].tolist() == scores[3].tolist()


def test_hybrid_degenerate_queries():
    """Queries with no known term, no text at all, more hits requested than documents match, and a batch that mixes
    them; a filter that passes nothing."""
    idx, X, Q, _ = _text_case(n_docs=3000, vocab=500, dim=256, nq=6)
    k = 10
    knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
    rare = int(np.argmin(np.where(idx.df > 0, idx.df, 1 << 30)))           # a term very few documents carry
    qterms = [[idx.vocab + 3, -1], [], [rare], [rare, rare], [0, 1, 2], [rare, idx.vocab + 9]]
    with _engine(dim=256) as e:
        e.append(X)
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        rows, scores = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
        rows_t, scores_t = e.search_hybrid(None, qterms, 4.5, 0.0, k)
        for b in range(6):
            wr, ws = fusion.hybrid(idx, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k)
            assert rows[b, :len(wr)].tolist() == wr.tolist() and (rows[b, len(wr):] == -1).all(), b
            np.testing.assert_allclose(scores[b, :len(wr)], ws, rtol=2e-6, atol=0)
            tr, ts = bm25.topk(idx.score(qterms[b], boost=4.5), k)
            assert rows_t[b, :len(tr)].tolist() == tr.tolist() and (rows_t[b, len(tr):] == -1).all(), b
            assert scores_t[b, :len(tr)].tolist() == ts.tolist()
        assert (rows_t[0] == -1).all() and (rows_t[1] == -1).all()         # nothing matches -> no hits, not an error
        df_rare = int(idx.df[rare])
        rows_big, _ = e.search_hybrid(None, [[rare]], 4.5, 0.0, 128)       # more hits requested than documents match
        assert df_rare < 128 and (rows_big[0] >= 0).sum() == df_rare and (rows_big[0, df_rare:] == -1).all()
        e.set_row_filter(np.zeros(3000, dtype=np.uint8))
        rows_f, _ = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
        assert (rows_f == -1).all()
        e.set_row_filter(None)
        rows_v, scores_v = e.search_hybrid(Q, None, 0.0, 2.0, k)           # vector-only bool.should
        assert np.array_equal(rows_v, knn_rows)
        np.testing.assert_allclose(scores_v, np.float32(2.0) * knn_scores, rtol=1e-6)


def test_order_free_text_kernel_equals_the_ordered_walk():
    """hybrid_tile_fast_kernel takes all postings of a tile as one flat list with shared-memory atomics; it only runs for
    queries whose clause sums are exact in double whatever the order (checked per query from the per-term score ranges),
    so its scores must be BIT-identical to the ordered kernel's -- and to the oracle's.  10 tiles, 40 queries, k = 10 and
    100, with and without the knn clause, with a row filter."""
    idx, X, Q, qterms = _text_case(n_docs=40000, vocab=3000, dim=64, nq=40)
    alive = (np.arange(40000) % 5 != 1)
    with _engine(dim=64) as e:
        e.append(X)
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        for k in (10, 100):
            for q in (Q, None):
                for mask in (None, alive):
                    e.set_row_filter(mask)
                    e.set_hybrid_ordered(False)
                    rows_f, scores_f = e.search_hybrid(q, qterms, 4.5, 2.0, k)
                    assert e.last_stats["path"] & 0x100, "the order-free kernel should have run"
                    e.set_hybrid_ordered(True)
                    rows_o, scores_o = e.search_hybrid(q, qterms, 4.5, 2.0, k)
                    assert not (e.last_stats["path"] & 0x100)
                    assert np.array_equal(rows_f, rows_o)
                    assert np.array_equal(scores_f.view(np.uint32), scores_o.view(np.uint32))
        e.set_row_filter(None)
        e.set_hybrid_ordered(False)
        rows_t, scores_t = e.search_hybrid(None, qterms[:6], 4.5, 0.0, 10)
        for b in range(6):
            qt = [t for t in qterms[b] if 0 <= t < idx.vocab]
            wr, ws = bm25.topk(idx.score(qt, boost=4.5), 10)
            assert rows_t[b, :len(wr)].tolist() == wr.tolist() and scores_t[b, :len(wr)].tolist() == ws.tolist()


def test_sums_that_could_round_take_the_ordered_kernel():
    """A query mixing a weight of 1e-12 with a weight of 50 has addends 2^45 apart: a double sum of such floats may round,
    so its result depends on the order and the engine must walk the terms in query order (no order-free bit in
    stats.path).  Scores still bit-identical to the oracle's ordered sum."""
    idx, X, Q, qterms = _text_case(n_docs=9000, vocab=800, dim=64, nq=4)
    with _engine(dim=64) as e:
        e.append(X)
        e.bm25_build(idx.indptr, idx.doc, idx.tf, idx.doclen)
        terms = [[5, 40, 200, 7], [3, 9]]
        weights = [np.array([50.0, 1e-12, 3.0, 1e-12], dtype=np.float32), np.array([2.0, 1.0], dtype=np.float32)]
        rows, scores = e.search_hybrid(None, terms, 0.0, 0.0, 10, qweights=weights)
        assert not (e.last_stats["path"] & 0x100)
        for b in range(2):
            wr, ws = bm25.topk(fuzzy.score(idx, terms[b], weights[b]), 10)
            assert rows[b, :len(wr)].tolist() == wr.tolist() and scores[b, :len(wr)].tolist() == ws.tolist()
        rows, scores = e.search_hybrid(None, terms[1:], 0.0, 0.0, 10, qweights=weights[1:])     # this one alone is order-free
        assert e.last_stats["path"] & 0x100
        wr, ws = bm25.topk(fuzzy.score(idx, terms[1], weights[1]), 10)
        assert rows[0, :len(wr)].tolist() == wr.tolist() and scores[0, :len(wr)].tolist() == ws.tolist()


def test_maxscore_pruning_keeps_results_exact():
    """147 tiles x 64 queries is ten times what the GPU holds at once, so most tiles start with a pruning bound and
    score the postings of the frequent (non-essential) terms for marked docs only.  The emitted ids and float32 scores
    must still be the oracle's -- text-only, fused with the knn clause, top-10 and top-100 -- and identical to the ordered
    kernel's, which never prunes."""
    n_docs, vocab, dim, nq = 600_000, 8000, 64, 64
    indptr, doc, tf, doclen = synth.text_corpus(n_docs, vocab=vocab, seed=81, median_len=50, max_len=200)
    idx = bm25.BM25Index(indptr, doc, tf, doclen)
    X = synth.embeddings(n_docs, dim, 82)
    Q = synth.embeddings(nq, dim, 83)
    qterms = synth.text_queries(nq, vocab=vocab, seed=84)
    with _engine(dim=dim, capacity_rows=n_docs) as e:
        e.append(X)
        e.bm25_build(indptr, doc, tf, doclen)
        e.set_hybrid_maxscore(True)
        for k in (10, 100):
            knn_rows, _, knn_scores = knn.knn_exact(X, Q, k)
            rows_b, scores_b = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
            assert e.last_stats["path"] & 0x100
            rows_t, scores_t = e.search_hybrid(None, qterms, 4.5, 0.0, k)
            e.set_hybrid_maxscore(False)
            rows_n, scores_n = e.search_hybrid(Q, qterms, 4.5, 2.0, k)            # the default: no term split
            assert np.array_equal(rows_b, rows_n) and np.array_equal(scores_b.view(np.uint32), scores_n.view(np.uint32))
            e.set_hybrid_maxscore(True)
            e.set_hybrid_ordered(True)
            rows_o, scores_o = e.search_hybrid(Q, qterms, 4.5, 2.0, k)
            e.set_hybrid_ordered(False)
            assert np.array_equal(rows_b, rows_o) and np.array_equal(scores_b.view(np.uint32), scores_o.view(np.uint32))
            for b in range(0, nq, 1 if k == 10 else 8):
                wr, ws = fusion.hybrid(idx, qterms[b], knn_rows[b], knn_scores[b], 4.5, 2.0, k)
                assert rows_b[b, :len(wr)].tolist() == wr.tolist(), (k, b)
                np.testing.assert_allclose(scores_b[b, :len(wr)], ws, rtol=2e-6, atol=0)
                tr, ts = bm25.topk(idx.score(qterms[b], boost=4.5), k)
                assert rows_t[b, :len(tr)].tolist() == tr.tolist() and scores_t[b, :len(tr)].tolist() == ts.tolist(), (k, b)


def test_64_different_patient_filters_in_one_call_equal_64_filtered_calls():
    """Every production query carries its own `term patientId` filter (app/main.py:1599-1604).  One batched call with a row
    list per query must return exactly what 64 single calls with that filter set return -- kNN (pre-filter) and hybrid
    (knn clause = the global k nearest, post-filter, and the pre-filter variant) -- and what the oracle returns."""
    n_docs, vocab, dim, nq, k = 30000, 2000, 128, 64, 10
    indptr, doc, tf, doclen = synth.text_corpus(n_docs, vocab=vocab, seed=91, median_len=40, max_len=160)
    idx = bm25.BM25Index(indptr, doc, tf, doclen)
    X = synth.embeddings(n_docs, dim, 92)
    Q = synth.embeddings(nq, dim, 93)
    qterms = synth.text_queries(nq, vocab=vocab, seed=94)
    rng = np.random.default_rng(95)
    patient = rng.integers(0, 200, size=n_docs)                      # ~150 rows per patient
    lists = [np.flatnonzero(patient == (b * 3) % 200) for b in range(nq)]
    lists[5] = lists[5][:4]                                          # fewer rows than k
    lists[6] = np.zeros(0, dtype=np.int64)                           # a patient without documents
    lists[7] = rng.permutation(lists[7])                             # the list need not be sorted
    not_tomb = np.ones(n_docs, dtype=bool)
    not_tomb[int(lists[0][0])] = False
    knn_g, _, knn_gs = knn.knn_exact(X, Q, k, alive=not_tomb)
    with _engine(dim=dim, capacity_rows=n_docs) as e:
        e.append(X)
        e.bm25_build(indptr, doc, tf, doclen)
        e.tombstone(int(lists[0][0]))
        r_knn, s_knn, key_knn = e.search_knn_filtered(Q, k, lists, want_keys=True)
        r_post, s_post = e.search_hybrid_filtered(Q, qterms, 4.5, 2.0, k, lists)
        r_pre, s_pre = e.search_hybrid_filtered(Q, qterms, 4.5, 2.0, k, lists, knn_pre=True)
        r_txt, s_txt = e.search_hybrid_filtered(None, qterms, 4.5, 0.0, k, lists)
        for b in range(nq):
            mask = np.zeros(n_docs, dtype=bool)
            mask[lists[b]] = True
            alive = mask.copy()
            alive[int(lists[0][0])] = False                           # the tombstoned vector never matches a knn clause
            # single calls with the same filter set for the whole call
            e.set_row_filter(mask)
            e.set_knn_prefilter(True)
            r1, s1, k1 = e.search_knn(Q[b:b + 1], k, want_keys=True)
            e.set_knn_prefilter(False)
            h1, hs1 = e.search_hybrid(Q[b:b + 1], [qterms[b]], 4.5, 2.0, k)
            t1, ts1 = e.search_hybrid(None, [qterms[b]], 4.5, 0.0, k)
            e.set_row_filter(None)
            assert r_knn[b].tolist() == r1[0].tolist() and s_knn[b].tolist() == s1[0].tolist(), b
            nv = int((r1[0] >= 0).sum())
            assert key_knn[b, :nv].tolist() == k1[0, :nv].tolist()
            assert r_post[b].tolist() == h1[0].tolist() and s_post[b].tolist() == hs1[0].tolist(), b
            assert r_txt[b].tolist() == t1[0].tolist() and s_txt[b].tolist() == ts1[0].tolist(), b
            # the oracle: kNN restricted to the list; fusion with the global / the restricted k nearest
            wr, _, ws = knn.knn_exact(X, Q[b:b + 1], k, alive=alive)
            kk = wr.shape[1]
            assert r_knn[b, :kk].tolist() == wr[0].tolist() and (r_knn[b, kk:] == -1).all()
            fr, fs = fusion.hybrid(idx, qterms[b], knn_g[b], knn_gs[b], 4.5, 2.0, k, alive=mask)
            assert r_post[b, :len(fr)].tolist() == fr.tolist(), b
            np.testing.assert_allclose(s_post[b, :len(fr)], fs, rtol=2e-6, atol=0)
            pr, ps = fusion.hybrid(idx, qterms[b], wr[0], ws[0], 4.5, 2.0, k, alive=mask)
            assert r_pre[b, :len(pr)].tolist() == pr.tolist(), b
            np.testing.assert_allclose(s_pre[b, :len(pr)], ps, rtol=2e-6, atol=0)
