"""GPU parity of the device-side text ingest (rass_text_add_rows / rass_text_commit, csrc/postings.cu) against the CSR a
from-scratch host build gives for the same token streams, and of the searches that follow.

The reference's side of this is a bulk request's effect on the `text` / `keyword` fields (app/main.py:1258-1269):
Lucene inverts the documents into a segment and merges segments; whatever the segmentation, the searchable index is the
term -> (ascending doc, term frequency) lists of all documents plus one norm byte per document and field, which is what
oracle/bm25.py scores.  So the check is bit equality of (indptr, doc, tf, doclen, norm) with the numpy construction over
ALL rows, after every commit, for any split of the rows into bulks.
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import smallfloat, synth

pytestmark = pytest.mark.gpu


def _engine(**kw):
    import rassengine_b200 as rb
    return rb.Engine(**kw)


def _host_csr(rows_tokens, vocab_sizes, n_rows):
    """rows_tokens[f] = {row: int32 local term ids}.  -> (indptr, doc, tf, doclen [F, N]) in the client's layout: field
    f owns the global ids [sum(vocab[:f]), ...)."""
    F = len(vocab_sizes)
    indptrs, docs, tfs = [], [], []
    doclen = np.zeros((F, n_rows), dtype=np.uint32)
    nnz = 0
    for f in range(F):
        per_term = [dict() for _ in range(vocab_sizes[f])]
        for r in sorted(rows_tokens[f]):
            ids = rows_tokens[f][r]
            doclen[f, r] = len(ids)
            for t in ids:
                per_term[int(t)][r] = per_term[int(t)].get(r, 0) + 1
        ip = np.zeros(vocab_sizes[f] + 1, dtype=np.int64)
        for t, d in enumerate(per_term):
            ip[t + 1] = ip[t] + len(d)
            docs.extend(d.keys())              # insertion order = ascending rows
            tfs.extend(min(c, 65535) for c in d.values())
        indptrs.append(ip[:-1] + nnz)
        nnz += int(ip[-1])
    indptr = np.concatenate(indptrs + [np.array([nnz], dtype=np.int64)])
    return indptr, np.asarray(docs, dtype=np.int32), np.asarray(tfs, dtype=np.uint16), doclen


def _check_export(e, rows_tokens, vocab_sizes, n_rows):
    indptr, doc, tf, doclen, norm = e.text_export()
    w_indptr, w_doc, w_tf, w_doclen = _host_csr(rows_tokens, vocab_sizes, n_rows)
    assert np.array_equal(indptr, w_indptr)
    assert np.array_equal(doc, w_doc)
    assert np.array_equal(tf, w_tf)
    assert np.array_equal(doclen, w_doclen)
    assert np.array_equal(norm, smallfloat.encode_lengths(w_doclen.reshape(-1)).reshape(w_doclen.shape))
    ip2, dc = e.text_stats()
    assert np.array_equal(ip2, w_indptr) and dc.tolist() == [int(np.count_nonzero(w_doclen[f])) for f in range(len(dc))]


def _zipf_tokens(rng, vocab, n):
    p = 1.0 / np.arange(1, vocab + 1) ** 1.07
    return rng.choice(vocab, size=n, p=p / p.sum()).astype(np.int32)


def test_segments_and_commits_equal_a_from_scratch_csr():
    """Three fields (one appears late, one is sparse), bulks of uneven size, vocabularies that grow between commits,
    empty rows, a row with 70 000 repeats of one term (tf clamps at 65 535), several bulks per commit."""
    rng = np.random.default_rng(5)
    with _engine(dim=256) as e:
        rows_tokens = [dict(), dict(), dict()]
        vocab = [0, 0, 0]
        n_rows = 0
        plan = [(300, 1), (1, 1), (2500, 2), (40, 1), (5000, 3), (700, 3)]      # (rows in the bulk, fields alive)
        for step, (n_bulk, alive) in enumerate(plan):
            grow = [400 + 250 * step, 60 + 10 * step, 30][:alive]
            for f in range(alive):
                vocab[f] = max(vocab[f], grow[f])
                rows, chunks = [], []
                for r in range(n_rows, n_rows + n_bulk):
                    if f == 1 and r % 3:                      # sparse field
                        continue
                    if f == 0 and r % 17 == 5:                # a document without text
                        continue
                    n = int(rng.integers(1, 60))
                    ids = _zipf_tokens(rng, vocab[f], n)
                    if f == 0 and r == 310:
                        ids = np.full(70000, 7, dtype=np.int32)
                    rows_tokens[f][r] = ids
                    rows.append(r)
                    chunks.append(ids)
                if rows:
                    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
                    indptr[1:] = np.cumsum([c.size for c in chunks])
                    e.text_add_rows(f, rows, indptr, np.concatenate(chunks))
            n_rows += n_bulk
            if step in (1, 3):                                # two bulks ride in one commit
                continue
            F = alive
            e.text_commit(vocab[:F], n_rows)
            _check_export(e, rows_tokens[:F], vocab[:F], n_rows)
        assert e.text_size()["F"] == 3 and e.text_size()["N"] == n_rows


def test_incremental_index_scores_like_a_host_built_one():
    """The same corpus ingested as token streams in 7 bulks and built from host arrays in one go: hybrid and text-only
    searches return the same rows and bit-identical float scores."""
    rng = np.random.default_rng(11)
    n_docs, dim, V = 30000, 256, 3000
    X = synth.embeddings(n_docs, dim, 12)
    toks = [_zipf_tokens(rng, V, int(rng.integers(5, 120))) for _ in range(n_docs)]
    rows_tokens = [{r: t for r, t in enumerate(toks)}]
    csr = _host_csr(rows_tokens, [V], n_docs)
    Q = synth.embeddings(16, dim, 13)
    qterms = [[int(t) for t in _zipf_tokens(rng, V, int(rng.integers(2, 9)))] for _ in range(16)]
    with _engine(dim=dim) as a, _engine(dim=dim) as b:
        a.append(X)
        b.append(X)
        a.bm25_build_fields(csr[0], csr[1], csr[2], np.zeros(V, np.int32), csr[3])
        cuts = [0, 1, 4000, 4001, 11000, 19000, 29999, n_docs]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            indptr = np.zeros(hi - lo + 1, dtype=np.int64)
            indptr[1:] = np.cumsum([t.size for t in toks[lo:hi]])
            b.text_add_rows(0, np.arange(lo, hi), indptr, np.concatenate(toks[lo:hi]))
            if hi in (4001, 19000):
                b.text_commit([V], hi)             # searchable in between, then grown again
                _check_export(b, [{r: toks[r] for r in range(hi)}], [V], hi)
        b.text_commit([V], n_docs)
        for w_text, w_knn in ((4.5, 2.0), (1.0, 0.0)):
            ra, sa = a.search_hybrid(Q if w_knn else None, qterms, w_text, w_knn, 10)
            rb_, sb = b.search_hybrid(Q if w_knn else None, qterms, w_text, w_knn, 10)
            assert np.array_equal(ra, rb_)
            assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32))


def test_rewritten_and_cleared_rows_are_replaced_at_the_commit():
    """Index requests with a known _id: rows rewritten with other text, rows that lose the field, rows rewritten twice
    before one commit, all mixed with fresh rows and a growing vocabulary, in two fields.  After every commit the index
    equals the from-scratch build of the rows' CURRENT contents (the re-sorting commit)."""
    rng = np.random.default_rng(29)
    with _engine(dim=256) as e:
        state = [dict(), dict()]
        vocab = [300, 40]
        n_rows = 0

        def send(f, changes):
            rows = sorted(changes)
            indptr = np.zeros(len(rows) + 1, dtype=np.int64)
            indptr[1:] = np.cumsum([changes[r].size for r in rows])
            cat = np.concatenate([changes[r] for r in rows]) if indptr[-1] else np.zeros(0, np.int32)
            e.text_add_rows(f, rows, indptr, cat)
            for r in rows:
                if changes[r].size:
                    state[f][r] = changes[r]
                else:
                    state[f].pop(r, None)

        for step in range(6):
            vocab = [vocab[0] + 150, vocab[1] + 5]
            fresh = int(rng.integers(200, 900))
            for f in range(2):
                ch = {r: _zipf_tokens(rng, vocab[f], int(rng.integers(1, 50))) for r in range(n_rows, n_rows + fresh)
                      if not (f == 1 and r % 2)}
                send(f, ch)
            n_rows += fresh
            if step >= 1:
                for f in range(2):
                    old = rng.choice(n_rows - fresh, size=60, replace=False)
                    ch = {}
                    for i, r in enumerate(old):
                        ch[int(r)] = (np.zeros(0, np.int32) if i % 5 == 0
                                      else _zipf_tokens(rng, vocab[f], int(rng.integers(1, 50))))
                    send(f, ch)
                    if step == 3:      # the same rows once more before the commit: the later segment wins
                        send(f, {int(r): _zipf_tokens(rng, vocab[f], 7) for r in old[:20]})
            e.text_commit(vocab, n_rows)
            _check_export(e, state, vocab, n_rows)
        # and the plain appending commit still follows a re-sorting one
        send(0, {n_rows: np.array([1, 1, 2], np.int32), n_rows + 1: np.array([5], np.int32)})
        n_rows += 2
        e.text_commit(vocab, n_rows)
        _check_export(e, state, vocab, n_rows)


def test_ingest_continues_a_host_built_index_and_rejects_bad_streams():
    rng = np.random.default_rng(17)
    V0, n0 = 500, 2000
    toks = {r: _zipf_tokens(rng, V0, int(rng.integers(1, 40))) for r in range(n0)}
    csr = _host_csr([toks], [V0], n0)
    import rassengine_b200 as rb
    with _engine(dim=256) as e:
        e.bm25_build_fields(csr[0], csr[1], csr[2], np.zeros(V0, np.int32), csr[3])
        V1, n1 = 800, 3500                                   # more terms, more rows
        new = {r: _zipf_tokens(rng, V1, int(rng.integers(1, 40))) for r in range(n0, n1)}
        rows = sorted(new)
        indptr = np.zeros(len(rows) + 1, dtype=np.int64)
        indptr[1:] = np.cumsum([new[r].size for r in rows])
        e.text_add_rows(0, rows, indptr, np.concatenate([new[r] for r in rows]))
        e.text_commit([V1], n1)
        toks.update(new)
        _check_export(e, [toks], [V1], n1)
        one = np.array([0, 3], dtype=np.int64)
        with pytest.raises(rb.RassError) as err:             # rows of a bulk are ascending and distinct
            e.text_add_rows(0, [n1 + 5, n1 + 5], np.array([0, 1, 2], np.int64), np.array([1, 2], np.int32))
        assert "ascending" in str(err.value)
        with pytest.raises(rb.RassError):                    # negative term id
            e.text_add_rows(0, [n1], one, np.array([1, -2, 3], np.int32))
        with pytest.raises(rb.RassError):                    # vocabularies only grow
            e.text_commit([V1 - 1], n1)
        _check_export(e, [toks], [V1], n1)                   # nothing of that stuck
        # a rewrite of a row the host-built index holds
        toks[100] = np.array([3, 3, 9], np.int32)
        e.text_add_rows(0, [100], one, toks[100])
        e.text_commit([V1], n1)
        _check_export(e, [toks], [V1], n1)


def test_client_bulks_go_through_the_device_ingest():
    """B200Client: bulks of chunk documents become searchable through device-side segments (no host CSR rebuild),
    a re-indexed document through the re-sorting commit; both give the hits of a client that indexed everything at
    once."""
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    rng = np.random.default_rng(23)
    n_docs, dim = 600, 64
    raw = rng.standard_normal((n_docs, dim)).astype(np.float32)
    texts = [" ".join(synth.token(int(t)) for t in _zipf_tokens(rng, 400, int(rng.integers(4, 50)))) for _ in range(n_docs)]
    docs = [{"doc_id": f"d{i}", "doc_type": "unstructured", "resourceType": "DocumentReference", "file_path": f"/p{i % 5}",
             "file_type": "txt", "patientId": f"pat-{i % 5}", "unstructuredText": texts[i]} for i in range(n_docs)]
    q_emb = rng.standard_normal((1, dim)).astype(np.float32)
    qtext = " ".join(synth.token(t) for t in (3, 17, 250))

    def make(bulks, rewrite=None):
        client = B200Client()
        name = ix.get_index_name("u")
        ix.ensure_index_exists(client, name, ix.index_body(dim))
        idxr = ix.B200Indexer(client, name)
        hits = None
        for lo, hi in bulks:
            ok, errors = ix.store_chunks(client, name, docs[lo:hi], raw[lo:hi])
            assert ok == hi - lo and not errors
            hits = idxr.hybrid_search(qtext, q_emb, k=8)       # a refresh per bulk, like OpenSearch
        if rewrite is not None:
            ok, errors = ix.store_chunks(client, name, [docs[rewrite]], raw[rewrite:rewrite + 1])
            assert ok == 1 and not errors
            hits = idxr.hybrid_search(qtext, q_emb, k=8)
        return client, name, [(h[0]["doc_id"], h[1]) for h in hits]

    c1, n1, once = make([(0, n_docs)])
    c2, n2, many = make([(0, 1), (1, 130), (130, 131), (131, 420), (420, n_docs)])
    assert once == many
    t2 = c2._indices[n2].text
    assert t2.device_commits == 5 and t2.host_rebuilds == 0
    c3, n3, again = make([(0, 300), (300, n_docs)], rewrite=7)  # same content rewritten: same hits
    assert again == once
    t3 = c3._indices[n3].text
    assert t3.device_commits == 3 and t3.host_rebuilds == 0


def test_keyword_fields_omit_norms_on_the_device():
    """A `keyword` field counts every document that has a value as length 1, however many values it sent
    (rass_text_omit_norms): norm bytes, docCount / sumTotalTermFreq and the scores equal a host-built index that was
    handed lengths of 0 / 1 (what the client's rebuild path passes)."""
    rng = np.random.default_rng(31)
    n, V = 4000, 60
    toks = {r: rng.integers(0, V, size=int(rng.integers(1, 4))).astype(np.int32) for r in range(n) if r % 4}
    indptr, doc, tf, doclen = _host_csr([toks], [V], n)
    dl01 = (doclen > 0).astype(np.uint32)
    rows = sorted(toks)
    ip = np.zeros(len(rows) + 1, dtype=np.int64)
    ip[1:] = np.cumsum([toks[r].size for r in rows])
    qterms = [[int(t) for t in rng.integers(0, V, size=3)] for _ in range(8)]
    with _engine(dim=256) as a, _engine(dim=256) as b:
        X = synth.embeddings(n, 256, 32)
        a.append(X)
        b.append(X)
        a.bm25_build_fields(indptr, doc, tf, np.zeros(V, np.int32), dl01)
        b.text_omit_norms(0, True)
        b.text_add_rows(0, rows, ip, np.concatenate([toks[r] for r in rows]))
        b.text_commit([V], n)
        ia, _, _, _, na = a.text_export()
        ib, db, tb, _, nb = b.text_export()
        assert np.array_equal(ia, ib) and np.array_equal(db, doc) and np.array_equal(tb, tf) and np.array_equal(na, nb)
        assert [x.tolist() for x in a.text_stats()] == [x.tolist() for x in b.text_stats()]
        ra, sa = a.search_hybrid(None, qterms, 2.0, 0.0, 10)
        rb_, sb = b.search_hybrid(None, qterms, 2.0, 0.0, 10)
        assert np.array_equal(ra, rb_) and np.array_equal(sa.view(np.uint32), sb.view(np.uint32))
