"""CPU tests of the host side: DSL parsing of the exact bodies the reference emits, the text index, the client's
bookkeeping and error conventions, and the sharded search control flow over gloo (world size 2)."""
import os
import sys

import numpy as np
import pytest

from oracle import bm25, knn, synth
from rassengine_b200 import dsl, text
from rassengine_b200.sharded import shard_bounds


def test_parse_knn_bodies_as_the_reference_builds_them():
    vec = [0.1] * 1024
    body = {"size": 3, "query": {"knn": {"embedding": {"vector": vec, "k": 3}}}, "terminate_after": 3}   # main.py:1538-1542
    p = dsl.parse_search_body(body)
    assert (p.kind, p.size, p.knn_k, p.knn_boost, p.knn_field) == ("knn", 3, 3, 1.0, "embedding") and p.vector is vec
    wrapped = {"size": 3, "terminate_after": 3, "query": {"bool": {"must": [body["query"]],
               "filter": [{"term": {"patientId": "p-1"}}]}}}                                             # main.py:1543-1550
    p = dsl.parse_search_body(wrapped)
    assert p.kind == "knn" and p.filters == [("patientId", "p-1")]


def test_parse_hybrid_body_as_the_reference_builds_it():
    fields = ["unstructuredText^3", "patientName^3", "conditionNote^2", "observationValue"]
    kw = ["patientGender^3", "encounterStatus"]
    body = {"size": 10, "terminate_after": 10, "query": {"bool": {"should": [
        {"multi_match": {"query": "what is diabetes", "fields": fields, "type": "best_fields", "operator": "or",
                         "fuzziness": "AUTO", "boost": 1.5}},
        {"multi_match": {"query": "what is diabetes", "fields": kw, "type": "best_fields", "operator": "or", "boost": 1.0}},
        {"knn": {"embedding": {"vector": [0.0] * 8, "k": 10, "boost": 2.0}}}],
        "minimum_should_match": 1}}}                                                                      # main.py:1574-1605
    p = dsl.parse_search_body(body)
    assert p.kind == "hybrid" and p.knn_boost == 2.0 and p.knn_k == 10
    assert p.text[0].boost == 1.5 and dict(p.text[0].fields)["unstructuredText"] == 3.0
    assert dict(p.text[0].fields)["observationValue"] == 1.0 and p.text[1].boost == 1.0


@pytest.mark.parametrize("body", [
    {"size": 3, "query": {"match_phrase": {"unstructuredText": "x"}}},
    {"size": 0, "aggs": {"a": {"terms": {"field": "resourceType"}}}},
    {"size": 3, "query": {"bool": {"must": [{"multi_match": {"query": "x", "fields": ["a"], "type": "phrase_prefix"}}]}}},
    {"size": 3, "query": {"bool": {"filter": [{"range": {"a": {"gte": 1}}, "term": {"b": 2}}], "should": [
        {"knn": {"embedding": {"vector": [0.0], "k": 3}}}]}}},       # two query types in one filter node
    {"size": 3, "sort": [{"encounterStart": "desc"}], "query": {"match_all": {}}},
])
def test_unsupported_dsl_raises_not_implemented(body):
    # (a range / match_phrase bool.filter is NOT in this list any more: it stays in the plan as a host-evaluated row
    # filter, tests/test_hostquery_cpu.py)
    with pytest.raises(NotImplementedError):
        dsl.parse_search_body(body)


def test_text_field_postings_equal_oracle_index():
    indptr, doc, tf, doclen = synth.text_corpus(400, vocab=120, seed=3, median_len=30, max_len=90)
    texts = synth.docs_as_text(indptr, doc, tf, 400)
    f = text.TextField()
    for r, t in enumerate(texts):
        f.set_row(r, t)
    ip, d, t_, dl = f.postings(400)
    # vocabulary ids are assigned in first-seen order: compare per token through the dictionary
    ref = bm25.BM25Index(indptr, doc, tf, doclen)
    assert np.array_equal(dl, doclen)
    for term_id in range(0, 120, 7):
        tid = f.vocab.get(synth.token(term_id))
        lo, hi = indptr[term_id], indptr[term_id + 1]
        if tid is None:
            assert lo == hi
            continue
        assert np.array_equal(d[ip[tid]:ip[tid + 1]], doc[lo:hi]) and np.array_equal(t_[ip[tid]:ip[tid + 1]], tf[lo:hi])
    assert text.analyze("Hello, World-2 FOO_bar") == ["hello", "world", "2", "foo_bar"]
    import json
    for case in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "analyzer_cases.json")))["cases"]:
        assert text.analyze(case["text"]) == case["tokens"], case
    assert f.query_terms("t00003 nope")[1] == -1
    f.set_row(5, None)
    assert f.postings(400)[3][5] == 0 and ref.doclen[5] > 0


def test_client_without_gpu_follows_error_conventions():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rassengine_b200.client import B200Client, bulk, NotFoundError
    from rassengine_b200 import indexer as ix
    c = B200Client(hosts=[{"host": "x", "port": 1}])
    ix.ensure_index_exists(c, "idx", ix.index_body(8))
    assert c.indices.exists("idx") and c.count(index="idx")["count"] == 0
    with pytest.raises(NotFoundError):
        c.count(index="nope")
    # no device: the bulk call reports per-doc errors instead of raising (helpers.bulk style), searches give []
    ok, errors = bulk(c, [{"_op_type": "index", "_index": "idx", "_id": "a", "_source": {"embedding": [0.0] * 8}}])
    assert ok == 0 and len(errors) == 1 and "no CPU fallback" in errors[0]["index"]["error"]
    idxr = ix.B200Indexer(c, "idx")
    assert idxr.semantic_search(np.ones((1, 8), dtype=np.float32), k=3) == []
    assert not idxr.has_any_data()


def test_shard_bounds_cover_rows_once():
    for n, g in ((10_000_000, 8), (100_000_000, 4), (7, 3), (5, 8)):
        spans = [shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


# ---- world_size-2 gloo run of the sharded search control flow with a CPU test double for the engine ----------
_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from oracle import knn, synth
from rassengine_b200.sharded import ShardedIndex, shard_bounds

class OracleOps:                          # CPU test double for the device ops: answers from the oracle
    def __init__(self, X, base): self.X, self.base = X, base
    def search(self, q, k, packed, scores):
        r, key, s = knn.knn_exact(self.X, q.numpy(), k)
        packed[1].copy_(torch.from_numpy(r + self.base)); scores.copy_(torch.from_numpy(s))
        packed[0].copy_(torch.from_numpy(key).view(torch.int64))
    def merge(self, gathered, world, B, k, out_rows, out_scores, out_keys=None, stride=None, stream=None, raw=False):
        assert not raw
        g = gathered.view(world, -1)[:, :2 * B * k].reshape(world, 2, B, k)
        keys = g[:, 0].contiguous().view(torch.float64).numpy(); rows = g[:, 1].numpy()
        for b in range(B):
            kk = keys[:, b].reshape(-1); rr = rows[:, b].reshape(-1)
            order = np.lexsort((rr, -kk))[:k]
            out_rows[b] = torch.from_numpy(rr[order])
            out_scores[b] = torch.from_numpy((1 / (2 - kk[order])).astype(np.float32))
    uncertified = 0                       # what the engine would report for the batch in flight
    def search_async(self, q, k, packed, scores, slot, flag):
        self.search(q, k, packed, scores); flag[0] = self.uncertified
    def search_wait(self, slot):
        return self.uncertified == 0, {}

dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['PORT']}", rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
X = synth.embeddings(4000, 64, 1); Q = torch.from_numpy(synth.embeddings(5, 64, 2)); k = 7
lo, hi = shard_bounds(4000, 2, rank)
idx = ShardedIndex(dim=64, ops=OracleOps(X[lo:hi], lo))
assert idx.world == 2 and idx.rank == rank
rows, scores = idx.search_dev(Q, k)      # the product's control flow: local search -> ONE all_gather -> merge
want_rows, _, want_scores = knn.knn_exact(X, Q.numpy(), k)
assert np.array_equal(rows.numpy(), want_rows), (rank, rows, want_rows)
np.testing.assert_allclose(scores.numpy(), want_scores, rtol=1e-6)
assert idx.merge_launches == 1
# pipelined flavour: two batches in flight, results identical; a batch with an uncertified query on ONE rank is
# repeated through the blocking path on EVERY rank (they all read the gathered counts, no extra collective)
Q2 = torch.from_numpy(synth.embeddings(5, 64, 3))
want2 = knn.knn_exact(X, Q2.numpy(), k)[0]
t0 = idx.search_dev_async(Q, k); t1 = idx.search_dev_async(Q2, k)
try:
    idx.search_dev_async(Q, k); raise SystemExit("a third ticket must be refused")
except RuntimeError:
    pass
r0, _ = idx.wait(t0); r1, _ = idx.wait(t1)
assert np.array_equal(r0.numpy(), want_rows) and np.array_equal(r1.numpy(), want2)
before = idx.merge_launches
idx.ops.uncertified = 1 if rank == 1 else 0
r0, _ = idx.wait(idx.search_dev_async(Q, k))
idx.ops.uncertified = 0
assert np.array_equal(r0.numpy(), want_rows) and idx.merge_launches == before + 2      # async attempt + blocking repeat

# hybrid: global knn list first, then every rank fuses its own rows, then a second all-gather + merge
from oracle import bm25, fusion
ip, dc, tf, dl = synth.text_corpus(4000, vocab=300, seed=3, median_len=30, max_len=90)
full = bm25.BM25Index(ip, dc, tf, dl)
qterms = synth.text_queries(5, vocab=300, seed=4)

class HybridOps(OracleOps):
    def bm25_build(self, indptr, doc, tf_, doclen, doc_count, sum_ttf, df):
        self.stats = (doc_count, sum_ttf, np.asarray(df))
    def fuse(self, B, qt, w_text, knn_rows, knn_scores, w_knn, k, packed, scores, qweights=None, qflags=None):
        n_local = self.X.shape[0]
        for b in range(B):
            text = full.score(qt[b], boost=w_text)[self.base:self.base + n_local]
            loc = [(int(r) - self.base, float(s)) for r, s in zip(knn_rows[b].tolist(), knn_scores[b].tolist())
                   if self.base <= r < self.base + n_local]
            fused = text.astype(np.float64); matched = text > 0
            for r, s in loc:
                fused[r] += np.float64(np.float32(w_knn) * np.float32(s)); matched[r] = True
            final = fused.astype(np.float32); docs = np.flatnonzero(matched)
            order = docs[np.lexsort((docs, -final[docs].astype(np.float64)))[:k]]
            rr = np.full(k, -1, dtype=np.int64); kk = np.zeros(k); rr[:order.size] = order + self.base
            kk[:order.size] = final[order].astype(np.float64)
            packed[1][b] = torch.from_numpy(rr); packed[0][b] = torch.from_numpy(kk).view(torch.int64)
            scores[b] = torch.from_numpy(kk.astype(np.float32))
    def merge(self, gathered, world, B, k, out_rows, out_scores, out_keys=None, stride=None, stream=None, raw=False):
        if not raw:                       # the knn list: the base class's cosine merge
            return OracleOps.merge(self, gathered, world, B, k, out_rows, out_scores, out_keys, stride, stream)
        g = gathered.view(world, 2, B, k)
        keys = g[:, 0].contiguous().view(torch.float64).numpy(); rows = g[:, 1].numpy()
        for b in range(B):
            kk = keys[:, b].reshape(-1); rr = rows[:, b].reshape(-1)
            ok = rr >= 0; kk, rr = kk[ok], rr[ok]
            order = np.lexsort((rr, -kk))[:k]
            out_rows[b, :order.size] = torch.from_numpy(rr[order]); out_rows[b, order.size:] = -1
            if out_keys is not None:
                out_keys[b, :order.size] = torch.from_numpy(kk[order]); out_keys[b, order.size:] = 0
            out_scores[b, :order.size] = torch.from_numpy(kk[order].astype(np.float32))     # raw fused scores

hidx = ShardedIndex(dim=64, ops=HybridOps(X[lo:hi], lo))
term_of = np.repeat(np.arange(300), np.diff(ip)); mine = (dc >= lo) & (dc < hi)
lip = np.zeros(301, dtype=np.int64); np.add.at(lip, term_of[mine] + 1, 1); lip = np.cumsum(lip)
hidx.bm25_build(lip, (dc[mine] - lo).astype(np.int32), tf[mine], dl[lo:hi])
assert hidx.ops.stats[0] == full.doc_count and hidx.ops.stats[1] == full.sum_ttf       # summed over the ranks
assert np.array_equal(hidx.ops.stats[2], full.df)
hr, hs = hidx.search_hybrid_dev(Q, qterms, 4.5, 2.0, k)
for b in range(5):
    wr, ws = fusion.hybrid(full, qterms[b], want_rows[b], want_scores[b], 4.5, 2.0, k)
    assert hr[b, :len(wr)].tolist() == wr.tolist(), (rank, b)
    np.testing.assert_allclose(hs[b, :len(wr)].numpy(), ws, rtol=1e-6)
assert hidx.merge_launches == 2          # one merge for the global knn list, one for the fused lists
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_search_two_ranks_gloo(tmp_path):
    import socket
    import subprocess
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), REPO=root, OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


def test_hostquery_building_blocks_agree_with_the_oracle():
    """The host-side (N4) evaluator restates SmallFloat norms, the edit distance and date math on its own; they must
    agree with the oracle's independent restatements."""
    import datetime as dt
    from oracle import fuzzy, smallfloat
    from rassengine_b200 import hostquery as hq
    assert np.array_equal(hq.LENGTH_TABLE.astype(np.float32), smallfloat.LENGTH_TABLE)
    lens = np.array([0, 1, 23, 24, 25, 39, 40, 41, 47, 48, 511, 512, 100000], dtype=np.int64)
    assert np.array_equal(hq.norm_bytes(lens), smallfloat.encode_lengths(lens))
    for a, b in (("paitent", "patient"), ("chest", "chets"), ("ca", "abc"), ("smith", "smyth"), ("a", "abcd")):
        assert hq._osa(a, b, 10) == fuzzy.osa_distance(a, b)
    now = dt.datetime(2026, 3, 31, 12, 0, tzinfo=dt.timezone.utc)
    assert hq._parse_date("now", now) == now
    assert hq._parse_date("now-1y", now) == now.replace(year=2025)
    assert hq._parse_date("now-2M", now) == now.replace(month=1, day=28)
    assert hq._parse_date("now-3d", now) == now - dt.timedelta(days=3)
    assert hq._parse_date("2024-02-29", now) == dt.datetime(2024, 2, 29, tzinfo=dt.timezone.utc)
    assert hq._parse_date("2024-02-29T10:00:00Z", now).hour == 10
    assert hq._parse_date("not a date", now) is None


def test_text_index_multi_field_postings_equal_oracle_fields():
    """TextIndex (host side of rass_bm25_build_fields): one CSR over all analysed / keyword fields whose per-field slices
    carry the same postings, lengths and statistics as the oracle's per-field indices; fuzzy term lists agree too."""
    from oracle import fuzzy, multifield
    from rassengine_b200.text import TextIndex
    docs = [{"unstructuredText": "chest pain today chest"}, {"patientName": "John Smith", "patientGender": "male"},
            {"unstructuredText": "pain in the chest wall", "patientGender": "female"}, {},
            {"patientName": ["Jon", "Smyth Smith"], "conditionNote": "pain"}]
    types = {"unstructuredText": "text", "patientName": "text", "patientGender": "keyword", "conditionNote": "text",
             "neverUsed": "text"}
    ti = TextIndex(types)
    for r, d in enumerate(docs):
        ti.set_doc(r, d, fresh=True)
    indptr, doc, tf, term_field, doclen = ti.postings(len(docs))
    fields = multifield.build(docs, types)
    assert set(ti.order) == set(fields) and "neverUsed" not in ti.fields
    terms = ti.terms_in_id_order()
    for fid, name in enumerate(ti.order):
        f = fields[name]
        base, n = ti.base[name], len(ti.fields[name].vocab)
        assert (term_field[base:base + n] == fid).all()
        assert np.array_equal(doclen[fid], f.index.doclen)
        for t in range(n):
            ot = f.terms.index(terms[base + t])
            lo, hi = indptr[base + t], indptr[base + t + 1]
            olo, ohi = f.index.indptr[ot], f.index.indptr[ot + 1]
            assert np.array_equal(doc[lo:hi], f.index.doc[olo:ohi]) and np.array_equal(tf[lo:hi], f.index.tf[olo:ohi])
        assert ti.fields[name].doc_count == f.index.doc_count
    # overwrite: a field the document no longer carries is cleared
    ti.set_doc(1, {"patientGender": "other"})
    indptr2, doc2, *_ = ti.postings(len(docs))
    assert 1 not in doc2[indptr2[ti.base["patientName"]]:indptr2[ti.base["patientName"] + len(ti.fields["patientName"].vocab)]]
    # fuzzy rewrite on the host side (dictionary scan replaced by the oracle's distance) equals the oracle's term list
    name_f = ti.fields["unstructuredText"]
    ti.postings(len(docs))
    f = multifield.build(docs[:1] + [{}] + docs[2:], types)["unstructuredText"]        # document 1 lost nothing here

    def expand(tok, me):
        ids = [(i, fuzzy.osa_distance(tok, t)) for i, t in enumerate(name_f.terms_in_id_order())
               if abs(len(t) - len(tok)) <= me]
        ids = [(i, d) for i, d in ids if d <= me]
        return np.array([i for i, _ in ids], dtype=np.int64), np.array([d for _, d in ids], dtype=np.int64)

    ids, ws = name_f.fuzzy_weighted_terms("chets pian wall", 4.5, expand)
    oids, ows = fuzzy.weighted_terms(f.index, f.terms, ["chets", "pian", "wall"], 4.5)
    assert [name_f.terms_in_id_order()[i] for i in ids] == [f.terms[i] for i in oids]
    np.testing.assert_array_equal(np.asarray(ws, dtype=np.float32), ows)


def test_dsl_text_only_and_fuzziness_shapes():
    must = {"size": 3, "query": {"bool": {"must": [{"multi_match": {"query": "x y", "fields": ["conditionNote^3", "unstructuredText^2"],
                                                                 "type": "best_fields", "operator": "or", "fuzziness": "AUTO"}}],
                                          "filter": [{"term": {"patientId": "p1"}}]}}, "terminate_after": 3}
    p = dsl.parse_search_body(must)
    assert p.kind == "hybrid" and p.vector is None and p.text[0].fuzziness == "AUTO" and p.filters == [("patientId", "p1")]
    assert dict(p.text[0].fields) == {"conditionNote": 3.0, "unstructuredText": 2.0}
    p = dsl.parse_search_body({"query": {"bool": {"should": [{"multi_match": {"query": "x", "fields": ["a"]}}]}}})
    assert p.kind == "hybrid" and p.text[0].fuzziness is None and p.size == 10
    for bad in ({"query": {"bool": {"should": [{"multi_match": {"query": "x", "fields": ["a"], "fuzziness": 2}}]}}},
                {"query": {"bool": {"must": [{"multi_match": {"query": "x", "fields": ["a"]}}, {"range": {"d": {"gte": 1}}}]}}},
                {"_source": ["patientId"], "query": {"match_all": {}}}):
        with pytest.raises(NotImplementedError):
            dsl.parse_search_body(bad)


def test_store_fhir_docs_entry_point_keeps_the_reference_shape():
    """store_fhir_docs_in_opensearch(structured_docs, unstructured_docs, client, index_name): structured documents in
    one bulk call, chunks normalised like app/main.py:1250-1251 and flushed in BATCH_SIZE bulks with `_id = doc_id` and
    `_routing = patientId`; async, returns None, swallows errors.  A recording client stands in for the engine."""
    import asyncio
    import inspect
    from rassengine_b200 import indexer as ix

    class Indices:
        def __init__(self):
            self.created = []

        def exists(self, name):
            return name in self.created

        def create(self, index, body=None):
            self.created.append(index)

    class Recorder:
        def __init__(self, fail=False):
            self.indices, self.calls, self.fail = Indices(), [], fail

        def bulk_actions(self, actions):
            if self.fail:
                raise RuntimeError("boom")
            self.calls.append(list(actions))
            return len(actions), []

    assert inspect.iscoroutinefunction(ix.store_fhir_docs_in_opensearch)
    assert list(inspect.signature(ix.store_fhir_docs_in_opensearch).parameters)[:4] == \
        ["structured_docs", "unstructured_docs", "client", "index_name"]
    structured = [{"doc_id": f"s{i}", "doc_type": "structured", "patientId": "pat-1"} for i in range(3)]
    n = ix.BATCH_SIZE + 5
    chunks = [{"doc_id": f"c{i}", "doc_type": "unstructured", "patientId": "pat-2", "unstructuredText": f"text {i}"}
              for i in range(n)]
    raw = np.random.default_rng(3).standard_normal((n, 8)).astype(np.float32) * 7
    seen = {}

    async def embed(texts, batch_size):
        seen["texts"], seen["batch_size"] = list(texts), batch_size
        return raw

    c = Recorder()
    assert asyncio.run(ix.store_fhir_docs_in_opensearch(structured, chunks, c, "idx", embed=embed)) is None
    assert c.indices.created == ["idx"]
    assert seen == {"texts": [d["unstructuredText"] for d in chunks], "batch_size": ix.BATCH_SIZE}
    assert [len(b) for b in c.calls] == [3, ix.BATCH_SIZE, 5]
    first = c.calls[0][0]
    assert first == {"_op_type": "index", "_index": "idx", "_id": "s0", "_source": structured[0], "_routing": "pat-1"}
    want = raw / (np.linalg.norm(raw, axis=1, keepdims=True) + 1e-9)
    got = np.stack([a["_source"]["embedding"] for b in c.calls[1:] for a in b])
    assert got.dtype == np.float32 and np.array_equal(got, want.astype(np.float32))
    assert [a["_id"] for b in c.calls[1:] for a in b] == [d["doc_id"] for d in chunks]
    assert "embedding" not in chunks[0]                  # the caller's dicts are left alone
    # plain (non-async) embedders work too; failures are logged, never raised (app/main.py:1239-1240, 1275-1276)
    c2 = Recorder()
    asyncio.run(ix.store_fhir_docs_in_opensearch([], chunks[:2], c2, "idx", embed=lambda t, batch_size: raw[:2]))
    assert [len(b) for b in c2.calls] == [2]
    asyncio.run(ix.store_fhir_docs_in_opensearch(structured, chunks, Recorder(fail=True), "idx", embed=embed))
    asyncio.run(ix.store_fhir_docs_in_opensearch(structured, chunks, c2, "idx"))       # no embedder: logged, structured stored
    assert len(c2.calls) == 2 and len(c2.calls[1]) == 3
    asyncio.run(ix.store_fhir_docs_in_opensearch(structured, chunks, None, "idx"))


def test_text_field_pending_stream_is_sorted_deduplicated_and_incremental():
    """What TextIndex.sync_device hands to rass_text_add_rows: rows ascending and distinct, the LAST write to a row
    wins, a row that lost the field travels as an empty token list; and the dictionary blob of the fuzzy scan grows by
    the new terms only, keyword terms as empty strings."""
    from rassengine_b200.text import TextField, TextIndex
    f = TextField()
    f.set_row_tokens(0, ["a", "b", "a"])
    f.set_row_tokens(2, ["c"])
    rows, indptr, terms = f.take_pending()
    assert rows.tolist() == [0, 2] and indptr.tolist() == [0, 3, 4] and terms.tolist() == [0, 1, 0, 2]
    assert f.take_pending()[0].size == 0                                   # nothing new
    f.set_row_tokens(5, ["d"])
    f.set_row_tokens(2, ["a", "e"])                                        # rewrite of an older row
    f.set_row_tokens(5, ["b"])                                             # written twice before the sync
    f.set_row_tokens(0, [])                                                # loses the field
    f.set_row_tokens(7, [])                                                # never had it: not a change
    rows, indptr, terms = f.take_pending()
    assert rows.tolist() == [0, 2, 5] and indptr.tolist() == [0, 0, 2, 3] and terms.tolist() == [0, 4, 1]
    assert sorted(f.row_terms) == [2, 5]
    blob, lens = f.encoded_terms()
    assert bytes(blob) == b"abcde" and lens.tolist() == [1, 1, 1, 1, 1]
    f.set_row_tokens(9, ["été", "a"])
    blob, lens = f.encoded_terms()
    assert bytes(blob) == "abcdeété".encode() and lens.tolist() == [1, 1, 1, 1, 1, 5]
    ti = TextIndex({"note": "text", "code": "keyword"})
    ti.set_doc(0, {"note": "Chest pain, chest", "code": "I20.9"}, fresh=True)
    ti.set_doc(1, {"code": ["I20.9", "R07.4"]}, fresh=True)
    blob, off = ti.vocab_blob()
    assert ti.order == ["note", "code"] and blob == b"chestpain" and off.tolist() == [0, 5, 9, 9, 9]
    assert ti.terms_in_id_order() == ["chest", "pain", "I20.9", "R07.4"]
