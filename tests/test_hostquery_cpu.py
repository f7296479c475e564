"""SURVEY.md 8f N4 without a GPU: the query shapes the client evaluates host-side (phrase / phrase_prefix, range with date
math, sort, collapse, `_source` filtering, terms aggregations, the patient-name lookup) need no kernel, so the same
documents, bodies and expectations as tests/test_gpu_hostquery.py run here against an index whose rows were registered
without an engine (the documents of this fixture carry no vectors).  Scores against oracle/multifield.py."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import multifield

_spec_ = importlib.util.spec_from_file_location(
    "gpu_hostquery_cases", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_hostquery.py"))
cases = importlib.util.module_from_spec(_spec_)
_spec_.loader.exec_module(cases)
DOCS, _spec, _ranked = cases.DOCS, cases._spec, cases._ranked


@pytest.fixture(scope="module")
def setup():
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    client = B200Client()
    name = ix.get_index_name("n4cpu")
    ix.ensure_index_exists(client, name, ix.index_body(16))
    idx = client._get(name)
    with idx.lock:                       # what _index_batch_locked does after the device append, minus the device
        for row, d in enumerate(DOCS):
            idx.sources.append(dict(d))
            idx.has_vec.append(False)
            idx.ids.append(d["doc_id"])
            idx.row_of[d["doc_id"]] = row
            idx.text.set_doc(row, d, fresh=True)
            idx._kw_add(row, d)
        idx.n_docs = len(DOCS)
    assert idx.engine is None
    types = {f.split("^")[0]: "text" for f in ix.TEXT_FIELDS}
    types.update({f.split("^")[0]: "keyword" for f in ix.KEYWORD_FIELDS})
    types["patientId"] = "keyword"
    yield client, name, ix.B200Indexer(client, name), multifield.build(DOCS, types)
    client.close()


def _phrase_clause(fields, query, specs, cb, prefix=False):
    best = np.zeros(len(DOCS), dtype=np.float32)
    for fname, fb in specs:
        f = fields.get(fname)
        if f is None:
            continue
        toks = multifield.field_tokens(DOCS, fname, f.kind)
        best = np.maximum(best, multifield.phrase_score(f, toks, query, np.float32(np.float32(cb) * np.float32(fb)),
                                                        prefix=prefix))
    return best


def test_count_and_match_all_without_engine(setup):
    client, name, idxr, _ = setup
    assert client.count(index=name)["count"] == len(DOCS) and idxr.has_any_data()
    resp = client.search(index=name, body={"size": 3, "sort": [{"doc_id": {"order": "asc"}}], "query": {"match_all": {}}})
    assert [h["_id"] for h in resp["hits"]["hits"]] == ["c1", "c2", "c3"] and resp["hits"]["total"]["value"] == len(DOCS)
    # keyword sort, descending, documents without the field last, `from` paging
    resp = client.search(index=name, body={"size": 3, "from": 1, "sort": [{"patientGender": "desc"}, {"doc_id": "asc"}],
                                           "query": {"match_all": {}}})
    assert [h["_id"] for h in resp["hits"]["hits"]] == ["p1", "p2", "p3"]          # other(p4) | male p1 p2 | female p3
    assert all(h["_score"] is None for h in resp["hits"]["hits"])


def test_phrase_family_scores_match_oracle(setup):
    client, name, idxr, fields = setup
    from rassengine_b200 import indexer as ix
    total = _phrase_clause(fields, "chest pain", _spec(ix.TEXT_FIELDS), 2.0).astype(np.float64) + \
        _phrase_clause(fields, "chest pain", _spec(ix.KEYWORD_FIELDS), 1.0).astype(np.float64)
    want = _ranked(total.astype(np.float32))[:10]
    hits = idxr.exact_match_search("chest pain", k=10)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) >= 4
    np.testing.assert_allclose([h[1] for h in hits], total.astype(np.float32)[want], rtol=1e-6)
    assert [h[0]["doc_id"] for h in idxr.exact_match_search("pain in chest", k=10)] == ["c2"]
    assert idxr.exact_match_search("pain chest", k=10) == []          # order matters
    s = _phrase_clause(fields, "john smith", _spec(["patientName^4", "patientId^4", "patientGender^3", "patientTelecom^3",
                                                    "practitionerName^3", "organizationName^3"]), 1.0)
    hits = idxr.entity_specific_search("john smith", k=5)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in _ranked(s)] == ["p1"]
    assert idxr.entity_specific_search("male", k=5, patient_id="pat-2")[0][0]["doc_id"] == "p2"
    s = _phrase_clause(fields, "chest pa", _spec(ix.B200Indexer.STRUCTURED_FIELDS), 1.0, prefix=True)
    hits = idxr.structured_search("chest pa", k=10)
    want = _ranked(s)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) == 3
    np.testing.assert_allclose([h[1] for h in hits], s[want], rtol=1e-6)


def test_fuzzy_best_fields_next_to_aggregations(setup):
    """comparison_search carries `aggs`, so its fuzzy multi_match is evaluated host-side; without an engine the
    dictionary scan is the host mirror of fuzzy_scan_kernel."""
    client, name, idxr, fields = setup
    cmp_fields = ["conditionCodeText^2", "observationValue", "observationUnit", "medRequestMedicationDisplay",
                  "procedureCodeText", "allergyCodeText"]
    s = multifield.clause_score(fields, "diabetis chest", _spec(cmp_fields), 1.0, True, len(DOCS))
    want = _ranked(s)
    hits = idxr.comparison_search("diabetis chest", k=10)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) == 4
    np.testing.assert_allclose([h[1] for h in hits], s[want], rtol=2e-6)
    hits = idxr.comparison_search("diabetis chest", k=10, patient_id="pat-1")
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want if DOCS[r]["patientId"] == "pat-1"]


def test_patient_name_lookup(setup):
    client, name, *_ = setup
    from rassengine_b200.indexer import resolve_patient_ids
    ids = resolve_patient_ids(client, name, "John Smith", top_k=5)
    assert ids[0] == "pat-1" and set(ids) == {"pat-1", "pat-2", "pat-4"}
    assert resolve_patient_ids(client, name, "Maria Garcia", top_k=5) == ["pat-3"]
    assert resolve_patient_ids(client, name, "Nobody Here", top_k=5) == []
    hits = client.search(index=name, body={"size": 3, "_source": ["patientId"], "collapse": {"field": "patientId"},
                                           "query": {"match": {"patientName": {"query": "smith"}}}})["hits"]["hits"]
    assert all(set(h["_source"]) == {"patientId"} for h in hits) and len({h["_source"]["patientId"] for h in hits}) == 3


def test_aggregations_collapse_sort_and_ranges(setup):
    client, name, idxr, _ = setup
    aggs = idxr.aggregate_search("anything")
    assert [(b["key"], b["doc_count"]) for b in aggs["by_patient"]["buckets"]] == \
        [("pat-1", 4), ("pat-3", 3), ("pat-2", 2), ("pat-4", 1)]
    assert [(b["key"], b["doc_count"]) for b in aggs["by_condition"]["buckets"]] == \
        [("Chest pain", 2), ("Chronic chest pain syndrome", 1), ("Type 2 diabetes mellitus", 1)]
    assert aggs["by_resource"]["buckets"] == []                      # resourceType.keyword is not in the mapping
    aggs = idxr.aggregate_search("anything", patient_id="pat-1")
    assert [(b["key"], b["doc_count"]) for b in aggs["by_patient"]["buckets"]] == [("pat-1", 4)]
    hits = idxr.document_fetch_search("ignored", k=5, patient_id="pat-1")
    assert [(h[0]["doc_id"], h[1]) for h in hits] == [("p1", 0.0)]
    assert idxr.document_fetch_search("ignored", k=5) == []
    body = {"size": 5, "sort": [{"conditionOnsetDateTime": {"order": "desc"}}],
            "query": {"bool": {"must": [
                {"multi_match": {"query": "chest pain diabetes", "fields": idxr.text_fields + idxr.keyword_fields,
                                 "type": "best_fields", "operator": "or"}},
                {"bool": {"should": [{"range": {f: {"gte": "now-1y", "lte": "now"}}} for f in idxr.DATE_FIELDS],
                          "minimum_should_match": 1}}]}}}
    hits = client.search(index=name, body=body)["hits"]["hits"]
    assert [h["_id"] for h in hits] == ["c4", "c1", "c2"] and all(h["_score"] is None for h in hits)
    assert idxr.temporal_search("chest pain diabetes", k=5) == []     # the reference's float(None) -> caught -> []
    # must_not + terms + exists + numeric-looking term values
    body = {"size": 10, "sort": [{"doc_id": {"order": "asc"}}],
            "query": {"bool": {"filter": [{"terms": {"patientId": ["pat-1", "pat-3"]}}, {"exists": {"field": "conditionCodeText"}}],
                               "must_not": [{"term": {"conditionClinicalStatus": "resolved"}}]}}}
    assert [h["_id"] for h in client.search(index=name, body=body)["hits"]["hits"]] == ["c1", "c2", "c4"]
    with pytest.raises(NotImplementedError):
        client.search(index=name, body={"query": {"more_like_this": {"fields": ["x"], "like": "y"}}})


# ---- random bool trees of keyword filters against plain python set logic -------------------------------------------
from hypothesis import given, settings, strategies as st   # noqa: E402

_FILTER_FIELDS = {"patientId": ["pat-1", "pat-2", "pat-3", "pat-4", "pat-9"],
                  "doc_type": ["structured", "unstructured", "other"],
                  "resourceType": ["Patient", "Condition", "Practitioner"],
                  "conditionClinicalStatus": ["active", "resolved"]}


def _leaf():
    def term(f):
        return st.sampled_from(_FILTER_FIELDS[f]).map(lambda v: {"term": {f: v}})

    def terms(f):
        return st.lists(st.sampled_from(_FILTER_FIELDS[f]), min_size=1, max_size=3).map(lambda vs: {"terms": {f: vs}})
    fields = st.sampled_from(sorted(_FILTER_FIELDS))
    return st.one_of(fields.flatmap(term), fields.flatmap(terms), fields.map(lambda f: {"exists": {"field": f}}))


def _tree():
    def bool_of(children):
        return st.fixed_dictionaries({}, optional={
            "must": st.lists(children, min_size=1, max_size=2), "filter": st.lists(children, min_size=1, max_size=2),
            "must_not": st.lists(children, min_size=1, max_size=2), "should": st.lists(children, min_size=1, max_size=3),
        }).filter(lambda d: d).map(lambda d: {"bool": d})
    return st.recursive(_leaf(), bool_of, max_leaves=8)


def _naive(node, doc):
    (kind, body), = node.items()
    if kind == "term":
        (f, v), = body.items()
        return doc.get(f) == v
    if kind == "terms":
        (f, vs), = body.items()
        return doc.get(f) in vs
    if kind == "exists":
        return doc.get(body["field"]) is not None
    ok = all(_naive(c, doc) for c in body.get("must", []) + body.get("filter", []))
    ok = ok and not any(_naive(c, doc) for c in body.get("must_not", []))
    if "should" in body:
        need = 0 if ("must" in body or "filter" in body) else 1          # Lucene's default minimum_should_match
        ok = ok and sum(_naive(c, doc) for c in body["should"]) >= need
    return ok


@pytest.mark.filterwarnings("ignore::hypothesis.errors.HypothesisWarning")
@settings(max_examples=150, deadline=None, derandomize=True)
@given(_tree())
def test_random_bool_trees_of_keyword_filters_match_set_logic(setup, query):
    client, name, *_ = setup
    body = {"size": 50, "sort": [{"doc_id": "asc"}], "query": query}
    got = [h["_id"] for h in client.search(index=name, body=body)["hits"]["hits"]]
    want = sorted(d["doc_id"] for d in DOCS if _naive(query, d))
    assert got == want, query


def test_ner_filter_clause_becomes_a_row_list_and_the_plan_stays_on_the_gpu_path(setup):
    """ner_preprocess (app/main.py:2589-2609) hands every search {"bool": {"must": [match_phrase | range ...]}} as
    filter_clause.  The DSL parser keeps such a clause in the plan (so knn / hybrid still run on the device) and the
    client turns that sub-tree alone into a row list."""
    from rassengine_b200.dsl import parse_search_body
    client, name, idxr, _ = setup
    ner = {"bool": {"must": [{"match_phrase": {"conditionCodeText": "chest pain"}},
                             {"range": {"conditionOnsetDateTime": {"gte": cases._days_ago(400), "lte": cases._days_ago(1)}}}]}}
    body = {"size": 3, "query": {"bool": {"should": [
        {"multi_match": {"query": "pain", "fields": ["conditionNote^2"], "type": "best_fields", "operator": "or",
                         "fuzziness": "AUTO", "boost": 1.5}},
        {"knn": {"embedding": {"vector": [0.0] * 16, "k": 3, "boost": 2.0}}}],
        "minimum_should_match": 1, "filter": [ner, {"term": {"patientId": "pat-1"}}]}}, "terminate_after": 3}
    plan = parse_search_body(body)
    assert plan.kind == "hybrid" and plan.filters == [("patientId", "pat-1")] and plan.host_filters == [ner]
    idx = client._get(name)
    rows = idx._filter_rows(plan.filters, plan.host_filters)
    assert [DOCS[r]["doc_id"] for r in rows.tolist()] == ["c1", "c2"]        # c3: other patient and too old
    assert idx._passes(4, plan) and not idx._passes(6, plan) and not idx._passes(0, plan)
    assert idx._filter_rows([], [ner]).tolist() == [4, 5]                        # cached per index version
    # an unparseable date (free text from the NER, app/main.py:2596-2601) matches nothing instead of raising
    bad = {"bool": {"must": [{"range": {"conditionOnsetDateTime": {"gte": "last tuesday", "lte": "last tuesday"}}}]}}
    assert idx._filter_rows([], [bad]).size == 0


def test_phrase_longer_than_the_field_value_is_no_match_not_an_error(setup):
    """A repeated-token phrase against a shorter field value: 'pain pain pain pain' vs 'Chest pain' must score 0
    (hostquery._phrase used to raise a broadcast error that the indexer's `except` turned into [])."""
    client, name, idxr, _ = setup
    body = {"size": 5, "query": {"multi_match": {"query": "pain pain pain pain", "fields": ["conditionCodeText"],
                                                 "type": "phrase"}}}
    assert client.search(index=name, body=body)["hits"]["hits"] == []
    body["query"]["multi_match"]["query"] = "chest pain"
    assert {h["_id"] for h in client.search(index=name, body=body)["hits"]["hits"]} == {"c1", "c2", "c3"}


def test_legacy_bm25_boost_is_the_same_scorer_with_the_boost_times_one_plus_k1():
    """B200Client(bm25_legacy_boost=True): Lucene-misc's LegacyBM25Similarity (what Elasticsearch 7 kept its old scores
    with) multiplies every query boost by float32(1 + k1) = 2.2f before the BM25 scorer is built; everything after is
    the same arithmetic.  So the phrase and the fuzzy best_fields searches of such a client equal the oracle called with
    that boost -- and the ranking of a pure text query does not move (oracle/SEMANTICS.md)."""
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    f22 = np.float32(1.0) + np.float32(1.2)
    client = B200Client(bm25_legacy_boost=True)
    name = ix.get_index_name("legacy")
    ix.ensure_index_exists(client, name, ix.index_body(16))
    idx = client._get(name)
    with idx.lock:
        for row, d in enumerate(DOCS):
            idx.sources.append(dict(d))
            idx.has_vec.append(False)
            idx.ids.append(d["doc_id"])
            idx.row_of[d["doc_id"]] = row
            idx.text.set_doc(row, d, fresh=True)
            idx._kw_add(row, d)
        idx.n_docs = len(DOCS)
    types = {f.split("^")[0]: "text" for f in ix.TEXT_FIELDS}
    types.update({f.split("^")[0]: "keyword" for f in ix.KEYWORD_FIELDS})
    types["patientId"] = "keyword"
    fields = multifield.build(DOCS, types)
    idxr = ix.B200Indexer(client, name)

    def phrase(specs, cb):
        best = np.zeros(len(DOCS), dtype=np.float32)
        for fname, fb in specs:
            f = fields.get(fname)
            if f is not None:
                toks = multifield.field_tokens(DOCS, fname, f.kind)
                bo = np.float32(np.float32(np.float32(cb) * np.float32(fb)) * f22)
                best = np.maximum(best, multifield.phrase_score(f, toks, "chest pain", bo, prefix=False))
        return best

    total = phrase(_spec(ix.TEXT_FIELDS), 2.0).astype(np.float64) + phrase(_spec(ix.KEYWORD_FIELDS), 1.0).astype(np.float64)
    want = _ranked(total.astype(np.float32))[:10]
    hits = idxr.exact_match_search("chest pain", k=10)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) >= 4
    np.testing.assert_allclose([h[1] for h in hits], total.astype(np.float32)[want], rtol=1e-6)
    # against the default client: same order, scores 2.2 times larger (up to the rounding of the boost product)
    base = B200Client()
    ix.ensure_index_exists(base, name, ix.index_body(16))
    bidx = base._get(name)
    with bidx.lock:
        for row, d in enumerate(DOCS):
            bidx.sources.append(dict(d))
            bidx.has_vec.append(False)
            bidx.ids.append(d["doc_id"])
            bidx.row_of[d["doc_id"]] = row
            bidx.text.set_doc(row, d, fresh=True)
            bidx._kw_add(row, d)
        bidx.n_docs = len(DOCS)
    ref = ix.B200Indexer(base, name).comparison_search("diabetis chest", k=10)
    got = idxr.comparison_search("diabetis chest", k=10)
    assert [h[0]["doc_id"] for h in got] == [h[0]["doc_id"] for h in ref] and len(got) == 4
    np.testing.assert_allclose([h[1] for h in got], [float(f22) * h[1] for h in ref], rtol=1e-6)
    client.close()
    base.close()
