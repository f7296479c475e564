"""the remaining OpenSearchIndexer methods and the patient-name lookup against the same client.
They are keyword bookkeeping evaluated host-side (rassengine_b200/hostquery.py); scores are checked against the
oracle restatements (oracle/multifield.py), structure (collapse, sort, aggregations, filters) against literals."""
import datetime as dt

import numpy as np
import pytest

from oracle import multifield

pytestmark = pytest.mark.gpu

TODAY = dt.datetime.now(dt.timezone.utc)


def _days_ago(n):
    return (TODAY - dt.timedelta(days=n)).strftime("%Y-%m-%d")


DOCS = [
    {"doc_id": "p1", "doc_type": "structured", "resourceType": "Patient", "patientId": "pat-1",
     "patientName": "John Smith", "patientGender": "male", "patientTelecom": "555 0101"},
    {"doc_id": "p2", "doc_type": "structured", "resourceType": "Patient", "patientId": "pat-2",
     "patientName": "Jon Smyth", "patientGender": "male"},
    {"doc_id": "p3", "doc_type": "structured", "resourceType": "Patient", "patientId": "pat-3",
     "patientName": "Maria Garcia Smith", "patientGender": "female"},
    {"doc_id": "p4", "doc_type": "structured", "resourceType": "Patient", "patientId": "pat-4",
     "patientName": "Smith John", "patientGender": "other"},
    {"doc_id": "c1", "doc_type": "structured", "resourceType": "Condition", "patientId": "pat-1",
     "conditionCodeText": "Chest pain", "conditionNote": "patient reports chest pain since monday",
     "conditionOnsetDateTime": _days_ago(30), "conditionClinicalStatus": "active"},
    {"doc_id": "c2", "doc_type": "structured", "resourceType": "Condition", "patientId": "pat-1",
     "conditionCodeText": "Chronic chest pain syndrome", "conditionNote": "pain in chest after exercise",
     "conditionOnsetDateTime": _days_ago(200), "conditionClinicalStatus": "active"},
    {"doc_id": "c3", "doc_type": "structured", "resourceType": "Condition", "patientId": "pat-2",
     "conditionCodeText": "Chest pain", "conditionNote": "chest wall pain",
     "conditionOnsetDateTime": _days_ago(900), "conditionClinicalStatus": "resolved"},
    {"doc_id": "c4", "doc_type": "structured", "resourceType": "Condition", "patientId": "pat-3",
     "conditionCodeText": "Type 2 diabetes mellitus", "conditionNote": "diet controlled",
     "conditionOnsetDateTime": _days_ago(10), "conditionClinicalStatus": "active"},
    {"doc_id": "n1", "doc_type": "unstructured", "patientId": "pat-1",
     "unstructuredText": "john smith came in with chest pain and shortness of breath"},
    {"doc_id": "n2", "doc_type": "unstructured", "patientId": "pat-3",
     "unstructuredText": "maria reports her diabetes is well controlled no chest complaints no pain"},
]
DIM = 16


@pytest.fixture(scope="module")
def setup():
    from rassengine_b200.client import B200Client
    from rassengine_b200 import indexer as ix
    client = B200Client()
    name = ix.get_index_name("n4")
    ix.ensure_index_exists(client, name, ix.index_body(DIM))
    structured = [d for d in DOCS if d["doc_type"] == "structured"]
    chunks = [d for d in DOCS if d["doc_type"] == "unstructured"]
    emb = np.random.default_rng(5).standard_normal((len(chunks), DIM)).astype(np.float32)
    assert ix.store_structured(client, name, structured) == (len(structured), [])
    assert ix.store_chunks(client, name, chunks, emb) == (len(chunks), [])
    types = {f.split("^")[0]: "text" for f in ix.TEXT_FIELDS}
    types.update({f.split("^")[0]: "keyword" for f in ix.KEYWORD_FIELDS})
    types["patientId"] = "keyword"
    yield client, name, ix.B200Indexer(client, name), multifield.build(DOCS, types), types, emb
    client.close()


def _spec(lst):
    return [(f.split("^")[0], float(f.split("^")[1]) if "^" in f else 1.0) for f in lst]


def _ranked(score):
    rows = np.flatnonzero(score > 0)
    return rows[np.lexsort((rows, -score[rows].astype(np.float64)))]


def test_patient_name_lookup(setup):
    client, name, *_ = setup
    from rassengine_b200.indexer import resolve_patient_ids
    # exact keyword + phrase + fuzzy all-terms: John Smith first; Jon Smyth is one edit per token away; "Smith John"
    # and "Maria Garcia Smith" only satisfy the fuzzy all-terms / partial clauses
    ids = resolve_patient_ids(client, name, "John Smith", top_k=5)
    assert ids[0] == "pat-1" and set(ids) == {"pat-1", "pat-2", "pat-4"}
    assert resolve_patient_ids(client, name, "Maria Garcia", top_k=5) == ["pat-3"]
    assert resolve_patient_ids(client, name, "Nobody Here", top_k=5) == []
    hits = client.search(index=name, body={"size": 3, "_source": ["patientId"], "collapse": {"field": "patientId"},
                                           "query": {"match": {"patientName": {"query": "smith"}}}})["hits"]["hits"]
    assert all(set(h["_source"]) == {"patientId"} for h in hits) and len({h["_source"]["patientId"] for h in hits}) == 3


def test_phrase_family_scores_match_oracle(setup):
    client, name, idxr, fields, types, _ = setup
    from rassengine_b200 import indexer as ix
    n = len(DOCS)

    def phrase_clause(query, specs, cb, prefix=False):
        best = np.zeros(n, dtype=np.float32)
        for fname, fb in specs:
            f = fields.get(fname)
            if f is None:
                continue
            toks = multifield.field_tokens(DOCS, fname, f.kind)
            best = np.maximum(best, multifield.phrase_score(f, toks, query, np.float32(np.float32(cb) * np.float32(fb)),
                                                            prefix=prefix))
        return best

    # exact_match_search: phrase over text fields (boost 2) + phrase over keyword fields
    total = phrase_clause("chest pain", _spec(ix.TEXT_FIELDS), 2.0).astype(np.float64) + \
        phrase_clause("chest pain", _spec(ix.KEYWORD_FIELDS), 1.0).astype(np.float64)
    want = _ranked(total.astype(np.float32))[:10]
    hits = idxr.exact_match_search("chest pain", k=10)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want]
    np.testing.assert_allclose([h[1] for h in hits], total.astype(np.float32)[want], rtol=1e-6)
    assert "c2" in [h[0]["doc_id"] for h in hits]                     # "Chronic chest pain syndrome" has the phrase
    c2_note_only = idxr.exact_match_search("pain in chest", k=10)
    assert [h[0]["doc_id"] for h in c2_note_only] == ["c2"]
    assert idxr.exact_match_search("pain chest", k=10) == []          # order matters
    # entity_specific_search: phrase over the entity fields
    s = phrase_clause("john smith", _spec(["patientName^4", "patientId^4", "patientGender^3", "patientTelecom^3",
                                           "practitionerName^3", "organizationName^3"]), 1.0)
    hits = idxr.entity_specific_search("john smith", k=5)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in _ranked(s)] == ["p1"]
    assert idxr.entity_specific_search("male", k=5, patient_id="pat-2")[0][0]["doc_id"] == "p2"
    # structured_search: phrase_prefix ("chest pa" -> chest + pa*), structured documents only
    s = phrase_clause("chest pa", _spec(ix.B200Indexer.STRUCTURED_FIELDS), 1.0, prefix=True)
    hits = idxr.structured_search("chest pa", k=10)
    want = _ranked(s)
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) == 3
    np.testing.assert_allclose([h[1] for h in hits], s[want], rtol=1e-6)


def test_best_fields_methods_gpu_and_host_agree_with_oracle(setup):
    client, name, idxr, fields, types, _ = setup
    n = len(DOCS)
    note_fields = ["conditionNote^3", "observationNote^3", "encounterNote^3", "medRequestNote^3", "procedureNote^3",
                   "allergyNote^3", "unstructuredText^2"]
    s = multifield.clause_score(fields, "chest pian", _spec(note_fields), 1.0, True, n)      # "pian" -> pain (1 edit)
    want = _ranked(s)
    hits = idxr.explanatory_search("chest pian", k=10)                # one scoring clause under must: the GPU text path
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) >= 4
    np.testing.assert_allclose([h[1] for h in hits], s[want], rtol=2e-6)
    cmp_fields = ["conditionCodeText^2", "observationValue", "observationUnit", "medRequestMedicationDisplay",
                  "procedureCodeText", "allergyCodeText"]
    s = multifield.clause_score(fields, "diabetis chest", _spec(cmp_fields), 1.0, True, n)
    want = _ranked(s)
    hits = idxr.comparison_search("diabetis chest", k=10)             # aggs next to the query: evaluated host-side
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want] and len(hits) == 4
    np.testing.assert_allclose([h[1] for h in hits], s[want], rtol=2e-6)
    hits = idxr.comparison_search("diabetis chest", k=10, patient_id="pat-1")
    assert [h[0]["doc_id"] for h in hits] == [DOCS[r]["doc_id"] for r in want if DOCS[r]["patientId"] == "pat-1"]


def test_aggregations_collapse_sort_and_filters(setup):
    client, name, idxr, *_ = setup
    aggs = idxr.aggregate_search("anything")
    assert [(b["key"], b["doc_count"]) for b in aggs["by_patient"]["buckets"]] == \
        [("pat-1", 4), ("pat-3", 3), ("pat-2", 2), ("pat-4", 1)]
    assert [(b["key"], b["doc_count"]) for b in aggs["by_condition"]["buckets"]] == \
        [("Chest pain", 2), ("Chronic chest pain syndrome", 1), ("Type 2 diabetes mellitus", 1)]
    assert aggs["by_resource"]["buckets"] == []                      # resourceType.keyword is not in the mapping
    aggs = idxr.aggregate_search("anything", patient_id="pat-1")
    assert [(b["key"], b["doc_count"]) for b in aggs["by_patient"]["buckets"]] == [("pat-1", 4)]
    # document_fetch_search: filter only (score 0) collapsed on the patient -> the first document of the patient
    hits = idxr.document_fetch_search("ignored", k=5, patient_id="pat-1")
    assert [(h[0]["doc_id"], h[1]) for h in hits] == [("p1", 0.0)]
    assert idxr.document_fetch_search("ignored", k=5) == []
    # temporal_search: the sorted search itself works (newest condition first, only dates within a year) ...
    body = {"size": 5, "sort": [{"conditionOnsetDateTime": {"order": "desc"}}],
            "query": {"bool": {"must": [
                {"multi_match": {"query": "chest pain diabetes", "fields": idxr.text_fields + idxr.keyword_fields,
                                 "type": "best_fields", "operator": "or"}},
                {"bool": {"should": [{"range": {f: {"gte": "now-1y", "lte": "now"}}} for f in idxr.DATE_FIELDS],
                          "minimum_should_match": 1}}]}}}
    hits = client.search(index=name, body=body)["hits"]["hits"]
    assert [h["_id"] for h in hits] == ["c4", "c1", "c2"] and all(h["_score"] is None for h in hits)
    # ... and the indexer method keeps the reference's behaviour: float(None) -> caught -> []
    assert idxr.temporal_search("chest pain diabetes", k=5) == []
    # hybrid_structured_search: phrase_prefix + knn restricted to structured documents of the patient
    q = np.random.default_rng(8).standard_normal((1, DIM)).astype(np.float32)
    hits = idxr.hybrid_structured_search("chest pa", q, k=5, patient_id="pat-1")
    assert [h[0]["doc_id"] for h in hits] == ["c1", "c2"] and all(h[0]["doc_type"] == "structured" for h in hits)
    with pytest.raises(NotImplementedError):
        client.search(index=name, body={"query": {"more_like_this": {"fields": ["x"], "like": "y"}}})


def test_ner_filter_clause_keeps_hybrid_and_knn_on_the_gpu(setup):
    """filter_clause from ner_preprocess is {"bool": {"must": [match_phrase | range]}} (app/main.py:2589-2609).  The
    hybrid / knn search it restricts must run on the device (row list from the host filter sub-tree) and return what the
    complete host evaluation of the same body returns."""
    client, name, idxr, fields, types, emb = setup
    idx = client._get(name)
    ner = {"bool": {"must": [{"match_phrase": {"unstructuredText": "chest pain"}}]}}
    q = emb[0:1] + 0.01
    before = idx.engine.last_stats.copy() if idx.engine.last_stats else {}
    got = idxr.hybrid_search("chest pain breath", q, k=3, filter_clause=ner, patient_id="pat-1")
    assert idx.engine.last_stats != before and idx.engine.last_stats["launches"] > 0       # the engine ran the search
    assert [d["doc_id"] for d, _ in got] == ["n1"]
    body = idxr._hybrid_body("chest pain breath", q, 3, 1.5, 1.0, 2.0, ner, "pat-1")
    hits, _, _ = idx.host_search(body)
    assert [h["_id"] for h in hits] == [d["doc_id"] for d, _ in got]
    np.testing.assert_allclose([h["_score"] for h in hits], [s for _, s in got], rtol=1e-6)
    # knn + NER filter (semantic_search): post-filter of the k nearest, like the nmslib engine
    got = idxr.semantic_search(q, k=2, filter_clause={"bool": {"must": [{"match_phrase": {"unstructuredText": "diabetes"}}]}})
    assert [d["doc_id"] for d, _ in got] == ["n2"]
    # size above RASS_MAX_K is refused, not silently truncated
    with pytest.raises(Exception):
        client.search(index=name, body={"size": 200, "query": {"knn": {"embedding": {"vector": q[0].tolist(), "k": 200}}}})
