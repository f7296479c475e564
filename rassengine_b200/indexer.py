"""B200Indexer: the reference's `OpenSearchIndexer` surface (app/main.py:1395-2150) for the retrieval hot path, plus
`ensure_index_exists` (app/main.py:350-579) and the vector half of `store_fhir_docs_in_opensearch`
(app/main.py:1211-1282).  Same method names, argument meaning, return shape ([( _source, _score )]) and error
behaviour (errors are logged and become `[]`), so `ask()` can construct it in place of OpenSearchIndexer.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from .client import B200Client, bulk

logger = logging.getLogger("rassengine_b200")

EMBED_DIM = int(os.getenv("EMBED_DIM", "1024"))          # app/main.py:80
TOP_K = int(os.getenv("TOP_K", "3"))                     # app/main.py:88
BATCH_SIZE = int(os.getenv("BATCH_SIZE", "64"))          # app/main.py:78
SHARD_COUNT = int(os.getenv("SHARD_COUNT", "1"))         # app/main.py:89
REPLICA_COUNT = int(os.getenv("REPLICA_COUNT", "0"))     # app/main.py:90
OPENSEARCH_INDEX_NAME = os.getenv("OPENSEARCH_INDEX_NAME", "rass-index")


def get_index_name(user_id: str) -> str:
    return f"{OPENSEARCH_INDEX_NAME}-{user_id}"          # app/main.py:346-347


# The boosted field lists of OpenSearchIndexer.__init__ (app/main.py:1403-1456): interface data of the two multi_match
# clauses.  Chunk documents only carry unstructuredText; structured FHIR documents carry the others.
TEXT_FIELDS = [
    "unstructuredText^3", "patientName^3", "patientAddress^3", "patientTelecom^3", "conditionCodeText^2",
    "conditionNote^2", "observationCodeText", "observationValue", "observationReferenceRange", "observationNote^2",
    "encounterType", "encounterReasonCode", "encounterLocation", "encounterNote", "medRequestMedicationDisplay",
    "medRequestNote", "procedureCodeText", "procedureNote", "allergyCodeText", "allergyNote^2", "practitionerName^3",
    "practitionerAddress", "practitionerTelecom", "organizationName^3", "organizationAddress", "organizationTelecom",
]
DATE_FIELDS = ["patientDOB", "conditionOnsetDateTime", "conditionRecordedDate", "observationEffectiveDateTime",
               "observationIssued", "encounterStart", "encounterEnd", "medRequestAuthoredOn",
               "procedurePerformedDateTime", "allergyOnsetDateTime"]                           # app/main.py:1457-1468
KEYWORD_FIELDS = [
    "patientGender^3", "patientMaritalStatus^2", "patientLanguage^3", "conditionCategory^2", "conditionClinicalStatus",
    "conditionVerificationStatus", "conditionSeverity", "observationUnit", "observationInterpretation",
    "encounterStatus", "encounterClass", "encounterServiceProvider", "medRequestIntent", "medRequestStatus",
    "medRequestPriority", "procedureStatus", "allergyClinicalStatus", "allergyVerificationStatus", "allergyType",
    "allergyCategory", "allergyCriticality", "practitionerGender", "practitionerSpecialty", "organizationType",
]


def index_body(dim: int = EMBED_DIM) -> dict:
    """The parts of the reference mapping the engine acts on (app/main.py:354-572): index.knn, shard counts, the
    knn_vector field, and the type (text / keyword) of every field the two multi_match clauses search."""
    props = {"doc_id": {"type": "keyword"}, "doc_type": {"type": "keyword"}, "patientId": {"type": "keyword"},
             "resourceType": {"type": "keyword"}, "file_path": {"type": "keyword"}, "file_type": {"type": "keyword"}}
    props.update({f.split("^")[0]: {"type": "text", "fields": {"keyword": {"type": "keyword", "ignore_above": 256}}}
                  for f in TEXT_FIELDS})
    props["unstructuredText"] = {"type": "text"}
    props.update({f: {"type": "date", "format": "yyyy-MM-dd||strict_date_optional_time||epoch_millis"}
                  for f in DATE_FIELDS})
    props.update({f.split("^")[0]: {"type": "keyword"} for f in KEYWORD_FIELDS})
    props["embedding"] = {"type": "knn_vector", "dimension": dim,
                          "method": {"name": "hnsw", "engine": "nmslib", "space_type": "cosinesimil",
                                     "parameters": {"m": 48, "ef_construction": 400}}}
    return {"settings": {"index": {"knn": True, "number_of_shards": SHARD_COUNT, "number_of_replicas": REPLICA_COUNT}},
            "mappings": {"properties": props}}


def ensure_index_exists(client, index_name: str, body: dict | None = None) -> None:
    """Idempotent create; errors are printed and swallowed like the reference (app/main.py:578-579)."""
    if not client:
        return
    try:
        if not client.indices.exists(index_name):
            client.indices.create(index=index_name, body=body or index_body())
    except Exception as exc:
        print(f"[ERROR] ensure_index_exists: {exc}")


def store_structured(client, index_name: str, docs: List[Dict]) -> Tuple[int, list]:
    """Structured half of store_fhir_docs_in_opensearch (app/main.py:1222-1240): FHIR resource documents without a
    vector, one bulk call, `_id = doc_id`.  They share the index with the chunks and match the text / keyword
    clauses of the hybrid query through their own fields."""
    if not client or not docs:
        return 0, []
    ensure_index_exists(client, index_name)
    actions = [{"_op_type": "index", "_index": index_name, "_id": d["doc_id"], "_source": d,
                "_routing": d.get("patientId")} for d in docs]
    ok, errors = bulk(client, actions)
    if errors:
        logger.error("bulk index errors: %s", errors[:3])
    return ok, errors


def store_chunks(client, index_name: str, docs: List[Dict], embeddings: np.ndarray, as_lists: bool = True,
                 flush: int | None = None) -> Tuple[int, list]:
    """Vector half of store_fhir_docs_in_opensearch (app/main.py:1247-1279): L2-normalise in fp32 exactly as the
    reference does, attach the embedding, bulk-index in flushes of BATCH_SIZE with `_id = doc_id`.

    as_lists=True sends each vector as a python list, byte for byte what the reference does (`.tolist()`,
    app/main.py:1256).  as_lists=False is the fast ingest path: the float32 rows go to the
    engine as numpy views -- same stored values, no per-element python objects -- and `flush` may be raised far
    above BATCH_SIZE because a flush is one pinned async copy, not an HTTP request."""
    if not client or not docs:
        return 0, []
    ensure_index_exists(client, index_name)
    embeddings = np.asarray(embeddings, dtype=np.float32)
    norms = np.linalg.norm(embeddings, axis=1, keepdims=True)
    embeddings = embeddings / (norms + 1e-9)
    flush = flush or BATCH_SIZE
    ok, errors, actions = 0, [], []
    for i, doc in enumerate(docs):
        d = dict(doc)
        d["embedding"] = embeddings[i].tolist() if as_lists else embeddings[i]
        actions.append({"_op_type": "index", "_index": index_name, "_id": d["doc_id"], "_source": d,
                        "_routing": d.get("patientId")})
        if len(actions) >= flush:
            s, e = bulk(client, actions)
            ok, actions = ok + s, []
            errors.extend(e)
    if actions:
        s, e = bulk(client, actions)
        ok += s
        errors.extend(e)
    if errors:
        logger.error("bulk index errors: %s", errors[:3])
    return ok, errors


async def store_fhir_docs_in_opensearch(structured_docs: List[Dict], unstructured_docs: List[Dict], client,
                                        index_name: str, embed=None) -> None:
    """The reference's ingest entry point under its own name, signature and error behaviour
    (app/main.py:1211-1282; app/embedding_gen.py:1061-1132): structured documents first, then the chunks with their
    embeddings, nothing raised to the caller.  `embed` stands for the reference's `embed_texts_in_batches`
    (app/main.py:1247-1248, the Ollama client -- outside the retrieval path): an async or plain callable
    `(texts, batch_size=...) -> float32 [n, dim]`; a host app passes its own or assigns `indexer.embed_texts_in_batches`.
    The float32 rows go to the engine as numpy views (no `.tolist()` / JSON detour): the stored
    values are the ones the reference would store."""
    if not client:
        print("[store_fhir_docs_in_opensearch] No OS client.")
        return
    ensure_index_exists(client, index_name)
    if structured_docs:
        try:
            ok, errors = store_structured(client, index_name, structured_docs)
            logger.info("Indexed %d structured docs, errors: %s", ok, errors)
        except Exception as exc:
            logger.error("Structured docs indexing error: %s", exc)
    if not unstructured_docs:
        return
    try:
        fn = embed or embed_texts_in_batches
        if fn is None:
            raise RuntimeError("no embedding function: pass embed= or assign indexer.embed_texts_in_batches")
        out = fn([d["unstructuredText"] for d in unstructured_docs], batch_size=BATCH_SIZE)
        if hasattr(out, "__await__"):
            out = await out
        ok, errors = store_chunks(client, index_name, unstructured_docs, np.asarray(out, dtype=np.float32),
                                  as_lists=False)
        logger.info("Indexed %d unstructured docs, errors: %s", ok, errors)
    except Exception as exc:
        logger.error("Unstructured docs indexing error: %s", exc)


embed_texts_in_batches = None     # the host application's embedding client (app/main.py:240-262), assigned by it


class B200Indexer:
    text_fields = TEXT_FIELDS
    keyword_fields = KEYWORD_FIELDS

    def __init__(self, client: B200Client, index_name: str):
        self.client = client
        self.index_name = index_name

    # -- app/main.py:1470-1478 ------------------------------------------------------------------------
    def has_any_data(self) -> bool:
        if not self.client:
            return False
        try:
            return self.client.count(index=self.index_name)["count"] > 0
        except Exception:
            return False

    @staticmethod
    def _unit_vector(query_emb: np.ndarray) -> list:
        norms = np.linalg.norm(query_emb, axis=1, keepdims=True)      # app/main.py:1536-1537
        return (query_emb / (norms + 1e-9))[0].tolist()

    @staticmethod
    def _filters(filter_clause, patient_id):
        out = []
        if filter_clause:
            out.append(filter_clause)
        if patient_id:
            out.append({"term": {"patientId": patient_id}})
        return out

    def _run(self, body: dict, patient_id, what: str) -> List[Tuple[Dict, float]]:
        try:
            resp = self.client.search(index=self.index_name, body=body, routing=patient_id)
            return [(hit["_source"], float(hit["_score"])) for hit in resp["hits"]["hits"]]
        except Exception as e:
            logger.error(f"{what} error: {e}")
            return []

    # -- app/main.py:1527-1560 ------------------------------------------------------------------------
    def semantic_search(self, query_emb: np.ndarray, k: int = TOP_K, filter_clause: Optional[Dict] = None,
                        patient_id: Optional[str] = None, query: Optional[str] = None) -> List[Tuple[Dict, float]]:
        """`query` is accepted and ignored: ask() passes it to every embedding-taking method (app/main.py:2879-2885)."""
        if query_emb.size == 0:
            return []
        body = {"size": k, "query": {"knn": {"embedding": {"vector": self._unit_vector(query_emb), "k": k}}},
                "terminate_after": k}
        flt = self._filters(filter_clause, patient_id)
        if flt:
            body["query"] = {"bool": {"must": [body["query"]], "filter": flt}}
        return self._run(body, patient_id, "Semantic search")

    # -- app/main.py:1562-1615 ------------------------------------------------------------------------
    def _hybrid_body(self, query: str, query_emb: np.ndarray, k: int, w_text: float, w_kw: float, w_knn: float,
                     filter_clause, patient_id) -> dict:
        should = [
            {"multi_match": {"query": query, "fields": self.text_fields, "type": "best_fields", "operator": "or",
                             "fuzziness": "AUTO", "boost": w_text}},
            {"multi_match": {"query": query, "fields": self.keyword_fields, "type": "best_fields", "operator": "or",
                             "boost": w_kw}},
            {"knn": {"embedding": {"vector": self._unit_vector(query_emb), "k": k, "boost": w_knn}}},
        ]
        bool_query = {"should": should, "minimum_should_match": 1}
        flt = self._filters(filter_clause, patient_id)
        if flt:
            bool_query["filter"] = flt
        return {"size": k, "query": {"bool": bool_query}, "terminate_after": k}

    def hybrid_search(self, query: str, query_emb: np.ndarray, k: int = TOP_K, filter_clause: Optional[Dict] = None,
                      patient_id: Optional[str] = None) -> List[Tuple[Dict, float]]:
        if not query.strip() or query_emb.size == 0:
            return []
        body = self._hybrid_body(query, query_emb, k, 1.5, 1.0, 2.0, filter_clause, patient_id)
        return self._run(body, patient_id, "Hybrid search")

    # -- app/main.py:1969-2027 (same clause shape, boosts 1.0 / 0.5 / 1.5; the malformed date clause is dropped) -
    def multi_intent_search(self, query: str, query_emb: np.ndarray, k: int = TOP_K,
                            filter_clause: Optional[Dict] = None, patient_id: Optional[str] = None):
        if not query.strip() or query_emb.size == 0:
            return []
        body = self._hybrid_body(query, query_emb, k, 1.0, 0.5, 1.5, filter_clause, patient_id)
        return self._run(body, patient_id, "Multi-intent search")

    # -- the north-star core: search(query_emb, query_text, top_k) -> _id/_score hits --------------------
    def search(self, query_emb: np.ndarray, query_text: Optional[str] = None, top_k: int = TOP_K) -> List[Dict]:
        query_emb = np.asarray(query_emb, dtype=np.float32).reshape(1, -1)
        if query_text and query_text.strip():
            body = self._hybrid_body(query_text, query_emb, top_k, 1.5, 1.0, 2.0, None, None)
        else:
            body = {"size": top_k,
                    "query": {"knn": {"embedding": {"vector": self._unit_vector(query_emb), "k": top_k}}}}
        try:
            return self.client.search(index=self.index_name, body=body)["hits"]["hits"]
        except Exception as e:
            logger.error(f"search error: {e}")
            return []

    # -- the keyword-side family: same bodies as the reference, answered host-side by the client ------
    DATE_FIELDS = DATE_FIELDS
    STRUCTURED_FIELDS = ["patientName^3", "patientGender^3", "patientTelecom^3", "conditionCodeText^2",
                         "conditionClinicalStatus", "conditionSeverity", "observationCodeText", "observationValue",
                         "observationUnit", "encounterStatus", "encounterClass", "medRequestMedicationDisplay",
                         "medRequestStatus", "procedureCodeText", "procedureStatus", "allergyCodeText",
                         "allergyClinicalStatus", "practitionerName^3", "organizationName^3"]   # app/main.py:1722-1742

    def _bool_body(self, bool_query: dict, k: int, filter_clause, patient_id, extra_filters=(), **top) -> dict:
        flt = self._filters(filter_clause, patient_id) + list(extra_filters)
        if flt:
            bool_query["filter"] = flt
        return {"size": k, "query": {"bool": bool_query}, "terminate_after": k, **top}

    def exact_match_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:1480-1525: phrase match over the text fields (boost 2) and the keyword fields."""
        if not query.strip():
            return []
        should = [{"multi_match": {"query": query, "fields": self.text_fields, "type": "phrase", "boost": 2.0}},
                  {"multi_match": {"query": query, "fields": self.keyword_fields, "type": "phrase"}}]
        body = self._bool_body({"should": should, "minimum_should_match": 1}, k, filter_clause, patient_id)
        return self._run(body, patient_id, "Exact match search")

    def structured_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:1617-1705 as intended (the reference reads an undefined `structured_fields` and raises
        NameError): phrase_prefix over the structured fields, restricted to structured documents."""
        if not query.strip():
            return []
        must = [{"multi_match": {"query": query, "fields": self.STRUCTURED_FIELDS, "type": "phrase_prefix",
                                 "operator": "and"}}]
        body = {"size": k, "query": {"bool": {"must": must, "filter": [{"term": {"doc_type": "structured"}}]}},
                "terminate_after": k}
        return self._run(body, patient_id, "Structured search")

    def hybrid_structured_search(self, query: str, query_emb: np.ndarray, k: int = TOP_K, filter_clause=None,
                                 patient_id=None):
        """app/main.py:1707-1775: phrase_prefix (boost 1.5) + knn (boost 2.0) over structured documents.  (The
        reference appends to a filter list that only exists when a filter or patient was given.)"""
        if not query.strip() or query_emb.size == 0:
            return []
        should = [{"multi_match": {"query": query, "fields": self.STRUCTURED_FIELDS, "type": "phrase_prefix",
                                   "operator": "and", "boost": 1.5}},
                  {"knn": {"embedding": {"vector": self._unit_vector(query_emb), "k": k, "boost": 2.0}}}]
        body = self._bool_body({"should": should, "minimum_should_match": 1}, k, filter_clause, patient_id,
                               extra_filters=[{"term": {"doc_type": "structured"}}])
        return self._run(body, patient_id, "Hybrid structured search")

    def aggregate_search(self, query: str, filter_clause=None, patient_id=None) -> Dict:
        """app/main.py:1777-1810: three terms aggregations, optionally under a filter; returns resp["aggregations"]."""
        body = {"size": 0, "aggs": {
            "by_condition": {"terms": {"field": "conditionCodeText.keyword", "size": 5}},
            "by_resource": {"terms": {"field": "resourceType.keyword", "size": 5}},
            "by_patient": {"terms": {"field": "patientId", "size": 5}}}}
        flt = self._filters(filter_clause, patient_id)
        if flt:
            body["query"] = {"bool": {"filter": flt}}
        try:
            return self.client.search(index=self.index_name, body=body, routing=patient_id)["aggregations"]
        except Exception as e:
            logger.error(f"Aggregate search error: {e}")
            return {}

    def comparison_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:1812-1866: best_fields AUTO over the comparable fields (+ an aggregation nobody reads)."""
        if not query.strip():
            return []
        fields = ["conditionCodeText^2", "observationValue", "observationUnit", "medRequestMedicationDisplay",
                  "procedureCodeText", "allergyCodeText"]
        should = [{"multi_match": {"query": query, "fields": fields, "type": "best_fields", "operator": "or",
                                   "fuzziness": "AUTO"}}]
        body = self._bool_body({"should": should, "minimum_should_match": 1}, k, filter_clause, patient_id,
                               aggs={"by_field": {"terms": {"field": "conditionCodeText.keyword", "size": 3}}})
        return self._run(body, patient_id, "Comparison search")

    def temporal_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:1868-1922: text match AND any date field within the last year, newest condition first.  A
        sorted search carries `_score: null`, so the reference's `float(hit["_score"])` raises and the method returns
        [] whenever something matched; that behaviour is kept (use client.search with the same body for the hits)."""
        if not query.strip():
            return []
        dates = {"bool": {"should": [{"range": {f: {"gte": "now-1y", "lte": "now"}}} for f in self.DATE_FIELDS],
                          "minimum_should_match": 1}}
        must = [{"multi_match": {"query": query, "fields": self.text_fields + self.keyword_fields,
                                 "type": "best_fields", "operator": "or"}}, dates]
        body = self._bool_body({"must": must}, k, filter_clause, patient_id,
                               sort=[{"conditionOnsetDateTime": {"order": "desc"}}])
        return self._run(body, patient_id, "Temporal search")

    def explanatory_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:1924-1967: best_fields AUTO over the note fields (GPU text path: one scoring clause under must)."""
        if not query.strip():
            return []
        fields = ["conditionNote^3", "observationNote^3", "encounterNote^3", "medRequestNote^3", "procedureNote^3",
                  "allergyNote^3", "unstructuredText^2"]
        must = [{"multi_match": {"query": query, "fields": fields, "type": "best_fields", "operator": "or",
                                 "fuzziness": "AUTO"}}]
        body = self._bool_body({"must": must}, k, filter_clause, patient_id)
        return self._run(body, patient_id, "Explanatory search")

    def entity_specific_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:2029-2075: phrase match over the entity fields."""
        if not query.strip():
            return []
        fields = ["patientName^4", "patientId^4", "patientGender^3", "patientTelecom^3", "practitionerName^3",
                  "organizationName^3"]
        must = [{"multi_match": {"query": query, "fields": fields, "type": "phrase", "operator": "and"}}]
        body = self._bool_body({"must": must}, k, filter_clause, patient_id)
        return self._run(body, patient_id, "Entity-specific search")

    def document_fetch_search(self, query: str, k: int = TOP_K, filter_clause=None, patient_id=None):
        """app/main.py:2079-2110: the patient's documents collapsed on patientId (one hit per patient)."""
        if not patient_id:
            return []
        flt = [{"term": {"patientId": patient_id}}] + ([filter_clause] if filter_clause else [])
        body = {"size": k, "query": {"bool": {"filter": flt}}, "collapse": {"field": "patientId"},
                "terminate_after": k}
        return self._run(body, patient_id, "Document fetch search")


def resolve_patient_ids(client, index_name: str, name: str, top_k: int = 5) -> List[str]:
    """The lookup of resolve_patient_ids_from_name (app/main.py:2709-2744): exact keyword, phrase or fuzzy all-terms
    match on patientName, one hit per patientId."""
    body = {"size": top_k, "_source": ["patientId"], "collapse": {"field": "patientId"},
            "query": {"bool": {"should": [
                {"term": {"patientName.keyword": name}},
                {"match_phrase": {"patientName": name}},
                {"match": {"patientName": {"query": name, "operator": "and", "fuzziness": "AUTO"}}}],
                "minimum_should_match": 1}}}
    try:
        resp = client.search(index=index_name, body=body)
        return [h["_source"]["patientId"] for h in resp.get("hits", {}).get("hits", []) if h["_source"].get("patientId")]
    except Exception as e:
        logger.error(f"Patient ID resolution failed: {e}")
        return []
