"""Query-DSL -> plan.  Parses exactly the body shapes RASSEngine's OpenSearchIndexer emits for the hot path:

  kNN       {"size": k, "query": {"knn": {"embedding": {"vector": v, "k": k}}}, "terminate_after": k}
            optionally wrapped as {"bool": {"must": [knn], "filter": [...]}}            app/main.py:1538-1550
  hybrid    {"size": k, "query": {"bool": {"should": [multi_match, multi_match, knn],
             "minimum_should_match": 1, "filter": [...]}}, "terminate_after": k}        app/main.py:1574-1605
             (multi_intent_search emits the same shape with other boosts            app/main.py:1982-2010)

Filters other than `term` (the NER filter_clause: match_phrase / range under a bool, app/main.py:2589-2609) stay in the
plan verbatim: the client turns that sub-tree alone into a row list on the host and the search itself stays on the
GPU.  Anything else (phrase, phrase_prefix, sort, aggs, collapse ...) raises NotImplementedError here and is evaluated
host-side by hostquery.py; shapes neither understands raise NotImplementedError to the caller, which
the reference's `except Exception: return []` (app/main.py:1558-1560, 1613-1615) turns into "no results".
`terminate_after` is ignored on purpose: the engine returns the true global top-k (reference defects).
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class TextClause:
    query: str
    fields: list[tuple[str, float]]      # (field name, field boost)
    boost: float = 1.0
    fuzziness: str | None = None         # "AUTO" on the reference's text_fields clause (app/main.py:1583)


@dataclass
class Plan:
    size: int = 10
    vector: list | None = None
    knn_k: int = 0
    knn_boost: float = 1.0
    knn_field: str = "embedding"
    text: list[TextClause] = field(default_factory=list)
    filters: list[tuple[str, object]] = field(default_factory=list)   # (field, value) term filters
    # any other bool.filter clause, verbatim: ner_preprocess hands {"bool": {"must": [match_phrase | range ...]}}
    # (app/main.py:2589-2609) to every search method; only THIS sub-tree is evaluated host-side, into a row list
    host_filters: list[dict] = field(default_factory=list)
    kind: str = "knn"                    # "knn" | "hybrid" | "match_all"


def _parse_field(spec: str) -> tuple[str, float]:
    if "^" in spec:
        name, b = spec.rsplit("^", 1)
        return name, float(b)
    return spec, 1.0


def _parse_knn(node: dict, plan: Plan):
    if len(node) != 1:
        raise NotImplementedError("knn clause with several fields")
    (fld, spec), = node.items()
    plan.knn_field = fld
    plan.vector = spec["vector"]
    plan.knn_k = int(spec.get("k", plan.size))
    plan.knn_boost = float(spec.get("boost", 1.0))


def _parse_filters(nodes, plan: Plan):
    if isinstance(nodes, dict):
        nodes = [nodes]
    for f in nodes or []:
        if not isinstance(f, dict) or len(f) != 1:
            raise NotImplementedError(f"unsupported filter clause: {f!r}")
        if set(f) == {"term"} and len(f["term"]) == 1:
            (fld, val), = f["term"].items()
            if isinstance(val, dict):
                val = val.get("value")
            plan.filters.append((fld, val))
        else:
            plan.host_filters.append(f)      # match_phrase / range / nested bool ...: rows via hostquery, scan on GPU


def parse_search_body(body: dict) -> Plan:
    if not isinstance(body, dict):
        raise ValueError("search body must be a dict")
    for key in ("aggs", "aggregations", "sort", "collapse", "from", "_source"):
        if key in body:
            raise NotImplementedError(f"'{key}' is outside the retrieval hot path (evaluated host-side)")
    plan = Plan(size=int(body.get("size", 10)))
    q = body.get("query")
    if q is None or q == {"match_all": {}}:
        plan.kind = "match_all"
        return plan
    if "knn" in q and len(q) == 1:
        _parse_knn(q["knn"], plan)
        return plan
    if "bool" not in q or len(q) != 1:
        raise NotImplementedError(f"unsupported query: {list(q)}")
    b = q["bool"]
    unknown = set(b) - {"must", "should", "filter", "minimum_should_match"}
    if unknown:
        raise NotImplementedError(f"unsupported bool keys: {sorted(unknown)}")
    _parse_filters(b.get("filter"), plan)
    must = b.get("must") or []
    if isinstance(must, dict):
        must = [must]
    should = b.get("should") or []
    if must:
        if should or len(must) != 1:
            raise NotImplementedError("bool.must with several clauses is evaluated host-side")
        if "knn" in must[0]:
            _parse_knn(must[0]["knn"], plan)
            return plan
        # a single scoring multi_match under must (explanatory_search, app/main.py:1924-1967) scores and matches
        # exactly like the same clause under should with minimum_should_match 1
        should = must
    if not should:
        raise NotImplementedError("bool query without must/should")
    if int(b.get("minimum_should_match", 1)) != 1:
        raise NotImplementedError("minimum_should_match != 1")
    plan.kind = "hybrid"
    for c in should:
        if "knn" in c and len(c) == 1:
            _parse_knn(c["knn"], plan)
        elif "multi_match" in c and len(c) == 1:
            m = c["multi_match"]
            if m.get("type", "best_fields") != "best_fields" or m.get("operator", "or") != "or":
                raise NotImplementedError("multi_match type/operator outside best_fields/or")
            fz = m.get("fuzziness")
            if fz is not None and str(fz).upper() not in ("AUTO", "0"):
                raise NotImplementedError(f"fuzziness {fz!r} (the reference only sends AUTO)")
            plan.text.append(TextClause(str(m.get("query", "")), [_parse_field(f) for f in m.get("fields", [])],
                                        float(m.get("boost", 1.0)),
                                        "AUTO" if fz is not None and str(fz).upper() == "AUTO" else None))
        else:
            raise NotImplementedError(f"unsupported should clause: {list(c)}")
    return plan
