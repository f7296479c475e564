"""multi_match over several fields + bool.should over several clauses (test infrastructure, parity UNPINNED).

Reference call sites: the two `multi_match` clauses of hybrid_search / multi_intent_search over the 26 `text_fields`
and 24 `keyword_fields` of OpenSearchIndexer (app/main.py:1403-1456, 1576-1594, 1984-2002); structured FHIR documents
and text chunks share one index (app/main.py:1223-1240), so a document can carry several of those fields.

Third-party semantics restated (OpenSearch 2.11 MultiMatchQuery, Lucene 9.7 DisjunctionMaxQuery / BooleanQuery):
  * `best_fields`, tie_breaker 0: clause score = max over fields of the field's score (each a float);
    a field's score = sum over its term queries (double accumulation, cast to float), boost = clause boost *
    field boost (* fuzzy term boost) in float; statistics (docCount, avgdl, df) are per field;
  * a `text` field is analysed (oracle/analyzer.py), `fuzziness: AUTO` applies (oracle/fuzzy.py);
  * a `keyword` field indexes the whole value as one term, the query string is NOT analysed, fuzziness does not
    apply (the reference sends none on that clause), norms are omitted: every value counts as length 1;
  * `bool.should`: S_text(d) = sum over clauses of the float clause scores, accumulated in double.
"""
from __future__ import annotations

import numpy as np

from . import fuzzy
from .analyzer import analyze
from .bm25 import BM25Index


class Field:
    def __init__(self, kind: str, docs_tokens: list[list[str]]):
        """docs_tokens[d] = the tokens of the field in document d ([] = the document lacks the field)."""
        self.kind = kind
        self.terms = sorted({t for toks in docs_tokens for t in toks})
        tid = {t: i for i, t in enumerate(self.terms)}
        ids = [[tid[t] for t in toks] for toks in docs_tokens]
        self.index = BM25Index.from_token_ids(ids, len(self.terms))
        if kind == "keyword":      # omitted norms
            dl = (self.index.doclen > 0).astype(np.uint32)
            self.index = BM25Index(self.index.indptr, self.index.doc, self.index.tf, dl)


def build(docs: list[dict], types: dict[str, str]) -> dict[str, Field]:
    out = {}
    for name, kind in types.items():
        toks = []
        for d in docs:
            v = d.get(name)
            vals = v if isinstance(v, (list, tuple)) else [v]
            if kind == "keyword":
                toks.append([str(x) for x in vals if x is not None])
            else:
                toks.append([t for x in vals if isinstance(x, str) for t in analyze(x)])
        if any(toks):
            out[name] = Field(kind, toks)
    return out


def clause_score(fields: dict[str, Field], query: str, specs, clause_boost: float, fuzziness: bool, n_docs: int):
    """float32 dense score of one multi_match(best_fields) clause; specs = [(field name, field boost)]."""
    best = np.zeros(n_docs, dtype=np.float32)
    for name, fb in specs:
        f = fields.get(name)
        if f is None:
            continue
        bo = np.float32(np.float32(clause_boost) * np.float32(fb))
        if f.kind == "keyword":
            tokens, fz = [query], False
        else:
            tokens, fz = analyze(query), fuzziness
        if fz:
            ids, ws = fuzzy.weighted_terms(f.index, f.terms, tokens, bo)
        else:
            ids, ws = [], []
            tid = {t: i for i, t in enumerate(f.terms)}
            for tok in tokens:
                t = tid.get(tok)
                if t is not None:
                    ids.append(t)
                    ws.append(np.float32(bo * f.index.idf(t)))
        if ids:
            best = np.maximum(best, fuzzy.score(f.index, ids, ws))
    return best


def text_total(fields: dict[str, Field], clauses, n_docs: int) -> np.ndarray:
    """float64 dense sum over clauses; clauses = [(query, specs, clause boost, fuzziness)]."""
    total = np.zeros(n_docs, dtype=np.float64)
    for query, specs, cb, fz in clauses:
        total += clause_score(fields, query, specs, cb, fz, n_docs).astype(np.float64)
    return total
