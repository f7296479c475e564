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


def phrase_score(f: Field, docs_tokens: list[list[str]], query: str, boost: float, prefix: bool = False) -> np.ndarray:
    """multi_match type phrase / phrase_prefix on one field (Lucene PhraseQuery / MultiPhraseQuery, slop 0): a document
    scores BM25 of its phrase frequency with weight boost * (float) sum of the idf of EVERY term of the query (all
    prefix expansions included, at most 50 in term order); a one-position query degenerates to a disjunction of term
    queries.  docs_tokens = the field's token list per document (positions)."""
    n = len(docs_tokens)
    out = np.zeros(n, dtype=np.float32)
    tokens = [query] if f.kind == "keyword" else analyze(query)
    if not tokens:
        return out
    tid = {t: i for i, t in enumerate(f.terms)}
    positions = []
    for i, tok in enumerate(tokens):
        if prefix and i == len(tokens) - 1:
            alts = [t for t in f.terms if t.startswith(tok)][:50]
        else:
            alts = [tok] if tok in tid else []
        if not alts:
            return out
        positions.append(alts)
    bo = np.float32(boost)
    one = np.float32(1.0)
    if len(positions) == 1:
        ids = [tid[t] for t in positions[0]]
        return fuzzy.score(f.index, ids, [np.float32(bo * f.index.idf(t)) for t in ids])
    idf = np.float32(sum(float(f.index.idf(tid[t])) for alts in positions for t in alts))
    w = np.float32(bo * idf)
    for d, toks in enumerate(docs_tokens):
        freq = 0
        for s in range(len(toks) - len(positions) + 1):
            if all(toks[s + j] in positions[j] for j in range(len(positions))):
                freq += 1
        if freq:
            out[d] = w - w / (one + np.float32(freq) * f.index.inv[f.index.norm[d]])
    return out


def field_tokens(docs: list[dict], name: str, kind: str) -> list[list[str]]:
    out = []
    for d in docs:
        v = d.get(name)
        vals = v if isinstance(v, (list, tuple)) else [v]
        out.append([str(x) for x in vals if x is not None] if kind == "keyword"
                   else [t for x in vals if isinstance(x, str) for t in analyze(x)])
    return out
