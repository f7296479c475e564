"""`fuzziness: AUTO` restatement (test infrastructure, parity UNPINNED).

Reference call site: the first `multi_match` of hybrid_search / multi_intent_search carries
`"fuzziness": "AUTO"` (app/main.py:1577-1585, 1985-1993).  What OpenSearch 2.11 / Lucene 9.7 do with it is third
party and not vendored; restated from the published algorithm:

  * AUTO = AUTO:3,6 -> max edits 0 for tokens of 1-2 characters, 1 for 3-5, 2 for longer ones; prefix_length 0,
    max_expansions 50, transpositions on (an adjacent swap is ONE edit; Lucene's Levenshtein automata with
    transpositions accept the optimal-string-alignment distance, restated here as a plain DP).
  * every analysed query token becomes a FuzzyQuery whose default rewrite (TopTermsBlendedFreqScoringRewrite) keeps
    the 50 best dictionary terms ordered by (boost desc, term asc) with
        boost = 1                                   for the token itself
        boost = 1 - edits / min(len(term), len(token))   otherwise        (float)
    and, when more than one term survives, scores each of them with the LARGEST document frequency among them
    ("blended" statistics): idf = (float) ln(1 + (docCount - maxDf + .5) / (maxDf + .5)).
  * each surviving term is a boosted TermQuery: weight = (clause boost * field boost * term boost) * idf in float,
    score = weight - weight / (1 + tf * inv[norm]); the token's score is the sum over its terms, the clause's score
    the sum over tokens (double accumulation, cast to float) -- same as oracle/bm25.py with explicit weights.
"""
from __future__ import annotations

import math

import numpy as np

from .bm25 import BM25Index

MAX_EXPANSIONS = 50


def auto_max_edits(n_chars: int) -> int:
    return 0 if n_chars <= 2 else (1 if n_chars <= 5 else 2)


def osa_distance(a: str, b: str) -> int:
    """Optimal string alignment distance: insert / delete / substitute / swap of two adjacent characters."""
    la, lb = len(a), len(b)
    prev2 = None
    prev = list(range(lb + 1))
    for i in range(1, la + 1):
        cur = [i] + [0] * lb
        for j in range(1, lb + 1):
            cost = 0 if a[i - 1] == b[j - 1] else 1
            v = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + cost)
            if i > 1 and j > 1 and a[i - 1] == b[j - 2] and a[i - 2] == b[j - 1]:
                v = min(v, prev2[j - 2] + 1)
            cur[j] = v
        prev2, prev = prev, cur
    return prev[lb]


def expand(vocab_terms: list[str], token: str, max_expansions: int = MAX_EXPANSIONS):
    """-> [(term_id, edits, boost float32)] ordered (boost desc, term asc), at most max_expansions entries."""
    me = auto_max_edits(len(token))
    out = []
    for tid, term in enumerate(vocab_terms):
        if term == token:
            out.append((tid, 0, np.float32(1.0)))
            continue
        if me == 0 or abs(len(term) - len(token)) > me:
            continue
        ed = osa_distance(token, term)
        if ed <= me:
            boost = np.float32(1.0) - np.float32(ed) / np.float32(min(len(term), len(token)))
            out.append((tid, ed, np.float32(boost)))
    out.sort(key=lambda e: (-float(e[2]), vocab_terms[e[0]]))
    return out[:max_expansions]


def weighted_terms(index: BM25Index, vocab_terms: list[str], tokens: list[str], boost: float):
    """The boosted term queries of the whole clause: (term ids, float32 weights), tokens in query order, a token's
    terms in (boost desc, term asc) order.  Terms without postings are dropped before the statistics are blended
    (a FuzzyTermsEnum only sees terms that exist in the segment)."""
    ids, ws = [], []
    bo = np.float32(boost)
    for tok in tokens:
        ex = [(t, ed, b) for t, ed, b in expand(vocab_terms, tok) if t < index.vocab and index.df[t] > 0]
        if not ex:
            continue
        max_df = max(int(index.df[t]) for t, _, _ in ex)
        idf = np.float32(math.log(1.0 + (index.doc_count - max_df + 0.5) / (max_df + 0.5)))
        for t, _, b in ex:
            ids.append(int(t))
            ws.append(np.float32(np.float32(bo * b) * idf))
    return ids, np.asarray(ws, dtype=np.float32)


def score(index: BM25Index, term_ids, weights) -> np.ndarray:
    """Dense float32 clause score for explicit (term, weight) pairs; arithmetic of BM25Index.score."""
    acc = np.zeros(index.n_docs, dtype=np.float64)
    one = np.float32(1.0)
    for t, w in zip(term_ids, np.asarray(weights, dtype=np.float32)):
        lo, hi = index.indptr[t], index.indptr[t + 1]
        if hi == lo:
            continue
        d = index.doc[lo:hi]
        tf = index.tf[lo:hi].astype(np.float32)
        s = w - w / (one + tf * index.inv[index.norm[d]])
        acc[d] += np.where(s > 0, s, np.float32(0)).astype(np.float64)
    return acc.astype(np.float32)
