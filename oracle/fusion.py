"""Hybrid fusion oracle (test infrastructure).

Reference: OpenSearchIndexer.hybrid_search, app/main.py:1562-1615.  The fusion
IS the query dict at app/main.py:1574-1598: a Lucene bool.should (boosted SUM,
minimum_should_match 1) of
    multi_match(text_fields,    boost 1.5)    app/main.py:1576-1585
    multi_match(keyword_fields, boost 1.0)    app/main.py:1586-1594
    knn(embedding, k, boost 2.0)              app/main.py:1595
and for multi_intent_search (app/main.py:1982-2003) the same shape with boosts
1.0 / 0.5 / 1.5.  There is no score normalisation in the reference.

For chunk documents only `unstructuredText^3` exists, hence
    S(d) = float( double(BM25_{boost=1.5*3}(q, d)) + double(float(2.0 * knn_score(d)) if d in kNN_k(q)) )
and a doc matches when BM25 > 0 or d in kNN_k(q).  Ranking: (S desc, row asc).
Third-party (Lucene BooleanScorer sums clause scores in double then casts to
float; the knn clause only matches the k nearest docs) -- UNPINNED.
"""
from __future__ import annotations

import numpy as np

from .bm25 import BM25Index


def hybrid(index: BM25Index, qterms, knn_rows: np.ndarray, knn_score32: np.ndarray,
           w_text: float, w_knn: float, k: int, alive: np.ndarray | None = None, text32: np.ndarray | None = None,
           text64: np.ndarray | None = None):
    """Returns rows[int64 <=k], score32[<=k].  text32: the text clause's dense float32 score when it is not the plain
    `or` of qterms (the fuzzy rewrite of oracle/fuzzy.py); text64: the double sum of several clauses' float scores
    (oracle/multifield.py)."""
    if text64 is not None:
        fused = np.array(text64, dtype=np.float64)
        matched = fused > 0
    else:
        text = index.score(qterms, boost=w_text) if text32 is None else np.asarray(text32, dtype=np.float32)
        fused = text.astype(np.float64)
        matched = text > 0
    contrib = (np.float32(w_knn) * np.asarray(knn_score32, dtype=np.float32)).astype(np.float32)
    for r, c in zip(np.asarray(knn_rows, dtype=np.int64), contrib):
        if r < 0:
            continue
        fused[r] += np.float64(c)
        matched[r] = True
    if alive is not None:
        matched &= np.asarray(alive, dtype=bool)
    final = fused.astype(np.float32)
    docs = np.flatnonzero(matched)
    order = np.lexsort((docs, -final[docs].astype(np.float64)))[:k]
    return docs[order].astype(np.int64), final[docs[order]]
