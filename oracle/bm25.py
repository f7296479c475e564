"""Lucene-9 BM25 restatement over CSR postings (test infrastructure).

Reference call site: the `multi_match` clauses of hybrid_search
(app/main.py:1577-1594).  For chunk documents only `unstructuredText` carries
text (app/main.py:1120-1130, 1196-1206), so best_fields' max-over-fields is the
single field score times its field boost (^3, app/main.py:1404).

Third-party semantics (Lucene 9.7 BM25Similarity, UNPINNED):
  k1 = 1.2f, b = 0.75f
  idf    = (float) ln(1 + (docCount - df + 0.5) / (df + 0.5))          (double math)
  avgdl  = (float) (sumTotalTermFreq / (double) docCount)
  inv[i] = 1f / (k1 * ((1 - b) + b * LENGTH_TABLE[i] / avgdl))         (float math)
  weight = boost * idf                                                 (float)
  score(tf, norm) = weight - weight / (1f + tf * inv[norm])            (float)
  a bool `or` query sums its term scorers in double and casts to float; a query
  token that occurs twice is two clauses.
docCount is the number of documents that have the field (>= 1 token).
Not restated: fuzziness AUTO expansion, per-shard statistics.
"""
from __future__ import annotations

import math

import numpy as np

from .smallfloat import LENGTH_TABLE, encode_lengths

K1 = np.float32(1.2)
B = np.float32(0.75)


class BM25Index:
    """CSR postings: indptr[V+1] (int64), doc[nnz] (int32, ascending within a term),
    tf[nnz] (uint16), doclen[N] (token count per doc, uint32)."""

    def __init__(self, indptr, doc, tf, doclen):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.doc = np.ascontiguousarray(doc, dtype=np.int32)
        self.tf = np.ascontiguousarray(tf, dtype=np.uint16)
        self.doclen = np.ascontiguousarray(doclen, dtype=np.uint32)
        self.n_docs = int(self.doclen.size)
        self.vocab = int(self.indptr.size - 1)
        self.df = np.diff(self.indptr).astype(np.int64)
        self.doc_count = int(np.count_nonzero(self.doclen))
        self.sum_ttf = int(self.doclen.astype(np.int64).sum())
        self.avgdl = np.float32(self.sum_ttf / float(self.doc_count)) if self.doc_count else np.float32(0)
        self.norm = encode_lengths(self.doclen)
        self.inv = self._inv_table()

    @classmethod
    def from_token_ids(cls, docs: list[list[int]], vocab: int) -> "BM25Index":
        """Build postings from per-document term-id lists (small corpora / tests)."""
        per_term: list[dict[int, int]] = [dict() for _ in range(vocab)]
        doclen = np.zeros(len(docs), dtype=np.uint32)
        for d, toks in enumerate(docs):
            doclen[d] = len(toks)
            for t in toks:
                per_term[t][d] = per_term[t].get(d, 0) + 1
        indptr = np.zeros(vocab + 1, dtype=np.int64)
        doc, tf = [], []
        for t in range(vocab):
            items = sorted(per_term[t].items())
            indptr[t + 1] = indptr[t] + len(items)
            doc.extend(i for i, _ in items)
            tf.extend(min(c, 65535) for _, c in items)
        return cls(indptr, np.array(doc, dtype=np.int32), np.array(tf, dtype=np.uint16), doclen)

    def _inv_table(self) -> np.ndarray:
        one = np.float32(1.0)
        if self.doc_count == 0:
            return np.zeros(256, dtype=np.float32)
        with np.errstate(all="ignore"):
            t = (B * LENGTH_TABLE) / self.avgdl          # float: (b * L) / avgdl
            t = (one - B) + t                             # float
            t = K1 * t                                    # float
            return (one / t).astype(np.float32)

    def idf(self, term: int) -> np.float32:
        df = int(self.df[term])
        return np.float32(math.log(1.0 + (self.doc_count - df + 0.5) / (df + 0.5)))

    def term_weights(self, qterms, boost: float) -> np.ndarray:
        """float32 weight per query term occurrence = boost * idf (float multiply)."""
        bo = np.float32(boost)
        return np.array([bo * self.idf(int(t)) if 0 <= int(t) < self.vocab else np.float32(0) for t in qterms],
                        dtype=np.float32)           # tokens outside the dictionary match nothing (skipped in score)

    def score(self, qterms, boost: float = 1.0) -> np.ndarray:
        """Dense float32 score per doc for `or` over qterms (duplicates count twice);
        0 where no term matches."""
        acc = np.zeros(self.n_docs, dtype=np.float64)
        one = np.float32(1.0)
        for t, w in zip(qterms, self.term_weights(qterms, boost)):
            t = int(t)
            if t < 0 or t >= self.vocab:
                continue
            lo, hi = self.indptr[t], self.indptr[t + 1]
            if hi == lo:
                continue
            d = self.doc[lo:hi]
            tf = self.tf[lo:hi].astype(np.float32)
            inv = self.inv[self.norm[d]]
            s = w - w / (one + tf * inv)                  # all float32
            assert s.dtype == np.float32
            acc[d] += s.astype(np.float64)                # each doc once per term: no aliasing
        return acc.astype(np.float32)


def topk(scores: np.ndarray, k: int):
    """(score desc, doc asc) over docs with score > 0."""
    docs = np.flatnonzero(scores > 0)
    order = np.lexsort((docs, -scores[docs].astype(np.float64)))[:k]
    return docs[order].astype(np.int64), scores[docs[order]]
