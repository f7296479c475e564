"""Seeded synthetic inputs shaped like the reference's data (test infrastructure).

Embeddings: "mxbai-embed-large-shaped" = 1024-d fp32, row-normalised exactly as
app/main.py:1250-1251.  Text: lowercase tokens t00000.. drawn Zipf(s) so that the
reference chunker's `str.split()` (app/main.py:2160-2170) and the standard
analyzer agree.  Shapes/seeds follow SURVEY.md section 8d.
"""
from __future__ import annotations

import numpy as np

from .knn import normalize_rows

SEED_CORPUS = 1234
SEED_QUERIES = 5678
SEED_TEXT = 4242


def embeddings(n: int, d: int = 1024, seed: int = SEED_CORPUS) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    return normalize_rows(x).astype(np.float32)


def clustered_queries(X: np.ndarray, nq: int, seed: int = SEED_QUERIES, noise: float = 0.05) -> np.ndarray:
    """Queries = corpus rows + N(0, noise^2), renormalised: stresses near-ties."""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, X.shape[0], size=nq)
    q = X[pick] + noise * rng.standard_normal((nq, X.shape[1]), dtype=np.float32)
    return normalize_rows(q.astype(np.float32)).astype(np.float32)


def plant_duplicates(X: np.ndarray, n_pairs: int, seed: int = 99) -> list[tuple[int, int]]:
    """Overwrite n_pairs rows with exact copies of other rows (in place)."""
    rng = np.random.default_rng(seed)
    pairs = []
    for _ in range(n_pairs):
        a, b = (int(v) for v in rng.choice(X.shape[0], size=2, replace=False))
        X[b] = X[a]
        pairs.append((min(a, b), max(a, b)))
    return pairs


def _zipf_cdf(vocab: int, s: float, damp_top: int = 0, damp: float = 1.0) -> np.ndarray:
    p = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    if damp_top:
        p[:damp_top] *= damp
    return np.cumsum(p / p.sum())


def text_corpus(n_docs: int, vocab: int = 30000, seed: int = SEED_TEXT, median_len: int = 120,
                max_len: int = 512, sigma: float = 0.6, zipf_s: float = 1.07):
    """Returns CSR postings (indptr int64[V+1], doc int32[nnz], tf uint16[nnz]) and
    doclen uint32[n_docs].  Doc length ~ clipped lognormal (max = CHUNK_SIZE 512,
    app/main.py:79)."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(rng.lognormal(np.log(median_len), sigma, size=n_docs)), 1, max_len).astype(np.int64)
    total = int(lens.sum())
    cdf = _zipf_cdf(vocab, zipf_s)
    terms = np.searchsorted(cdf, rng.random(total), side="left").astype(np.int64)
    np.minimum(terms, vocab - 1, out=terms)
    docs = np.repeat(np.arange(n_docs, dtype=np.int64), lens)
    key = terms * n_docs + docs
    uniq, counts = np.unique(key, return_counts=True)
    t_of = uniq // n_docs
    d_of = (uniq - t_of * n_docs).astype(np.int32)
    indptr = np.zeros(vocab + 1, dtype=np.int64)
    np.add.at(indptr, t_of + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, d_of, np.minimum(counts, 65535).astype(np.uint16), lens.astype(np.uint32)


def text_queries(nq: int, vocab: int = 30000, seed: int = SEED_TEXT + 1, zipf_s: float = 1.07,
                 min_terms: int = 3, max_terms: int = 12) -> list[list[int]]:
    """Term-id lists; the 50 most frequent terms are down-weighted 10x (stop-word-like)."""
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(vocab, zipf_s, damp_top=50, damp=0.1)
    out = []
    for _ in range(nq):
        m = int(rng.integers(min_terms, max_terms + 1))
        t = np.minimum(np.searchsorted(cdf, rng.random(m), side="left"), vocab - 1)
        out.append([int(v) for v in t])
    return out


def token(term_id: int) -> str:
    return f"t{term_id:05d}"


def docs_as_text(indptr, doc, tf, n_docs: int) -> list[str]:
    """Materialise postings back into whitespace-joined documents (small corpora only)."""
    words: list[list[str]] = [[] for _ in range(n_docs)]
    for t in range(len(indptr) - 1):
        for j in range(indptr[t], indptr[t + 1]):
            words[int(doc[j])].extend([token(t)] * int(tf[j]))
    return [" ".join(w) for w in words]
