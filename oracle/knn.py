"""Exact kNN oracle (test infrastructure; see oracle/__init__.py).

Follows the reference call sites:
  * row normalisation before indexing      app/main.py:1250-1251
  * query normalisation before the search  app/main.py:1536-1537, 1572-1573
  * knn clause {"knn": {"embedding": {"vector": v, "k": k}}}   app/main.py:1538-1542
  * vector field: dim 1024, space_type cosinesimil              app/main.py:563-572

Third-party semantics restated (OpenSearch k-NN plugin, nmslib engine, UNPINNED):
  score = 1 / (1 + dist), dist = 1 - cos   =>  score = 1 / (2 - cos)
  (L2 extension: score = 1 / (1 + ||q - x||^2)).

The reference search is approximate (HNSW); this oracle is the exact scan it
approximates.  "Exact" is defined as: the dot products of the STORED fp32
values accumulated in fp64 (fp32*fp32 products are exact in fp64), cos formed in
fp64 from fp64 norms, ranking by (cos desc, row asc).  Zero-norm rows or queries
have cos = 0 by definition (the reference stores an all-zero row for empty
text, app/main.py:227-228 + 1251).
"""
from __future__ import annotations

import numpy as np

COSINE = 0
L2 = 1


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """x / (||x||_2 + 1e-9) row-wise in the array's own dtype (app/main.py:1250-1251)."""
    norms = np.linalg.norm(x, axis=1, keepdims=True)
    return x / (norms + 1e-9)


def _rank(keys_desc: np.ndarray, rows: np.ndarray, k: int):
    """Order by (key desc, row asc) and keep k."""
    order = np.lexsort((rows, -keys_desc))
    return order[:k]


def cos64(X: np.ndarray, q: np.ndarray, xnorm64: np.ndarray | None = None) -> np.ndarray:
    """fp64-accumulated cosine of every row of X (fp32) with q (fp32)."""
    X64 = np.asarray(X, dtype=np.float64)
    q64 = np.asarray(q, dtype=np.float64).reshape(-1)
    # not `X64 @ q64`: a BLAS gemv reduces a row differently depending on where it sits in the matrix, so two
    # identical rows could get keys one ulp apart and the (key desc, row asc) rule would not see the tie.  The
    # row-wise pairwise sum below depends on the row's values only.
    dot = np.empty(X64.shape[0], dtype=np.float64)
    for c0 in range(0, X64.shape[0], 16384):
        dot[c0:c0 + 16384] = (X64[c0:c0 + 16384] * q64[None, :]).sum(axis=1)
    if xnorm64 is None:
        xnorm64 = np.sqrt(np.einsum("ij,ij->i", X64, X64))
    qn = np.sqrt(np.dot(q64, q64))
    den = xnorm64 * qn
    out = np.zeros_like(dot)
    np.divide(dot, den, out=out, where=den > 0)
    return out


def l2sq64(X: np.ndarray, q: np.ndarray) -> np.ndarray:
    X64 = np.asarray(X, dtype=np.float64)
    q64 = np.asarray(q, dtype=np.float64).reshape(1, -1)
    diff = X64 - q64
    return np.einsum("ij,ij->i", diff, diff)


def score_from_cos(c64: np.ndarray) -> np.ndarray:
    """OpenSearch nmslib/cosinesimil score, evaluated in fp64, emitted as fp32."""
    return (1.0 / (2.0 - np.asarray(c64, dtype=np.float64))).astype(np.float32)


def score_from_l2sq(d64: np.ndarray) -> np.ndarray:
    return (1.0 / (1.0 + np.asarray(d64, dtype=np.float64))).astype(np.float32)


def knn_exact_full(X: np.ndarray, Q: np.ndarray, k: int, metric: int = COSINE,
                   alive: np.ndarray | None = None):
    """The definition: full fp64 scan.  Returns rows[int64 B,k'], key64[B,k'], score32[B,k'].

    key64 is cos (COSINE) or squared distance (L2).  k' = min(k, #alive rows).
    Slow (O(B*N*d) in fp64); use knn_exact for anything large.
    """
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, X.shape[1])
    n = X.shape[0]
    rows_all = np.arange(n, dtype=np.int64)
    if alive is not None:
        rows_all = rows_all[np.asarray(alive, dtype=bool)]
    kk = min(k, rows_all.size)
    out_rows = np.full((Q.shape[0], kk), -1, dtype=np.int64)
    out_key = np.zeros((Q.shape[0], kk), dtype=np.float64)
    Xa = X[rows_all]
    xn = np.sqrt(np.einsum("ij,ij->i", Xa.astype(np.float64), Xa.astype(np.float64))) if metric == COSINE else None
    for b in range(Q.shape[0]):
        if metric == COSINE:
            key = cos64(Xa, Q[b], xn)
            sel = _rank(key, rows_all, kk)
        else:
            key = l2sq64(Xa, Q[b])
            sel = _rank(-key, rows_all, kk)
        out_rows[b] = rows_all[sel]
        out_key[b] = key[sel]
    score = score_from_cos(out_key) if metric == COSINE else score_from_l2sq(out_key)
    return out_rows, out_key, score


def _row_norms64(X: np.ndarray, block: int = 1024) -> np.ndarray:
    """||x||_2 of every row, accumulated in fp64, in cache-sized row blocks (no n x d fp64 temporary)."""
    out = np.empty(X.shape[0], dtype=np.float64)
    for c0 in range(0, X.shape[0], block):
        x64 = X[c0:c0 + block].astype(np.float64)
        out[c0:c0 + block] = np.sqrt(np.einsum("ij,ij->i", x64, x64))
    return out


def merge_topk(parts, k: int, metric: int = COSINE):
    """Top-k of the union of per-row-range top-k lists: parts = [(rows int64 [B, k_i], key64 [B, k_i]), ...] with
    GLOBAL row ids (-1 = empty slot).  Ranking (cos desc | d^2 asc, row asc) -- what a coordinator does with the
    per-shard lists (app/main.py:357).  Returns rows [B, k'], key64 [B, k'], score32 [B, k'], k' = min(k, #entries)."""
    rows = np.concatenate([np.asarray(r, dtype=np.int64) for r, _ in parts], axis=1)
    keys = np.concatenate([np.asarray(v, dtype=np.float64) for _, v in parts], axis=1)
    B = rows.shape[0]
    kk = min(k, int((rows >= 0).sum(axis=1).min())) if rows.size else 0
    out_rows = np.full((B, kk), -1, dtype=np.int64)
    out_key = np.zeros((B, kk), dtype=np.float64)
    for b in range(B):
        ok = rows[b] >= 0
        r, v = rows[b][ok], keys[b][ok]
        sel = _rank(v if metric == COSINE else -v, r, kk)
        out_rows[b], out_key[b] = r[sel], v[sel]
    score = score_from_cos(out_key) if metric == COSINE else score_from_l2sq(out_key)
    return out_rows, out_key, score


def knn_exact_stream(chunks, Q: np.ndarray, k: int, metric: int = COSINE):
    """knn_exact over a corpus that arrives as (first_global_row, X_chunk) pieces (a 10M x 1024 corpus does not fit
    one numpy array comfortably): exact top-k per piece, merged by the same total order.  Identical to knn_exact on
    the concatenation -- the top-k of a union is the top-k of the per-part top-k lists."""
    parts = []
    for base, X in chunks:
        r, key, _ = knn_exact(X, Q, k, metric)
        parts.append((np.where(r >= 0, r + int(base), -1), key))
    return merge_topk(parts, k, metric)


def knn_exact(X: np.ndarray, Q: np.ndarray, k: int, metric: int = COSINE,
              alive: np.ndarray | None = None, chunk: int = 262144, guard: int = 8):
    """Same result as knn_exact_full, computed as an fp32 BLAS prefilter plus an
    fp64 re-evaluation of every row whose fp32 score is within a proven error
    band of the k-th best.

    Proof sketch: |fp32 score - fp64 score| <= eps for every row (eps below is a
    deliberately loose bound for d <= 4096 and |score| <= ~1).  Let b_k be the
    k-th best fp32 score.  At least k rows have exact score >= b_k - eps, so the
    exact k-th best s_k >= b_k - eps; a row in the exact top-k therefore has
    fp32 score >= b_k - 2*eps.  Re-ranking all such rows in fp64 is exact.
    """
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, X.shape[1])
    n, d = X.shape
    B = Q.shape[0]
    alive_mask = None if alive is None else np.asarray(alive, dtype=bool)
    n_alive = n if alive_mask is None else int(alive_mask.sum())
    kk = min(k, n_alive)
    out_rows = np.full((B, kk), -1, dtype=np.int64)
    out_key = np.zeros((B, kk), dtype=np.float64)
    if kk == 0:
        return out_rows, out_key, np.zeros((B, 0), dtype=np.float32)

    # fp32 surrogate: cosine -> dot(x, qhat) / ||x||;  L2 -> dot(x,q) - 0.5||x||^2 (monotone in -d^2)
    xn32 = _row_norms64(X)
    if metric == COSINE:
        qn = np.linalg.norm(Q.astype(np.float64), axis=1)
        Qs = (Q / np.where(qn > 0, qn, 1.0)[:, None]).astype(np.float32)
        inv = np.zeros(n, dtype=np.float32)
        np.divide(1.0, xn32, out=inv, where=xn32 > 0, casting="unsafe")
        scale = 1.0
    else:
        Qs = Q
        scale = float(max(1.0, xn32.max(initial=0.0)) * max(1.0, np.linalg.norm(Q.astype(np.float64), axis=1).max()))
    eps = 64.0 * d * 2.0 ** -24 * scale  # loose: n*u*|x||q| with a 64x cushion

    want = min(n_alive, kk * guard + 64)
    cand_rows = [[] for _ in range(B)]
    cand_sur = [[] for _ in range(B)]
    floor = np.full(B, -np.inf)  # best surrogate score any DROPPED row can have
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        S = Qs @ X[c0:c1].T  # fp32 sgemm
        if metric == COSINE:
            S *= inv[c0:c1][None, :]
        else:
            S -= (0.5 * xn32[c0:c1] ** 2).astype(np.float32)[None, :]
        if alive_mask is not None:
            S[:, ~alive_mask[c0:c1]] = -np.inf
        w = min(want, c1 - c0)
        part = np.argpartition(-S, w - 1, axis=1)[:, :w]
        kept = np.take_along_axis(S, part, axis=1)
        if w < c1 - c0:
            floor = np.maximum(floor, kept.min(axis=1))
        for b in range(B):
            cand_rows[b].append(part[b] + c0)
            cand_sur[b].append(kept[b])
    for b in range(B):
        rows = np.concatenate(cand_rows[b]).astype(np.int64)
        sur = np.concatenate(cand_sur[b])
        fin = np.isfinite(sur)
        rows, sur = rows[fin], sur[fin]
        order = np.argsort(-sur, kind="stable")
        rows, sur = rows[order], sur[order]
        b_k = sur[kk - 1]
        band = sur >= b_k - 2 * eps
        # a row dropped by a per-chunk partition has surrogate <= floor[b]; the band
        # must stay strictly above it, else fall back to the definition.
        if not (b_k - 2 * eps > floor[b]):
            r1, k1, _ = knn_exact_full(X, Q[b:b + 1], k, metric, alive)
            out_rows[b], out_key[b] = r1[0], k1[0]
            continue
        rows = rows[band]
        if metric == COSINE:
            key = cos64(X[rows], Q[b])
            sel = _rank(key, rows, kk)
        else:
            key = l2sq64(X[rows], Q[b])
            sel = _rank(-key, rows, kk)
        out_rows[b] = rows[sel]
        out_key[b] = key[sel]
    score = score_from_cos(out_key) if metric == COSINE else score_from_l2sq(out_key)
    return out_rows, out_key, score


def knn_fp32_baseline(X: np.ndarray, Q: np.ndarray, k: int, chunk: int = 1 << 20):
    """The reference-shaped CPU path that gets TIMED (not the arbiter): fp32 sgemm on
    unit rows + argpartition + sort, in row chunks to bound RAM.  Returns rows, score32."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    B = Q.shape[0]
    best_s = np.full((B, 0), -np.inf, dtype=np.float32)
    best_r = np.zeros((B, 0), dtype=np.int64)
    for c0 in range(0, X.shape[0], chunk):
        S = Q @ X[c0:c0 + chunk].T
        w = min(k, S.shape[1])
        part = np.argpartition(-S, w - 1, axis=1)[:, :w]
        ps = np.take_along_axis(S, part, axis=1)
        best_s = np.concatenate([best_s, ps], axis=1)
        best_r = np.concatenate([best_r, part + c0], axis=1)
        if best_s.shape[1] > k:
            keep = np.argpartition(-best_s, k - 1, axis=1)[:, :k]
            best_s = np.take_along_axis(best_s, keep, axis=1)
            best_r = np.take_along_axis(best_r, keep, axis=1)
    order = np.argsort(-best_s, axis=1, kind="stable")
    best_s = np.take_along_axis(best_s, order, axis=1)
    best_r = np.take_along_axis(best_r, order, axis=1)
    return best_r, (1.0 / (2.0 - best_s.astype(np.float64))).astype(np.float32)
