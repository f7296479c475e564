#!/usr/bin/env python
"""Generates the committed golden fixtures in this directory.

The reference holds no golden vectors for the retrieval path (its only test,
tests/test_main.py:15-35, checks HTTP status and key presence) and its executors
(OpenSearch/Lucene/nmslib) cannot run here, so these fixtures come from
  (a) hand-computed literals (knn_tiny, smallfloat table excerpts), and
  (b) the numpy oracle under oracle/, cross-checked in this script against an
      independent pure-Python scalar restatement (struct-rounded float32 maths).
Run from the repo root:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bm25, fusion, knn, smallfloat, synth  # noqa: E402


def f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


# ----------------------------------------------------------------------------------
# independent scalar BM25 (pure Python) -- Lucene 9 formula, float32 rounding by struct
# ----------------------------------------------------------------------------------
def scalar_bm25(docs: list[list[int]], qterms: list[int], boost: float) -> list[float]:
    n_with = sum(1 for d in docs if d)
    sum_ttf = sum(len(d) for d in docs)
    avgdl = f32(sum_ttf / float(n_with))
    k1, b = f32(1.2), f32(0.75)
    out = []
    for d in docs:
        enc = smallfloat.int_to_byte4(len(d))
        L = float(smallfloat.byte4_to_int(enc))
        inv = f32(1.0 / f32(k1 * f32(f32(1.0 - b) + f32(f32(b * L) / avgdl))))
        tot = 0.0
        for t in qterms:
            df = sum(1 for dd in docs if t in dd)
            if df == 0:
                continue
            tf = float(d.count(t))
            if tf == 0:
                continue
            idf = f32(math.log(1.0 + (n_with - df + 0.5) / (df + 0.5)))
            w = f32(f32(boost) * idf)
            s = f32(w - f32(w / f32(1.0 + f32(tf * inv))))
            tot += s
        out.append(f32(tot))
    return out


def make_bm25_micro():
    # 5 docs over a 12-term vocabulary; doc 3 has 45 tokens (> 39) so SmallFloat quantises it to 44
    docs = [
        [0, 1, 2, 3, 1],
        [4, 5, 1, 1, 1, 6],
        [7, 8, 9],
        [1, 2] * 20 + [10, 10, 10, 11, 0],
        [11, 3, 3, 5],
    ]
    queries = [[1, 3], [10, 1, 1], [9], [6, 7, 11, 0, 2]]
    idx = bm25.BM25Index.from_token_ids(docs, 12)
    cases = []
    for q in queries:
        for boost in (1.0, 4.5):
            want = scalar_bm25(docs, q, boost)
            got = idx.score(q, boost=boost)
            assert [f32(v) for v in got.tolist()] == want, (q, boost, got.tolist(), want)
            cases.append({"qterms": q, "boost": boost, "scores": want})
    assert smallfloat.byte4_to_int(smallfloat.int_to_byte4(45)) == 44
    return {"docs": docs, "vocab": 12, "cases": cases,
            "avgdl": float(idx.avgdl), "norm_bytes": idx.norm.tolist()}


def make_fusion_micro(bm):
    docs = bm["docs"]
    idx = bm25.BM25Index.from_token_ids(docs, bm["vocab"])
    # kNN set: doc 2 (kNN-only for query [1,3]: it has no query term), doc 0 (both), doc 4 (both)
    knn_rows = np.array([2, 0, 4], dtype=np.int64)
    knn_scores = np.array([0.9, 0.75, 0.6], dtype=np.float32)
    cases = []
    for w_text, w_knn in ((4.5, 2.0), (3.0, 1.5)):
        rows, sc = fusion.hybrid(idx, [1, 3], knn_rows, knn_scores, w_text, w_knn, k=5)
        # independent recomputation
        text = scalar_bm25(docs, [1, 3], w_text)
        tot = {}
        for d, t in enumerate(text):
            if t > 0:
                tot[d] = float(t)
        for r, s in zip(knn_rows.tolist(), knn_scores.tolist()):
            tot[r] = tot.get(r, 0.0) + f32(f32(w_knn) * f32(s))
        want = sorted(((-f32(v), d) for d, v in tot.items()))[:5]
        assert [d for _, d in want] == rows.tolist(), (want, rows)
        assert [-v for v, _ in want] == [f32(x) for x in sc.tolist()]
        cases.append({"qterms": [1, 3], "knn_rows": knn_rows.tolist(),
                      "knn_scores": [f32(x) for x in knn_scores.tolist()],
                      "w_text": w_text, "w_knn": w_knn, "k": 5,
                      "rows": rows.tolist(), "scores": [f32(x) for x in sc.tolist()]})
    return {"cases": cases}


def data_digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_knn_seeded(name: str, n: int, nq: int, k: int, d: int = 1024, clustered: bool = False, dup_pairs: int = 0):
    X = synth.embeddings(n, d, synth.SEED_CORPUS)
    pairs = synth.plant_duplicates(X, dup_pairs) if dup_pairs else []
    if clustered:
        Q = synth.clustered_queries(X, nq, synth.SEED_QUERIES)
    else:
        Q = synth.embeddings(nq, d, synth.SEED_QUERIES)
    rows, key, score = knn.knn_exact(X, Q, k)
    # cross-check a sample against the definition
    sample = list(range(0, nq, max(1, nq // 8)))
    r2, k2, s2 = knn.knn_exact_full(X, Q[sample], k)
    assert np.array_equal(rows[sample], r2) and np.array_equal(score[sample], s2)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), rows=rows.astype(np.int32), cos=key, score=score,
                        meta=np.array(json.dumps({"n": n, "nq": nq, "k": k, "d": d, "clustered": clustered,
                                                  "dup_pairs": pairs, "x_sha256": data_digest(X),
                                                  "q_sha256": data_digest(Q)})))
    print(name, "rows sha", data_digest(rows)[:16])


def main():
    bm = make_bm25_micro()
    fu = make_fusion_micro(bm)
    with open(os.path.join(HERE, "bm25_micro.json"), "w") as f:
        json.dump(bm, f, indent=1)
    with open(os.path.join(HERE, "fusion_micro.json"), "w") as f:
        json.dump(fu, f, indent=1)
    # seeded kNN cases: small + cfg-1, plus a clustered/duplicate stress case
    make_knn_seeded("knn_small", n=20000, nq=64, k=10)
    make_knn_seeded("knn_clustered_dups", n=20000, nq=64, k=10, clustered=True, dup_pairs=16)
    make_knn_seeded("knn_cfg1", n=100000, nq=1000, k=10)
    make_knn_seeded("knn_small_k100", n=20000, nq=16, k=100)


if __name__ == "__main__":
    main()
