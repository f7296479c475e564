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

from oracle import bm25, fusion, knn, multifield, smallfloat, synth  # noqa: E402


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


# ----------------------------------------------------------------------------------
# independent scalar multi-field scorer: per-field statistics, fuzziness AUTO, dis-max over fields, sum over clauses
# ----------------------------------------------------------------------------------
def scalar_osa(a: str, b: str) -> int:
    """Optimal string alignment distance by memoised recursion (oracle/fuzzy.py uses the row DP)."""
    memo = {}

    def go(i, j):
        if i == 0 or j == 0:
            return i + j
        if (i, j) not in memo:
            v = min(go(i - 1, j) + 1, go(i, j - 1) + 1, go(i - 1, j - 1) + (a[i - 1] != b[j - 1]))
            if i > 1 and j > 1 and a[i - 1] == b[j - 2] and a[i - 2] == b[j - 1]:
                v = min(v, go(i - 2, j - 2) + 1)
            memo[(i, j)] = v
        return memo[(i, j)]
    return go(len(a), len(b))


def scalar_field_score(docs_tokens: list[list[str]], keyword: bool, tokens: list[str], boost: float, fuzzy_auto: bool):
    """float score per document of one field of one clause; docs_tokens[d] = the field's tokens ([] = field absent)."""
    n_with = sum(1 for t in docs_tokens if t)
    lens = [min(len(t), 1) if keyword else len(t) for t in docs_tokens]
    avgdl = f32(sum(lens) / float(n_with))
    k1, b = f32(1.2), f32(0.75)
    df = {}
    for toks in docs_tokens:
        for t in set(toks):
            df[t] = df.get(t, 0) + 1
    idf_of = lambda d_f: f32(math.log(1.0 + (n_with - d_f + 0.5) / (d_f + 0.5)))
    weighted = []                                     # (term, float weight), one entry per term query
    for tok in tokens:
        edits = 0 if (not fuzzy_auto or len(tok) <= 2) else (1 if len(tok) <= 5 else 2)
        cands = []
        for term in df:
            if term == tok:
                cands.append((1.0, term))
            elif edits and abs(len(term) - len(tok)) <= edits:
                ed = scalar_osa(tok, term)
                if ed <= edits:
                    cands.append((f32(1.0 - f32(f32(ed) / f32(min(len(term), len(tok))))), term))
        if not cands:
            continue
        cands.sort(key=lambda c: (-c[0], c[1]))
        cands = cands[:50]
        idf = idf_of(max(df[t] for _, t in cands))    # blended: the largest df among the survivors
        for tb, term in cands:
            weighted.append((term, f32(f32(f32(boost) * tb) * idf) if fuzzy_auto else f32(f32(boost) * idf)))
    out = []
    for toks, ln in zip(docs_tokens, lens):
        L = float(smallfloat.byte4_to_int(smallfloat.int_to_byte4(ln)))
        inv = f32(1.0 / f32(k1 * f32(f32(1.0 - b) + f32(f32(b * L) / avgdl))))
        tot = 0.0
        for term, w in weighted:
            tf = float(toks.count(term))
            if tf:
                tot += f32(w - f32(w / f32(1.0 + f32(tf * inv))))
        out.append(f32(tot))
    return out


def make_multifield_micro():
    long_body = " ".join(["patient", "reports"] * 22 + ["severe", "chronic", "pian"])  # 47 tokens -> quantised to 46
    docs = [
        {"title": "type two diabetes", "body": "patient with diabetes mellitus on metformin", "tag": "active"},
        {"title": "diabetis follow up", "body": "diabetse controlled of late", "tag": "resolved"},
        {"title": "chest pain", "body": "chronic chest pain radiating to the arm", "tag": "active"},
        {"body": long_body, "tag": "type two"},
        {"title": "pain clinic", "tag": "chest pain"},
        {"title": "of", "body": "of on to"},
        {"body": "no complaints today"},
        {"title": "diabetes diabetes diabetes", "body": "pain", "tag": "inactive"},
        {"tag": "active"},
        {"title": "chronic chest pian", "body": "diabetic neuropathy with pain"},
    ]
    types = {"title": "text", "body": "text", "tag": "keyword"}
    text_specs, kw_specs = [["title", 3.0], ["body", 1.0]], [["tag", 2.0]]
    fields = multifield.build(docs, types)
    n = len(docs)
    toks = {f: [(([d[f]] if f in d else []) if k == "keyword" else d.get(f, "").split()) for d in docs]
            for f, k in types.items()}
    cases = []
    for query in ("diabetes pain", "active", "chronic chest pian", "type two", "of", "zzzz", "diabetes diabetes"):
        for w_text, w_kw in ((1.5, 1.0), (1.0, 0.5)):
            want = [0.0] * n
            for specs, cb, fz in ((text_specs, w_text, True), (kw_specs, w_kw, False)):
                best = [0.0] * n
                for fname, fb in specs:
                    kw = types[fname] == "keyword"
                    sc = scalar_field_score(toks[fname], kw, [query] if kw else query.split(),
                                            f32(f32(cb) * f32(fb)), fz and not kw)
                    best = [max(x, y) for x, y in zip(best, sc)]
                want = [x + y for x, y in zip(want, best)]
            got = multifield.text_total(fields, [(query, [tuple(x) for x in text_specs], w_text, True),
                                                 (query, [tuple(x) for x in kw_specs], w_kw, False)], n)
            assert got.tolist() == want, (query, got.tolist(), want)
            cases.append({"query": query, "w_text": w_text, "w_keyword": w_kw, "totals": want})
    # the cases exercise what they are meant to: a transposition counts as one edit, 2-character tokens do not expand,
    # a keyword value only matches as a whole string, the 47-token body is quantised
    by = {(c["query"], c["w_text"]): c["totals"] for c in cases}
    assert by[("chronic chest pian", 1.5)][2] > 0 and by[("chronic chest pian", 1.5)][9] > by[("chronic chest pian", 1.5)][2]
    assert [i for i, v in enumerate(by[("of", 1.5)]) if v > 0] == [1, 5]
    assert by[("type two", 1.5)][3] > 0 and by[("type two", 1.5)][0] > 0          # keyword value / analysed title
    act = by[("active", 1.5)]                                                      # "inactive" is 2 edits away but the
    assert act[7] == 0 and act[0] == act[2] == act[8] > 0                          # keyword clause is not fuzzy
    assert all(v == 0 for v in by[("zzzz", 1.5)])
    assert len(long_body.split()) == 47 and smallfloat.byte4_to_int(smallfloat.int_to_byte4(47)) == 46
    return {"docs": docs, "types": types, "text_fields": text_specs, "keyword_fields": kw_specs, "cases": cases}


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
    with open(os.path.join(HERE, "multifield_micro.json"), "w") as f:
        json.dump(make_multifield_micro(), f, indent=1)
    # seeded kNN cases: small + cfg-1, plus a clustered/duplicate stress case
    make_knn_seeded("knn_small", n=20000, nq=64, k=10)
    make_knn_seeded("knn_clustered_dups", n=20000, nq=64, k=10, clustered=True, dup_pairs=16)
    make_knn_seeded("knn_cfg1", n=100000, nq=1000, k=10)
    make_knn_seeded("knn_small_k100", n=20000, nq=16, k=100)


if __name__ == "__main__":
    main()
