"""Property tests (hypothesis, CPU only) of the oracle's building blocks and of the host-side mirrors the product keeps
of them: size-independent facts that hold for every input, next to the known answers in test_oracle.py."""
import datetime as dt

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import analyzer, bm25, fusion, fuzzy, knn, smallfloat
from rassengine_b200 import hostquery, text
from rassengine_b200.sharded import shard_bounds

FAST = settings(max_examples=200, deadline=None, derandomize=True)
SLOW = settings(max_examples=40, deadline=None, derandomize=True)
words = st.text(alphabet="abcdeio", min_size=0, max_size=9)


@FAST
@given(st.integers(0, 2**31 - 1))
def test_smallfloat_rounds_down_to_four_significant_bits(i):
    """SmallFloat.intToByte4 (Lucene norms): exact below 40, otherwise the largest representable value <= i, never
    more than 1/8 below it; the host mirror used by hostquery.py encodes to the same byte."""
    b = smallfloat.int_to_byte4(i)
    back = smallfloat.byte4_to_int(b)
    assert 0 <= b <= 255 and back <= i
    if i < 40:
        assert back == i
    else:
        assert i - back < max(1, (i - 24) / 8 + 1)
    if b < 255:
        assert smallfloat.byte4_to_int(b + 1) > i                     # the next byte would overshoot
    assert int(hostquery.norm_bytes(np.array([i]))[0]) == b
    assert int(smallfloat.encode_lengths(np.array([i]))[0]) == b


@FAST
@given(st.integers(0, 2**31 - 2))
def test_smallfloat_is_monotone(i):
    assert smallfloat.int_to_byte4(i) <= smallfloat.int_to_byte4(i + 1)


@FAST
@given(words, words)
def test_osa_distance_properties(a, b):
    """Optimal string alignment (what `fuzziness: AUTO` with transpositions accepts): a metric-like distance bounded by
    the lengths; the host mirror (capped) agrees with the oracle wherever the cap allows."""
    d = fuzzy.osa_distance(a, b)
    assert d == fuzzy.osa_distance(b, a)
    assert (d == 0) == (a == b)
    assert abs(len(a) - len(b)) <= d <= max(len(a), len(b))
    for cap in (1, 2):
        h = hostquery._osa(a, b, cap)
        assert (h == d) if d <= cap else (h > cap)


@FAST
@given(words.filter(lambda w: len(w) >= 2), st.integers(0, 7))
def test_adjacent_swap_is_one_edit(w, i):
    i %= len(w) - 1
    s = w[:i] + w[i + 1] + w[i] + w[i + 2:]
    assert fuzzy.osa_distance(w, s) == (0 if s == w else 1)


@FAST
@given(st.lists(words, min_size=0, max_size=40, unique=True), words)
def test_fuzzy_expansion_is_sorted_and_within_auto_edits(vocab, token):
    vocab = sorted(vocab)
    out = fuzzy.expand(vocab, token)
    me = fuzzy.auto_max_edits(len(token))
    assert len(out) <= fuzzy.MAX_EXPANSIONS
    keys = [(-float(b), vocab[t]) for t, _, b in out]
    assert keys == sorted(keys)
    for t, ed, boost in out:
        assert ed <= me and ed == fuzzy.osa_distance(token, vocab[t])
        assert 0.0 <= float(boost) <= 1.0 and (float(boost) == 1.0) == (ed == 0)
    got = {t for t, _, _ in out}
    if len(out) < fuzzy.MAX_EXPANSIONS:                              # nothing within reach was left out
        for t, term in enumerate(vocab):
            if t not in got:
                assert term != token and (me == 0 or fuzzy.osa_distance(token, term) > me)


_TRICKY = "abzAZ019_'\".,;:/-@ \u00e9\u00fc\u0301\u0130\u03a3\u05d0\u05d1\u30ab\u30ad\u3072\u60a3\u0e2a\u0e31\u200d\u2019\u00b7\u2044\u00b0"


@FAST
@given(st.one_of(st.text(max_size=60), st.text(alphabet=_TRICKY, max_size=40)))
def test_product_analyzer_equals_oracle_analyzer(s):
    """rassengine_b200.analysis.analyze (a character scanner) and oracle.analyzer.analyze (one regular expression over
    word-break classes) are written independently; they must tokenise alike -- arbitrary text, and text drawn from
    the characters the UAX#29 rules turn on (mid-letter / mid-number punctuation, combining marks, Hebrew, kana, Han,
    Thai)."""
    got = text.analyze(s)
    assert got == analyzer.analyze(s)
    for tok in got:
        assert tok and len(tok) <= 255
        assert all(len(c.lower()) != 1 or c.lower() == c for c in tok)
    if s.isascii():      # the ASCII fast path states the same rules as the general scanner
        from rassengine_b200 import analysis
        assert analysis._analyze_unicode(s) == got


@SLOW
@given(st.integers(1, 60), st.integers(1, 8), st.integers(1, 12), st.integers(0, 2**31 - 1))
def test_knn_chunked_oracle_equals_the_definition(n, nq, k, seed):
    """knn_exact (chunked, what the GPU tests compare against) == knn_exact_full (the plain definition), duplicates and
    an all-zero row included; ids sorted by (key desc, row asc), fewer than k rows -> -1 padding."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 16)).astype(np.float32)
    if n > 3:
        X[n - 1] = X[0]
        X[n // 2] = 0
    Q = rng.standard_normal((nq, 16)).astype(np.float32)
    for metric in (knn.COSINE, knn.L2):
        r1, k1, s1 = knn.knn_exact(X, Q, k, metric=metric)
        r2, k2, s2 = knn.knn_exact_full(X, Q, k, metric=metric)
        assert np.array_equal(r1, r2) and np.array_equal(s1, s2)
        for b in range(nq):
            valid = r1[b] >= 0
            assert valid.sum() == min(k, n) and len(set(r1[b][valid].tolist())) == valid.sum()
            keys = k1[b][valid] if metric == knn.COSINE else -k1[b][valid]      # L2 reports the squared distance
            assert np.all(keys[:-1] >= keys[1:])
            ties = np.flatnonzero(keys[:-1] == keys[1:])
            assert np.all(r1[b][valid][ties] < r1[b][valid][ties + 1])


@SLOW
@given(st.integers(2, 40), st.integers(0, 2**31 - 1), st.floats(0.5, 5.0), st.floats(0.5, 3.0))
def test_fusion_is_a_boosted_sum_over_the_matching_docs(n, seed, w_text, w_knn):
    """bool.should: a doc matches when a clause matches; S = float(double(text) + double(float(w_knn * knn))); ranking
    (S desc, row asc); with no kNN rows the fusion is the BM25 ranking itself."""
    rng = np.random.default_rng(seed)
    docs = [rng.integers(0, 12, size=int(rng.integers(0, 30))).tolist() for _ in range(n)]
    idx = bm25.BM25Index.from_token_ids(docs, 12)
    q = rng.integers(0, 12, size=3).tolist()
    kk = min(5, n)
    knn_rows = rng.choice(n, size=kk, replace=False).astype(np.int64)
    knn_scores = np.sort(rng.uniform(0.4, 1.0, size=kk).astype(np.float32))[::-1]
    rows, sc = fusion.hybrid(idx, q, knn_rows, knn_scores, w_text, w_knn, k=n)
    text32 = idx.score(q, boost=w_text)
    want = {d: float(text32[d]) for d in range(n) if text32[d] > 0}
    for r, s in zip(knn_rows.tolist(), knn_scores.tolist()):
        want[r] = want.get(r, 0.0) + float(np.float32(np.float32(w_knn) * np.float32(s)))
    assert set(rows.tolist()) == set(want)
    assert [np.float32(want[r]) for r in rows.tolist()] == sc.tolist()
    order = sorted(want, key=lambda d: (-float(np.float32(want[d])), d))
    assert rows.tolist() == order
    rows0, sc0 = fusion.hybrid(idx, q, np.empty(0, dtype=np.int64), np.empty(0, dtype=np.float32), w_text, w_knn, k=n)
    tr, ts = bm25.topk(text32, n)
    assert rows0.tolist() == tr.tolist() and sc0.tolist() == ts.tolist()


@FAST
@given(st.integers(0, 10**9), st.integers(1, 16))
def test_shard_bounds_partition_the_rows(n, world):
    cuts = [shard_bounds(n, world, r) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    for (lo, hi), (lo2, _) in zip(cuts, cuts[1:]):
        assert lo <= hi == lo2
    sizes = [hi - lo for lo, hi in cuts]
    assert max(sizes) <= -(-n // world)


@FAST
@given(st.datetimes(min_value=dt.datetime(1971, 1, 1), max_value=dt.datetime(2090, 12, 31)),
       st.sampled_from("yMwdhms"), st.integers(0, 30), st.sampled_from("+-"))
def test_date_math_never_raises_and_moves_the_right_way(now, unit, n, sign):
    """`now-1y` style bounds of the reference's range clauses (app/main.py:1890-1900): any calendar day, leap days included."""
    now = now.replace(tzinfo=dt.timezone.utc)
    got = hostquery._parse_date(f"now{sign}{n}{unit}", now)
    assert got is not None
    if n == 0:
        assert abs((got - now).total_seconds()) <= 3 * 86400          # month/year arithmetic may clamp the day
    elif sign == "+":
        assert got > now - dt.timedelta(days=4)
    else:
        assert got < now + dt.timedelta(days=4)
    assert hostquery._parse_date("now", now) == now
    assert hostquery._parse_date("not a date", now) is None


def _scalar():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
    spec = importlib.util.spec_from_file_location("make_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@SLOW
@given(st.integers(1, 25), st.integers(0, 2**31 - 1), st.sampled_from([1.0, 1.5, 4.5]))
def test_bm25_oracle_equals_the_independent_scalar_scorer(n, seed, boost):
    """oracle/bm25.py (numpy, vectorised) against the pure-Python scalar restatement that made the golden fixtures
    (struct-rounded float32 arithmetic), on random corpora with empty documents, repeated query terms, absent terms and
    lengths on both sides of the SmallFloat quantisation."""
    mg = _scalar()
    rng = np.random.default_rng(seed)
    docs = [rng.integers(0, 9, size=int(rng.choice([0, 3, 17, 41, 90]))).tolist() for _ in range(n)]
    if not any(docs):
        docs[0] = [1, 2]
    q = rng.integers(0, 11, size=int(rng.integers(1, 5))).tolist()          # ids 9, 10 never occur
    idx = bm25.BM25Index.from_token_ids(docs, 11)
    assert [mg.f32(v) for v in idx.score(q, boost=boost).tolist()] == mg.scalar_bm25(docs, q, boost)


@SLOW
@given(st.integers(2, 14), st.integers(0, 2**31 - 1))
def test_multifield_oracle_equals_the_independent_scalar_scorer(n, seed):
    """oracle/multifield.py + oracle/fuzzy.py against the scalar scorer: random documents over two analysed fields and
    one keyword field, words drawn from a pool full of one- and two-edit neighbours."""
    from oracle import multifield
    mg = _scalar()
    rng = np.random.default_rng(seed)
    pool = ["pain", "pian", "paint", "gain", "of", "on", "diabetes", "diabetis", "diabetse", "dibetes", "chest", "chst",
            "active", "inactive", "ab", "abc", "abcd"]
    pick = lambda lo, hi: " ".join(pool[int(i)] for i in rng.integers(0, len(pool), size=int(rng.integers(lo, hi))))
    docs = []
    for _ in range(n):
        d = {}
        if rng.random() < 0.8:
            d["t"] = pick(1, 5)
        if rng.random() < 0.8:
            d["b"] = pick(1, 50)
        if rng.random() < 0.6:
            d["g"] = pick(1, 3)
        docs.append(d)
    if not any("t" in d for d in docs):
        docs[0]["t"] = "pain"
    types = {"t": "text", "b": "text", "g": "keyword"}
    fields = multifield.build(docs, types)
    toks = {f: [(([d[f]] if f in d else []) if k == "keyword" else d.get(f, "").split()) for d in docs]
            for f, k in types.items()}
    query = pick(1, 4)
    text_specs, kw_specs = [("t", 3.0), ("b", 1.0)], [("g", 2.0)]
    want = [0.0] * n
    for specs, cb, fz in ((text_specs, 1.5, True), (kw_specs, 1.0, False)):
        best = [0.0] * n
        for fname, fb in specs:
            if fname not in fields:
                continue
            kw = types[fname] == "keyword"
            sc = mg.scalar_field_score(toks[fname], kw, [query] if kw else query.split(),
                                       mg.f32(mg.f32(cb) * mg.f32(fb)), fz and not kw)
            best = [max(x, y) for x, y in zip(best, sc)]
        want = [x + y for x, y in zip(want, best)]
    got = multifield.text_total(fields, [(query, text_specs, 1.5, True), (query, kw_specs, 1.0, False)], n)
    assert got.tolist() == want, query


@SLOW
@given(st.integers(1, 30), st.integers(0, 2**31 - 1), st.sampled_from([1.0, 4.5]))
def test_host_fuzzy_rewrite_equals_the_oracle(n, seed, boost):
    """TextField.fuzzy_weighted_terms (product: ranks the device dictionary scan's hits, blends the statistics, builds
    the (term, weight) list the kernel consumes) against oracle/fuzzy.py.  The dictionary scan itself -- the device
    kernel in production, compared with the oracle in test_gpu_hybrid.py -- is played by a brute-force scan here.
    Term ids differ (insertion order vs sorted), so terms are compared by string."""
    rng = np.random.default_rng(seed)
    pool = ["pain", "pian", "paint", "gain", "of", "on", "diabetes", "diabetis", "diabetse", "dibetes", "chest", "chst",
            "active", "inactive", "ab", "abc", "abcd", "pains", "spain"]
    fld = text.TextField()
    docs_tokens = []
    for row in range(n):
        toks = [pool[int(i)] for i in rng.integers(0, len(pool), size=int(rng.integers(0, 12)))]
        docs_tokens.append(toks)
        fld.set_row_tokens(row, toks)
    fld.postings(n)
    by_id = fld.terms_in_id_order()

    def expand(tok, me):
        hits = [(t, fuzzy.osa_distance(tok, term)) for t, term in enumerate(by_id)]
        hits = [(t, e) for t, e in hits if e <= me]
        return np.array([t for t, _ in hits], dtype=np.int64), np.array([e for _, e in hits], dtype=np.int64)

    terms = sorted({t for toks in docs_tokens for t in toks})
    tid = {t: i for i, t in enumerate(terms)}
    idx = bm25.BM25Index.from_token_ids([[tid[t] for t in toks] for toks in docs_tokens], len(terms))
    query = " ".join(pool[int(i)] for i in rng.integers(0, len(pool), size=3)) + " zzzz q"
    ids, ws = fld.fuzzy_weighted_terms(query, boost, expand)
    o_ids, o_ws = fuzzy.weighted_terms(idx, terms, analyzer.analyze(query), boost)
    assert [by_id[t] for t in ids] == [terms[t] for t in o_ids]
    assert [np.float32(w) for w in ws] == list(o_ws)
    e_ids, e_ws = fld.exact_weighted_terms(analyzer.analyze(query), boost)
    want = [(t, np.float32(np.float32(boost) * idx.idf(tid[t]))) for t in analyzer.analyze(query) if t in tid]
    assert [(by_id[t], w) for t, w in zip(e_ids, e_ws)] == want


def _bf16(a):
    """float32 -> nearest-even bfloat16, returned as float32 (what store_convert_kernel / query_prep_kernel keep)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


@SLOW
@given(st.integers(1, 200), st.sampled_from([8, 64, 256, 1024]), st.integers(0, 2**31 - 1), st.sampled_from([1e-3, 1.0, 50.0]))
def test_certificate_error_allowance_bounds_the_bf16_scan_error(n, d, seed, scale):
    """The premise of the certificate in finish.cu: for every row, |scan key - exact key| <= eps with
    eps = rho_x (1 + rho_q) + rho_q + d 2^-23, rho = relative bf16 rounding residual of the rows / of the query
    (Cauchy-Schwarz), the last term covering the fp32 accumulation and the scale multiply.  Checked in numpy: bf16
    operands, float32 accumulation, key = dot * float(1 / ||x||), against the fp64 cosine of the stored fp32 values."""
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) * scale * rng.uniform(0.2, 3.0, size=(n, 1))).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    X64, q64 = X.astype(np.float64), q.astype(np.float64)
    xn = np.sqrt((X64 * X64).sum(axis=1))
    q_hat = (q64 / np.sqrt((q64 * q64).sum())).astype(np.float32)
    X16, q16 = _bf16(X), _bf16(q_hat)
    rho_x = float((np.sqrt(((X64 - X16.astype(np.float64)) ** 2).sum(axis=1)) / xn).max())
    qh64 = q_hat.astype(np.float64)
    rho_q = float(np.sqrt(((qh64 - q16.astype(np.float64)) ** 2).sum()) / np.sqrt((qh64 * qh64).sum()))
    eps = (rho_x * (1.0 + rho_q) + rho_q + d * 2.0 ** -23) * 1.0001
    acc = np.zeros(n, dtype=np.float32)
    for j in range(d):                                   # float32 accumulation, the plainest order
        acc = (acc + X16[:, j] * q16[j]).astype(np.float32)
    key = acc * (1.0 / xn).astype(np.float32)
    exact = knn.cos64(X, q)
    assert np.abs(key.astype(np.float64) - exact).max() <= eps
    assert rho_x <= 2.0 ** -8 and rho_q <= 2.0 ** -8      # bf16 keeps 8 significant bits


@SLOW
@given(st.integers(1, 200), st.sampled_from([8, 64, 256, 1024]), st.integers(0, 2**31 - 1), st.sampled_from([1e-2, 1.0, 20.0]))
def test_certificate_error_allowance_l2(n, d, seed, scale):
    """Same premise for the L2 metric: scan key = dot(bf16 x, bf16 q) - fl32(||x||^2 / 2), exact surrogate
    (||q||^2 - d^2) / 2 = x.q - ||x||^2 / 2; the allowance scales with max ||x|| * ||q|| and adds the rounding of the
    per-row offset (finish.cu)."""
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) * scale * rng.uniform(0.2, 3.0, size=(n, 1))).astype(np.float32)
    q = (rng.standard_normal(d) * rng.uniform(0.1, 4.0)).astype(np.float32)
    X64, q64 = X.astype(np.float64), q.astype(np.float64)
    n2 = (X64 * X64).sum(axis=1)
    xn, qn = np.sqrt(n2), float(np.sqrt((q64 * q64).sum()))
    X16, q16 = _bf16(X), _bf16(q)
    rho_x = float((np.sqrt(((X64 - X16.astype(np.float64)) ** 2).sum(axis=1)) / xn).max())
    rho_q = float(np.sqrt(((q64 - q16.astype(np.float64)) ** 2).sum()) / qn)
    xm = float(xn.max())
    e = rho_x * (1.0 + rho_q) + rho_q + d * 2.0 ** -23
    eps = (e * xm * qn * 1.0001 + 6.0e-8 * xm * xm) * 1.0001
    acc = np.zeros(n, dtype=np.float32)
    for j in range(d):
        acc = (acc + X16[:, j] * q16[j]).astype(np.float32)
    key = (acc + (-0.5 * n2).astype(np.float32)).astype(np.float32)
    exact = (X64 * q64[None, :]).sum(axis=1) - 0.5 * n2
    assert np.abs(key.astype(np.float64) - exact).max() <= eps
