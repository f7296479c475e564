"""Host-side evaluation of the query-DSL shapes that are NOT the retrieval hot path.

The hot path (knn, hybrid bool.should of multi_match + knn) runs on the GPU (client.py -> rass_search_*).  Everything
else the reference sends to OpenSearch is keyword bookkeeping with no vector math; it is evaluated here with numpy
over the host copy of the postings, so that the remaining `OpenSearchIndexer` methods and the patient-name lookup
work against the same client object:

  exact_match_search       multi_match type phrase                                   app/main.py:1480-1525
  hybrid_structured_search multi_match phrase_prefix + knn + filters                 app/main.py:1707-1775
  aggregate_search         size 0 + terms aggregations (+ filter)                    app/main.py:1777-1810
  comparison_search        multi_match best_fields AUTO + aggs                       app/main.py:1812-1866
  temporal_search          must[multi_match, bool.should[range]] + sort              app/main.py:1868-1922
  explanatory_search       must[multi_match best_fields AUTO]                        app/main.py:1924-1967
  entity_specific_search   multi_match type phrase, operator and                     app/main.py:2029-2075
  document_fetch_search    bool.filter + collapse                                    app/main.py:2079-2110
  resolve_patient_ids_from_name   should[term .keyword, match_phrase, match AUTO and] + collapse + _source
                                                                                     app/main.py:2709-2744

Scoring follows the same restated Lucene semantics as the device kernels (per-field BM25 with SmallFloat norms,
float ops one by one, clause sums in double; phrase = BM25 of the phrase frequency with the summed idf of its terms;
bool = sum of the matching scoring clauses, filter / must_not do not score).  Third party and unpinned, like the rest.
Every evaluator returns a dense (score float32 [rows], match bool [rows]) pair.
"""
from __future__ import annotations

import datetime as _dt
import math
import re

import numpy as np

from .text import TextField, analyze

K1 = np.float32(1.2)
B = np.float32(0.75)


# ---- Lucene SmallFloat.intToByte4 (4 significant bits above 23, rounding down), the norm byte of a field length --
def _int4_to_long(e: int) -> int:
    bits, shift = e & 7, (e >> 3) - 1
    return bits if shift == -1 else (bits | 8) << shift


LENGTH_TABLE = np.array([b if b < 24 else 24 + _int4_to_long(b - 24) for b in range(256)], dtype=np.int64)


def norm_bytes(lengths: np.ndarray) -> np.ndarray:
    """largest byte whose decoded length is <= the length (the table is strictly increasing)"""
    return (np.searchsorted(LENGTH_TABLE, np.asarray(lengths, dtype=np.int64), side="right") - 1).astype(np.uint8)


class FieldView:
    """Host statistics of one analysed / keyword field as of the last TextIndex.postings()."""

    def __init__(self, fld: TextField, keyword: bool, n_rows: int):
        self.fld = fld
        self.keyword = keyword
        self.indptr, self.doc, self.tf, doclen = fld.postings(n_rows)
        if keyword:
            doclen = (doclen > 0).astype(np.uint32)
        self.df = np.diff(self.indptr)
        self.doc_count = int(np.count_nonzero(doclen))
        self.norm = norm_bytes(doclen)
        one = np.float32(1.0)
        if self.doc_count:
            avgdl = np.float32(int(doclen.astype(np.int64).sum()) / float(self.doc_count))
            with np.errstate(all="ignore"):
                t = (B * LENGTH_TABLE.astype(np.float32)) / avgdl
                self.inv = (one / (K1 * ((one - B) + t))).astype(np.float32)
        else:
            self.inv = np.zeros(256, dtype=np.float32)

    def idf(self, df: int) -> np.float32:
        return np.float32(math.log(1.0 + (self.doc_count - df + 0.5) / (df + 0.5)))

    def term_scores(self, term: int, w: np.float32):
        """(rows, float32 scores) of one term query."""
        lo, hi = self.indptr[term], self.indptr[term + 1]
        rows = self.doc[lo:hi]
        tf = self.tf[lo:hi].astype(np.float32)
        s = w - w / (np.float32(1.0) + tf * self.inv[self.norm[rows]])
        return rows, s.astype(np.float32)


_DATE_MATH = re.compile(r"^now(?:([+-])(\d+)([yMwdhHms]))?(?:/[yMwdhHms])?$")


def _parse_date(v, now: _dt.datetime):
    """ISO-8601 dates / datetimes, epoch millis, and the date math the reference uses ("now", "now-1y")."""
    if isinstance(v, (int, float)):
        return _dt.datetime.fromtimestamp(v / 1e3, tz=_dt.timezone.utc)
    s = str(v).strip()
    m = _DATE_MATH.match(s)
    if m:
        if not m.group(1):
            return now
        n = int(m.group(2)) * (1 if m.group(1) == "+" else -1)
        unit = m.group(3)
        if unit == "y":
            return now.replace(year=now.year + n, day=min(now.day, 28) if now.month == 2 else now.day)
        if unit == "M":
            mm = now.month - 1 + n
            return now.replace(year=now.year + mm // 12, month=mm % 12 + 1, day=min(now.day, 28))
        secs = {"w": 604800, "d": 86400, "h": 3600, "H": 3600, "m": 60, "s": 1}[unit]
        return now + _dt.timedelta(seconds=n * secs)
    try:
        d = _dt.datetime.fromisoformat(s.replace("Z", "+00:00"))
    except ValueError:
        return None
    return d if d.tzinfo else d.replace(tzinfo=_dt.timezone.utc)


def _field_spec(spec: str):
    if "^" in spec:
        name, b = spec.rsplit("^", 1)
        return name, float(b)
    return spec, 1.0


class HostSearcher:
    def __init__(self, index):
        self.ix = index                       # client._Index
        self.n = len(index.sources)
        self._views = index.host_views        # per-field statistics, dropped by the index whenever a document changes
        self.now = _dt.datetime.now(_dt.timezone.utc)

    # -- field access -------------------------------------------------------------------------------------
    def view(self, name: str):
        """`X.keyword` is the keyword sub-field the mapping declares for text field X (app/main.py:375-380): the whole
        value as one term.  It is derived on demand from the stored sources."""
        v = self._views.get(name)
        if v is not None:
            return v
        ti = self.ix.text
        if name in ti.fields:
            v = FieldView(ti.fields[name], ti.types.get(name) == "keyword", self.n)
        elif name.endswith(".keyword") or name in self.ix.kw:
            base = name[:-8] if name.endswith(".keyword") else name
            fld = TextField()
            for row, src in enumerate(self.ix.sources):
                val = (src or {}).get(base)
                vals = val if isinstance(val, (list, tuple)) else [val]
                toks = [str(x) for x in vals if x is not None and not isinstance(x, (dict, list, tuple))]
                if toks:
                    fld.set_row_tokens(row, toks)
            if not fld.vocab:
                return None
            v = FieldView(fld, True, self.n)
        else:
            return None
        self._views[name] = v
        return v

    def _empty(self):
        return np.zeros(self.n, dtype=np.float32), np.zeros(self.n, dtype=bool)

    # -- term groups: one group = the alternatives of one query token (the token itself or its fuzzy expansions) ---
    def _token_groups(self, v: FieldView, name: str, text: str, fuzzy: bool, boost: np.float32):
        """[(term ids, float32 weights)] per token, in query order; a token without any term yields an empty group."""
        tokens = [text] if v.keyword else analyze(text)
        groups = []
        for tok in tokens:
            if fuzzy and not v.keyword:
                ids, ws = v.fld.fuzzy_weighted_terms([tok], float(boost), self._expander(v, name))
            else:
                ids, ws = v.fld.exact_weighted_terms([tok], float(boost))
            groups.append((ids, ws))
        return groups

    def _expander(self, v: FieldView, name: str):
        """Dictionary scan for fuzziness AUTO: on the device when the field is part of the uploaded dictionary."""
        ti, eng = self.ix.text, self.ix.engine
        if eng is not None and name in ti.base and not ti.dirty:
            base, n = ti.base[name], len(v.fld.vocab)

            def dev(tok, me):
                t, e = eng.fuzzy_expand(tok, me, base, base + n)
                return t - base, e
            return dev

        def host(tok, me):          # derived sub-fields (X.keyword) never take fuzziness; tiny dictionaries only
            terms = v.fld.terms_in_id_order()
            ids, eds = [], []
            for t, term in enumerate(terms):
                if abs(len(term) - len(tok)) <= me:
                    d = _osa(tok, term, me)
                    if d <= me:
                        ids.append(t)
                        eds.append(d)
            return np.asarray(ids, dtype=np.int64), np.asarray(eds, dtype=np.int64)
        return host

    def _field_terms_score(self, v: FieldView, groups, operator: str):
        """Sum over the tokens' matching terms; `and` requires every token to match."""
        acc = np.zeros(self.n, dtype=np.float64)
        hit_all = np.ones(self.n, dtype=bool)
        any_hit = np.zeros(self.n, dtype=bool)
        for ids, ws in groups:
            g = np.zeros(self.n, dtype=bool)
            for t, w in zip(ids, ws):
                rows, s = v.term_scores(int(t), np.float32(w))
                pos = s > 0
                acc[rows[pos]] += s[pos].astype(np.float64)
                g[rows[pos]] = True
            hit_all &= g
            any_hit |= g
        match = hit_all & any_hit if operator == "and" else any_hit
        score = acc.astype(np.float32)
        score[~match] = 0
        return score, match

    # -- phrase -------------------------------------------------------------------------------------------------
    def _phrase(self, v: FieldView, text: str, boost: np.float32, prefix: bool):
        tokens = [text] if v.keyword else analyze(text)
        if not tokens:
            return self._empty()
        vocab = v.fld.vocab
        positions: list[list[int]] = []
        for i, tok in enumerate(tokens):
            if prefix and i == len(tokens) - 1:
                alts = sorted(t for t in vocab if t.startswith(tok))[:50]      # max_expansions 50, term order
                ids = [vocab[t] for t in alts if v.df[vocab[t]] > 0]
            else:
                t = vocab.get(tok, -1)
                ids = [t] if t >= 0 and v.df[t] > 0 else []
            if not ids:
                return self._empty()
            positions.append(ids)
        if len(positions) == 1:
            # one position: a plain disjunction of term queries
            bo = np.float32(np.float32(boost) * v.fld.sim_boost)
            groups = [(positions[0], [np.float32(bo * v.idf(int(v.df[t]))) for t in positions[0]])]
            return self._field_terms_score(v, groups, "or")
        idf = np.float32(sum(float(v.idf(int(v.df[t]))) for ids in positions for t in ids))
        w = np.float32(np.float32(np.float32(boost) * v.fld.sim_boost) * idf)
        cand = None
        for ids in positions:
            rows = np.unique(np.concatenate([v.doc[v.indptr[t]:v.indptr[t + 1]] for t in ids]))
            cand = rows if cand is None else np.intersect1d(cand, rows, assume_unique=True)
            if cand.size == 0:
                return self._empty()
        score, match = self._empty()
        sets = [np.asarray(ids, dtype=np.int64) for ids in positions]
        for row in cand.tolist():
            seq = v.fld.row_terms[row]
            if seq.size < len(sets):         # shorter than the phrase (a repeated query token against a short field)
                continue
            ok = np.isin(seq[: seq.size - len(sets) + 1], sets[0])
            for j in range(1, len(sets)):
                ok &= np.isin(seq[j: seq.size - len(sets) + 1 + j], sets[j])
            freq = int(ok.sum())
            if freq:
                s = w - w / (np.float32(1.0) + np.float32(freq) * v.inv[v.norm[row]])
                score[row] = s
                match[row] = s > 0
        return score, match

    # -- query nodes ----------------------------------------------------------------------------------------------
    def eval(self, node):
        if not isinstance(node, dict) or len(node) != 1:
            raise NotImplementedError(f"unsupported query node: {node!r}")
        (kind, body), = node.items()
        fn = getattr(self, "_q_" + kind, None)
        if fn is None:
            raise NotImplementedError(f"query type {kind!r}")
        return fn(body)

    def _q_match_all(self, body):
        alive = np.array([s is not None for s in self.ix.sources], dtype=bool)
        return alive.astype(np.float32) * np.float32(body.get("boost", 1.0) if isinstance(body, dict) else 1.0), alive

    def _q_term(self, body):
        (fname, val), = body.items()
        boost = 1.0
        if isinstance(val, dict):
            boost = float(val.get("boost", 1.0))
            val = val.get("value")
        v = self.view(fname)
        if v is None:
            return self._empty()
        tok = str(val)
        ids, ws = v.fld.exact_weighted_terms([tok], boost)
        return self._field_terms_score(v, [(ids, ws)], "or")

    def _q_terms(self, body):
        (fname, vals), = [(k, x) for k, x in body.items() if k != "boost"]
        v = self.view(fname)
        if v is None:
            return self._empty()
        match = np.zeros(self.n, dtype=bool)
        for val in vals:
            t = v.fld.vocab.get(str(val), -1)
            if t >= 0:
                match[v.doc[v.indptr[t]:v.indptr[t + 1]]] = True
        return match.astype(np.float32) * np.float32(body.get("boost", 1.0)), match      # terms is constant-score

    def _q_exists(self, body):
        fname = body["field"]
        match = np.array([s is not None and s.get(fname) is not None for s in self.ix.sources], dtype=bool)
        return match.astype(np.float32), match

    def _column(self, fname: str, is_date: bool):
        """(values float64 [rows], present bool [rows]) of one source field -- dates as epoch seconds -- cached per index
        version, so a range filter costs one vectorised comparison per query instead of a Python pass over the rows."""
        key = ("column", fname, is_date)
        col = self._views.get(key)
        if col is None:
            vals = np.zeros(self.n, dtype=np.float64)
            ok = np.zeros(self.n, dtype=bool)
            for row, src in enumerate(self.ix.sources):
                raw = (src or {}).get(fname)
                if raw is None:
                    continue
                try:
                    if is_date:
                        d = _parse_date(raw, self.now)
                        x = None if d is None else d.timestamp()
                    else:
                        x = float(raw)
                except (TypeError, ValueError):
                    continue
                if x is not None:
                    vals[row], ok[row] = x, True
            col = self._views[key] = (vals, ok)
        return col

    def _q_range(self, body):
        (fname, cond), = body.items()
        is_date = self.ix.field_type(fname) == "date" or any(isinstance(x, str) for x in cond.values())
        if is_date:
            bounds = {op: _parse_date(val, self.now) for op, val in cond.items() if op in ("gt", "gte", "lt", "lte")}
            if any(b is None for b in bounds.values()):           # an unparseable bound (free text from the NER) matches nothing
                return self._empty()
            bounds = {op: b.timestamp() for op, b in bounds.items()}
        else:
            bounds = {op: float(val) for op, val in cond.items() if op in ("gt", "gte", "lt", "lte")}
        vals, match = self._column(fname, is_date)
        match = match.copy()
        if "gt" in bounds:
            match &= vals > bounds["gt"]
        if "gte" in bounds:
            match &= vals >= bounds["gte"]
        if "lt" in bounds:
            match &= vals < bounds["lt"]
        if "lte" in bounds:
            match &= vals <= bounds["lte"]
        return match.astype(np.float32) * np.float32(cond.get("boost", 1.0)), match      # constant score 1

    def _q_match(self, body):
        (fname, spec), = body.items()
        if not isinstance(spec, dict):
            spec = {"query": spec}
        v = self.view(fname)
        if v is None:
            return self._empty()
        fuzzy = str(spec.get("fuzziness", "")).upper() == "AUTO"
        groups = self._token_groups(v, fname, str(spec.get("query", "")), fuzzy, np.float32(spec.get("boost", 1.0)))
        return self._field_terms_score(v, groups, str(spec.get("operator", "or")).lower())

    def _q_match_phrase(self, body):
        (fname, spec), = body.items()
        if not isinstance(spec, dict):
            spec = {"query": spec}
        v = self.view(fname)
        if v is None:
            return self._empty()
        return self._phrase(v, str(spec.get("query", "")), np.float32(spec.get("boost", 1.0)), prefix=False)

    def _q_match_phrase_prefix(self, body):
        (fname, spec), = body.items()
        if not isinstance(spec, dict):
            spec = {"query": spec}
        v = self.view(fname)
        if v is None:
            return self._empty()
        return self._phrase(v, str(spec.get("query", "")), np.float32(spec.get("boost", 1.0)), prefix=True)

    def _q_multi_match(self, m):
        kind = m.get("type", "best_fields")
        if kind not in ("best_fields", "phrase", "phrase_prefix"):
            raise NotImplementedError(f"multi_match type {kind!r}")
        query, cb = str(m.get("query", "")), np.float32(m.get("boost", 1.0))
        fuzzy = str(m.get("fuzziness", "")).upper() == "AUTO"
        op = str(m.get("operator", "or")).lower()
        best, match = self._empty()
        for spec in m.get("fields", []):
            fname, fb = _field_spec(spec)
            v = self.view(fname)
            if v is None:
                continue
            bo = np.float32(cb * np.float32(fb))
            if kind == "best_fields":
                s, mt = self._field_terms_score(v, self._token_groups(v, fname, query, fuzzy, bo), op)
            else:
                s, mt = self._phrase(v, query, bo, prefix=(kind == "phrase_prefix"))
            best = np.maximum(best, s)          # tie_breaker 0: the best field's score
            match |= mt
        return best, match

    def _q_knn(self, body):
        (fname, spec), = body.items()
        eng = self.ix.engine
        q = np.asarray(spec["vector"], dtype=np.float32).reshape(1, -1)
        k = max(int(spec.get("k", 10)), 1)
        if k > 128:
            raise NotImplementedError("knn k above 128")
        with self.ix.lock:
            rows, scores = eng.search_knn(q, k)
        score, match = self._empty()
        for r, s in zip(rows[0], scores[0]):
            if r >= 0:
                score[int(r)] = np.float32(s) * np.float32(spec.get("boost", 1.0))
                match[int(r)] = True
        return score, match

    def _q_bool(self, b):
        unknown = set(b) - {"must", "should", "filter", "must_not", "minimum_should_match", "boost"}
        if unknown:
            raise NotImplementedError(f"unsupported bool keys: {sorted(unknown)}")

        def clauses(key):
            c = b.get(key) or []
            return [c] if isinstance(c, dict) else list(c)

        total = np.zeros(self.n, dtype=np.float64)
        match = np.array([s is not None for s in self.ix.sources], dtype=bool)
        for node in clauses("must"):
            s, m = self.eval(node)
            total += s.astype(np.float64)
            match &= m
        for node in clauses("filter"):
            match &= self.eval(node)[1]
        for node in clauses("must_not"):
            match &= ~self.eval(node)[1]
        should = clauses("should")
        if should:
            n_hit = np.zeros(self.n, dtype=np.int32)
            for node in should:
                s, m = self.eval(node)
                total += np.where(m, s, np.float32(0)).astype(np.float64)
                n_hit += m
            default_msm = 0 if (clauses("must") or clauses("filter")) else 1
            msm = int(b.get("minimum_should_match", default_msm))
            match &= n_hit >= msm
        score = (total.astype(np.float32) * np.float32(b.get("boost", 1.0))).astype(np.float32)
        score[~match] = 0
        return score, match

    # -- request ------------------------------------------------------------------------------------------------
    def search(self, body: dict):
        """-> (hits [(row, score or None, sort values or None)], aggregations dict or None)"""
        q = body.get("query") or {"match_all": {}}
        score, match = self.eval(q)
        rows = np.flatnonzero(match)
        sort = body.get("sort")
        tracked = True
        if sort:
            specs = sort if isinstance(sort, list) else [sort]
            keys = []
            for sp in reversed(specs):
                (fname, opt), = sp.items() if isinstance(sp, dict) else ((sp, {}),)
                desc = (opt.get("order", "asc") if isinstance(opt, dict) else str(opt)) == "desc"
                if fname == "_score":
                    vals = score[rows].astype(np.float64)
                    missing = np.zeros(rows.size, dtype=bool)
                else:
                    raw = [(self.ix.sources[r] or {}).get(fname) for r in rows]
                    is_date = self.ix.field_type(fname) == "date"
                    conv = [(_parse_date(x, self.now).timestamp() if is_date and x is not None and
                             _parse_date(x, self.now) is not None else
                             (float(x) if isinstance(x, (int, float)) else None)) for x in raw]
                    if not is_date and any(isinstance(x, str) for x in raw):
                        # keyword field: lexicographic order of the values (their rank among the distinct values)
                        rank = {v: float(i) for i, v in enumerate(sorted({x for x in raw if isinstance(x, str)}))}
                        conv = [rank.get(x) if isinstance(x, str) else None for x in raw]
                    missing = np.array([c is None for c in conv], dtype=bool)
                    vals = np.array([0.0 if c is None else c for c in conv], dtype=np.float64)
                keys.append(np.where(missing, np.inf, -vals if desc else vals))      # missing values sort last
            order = np.lexsort([rows] + keys)
            tracked = any((list(sp)[0] if isinstance(sp, dict) else sp) == "_score" for sp in specs)
        else:
            order = np.lexsort((rows, -score[rows].astype(np.float64)))
        rows = rows[order]
        collapse = body.get("collapse")
        if collapse:
            fname = collapse["field"]
            seen, kept = set(), []
            for r in rows.tolist():
                key = (self.ix.sources[r] or {}).get(fname)
                if key is None or key not in seen:
                    kept.append(r)
                    if key is not None:
                        seen.add(key)
            rows = np.asarray(kept, dtype=np.int64)
        aggs = None
        spec = body.get("aggs") or body.get("aggregations")
        if spec:
            aggs = {}
            matched = np.flatnonzero(match)
            for name, a in spec.items():
                if set(a) != {"terms"}:
                    raise NotImplementedError(f"aggregation {list(a)}")
                fname, size = a["terms"]["field"], int(a["terms"].get("size", 10))
                base = fname[:-8] if fname.endswith(".keyword") else fname
                declared = self.ix.field_type(fname) == "keyword" or (
                    fname.endswith(".keyword") and self.ix.has_keyword_subfield(base))
                counts: dict = {}
                if declared:
                    for r in matched.tolist():
                        val = (self.ix.sources[r] or {}).get(base)
                        for x in (val if isinstance(val, (list, tuple)) else [val]):
                            if x is not None:
                                counts[x] = counts.get(x, 0) + 1
                top = sorted(counts.items(), key=lambda kv: (-kv[1], str(kv[0])))
                aggs[name] = {"doc_count_error_upper_bound": 0,
                              "sum_other_doc_count": int(sum(c for _, c in top[size:])),
                              "buckets": [{"key": key, "doc_count": c} for key, c in top[:size]]}
        frm, size = int(body.get("from", 0)), int(body.get("size", 10))
        rows = rows[frm: frm + size]
        return [(int(r), float(score[r]) if tracked else None) for r in rows.tolist()], aggs, int(match.sum())


def _osa(a: str, b: str, cap: int) -> int:
    """Optimal-string-alignment distance (host mirror of fuzzy_scan_kernel, for derived sub-fields only)."""
    la, lb = len(a), len(b)
    prev2, prev = None, list(range(lb + 1))
    for i in range(1, la + 1):
        cur = [i] + [0] * lb
        for j in range(1, lb + 1):
            v = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[i - 1] != b[j - 1]))
            if i > 1 and j > 1 and a[i - 1] == b[j - 2] and a[i - 2] == b[j - 1]:
                v = min(v, prev2[j - 2] + 1)
            cur[j] = v
        if min(cur) > cap:
            return cap + 1
        prev2, prev = prev, cur
    return prev[lb]
