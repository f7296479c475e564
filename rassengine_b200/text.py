"""Host side of the text clause: analysis and the inverted index handed to rass_bm25_build.

The reference maps `unstructuredText` as {"type": "text"} with the default (standard) analyzer
(app/main.py:555-556): UAX#29 word segmentation + lower-casing, no stop words -- `analysis.analyze`; the reference's
own chunker splits on whitespace (app/main.py:2160-2170).  Postings are CSR (term -> ascending doc rows with term
frequencies) plus a token count per row, the layout the device BM25 kernel walks.
"""
from __future__ import annotations

import numpy as np

from .analysis import analyze  # noqa: F401  (re-exported: client.py, hostquery.py and the tests import it from here)


class TextField:
    """Growing term dictionary + per-row token-id arrays for one analysed field."""

    def __init__(self):
        self.vocab: dict[str, int] = {}
        self.row_terms: dict[int, np.ndarray] = {}   # row -> int32 term ids (with repeats)
        self.dirty = True
        self._fuzzy: dict = {}                       # token -> ranked expansion, as of the last postings()
        self._terms: list[str] = []                  # id -> term, extended when the vocabulary grew
        self._enc_blob = bytearray()                 # the terms' utf-8 bytes, back to back (encoded_terms)
        self._enc_lens: list[int] = []
        self.df = np.zeros(0, dtype=np.int64)        # document frequency per term, as of the last postings()
        self.doc_count = 0                           # documents with at least one token, as of the last postings()
        # device-side ingest (TextIndex.sync_device): rows added, rewritten or cleared since the last sync
        self.pending_rows: list[int] = []
        self.pending_ids: list[np.ndarray] = []
        self.rebuild = False                         # set by whoever needs the rebuild from host arrays instead
        # what the similarity multiplies every query boost by before it builds the scorer: 1 (Lucene's BM25Similarity,
        # what SURVEY.md 8c states) or float32(1 + k1) = 2.2f (LegacyBM25Similarity: B200Client(bm25_legacy_boost=True))
        self.sim_boost = np.float32(1.0)

    def set_row(self, row: int, text: str | None):
        self.set_row_tokens(row, analyze(text) if text else [])

    def set_row_tokens(self, row: int, toks: list[str]):
        if not toks:
            if self.row_terms.pop(row, None) is not None:       # the row loses the field: an empty rewrite
                self.dirty = True
                self.pending_rows.append(row)
                self.pending_ids.append(np.zeros(0, dtype=np.int32))
            return
        v = self.vocab
        ids = list(map(v.get, toks))                  # C-speed for the tokens the dictionary knows (nearly all)
        if None in ids:
            for i, t in enumerate(toks):
                if ids[i] is None:
                    ids[i] = v.setdefault(t, len(v))
        self._store(row, np.array(ids, dtype=np.int32))

    def set_row_ids(self, row: int, ids: np.ndarray):
        """Pre-tokenised input (synthetic corpora): term ids must be < declare_vocab()."""
        self._store(row, np.asarray(ids, dtype=np.int32))

    def _store(self, row: int, ids: np.ndarray):
        self.row_terms[row] = ids
        self.dirty = True
        self.pending_rows.append(row)
        self.pending_ids.append(ids)

    def take_pending(self):
        """(rows int64 [n] ascending and distinct, tok_indptr int64 [n + 1], tok_terms int32) of the rows added,
        rewritten or cleared since the last call (the last write to a row wins)."""
        rows = np.asarray(self.pending_rows, dtype=np.int64)
        ids = self.pending_ids
        if rows.size > 1 and not np.all(rows[1:] > rows[:-1]):
            last = {}
            for i, r in enumerate(self.pending_rows):
                last[r] = i
            keep = sorted(last.values(), key=lambda i: self.pending_rows[i])
            rows = rows[keep]
            ids = [self.pending_ids[i] for i in keep]
        indptr = np.zeros(rows.size + 1, dtype=np.int64)
        if rows.size:
            indptr[1:] = np.cumsum([a.size for a in ids])
        terms = np.concatenate(ids).astype(np.int32) if rows.size else np.zeros(0, np.int32)
        self.pending_rows, self.pending_ids = [], []
        return rows, indptr, terms

    def query_terms(self, text: str) -> list[int]:
        """Term ids of the query tokens; unknown tokens map to -1 (the kernel ignores them)."""
        return [self.vocab.get(t, -1) for t in analyze(text)]

    def postings(self, n_rows: int):
        """CSR over the current rows: indptr int64 [V+1], doc int32 [nnz] ascending per term, tf uint16 [nnz],
        doclen uint32 [n_rows]."""
        V = len(self.vocab)
        self._fuzzy = {}
        doclen = np.zeros(n_rows, dtype=np.uint32)
        if not self.row_terms:
            self.df, self.doc_count = np.zeros(V, dtype=np.int64), 0
            return np.zeros(V + 1, dtype=np.int64), np.zeros(0, np.int32), np.zeros(0, np.uint16), doclen
        rows = np.fromiter(self.row_terms.keys(), dtype=np.int64, count=len(self.row_terms))
        lens = np.fromiter((a.size for a in self.row_terms.values()), dtype=np.int64, count=rows.size)
        doclen[rows] = lens
        terms = np.concatenate(list(self.row_terms.values())).astype(np.int64)
        V = max(V, int(terms.max()) + 1)
        docs = np.repeat(rows, lens)
        indptr, d_of, tf = csr_from_pairs(terms, docs, V, n_rows)
        self.df, self.doc_count = np.diff(indptr), int(np.count_nonzero(doclen))
        return indptr, d_of, tf, doclen

    def terms_in_id_order(self) -> list[str]:
        if len(self._terms) > len(self.vocab):
            self._terms = []
        if len(self._terms) < len(self.vocab):
            # ids are handed out in insertion order (vocab.setdefault(t, len(vocab))), which is the dict's own order
            import itertools
            self._terms.extend(itertools.islice(self.vocab, len(self._terms), None))
        return self._terms

    def encoded_terms(self):
        """(utf-8 bytes of all terms back to back, int64 byte length per term), extended by the terms that are new
        since the last call -- what the device dictionary of the fuzzy scan is uploaded from."""
        terms = self.terms_in_id_order()
        n_old = len(self._enc_lens)
        if n_old > len(terms):
            self._enc_blob, self._enc_lens, n_old = bytearray(), [], 0
        if n_old < len(terms):
            enc = [t.encode("utf-8", "replace") for t in terms[n_old:]]
            self._enc_blob += b"".join(enc)
            self._enc_lens.extend(len(e) for e in enc)
        return self._enc_blob, np.asarray(self._enc_lens, dtype=np.int64)

    @staticmethod
    def auto_max_edits(n_chars: int) -> int:
        """fuzziness AUTO = AUTO:3,6: 0 edits below 3 characters, 1 for 3..5, 2 from 6 on."""
        return 0 if n_chars <= 2 else (1 if n_chars <= 5 else 2)

    def exact_weighted_terms(self, tokens: list[str], boost: float):
        """Plain term queries (no fuzziness): (term ids, float32 weights = float(boost) * idf), query order."""
        import math
        ids, ws = [], []
        bo = np.float32(np.float32(boost) * self.sim_boost)
        for tok in tokens:
            t = self.vocab.get(tok, -1)
            if t < 0 or t >= self.df.size or self.df[t] == 0:
                continue
            df = int(self.df[t])
            idf = np.float32(math.log(1.0 + (self.doc_count - df + 0.5) / (df + 0.5)))
            ids.append(int(t))
            ws.append(np.float32(bo * idf))
        return ids, ws

    def fuzzy_weighted_terms(self, text, boost: float, expand, max_expansions: int = 50):
        """The boosted term queries `multi_match(..., fuzziness: AUTO)` rewrites a query string to (Lucene FuzzyQuery +
        TopTermsBlendedFreqScoringRewrite, restated in oracle/fuzzy.py): for every token the <= 50 best dictionary
        terms by (similarity boost desc, term asc), scored with the largest document frequency among them.
        expand(token, max_edits) -> (term ids, edits) is the device dictionary scan.  Returns (term ids, float32
        weights = float(boost * term boost) * idf)."""
        import math
        ids: list[int] = []
        ws: list[np.float32] = []
        bo = np.float32(np.float32(boost) * self.sim_boost)
        for tok in (analyze(text) if isinstance(text, str) else text):
            hit = self._fuzzy.get(tok)
            if hit is None:
                me = self.auto_max_edits(len(tok))
                if me == 0 or len(tok) > 64:
                    t = self.vocab.get(tok, -1)
                    cand = [(t, 0)] if t >= 0 else []
                else:
                    tid, ed = expand(tok, me)
                    cand = list(zip(tid.tolist(), ed.tolist()))
                cand = [(t, e) for t, e in cand if t < self.df.size and self.df[t] > 0]
                scored, idf = [], np.float32(0)
                if cand:
                    terms_by_id = self.terms_in_id_order()
                    for t, e in cand:
                        b = np.float32(1.0) if e == 0 else \
                            np.float32(1.0) - np.float32(e) / np.float32(min(len(terms_by_id[t]), len(tok)))
                        scored.append((t, np.float32(b)))
                    scored.sort(key=lambda x: (-float(x[1]), terms_by_id[x[0]]))
                    scored = scored[:max_expansions]
                    max_df = max(int(self.df[t]) for t, _ in scored)
                    idf = np.float32(math.log(1.0 + (self.doc_count - max_df + 0.5) / (max_df + 0.5)))
                if len(self._fuzzy) > 50000:
                    self._fuzzy.clear()
                hit = self._fuzzy[tok] = (scored, idf)       # the dictionary scan and its ranking, per distinct token
            scored, idf = hit
            for t, b in scored:
                ids.append(int(t))
                ws.append(np.float32(np.float32(bo * b) * idf))
        return ids, ws


def csr_from_pairs(terms: np.ndarray, docs: np.ndarray, V: int, n_rows: int):
    key = terms * np.int64(n_rows) + docs
    uniq, counts = np.unique(key, return_counts=True)
    t_of = uniq // n_rows
    d_of = (uniq - t_of * n_rows).astype(np.int32)
    indptr = np.zeros(V + 1, dtype=np.int64)
    np.add.at(indptr, t_of + 1, 1)
    return np.cumsum(indptr), d_of, np.minimum(counts, 65535).astype(np.uint16)


class TextIndex:
    """Every analysed (`text`) and `keyword` field of an index, sharing one CSR on the device: field f owns the global
    term ids [base[f], base[f] + V_f).  A keyword field is a field whose "analyzer" emits the whole value as one token
    (no lower-casing), which is how Lucene indexes it; its norms are omitted, i.e. every value counts as length 1."""

    def __init__(self, field_types: dict[str, str], sim_boost=1.0):
        self.sim_boost = np.float32(sim_boost)    # handed to every field (TextField.sim_boost)
        self.types = dict(field_types)            # field name -> "text" | "keyword"
        self.fields: dict[str, TextField] = {}    # created when the first document carries the field
        self.order: list[str] = []                # field id -> name
        self.base: dict[str, int] = {}            # as of the last postings()
        self.dirty = True
        self.device_commits = 0                   # syncs that went through the device-side ingest / a host rebuild
        self.host_rebuilds = 0

    def field_id(self, name: str) -> int:
        return self.order.index(name)

    def tokens(self, name: str, value) -> list[str]:
        values = value if isinstance(value, (list, tuple)) else [value]
        out: list[str] = []
        for v in values:
            if v is None:
                continue
            if self.types.get(name) == "keyword":
                out.append(str(v))
            elif isinstance(v, str):
                out.extend(analyze(v))
        return out

    def set_doc(self, row: int, src: dict | None, fresh: bool = False):
        """(Re)index one row: every declared field the document carries; on an overwrite (fresh = False) the fields
        it no longer carries are cleared."""
        src = src or {}
        names = [n for n in src if n in self.types] if fresh else list(self.types)
        for name in names:
            toks = self.tokens(name, src[name]) if name in src else []
            fld = self.fields.get(name)
            if fld is None:
                if not toks:
                    continue
                fld = self.fields[name] = TextField()
                fld.sim_boost = self.sim_boost
                self.order.append(name)
            before = fld.dirty
            fld.dirty = False
            fld.set_row_tokens(row, toks)
            if fld.dirty:
                self.dirty = True
            fld.dirty = fld.dirty or before

    def postings(self, n_rows: int):
        """Combined CSR: indptr int64 [V+1], doc int32, tf uint16, term_field int32 [V], doclen uint32 [F, n_rows]."""
        indptrs, docs, tfs, fields, lens = [], [], [], [], []
        base = nnz = 0
        self.base = {}
        for fid, name in enumerate(self.order):
            fld = self.fields[name]
            ip, d, t, dl = fld.postings(n_rows)
            if self.types.get(name) == "keyword":
                dl = (dl > 0).astype(np.uint32)        # omitted norms: length 1 for every document that has a value
            self.base[name] = base
            indptrs.append(ip[:-1] + nnz)
            nnz += d.size
            docs.append(d)
            tfs.append(t)
            fields.append(np.full(ip.size - 1, fid, dtype=np.int32))
            lens.append(dl)
            base += ip.size - 1
        if not self.order:
            z = np.zeros
            return z(1, np.int64), z(0, np.int32), z(0, np.uint16), z(0, np.int32), z((1, n_rows), np.uint32)
        indptr = np.concatenate(indptrs + [np.array([nnz], dtype=np.int64)]).astype(np.int64)
        return (indptr, np.concatenate(docs), np.concatenate(tfs), np.concatenate(fields), np.stack(lens))

    def sync_device(self, engine, n_rows: int):
        """Bring the engine's postings up to date with the rows indexed so far.  New, rewritten and cleared rows travel
        as token-id streams and are inverted on the device (rass_text_add_rows + rass_text_commit: a segment per bulk,
        one merge pass, or a re-sort when rows were rewritten; a handle spread over several GPUs splits the streams by
        its row map; keyword fields omit norms).  The rebuild from host arrays (rass_bm25_build_fields) remains for a
        caller that sets `rebuild` on a field."""
        flds = [self.fields[n] for n in self.order]
        if any(f.rebuild for f in flds) or not self.order:
            indptr, doc, tf, term_field, doclen = self.postings(n_rows)
            engine.bm25_build_fields(indptr, doc, tf, term_field, doclen)
            for f in flds:
                f.pending_rows, f.pending_ids, f.rebuild = [], [], False
            self.host_rebuilds += 1
        else:
            for fid, (name, f) in enumerate(zip(self.order, flds)):
                if self.types.get(name) == "keyword":
                    engine.text_omit_norms(fid, True)      # however many values: a document counts as length 1
                if f.pending_rows:
                    engine.text_add_rows(fid, *f.take_pending())
            engine.text_commit([len(f.vocab) for f in flds], n_rows)
            # the host-side statistics the query rewriting needs (df per term, docCount per field)
            indptr, doc_count = engine.text_stats()
            base = 0
            self.base = {}
            for fid, (name, f) in enumerate(zip(self.order, flds)):
                V = len(f.vocab)
                self.base[name] = base
                f.df = np.diff(indptr[base:base + V + 1])
                f.doc_count = int(doc_count[fid])
                f._fuzzy = {}
                base += V
            self.device_commits += 1
        self.dirty = False

    def vocab_blob(self):
        """The device dictionary in global term-id order: (bytes, int64 offsets [V + 1]).  Only analysed fields are
        ever scanned (fuzziness applies to their tokens; a keyword field's whole value is matched exactly), so keyword
        terms are uploaded as empty strings -- a `doc_id` field alone has one term per document."""
        blobs, lens = [], []
        for name in self.order:
            f = self.fields[name]
            if self.types.get(name) == "keyword":
                lens.append(np.zeros(len(f.vocab), dtype=np.int64))
            else:
                b, ln = f.encoded_terms()
                blobs.append(bytes(b))
                lens.append(ln)
        off = np.zeros(sum(l.size for l in lens) + 1, dtype=np.int64)
        if lens:
            off[1:] = np.cumsum(np.concatenate(lens))
        return b"".join(blobs), off

    def terms_in_id_order(self) -> list[str]:
        out: list[str] = []
        for name in self.order:
            out.extend(self.fields[name].terms_in_id_order())
        return out
