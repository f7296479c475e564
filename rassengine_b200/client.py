"""An `opensearchpy.OpenSearch`-shaped client whose "server" is the B200 engine in this process.

It stands in for the module-global `os_client` of the reference (app/main.py:337-343; app/embedding_gen.py:121-133)
and answers the calls the hot path makes on it:

  client.indices.exists(index)                 app/main.py:352
  client.indices.create(index=, body=)         app/main.py:576   (reads the knn_vector field: dimension, space_type)
  client.count(index=)                         app/main.py:1475
  client.info()                                app/embedding_gen.py:129
  client.search(index=, body=, routing=)       app/main.py:1552, 1607
  helpers.bulk(client, actions)                app/main.py:1237, 1269, 1279

`_source` dicts and the `_id` <-> row map live on the host; vectors and postings live on the GPU.  Row ids are
append-ordered, so the engine's "row ascending" tie-break equals Lucene's doc-id order.
"""
from __future__ import annotations

import json
import os
import threading
import time

import numpy as np

from . import _capi as capi
from .dsl import Plan, parse_search_body
from .batcher import MicroBatcher
from .engine import Engine
from .text import TextIndex

TEXT_FIELD = "unstructuredText"      # the only analysed field chunk documents carry (app/main.py:1120-1130)
FILTER_FIELDS = ("patientId", "doc_type", "resourceType", "doc_id")   # keyword fields the hot path filters on
_SPACE = {"cosinesimil": capi.METRIC_COSINE, "l2": capi.METRIC_L2}
MAX_K = 128                          # RASS_MAX_K: the largest top-k one engine call returns
FILTER_LIST_MAX = 1 << 17            # filters passing at most this many rows are scored row by row (rass_search_*_filtered)


class NotFoundError(KeyError):
    """index_not_found_exception"""


class RequestError(ValueError):
    """resource_already_exists_exception / bad request"""


def resolve_devices(devices, device: int, body: dict | None) -> list[int]:
    """GPUs of one index, in this order of precedence: an explicit `devices` list, the RASS_B200_DEVICES environment
    variable ("0,1,2,3" or "all"), `settings.index.number_of_shards` of the create body (app/main.py:357: the reference
    passes SHARD_COUNT there) capped by the GPUs present, else the single `device`."""
    if devices:
        return [int(d) for d in devices]
    env = os.environ.get("RASS_B200_DEVICES", "").strip()
    if env:
        if env.lower() == "all":
            import ctypes
            n = ctypes.c_int(0)
            try:
                ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n))
            except OSError:
                n.value = 0
            return list(range(max(1, n.value)))
        return [int(x) for x in env.split(",") if x.strip() != ""]
    shards = 1
    try:
        shards = int((body or {}).get("settings", {}).get("index", {}).get("number_of_shards", 1))
    except (TypeError, ValueError):
        shards = 1
    if shards > 1:
        n_gpu = _gpu_count()
        if n_gpu > 1:
            return list(range(device, device + min(shards, n_gpu - device)))
    return [device]


def _gpu_count() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 1


class _Index:
    def __init__(self, name: str, body: dict | None, device: int, knn_filter: str = "post", devices=None,
                 sim_boost=1.0):
        self.name = name
        self.body = body or {}
        self.device = device
        self.devices = resolve_devices(devices, device, body)
        self.knn_filter = knn_filter
        props = self.body.get("mappings", {}).get("properties", {})
        self.vector_field = None
        self.dim = None
        self.metric = capi.METRIC_COSINE
        for fld, spec in props.items():
            if isinstance(spec, dict) and spec.get("type") == "knn_vector":
                self.vector_field = fld
                self.dim = int(spec["dimension"])
                space = spec.get("method", {}).get("space_type", "l2")
                if space not in _SPACE:
                    raise NotImplementedError(f"space_type {space!r}")
                self.metric = _SPACE[space]
        self.engine: Engine | None = None
        self.lock = threading.RLock()            # the engine takes one call at a time per handle
        self.batcher: MicroBatcher | None = None
        self.batch_window_s: float | None = None
        self.sources: list[dict | None] = []     # row -> _source WITHOUT the vector (it lives on the GPU)
        self.has_vec: list[bool] = []            # row -> the document carried a vector
        self.ids: list[str | None] = []          # row -> _id
        self.row_of: dict[str, int] = {}
        # analysed / keyword fields of the mapping (app/main.py:361-561); chunk documents only carry TEXT_FIELD
        types = {f: spec["type"] for f, spec in props.items()
                 if isinstance(spec, dict) and spec.get("type") in ("text", "keyword")}
        types.setdefault(TEXT_FIELD, "text")
        self.text = TextIndex(types, sim_boost)
        self._global_df = None            # df of every global term id, concatenated (rebuilt after a sync)
        self.n_docs = 0
        self.host_views: dict = {}               # hostquery.FieldView cache (N4 host-side queries)
        self.kw: dict[str, dict[object, list[int]]] = {f: {} for f in FILTER_FIELDS}   # field -> value -> rows

    # -- engine ------------------------------------------------------------------------------------------
    def _ensure_engine(self, dim: int | None = None) -> Engine:
        if self.engine is None:
            if self.dim is None:
                self.dim = dim or 1024
                self.vector_field = self.vector_field or "embedding"
            self.engine = Engine(dim=self.dim, metric=self.metric, device=self.device, devices=self.devices)
        return self.engine

    def close(self):
        if self.batcher is not None:
            self.batcher.close()
            self.batcher = None
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    def field_type(self, name: str) -> str | None:
        """Mapping type of a field; `X.keyword` resolves to the keyword sub-field of X when the mapping declares one."""
        props = self.body.get("mappings", {}).get("properties", {})
        if name in props and isinstance(props[name], dict):
            return props[name].get("type")
        if name.endswith(".keyword") and self.has_keyword_subfield(name[:-8]):
            return "keyword"
        return None

    def has_keyword_subfield(self, base: str) -> bool:
        spec = self.body.get("mappings", {}).get("properties", {}).get(base)
        return isinstance(spec, dict) and isinstance(spec.get("fields"), dict) and \
            spec["fields"].get("keyword", {}).get("type") == "keyword"

    def host_search(self, body: dict):
        """Query shapes outside the GPU hot path, evaluated on the host: -> (hits, aggs, total)."""
        from .hostquery import HostSearcher
        with self.lock:
            self._sync_text()                 # per-field statistics and the device dictionary (fuzziness) are current
            found, aggs, total = HostSearcher(self).search(body)
        src_filter = body.get("_source")
        vf = self.vector_field or "embedding"
        need_vec = src_filter is not False and not (isinstance(src_filter, (list, tuple)) and vf not in src_filter)
        hits = []
        for h in self._hits(found, with_vectors=need_vec):
            if isinstance(src_filter, (list, tuple)):
                h["_source"] = {k: v for k, v in (h["_source"] or {}).items() if k in src_filter}
            elif src_filter is False:
                h.pop("_source")
            hits.append(h)
        return hits, aggs, total

    def _knn(self, q: np.ndarray, k: int):
        """kNN of one query; coalesced with concurrent callers when the client was given a batch window."""
        if self.batch_window_s is None:
            with self.lock:
                return self.engine.search_knn(q, k)
        if self.batcher is None:
            with self.lock:                   # threads race here on the first request
                if self.batcher is None:
                    def locked(Q, kk):
                        with self.lock:
                            return self.engine.search_knn(Q, kk)
                    self.batcher = MicroBatcher(locked, max_batch=64, max_wait_s=self.batch_window_s)
        rows, scores = self.batcher.search(q, k)
        return rows[None, :], scores[None, :]

    def _strip(self, src: dict) -> dict:
        vf = self.vector_field or "embedding"
        return {k: v for k, v in src.items() if k != vf} if vf in src else src

    # -- indexing: "_op_type": "index" = insert or overwrite by _id --------------------------------------
    def index_batch(self, docs: list[tuple[str, dict]]):
        """docs: (_id, _source).  New ids are appended in one device call; known ids are overwritten in place.
        `_source[vector_field]` may be a list of floats (what the reference sends, app/main.py:1256) or a numpy row
        (the fast ingest path: no .tolist() / JSON detour)."""
        with self.lock:
            self._index_batch_locked(docs)

    def _index_batch_locked(self, docs: list[tuple[str, dict]]):
        self.host_views = {}
        vf = self.vector_field or "embedding"
        fresh: list[tuple[str, dict]] = []
        seen: dict[str, int] = {}
        for _id, src in docs:
            if _id in self.row_of:
                self._overwrite(self.row_of[_id], src)
            elif _id in seen:
                fresh[seen[_id]] = (_id, src)     # same id twice in one batch: last write wins
            else:
                seen[_id] = len(fresh)
                fresh.append((_id, src))
        if not fresh:
            return
        first_vec = next((s.get(vf) for _, s in fresh if s.get(vf) is not None), None)
        eng = self._ensure_engine(len(first_vec) if first_vec is not None else None)
        mat = np.zeros((len(fresh), self.dim), dtype=np.float32)
        has = np.zeros(len(fresh), dtype=bool)
        for i, (_, src) in enumerate(fresh):
            v = src.get(vf)
            if v is not None:
                if len(v) != self.dim:
                    raise RequestError(f"vector length {len(v)} != dimension {self.dim}")
                mat[i] = v
                has[i] = True
        first = eng.append(mat)
        for i, (_id, src) in enumerate(fresh):
            row = first + i
            if not has[i]:
                eng.tombstone(row)               # no embedding: the row never matches a knn clause
            self.sources.append(self._strip(src))
            self.has_vec.append(bool(has[i]))
            self.ids.append(_id)
            self.row_of[_id] = row
            self.text.set_doc(row, src, fresh=True)
            self._kw_add(row, src)
        self.n_docs += len(fresh)

    def _overwrite(self, row: int, src: dict):
        vf = self.vector_field or "embedding"
        eng = self._ensure_engine()
        v = src.get(vf)
        if v is not None:
            eng.overwrite(row, np.asarray(v, dtype=np.float32))
        else:
            eng.tombstone(row)
        self._kw_remove(row, self.sources[row] or {})
        self.sources[row] = self._strip(src)
        self.has_vec[row] = v is not None
        self.text.set_doc(row, src)
        self._kw_add(row, src)

    def _kw_add(self, row: int, src: dict):
        for f in FILTER_FIELDS:
            v = src.get(f)
            if isinstance(v, (str, int, bool)):
                self.kw[f].setdefault(v, []).append(row)

    def _kw_remove(self, row: int, src: dict):
        for f in FILTER_FIELDS:
            v = src.get(f)
            rows = self.kw[f].get(v) if isinstance(v, (str, int, bool)) else None
            if rows and row in rows:
                rows.remove(row)

    def _host_filter_rows(self, nodes) -> np.ndarray:
        """Rows matching every non-term bool.filter clause (the NER filter_clause, app/main.py:2589-2609): only this
        sub-tree is evaluated on the host (postings for match_phrase, cached columns for range); the knn / hybrid
        search it restricts stays on the GPU.  The result is cached per index version and clause."""
        from .hostquery import HostSearcher
        key = ("filter", json.dumps(nodes, sort_keys=True, default=str))
        hit = self.host_views.get(key)
        if hit is None:
            hs = HostSearcher(self)
            match = np.ones(len(self.sources), dtype=bool)
            for node in nodes:
                match &= hs.eval(node)[1]
            hit = self.host_views[key] = np.flatnonzero(match).astype(np.int64)
        return hit

    def _filter_rows(self, filters, host_filters=()) -> np.ndarray:
        """Rows passing every term filter, ascending.  The keyword fields the hot path filters on (patientId,
        doc_type ...) keep value -> rows lists, so a patient filter costs O(rows of the patient), not O(index)."""
        out = self._host_filter_rows(list(host_filters)) if host_filters else None
        for f, v in filters:
            if f in self.kw:
                rows = np.asarray(self.kw[f].get(v, []), dtype=np.int64)
            else:
                rows = np.asarray([r for r, src in enumerate(self.sources) if src is not None and src.get(f) == v],
                                  dtype=np.int64)
            out = np.unique(rows) if out is None else np.intersect1d(out, rows)
        return out if out is not None else np.arange(len(self.sources), dtype=np.int64)

    def _apply_filter(self, plan: Plan):
        self.engine.set_row_filter_rows(self._filter_rows(plan.filters, plan.host_filters), len(self.sources))

    def _postings_of(self, gids) -> int:
        """Postings the listed global term ids hold together, as of the last sync."""
        g = self._global_df
        if g is None or len(g) != sum(f.df.size for f in self.text.fields.values()):
            g = self._global_df = (np.concatenate([self.text.fields[n].df for n in self.text.order])
                                   if self.text.order else np.zeros(0, dtype=np.int64))
        a = np.asarray(gids, dtype=np.int64)
        a = a[(a >= 0) & (a < g.size)]
        return int(g[a].sum())

    def _sync_text(self):
        if self.text.dirty and self.engine is not None:
            self.text.sync_device(self.engine, len(self.sources))
            self._global_df = None
            self.engine.set_vocab_raw(*self.text.vocab_blob())

    # -- search ------------------------------------------------------------------------------------------
    def _hits(self, pairs, with_vectors: bool = True) -> list[dict]:
        """[(row, score)] -> hit dicts.  `_source` comes back whole, like OpenSearch returns it: the vectors of all
        hits are read back from the device store in one gather (float32 -> python floats, the same doubles the
        reference's JSON round trip yields)."""
        pairs = list(pairs)
        vf = self.vector_field or "embedding"
        with_vec = [r for r, _ in pairs if self.sources[r] is not None and self.has_vec[r]] if with_vectors else []
        vecs = {}
        if with_vec:
            with self.lock:
                got = self.engine.read_rows_list(with_vec)
            vecs = {r: got[i].tolist() for i, r in enumerate(with_vec)}
        out = []
        for r, s in pairs:
            src = self.sources[r]
            if r in vecs:
                src = dict(src)
                src[vf] = vecs[r]
            out.append({"_index": self.name, "_id": self.ids[r], "_score": None if s is None else float(s),
                        "_source": src})
        return out

    def _hit(self, row: int, score: float) -> dict:
        return self._hits([(row, score)])[0]

    def _passes(self, row: int, plan: Plan) -> bool:
        src = self.sources[row] or {}
        if not all(src.get(f) == v for f, v in plan.filters):
            return False
        if plan.host_filters:
            rows = self._host_filter_rows(plan.host_filters)
            i = int(np.searchsorted(rows, row))
            return i < rows.size and rows[i] == row
        return True

    def search(self, plan: Plan) -> list[dict]:
        if plan.size <= 0 or self.engine is None or not self.sources:
            return []
        if plan.kind == "knn" and not plan.host_filters and not (plan.filters and self.knn_filter == "pre"):
            return self._search(plan)         # takes the lock per engine call, so concurrent callers can coalesce
        with self.lock:
            return self._search(plan)

    def _search(self, plan: Plan) -> list[dict]:
        eng = self.engine
        if plan.kind == "match_all":
            rows = [r for r in range(len(self.sources)) if self.sources[r] is not None][: plan.size]
            return self._hits((r, 1.0) for r in rows)
        if plan.vector is not None and plan.knn_field != (self.vector_field or "embedding"):
            raise RequestError(f"field {plan.knn_field!r} is not a knn_vector")
        q = None
        if plan.vector is not None:
            q = np.asarray(plan.vector, dtype=np.float32).reshape(1, -1)
            if q.shape[1] != self.dim:
                raise RequestError(f"query vector length {q.shape[1]} != dimension {self.dim}")
        has_filter = bool(plan.filters or plan.host_filters)
        if max(plan.knn_k, plan.size) > MAX_K:
            raise RequestError(f"size / k above {MAX_K} is not supported (asked for {max(plan.knn_k, plan.size)})")
        if plan.kind == "knn":
            k = max(plan.knn_k, 1)
            if has_filter and self.knn_filter == "pre":
                # exact filtered kNN: the k best rows OF THE PATIENT come back instead of whichever
                # of the global k nearest happen to pass.  A selective filter (a patient's few hundred rows) is scored
                # row by row, no corpus pass; a broad one is folded into the scan as a pass mask.
                frows = self._filter_rows(plan.filters, plan.host_filters)
                if frows.size <= FILTER_LIST_MAX and len(eng.devices) == 1:
                    rows, scores = eng.search_knn_filtered(q, k, [frows])
                    hits = [(int(r), float(s) * plan.knn_boost) for r, s in zip(rows[0], scores[0]) if r >= 0]
                    return self._hits(hits[: plan.size])
                self._apply_filter(plan)
                eng.set_knn_prefilter(True)
                try:
                    rows, scores = eng.search_knn(q, k)
                finally:
                    eng.set_knn_prefilter(False)
                    eng.set_row_filter(None)
                hits = [(int(r), float(s) * plan.knn_boost) for r, s in zip(rows[0], scores[0]) if r >= 0]
                return self._hits(hits[: plan.size])
            rows, scores = self._knn(q, k)
            # OpenSearch applies bool.filter to the k nearest neighbours of the nmslib engine (post-filter)
            hits = [(int(r), float(s) * plan.knn_boost) for r, s in zip(rows[0], scores[0]) if r >= 0]
            hits = [(r, s) for r, s in hits if self._passes(r, plan)]
            return self._hits(hits[: plan.size])
        # hybrid: bool.should boosted sum
        k = max(plan.size, 1)
        # text clauses: multi_match best_fields = max over the fields of the field's term sum (every declared field the
        # index holds postings for), bool.should = sum over the clauses; term weights and structure marks go to the
        # device as (term, weight, flag) lists (bit 0 = last term of a field group, bit 1 = last term of a clause)
        ids, ws, flags = [], [], []
        have_text = False
        for clause in plan.text:
            first_of_clause = len(ids)
            for fname, fb in clause.fields:
                fld = self.text.fields.get(fname)
                if fld is None:
                    continue     # no document carries the field: nothing can match
                self._sync_text()
                have_text = True
                base = self.text.base[fname]
                bo = float(np.float32(clause.boost) * np.float32(fb))
                if self.text.types.get(fname) == "keyword":
                    t_ids, t_ws = fld.exact_weighted_terms([clause.query], bo)      # the whole string is the term
                elif clause.fuzziness == "AUTO":
                    def expand(tok, me, base=base, n=len(fld.vocab)):
                        t, e = eng.fuzzy_expand(tok, me, base, base + n)
                        return t - base, e
                    t_ids, t_ws = fld.fuzzy_weighted_terms(clause.query, bo, expand)
                else:
                    from .text import analyze
                    t_ids, t_ws = fld.exact_weighted_terms(analyze(clause.query), bo)
                if not t_ids:
                    continue
                ids.extend(base + t for t in t_ids)
                ws.extend(t_ws)
                flags.extend([0] * (len(t_ids) - 1) + [1])
            if len(ids) > first_of_clause:
                flags[-1] |= 2
        qterms = [ids] if have_text else None
        qweights = [ws] if have_text else None
        qflags = [flags] if have_text else None
        w_text = 0.0
        if q is not None and max(plan.knn_k, 1) != k:
            raise NotImplementedError("knn k different from size in a hybrid query")
        if q is None and qterms is None:
            return []
        # bool.filter: only rows that satisfy every filter may score.  A selective filter (every production query carries
        # its own `term patientId`, app/main.py:1599-1604) travels as the list of passing rows and those rows are scored
        # directly -- text clauses by posting lookups, the knn clause still the k nearest of the whole corpus; a broad
        # filter becomes a device-side pass mask for the tile kernel.
        if has_filter and self.batch_window_s is None and len(eng.devices) == 1:
            frows = self._filter_rows(plan.filters, plan.host_filters)
            # ... unless the lookups would cost more than walking the postings: a fuzzy query string expands to a few
            # hundred terms, and rows x terms binary searches then lose against the tile kernel over the terms' lists
            # (RASS_B200_FILTER_ROUTE=list|mask forces one of the two: the A/B switch of the measurement)
            n_terms = len(ids) if have_text else 0
            lookups = frows.size * max(n_terms, 1) * 16
            walk = (self._postings_of(ids) if have_text else 0) + len(self.sources) // 4
            route = os.environ.get("RASS_B200_FILTER_ROUTE") or ("list" if lookups <= walk else "mask")
            if frows.size <= FILTER_LIST_MAX and route == "list":
                rows, scores = eng.search_hybrid_filtered(q, qterms, w_text, plan.knn_boost, k, [frows],
                                                          qweights=qweights, qflags=qflags)
                return self._hits([(int(r), float(s)) for r, s in zip(rows[0], scores[0]) if r >= 0][: plan.size])
        if has_filter:
            self._apply_filter(plan)
        else:
            eng.set_row_filter(None)
        try:
            if q is not None and self.batch_window_s is not None:
                # the corpus pass of the knn clause is shared with concurrent requests (MicroBatcher); the text
                # clauses and the filter of THIS request are fused against its k nearest afterwards
                self.lock.release()
                try:
                    knn_rows, knn_scores = self._knn(q, k)
                finally:
                    self.lock.acquire()
                if has_filter:
                    self._apply_filter(plan)                # another request may have changed it meanwhile
                else:
                    eng.set_row_filter(None)
                rows, scores = eng.fuse_hybrid(qterms, w_text, knn_rows, knn_scores, plan.knn_boost, k,
                                               qweights=qweights, qflags=qflags)
            else:
                rows, scores = eng.search_hybrid(q, qterms, w_text, plan.knn_boost, k, qweights=qweights,
                                                 qflags=qflags)
        finally:
            if has_filter:
                eng.set_row_filter(None)
        return self._hits([(int(r), float(s)) for r, s in zip(rows[0], scores[0]) if r >= 0][: plan.size])


class IndicesClient:
    def __init__(self, client: "B200Client"):
        self._c = client

    def exists(self, index: str, **_) -> bool:
        return index in self._c._indices

    def create(self, index: str, body: dict | None = None, **_) -> dict:
        if index in self._c._indices:
            raise RequestError(f"resource_already_exists_exception: index [{index}] already exists")
        idx = _Index(index, body, self._c.device, self._c.knn_filter, self._c.devices, self._c.sim_boost)
        idx.batch_window_s = self._c.batch_window_s
        self._c._indices[index] = idx
        return {"acknowledged": True, "shards_acknowledged": True, "index": index}

    def delete(self, index: str, **_) -> dict:
        idx = self._c._indices.pop(index, None)
        if idx is None:
            raise NotFoundError(f"index_not_found_exception: no such index [{index}]")
        idx.close()
        return {"acknowledged": True}

    def refresh(self, index: str | None = None, **_) -> dict:
        return {"_shards": {"total": 1, "successful": 1, "failed": 0}}


class B200Client:
    """Drop-in for `OpenSearch(hosts=[...], ...)`; connection arguments are accepted and ignored."""

    def __init__(self, hosts=None, device: int = 0, knn_filter: str = "post", batch_window_ms: float | None = None,
                 devices=None, bm25_legacy_boost: bool = False, **_ignored):
        """knn_filter: "post" (default) applies bool.filter to the k nearest neighbours, as OpenSearch's nmslib engine
        does; "pre" returns the exact top-k among the rows passing the filter (device-side pass mask in the scan).
        bm25_legacy_boost: score text clauses the way Lucene's LegacyBM25Similarity does -- every boost times
        float32(1 + k1) = 2.2f before the scorer is built, otherwise the same arithmetic (oracle/SEMANTICS.md: which of
        the two a given OpenSearch release ships is part of what is unpinned; the default is the form SURVEY.md states)."""
        if knn_filter not in ("post", "pre"):
            raise ValueError("knn_filter must be 'post' or 'pre'")
        self.device = device
        # devices: spread every index over these GPUs in THIS process (one sharded engine handle); otherwise
        # RASS_B200_DEVICES or the create body's settings.index.number_of_shards decide (resolve_devices)
        self.devices = list(devices) if devices else None
        self.knn_filter = knn_filter
        self.sim_boost = np.float32(1.0) + np.float32(1.2) if bm25_legacy_boost else np.float32(1.0)
        # batch_window_ms: concurrent knn searches (threads) wait up to this long to share one corpus pass
        self.batch_window_s = None if batch_window_ms is None else batch_window_ms / 1e3
        self._indices: dict[str, _Index] = {}
        self.indices = IndicesClient(self)

    def _get(self, index: str) -> _Index:
        try:
            return self._indices[index]
        except KeyError:
            raise NotFoundError(f"index_not_found_exception: no such index [{index}]") from None

    def info(self, **_) -> dict:
        ver = capi.lib().rass_version().decode()
        return {"name": "rass-b200", "cluster_name": "rass-b200", "version": {"distribution": "rass-b200",
                "number": "2.11.1", "build_type": ver}, "tagline": "The OpenSearch Project: https://opensearch.org/"}

    def ping(self, **_) -> bool:
        return True

    def count(self, index: str | None = None, body=None, **_) -> dict:
        n = self._get(index).n_docs if index else sum(i.n_docs for i in self._indices.values())
        return {"count": n, "_shards": {"total": 1, "successful": 1, "skipped": 0, "failed": 0}}

    def index(self, index: str, body: dict, id: str | None = None, routing=None, **_) -> dict:
        idx = self._get(index)
        _id = id if id is not None else f"auto-{len(idx.ids)}"
        created = _id not in idx.row_of
        idx.index_batch([(_id, body)])
        return {"_index": index, "_id": _id, "result": "created" if created else "updated"}

    def bulk_actions(self, actions) -> tuple[int, list]:
        """helpers.bulk: consecutive actions for one index are ingested as one batch (one device append)."""
        ok, errors = 0, []
        run_index, run = None, []

        def flush():
            nonlocal ok, run
            if not run:
                return
            try:
                self._get(run_index).index_batch(run)
                ok += len(run)
            except Exception as e:  # per-batch error report, as helpers.bulk(raise_on_error=False) would give
                errors.extend({"index": {"_index": run_index, "_id": i, "error": str(e)}} for i, _ in run)
            run = []

        for a in actions:
            op = a.get("_op_type", "index")
            if op != "index":
                errors.append({op: {"_id": a.get("_id"), "error": "unsupported _op_type"}})
                continue
            name = a["_index"]
            if name != run_index:
                flush()
                run_index = name
            src = a.get("_source")
            if src is None:
                src = {k: v for k, v in a.items() if not k.startswith("_")}
            _id = a.get("_id")
            if _id is None:
                _id = f"auto-{len(self._get(name).ids) + len(run)}"
            run.append((str(_id), src))
        flush()
        return ok, errors

    def search(self, index: str | None = None, body: dict | None = None, routing=None, **_) -> dict:
        t0 = time.perf_counter()
        idx = self._get(index)
        aggs = None
        try:
            plan = parse_search_body(body or {})          # the hot path: knn / hybrid shapes, on the GPU
            hits = idx.search(plan)
            total = len(hits)
        except NotImplementedError:
            hits, aggs, total = idx.host_search(body or {})    # everything else: host-side (raises if unsupported)
        scores = [h["_score"] for h in hits if h.get("_score") is not None]
        resp = {"took": int((time.perf_counter() - t0) * 1e3), "timed_out": False,
                "_shards": {"total": 1, "successful": 1, "skipped": 0, "failed": 0},
                "hits": {"total": {"value": total, "relation": "eq"},
                         "max_score": max(scores, default=None), "hits": hits}}
        if aggs is not None:
            resp["aggregations"] = aggs
        return resp

    def close(self):
        for idx in self._indices.values():
            idx.close()
        self._indices.clear()

    # -- snapshot / restore (the reference relies on the OpenSearch container keeping its index on disk) ------
    def snapshot(self, directory: str) -> list[str]:
        """Writes every index as <name>.vec (rass_save: stored rows + tombstones) and <name>.json (mapping, _ids and
        vector-less _sources in row order).  Returns the index names written."""
        os.makedirs(directory, exist_ok=True)
        names = []
        for name, idx in self._indices.items():
            with idx.lock:
                meta = {"name": name, "body": idx.body, "ids": idx.ids, "sources": idx.sources,
                        "has_vec": idx.has_vec, "n_docs": idx.n_docs, "dim": idx.dim}
                with open(os.path.join(directory, name + ".json"), "w") as f:
                    json.dump(meta, f)
                if idx.engine is not None:
                    idx.engine.save(os.path.join(directory, name + ".vec"))
            names.append(name)
        return names

    def restore(self, directory: str) -> list[str]:
        """Loads every <name>.json / <name>.vec pair of `directory`; rows keep their ids, so results are identical
        to the snapshotted client's.  The postings are rebuilt from the restored text on the first hybrid query."""
        names = []
        for fn in sorted(os.listdir(directory)):
            if not fn.endswith(".json"):
                continue
            with open(os.path.join(directory, fn)) as f:
                meta = json.load(f)
            name = meta["name"]
            if name in self._indices:
                raise RequestError(f"resource_already_exists_exception: index [{name}] already exists")
            idx = _Index(name, meta["body"], self.device, self.knn_filter, self.devices, self.sim_boost)
            idx.batch_window_s = self.batch_window_s
            idx.ids, idx.sources, idx.has_vec = meta["ids"], meta["sources"], meta["has_vec"]
            idx.n_docs = meta["n_docs"]
            idx.row_of = {i: r for r, i in enumerate(idx.ids) if i is not None}
            vec = os.path.join(directory, name + ".vec")
            if os.path.exists(vec):
                idx.dim = idx.dim or meta.get("dim")
                idx._ensure_engine(idx.dim).load(vec)
            for row, src in enumerate(idx.sources):
                if src is not None:
                    idx.text.set_doc(row, src, fresh=True)
                    idx._kw_add(row, src)
            self._indices[name] = idx
            names.append(name)
        return names


def bulk(client: B200Client, actions, **_) -> tuple[int, list]:
    """opensearchpy.helpers.bulk(client, actions) -> (success_count, errors)   (app/main.py:1237)"""
    return client.bulk_actions(actions)
