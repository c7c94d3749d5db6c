"""Random-projection / E2-LSH index of the reference (src/algorithms/lsh.py) with the candidate
re-scoring on the B200 rerank kernel.

Split of work:
* ``LSHIndexer.build``  - hash tables stay host-side Python objects, as in the reference
  (lsh.py:95-138), but the N x T x H projections are one batched matrix product instead of a
  per-row loop.  Draw order of the random state, bit weights and key types are the reference's, so
  the same seed gives the same buckets.
* ``LSHSearcher``       - bucket lookup and vote ordering on the host (lsh.py:219-240), then ONE
  ``vdb_rerank_topk`` launch re-scores every query's candidates (lsh.py:242-283 did this per query
  in NumPy).  Distances follow lsh.py:242-250: cosine -> ``1 - v.q``, l2 -> Euclidean norm.
  Queries without any bucket hit fall back to an exact scan over all rows (lsh.py:232-233) on the
  flat kernel, or return (+inf, -1) padding when ``fallback_to_bruteforce`` is off."""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .base_algorithm import BaseAlgorithm
from .modular import BaseIndexer, BaseSearcher, IndexArtifact, register_indexer, register_searcher


def _normalize_rows(matrix: np.ndarray) -> np.ndarray:
    norms = np.linalg.norm(matrix, axis=1, keepdims=True)
    return np.divide(matrix, norms, out=np.zeros_like(matrix), where=norms > 0)


def _project(vectors: np.ndarray, projections: np.ndarray) -> np.ndarray:
    """[n, d] x [T, H, d] -> [n, T, H] float32 (one sgemm)."""
    t, h, d = projections.shape
    return (vectors @ projections.reshape(t * h, d).T).reshape(vectors.shape[0], t, h)


def _cosine_keys(proj: np.ndarray, bit_weights: np.ndarray) -> np.ndarray:
    """Sign bits weighted 1 << arange(H) -> uint64 key per (row, table)  (lsh.py:78-80)."""
    return ((proj >= 0).astype(np.uint64) * bit_weights[None, None, :]).sum(axis=2)


def _l2_codes(proj: np.ndarray, offsets: np.ndarray, bucket_width: float) -> np.ndarray:
    """floor((P v + b) / w) as int32 per (row, table, hash)  (lsh.py:82-84)."""
    return np.floor((proj + offsets[None, :, :]) / bucket_width).astype(np.int32)


class LSHIndexer(BaseIndexer):
    SUPPORTED_METRICS = {"cosine", "l2"}

    def __init__(self, name: str, dimension: int, metric: str = "cosine", num_tables: int = 8, hash_size: int = 16,
                 bucket_width: float = 4.0, seed: int = 42, **kwargs: Any) -> None:
        super().__init__(name, dimension, metric, num_tables=num_tables, hash_size=hash_size, bucket_width=bucket_width,
                         seed=seed, **kwargs)
        if metric not in self.SUPPORTED_METRICS:
            raise ValueError(f"LSHIndexer supports metrics {self.SUPPORTED_METRICS}, received '{metric}'")
        if hash_size <= 0:
            raise ValueError("hash_size must be positive")
        if num_tables <= 0:
            raise ValueError("num_tables must be positive")
        if metric == "l2" and bucket_width <= 0:
            raise ValueError("bucket_width must be positive for L2 LSH")
        self.num_tables, self.hash_size, self.bucket_width, self.seed = num_tables, hash_size, bucket_width, seed

    def build(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> IndexArtifact:
        if vectors.shape[1] != self.dimension:
            raise ValueError(f"Expected vectors with dimension {self.dimension}, received {vectors.shape[1]}")
        rng = np.random.RandomState(self.seed)                      # draw order as lsh.py:70-76,99-101
        projections = rng.normal(size=(self.num_tables, self.hash_size, self.dimension)).astype(np.float32)
        offsets = None
        if self.metric == "l2":
            offsets = rng.uniform(0.0, self.bucket_width, size=(self.num_tables, self.hash_size)).astype(np.float32)
        bit_weights = (1 << np.arange(self.hash_size, dtype=np.uint64))
        store = vectors.astype(np.float32, copy=True)
        if self.metric == "cosine":
            store = _normalize_rows(store)

        tables: List[Dict[Any, np.ndarray]] = []
        # the same tables as flat arrays for the device-side candidate generation (vdb_lsh_candidates): per table the
        # rows grouped by bucket (`ids`), and per bucket its (start, length) - looked up by key on the host
        csr_ids = np.empty((self.num_tables, store.shape[0]), dtype=np.int32)
        csr_slots: List[Any] = []
        proj = _project(store, projections)
        if self.metric == "cosine":
            keys = _cosine_keys(proj, bit_weights)                   # [n, T]
            for t in range(self.num_tables):
                order = np.argsort(keys[:, t], kind="stable")        # rows of a bucket stay in insertion order
                uniq, start = np.unique(keys[order, t], return_index=True)
                stops = np.append(start[1:], order.size)
                tables.append({int(k): order[a:b] for k, a, b in zip(uniq.tolist(), start.tolist(), stops.tolist())})
                csr_ids[t] = order
                csr_slots.append((uniq.astype(np.uint64), start.astype(np.int64), (stops - start).astype(np.int32)))
        else:
            codes = _l2_codes(proj, offsets, self.bucket_width)       # [n, T, H]
            for t in range(self.num_tables):
                uniq, inverse = np.unique(codes[:, t, :], axis=0, return_inverse=True)
                inverse = inverse.reshape(-1)
                order = np.argsort(inverse, kind="stable")
                start = np.searchsorted(inverse[order], np.arange(uniq.shape[0]))
                stops = np.append(start[1:], order.size)
                tables.append({tuple(u): order[a:b] for u, a, b in zip(uniq.tolist(), start.tolist(), stops.tolist())})
                csr_ids[t] = order
                csr_slots.append({tuple(u): (a, b - a) for u, a, b in zip(uniq.tolist(), start.tolist(), stops.tolist())})

        meta: Dict[str, Any] = {"metric": self.metric, "num_tables": self.num_tables, "hash_size": self.hash_size}
        if self.metric == "cosine":
            meta["normalize_queries"] = True
        else:
            meta["bucket_width"] = self.bucket_width
        data: Dict[str, Any] = {"tables": tables, "projections": projections, "vector_store": store,
                                "bit_weights": bit_weights, "csr_ids": csr_ids, "csr_slots": csr_slots}
        if offsets is not None:
            data["offsets"] = offsets
        return IndexArtifact(kind="lsh", data=data, metadata=meta)


register_indexer("LSHIndexer", LSHIndexer)


class LSHSearcher(BaseSearcher):
    def __init__(self, name: str, dimension: int, metric: str = "cosine", candidate_multiplier: float = 4.0,
                 max_candidates: Optional[int] = None, fallback_to_bruteforce: bool = True, **kwargs: Any) -> None:
        super().__init__(name, dimension, metric, candidate_multiplier=candidate_multiplier, max_candidates=max_candidates,
                         fallback_to_bruteforce=fallback_to_bruteforce, **kwargs)
        if candidate_multiplier <= 0:
            raise ValueError("candidate_multiplier must be positive")
        self.candidate_multiplier = candidate_multiplier
        self.max_candidates = max_candidates
        self.fallback_to_bruteforce = fallback_to_bruteforce
        self._flat = None

    def attach(self, artifact: IndexArtifact, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        from .. import engine
        if artifact.kind != "lsh":
            raise ValueError("LSHSearcher can only attach to artifacts produced by LSHIndexer")
        data = artifact.data
        self.tables = data["tables"]
        self.projections: np.ndarray = data["projections"]
        self.vector_store: np.ndarray = data["vector_store"]
        self.bit_weights: np.ndarray = data["bit_weights"]
        self.offsets: Optional[np.ndarray] = data.get("offsets")
        self.metric = artifact.metadata.get("metric", self.metric)
        if self.metric not in {"cosine", "l2"}:
            raise ValueError(f"Unsupported metric '{self.metric}' for LSHSearcher")
        self.normalize_queries = artifact.metadata.get("normalize_queries", False)
        self.hash_size = artifact.metadata.get("hash_size", self.vector_store.shape[1])
        self.num_tables = artifact.metadata.get("num_tables", len(self.tables))
        self.bucket_width = artifact.metadata.get("bucket_width", None)
        if self.metric == "l2" and (self.offsets is None or self.bucket_width is None):
            raise RuntimeError("L2 hashing requires offsets and bucket_width")
        # the store is already normalised for cosine: score it as inner product, report 1 - score
        self._reranker = engine.Reranker(self.vector_store, "l2" if self.metric == "l2" else "ip", self.params.get("device"))
        self._flags = engine._lib.OUT_SQRT if self.metric == "l2" else engine._lib.OUT_ONE_MINUS
        # candidate generation: "device" (bucket union + vote order by vdb_lsh_candidates) or "host" (the NumPy
        # restatement of the reference's Counter walk; kept as the checker of the device path)
        self.candidate_generation = str(self.params.get("candidate_generation", "device"))
        self._csr_slots = data.get("csr_slots")
        self._csr_ids = None
        if self.candidate_generation == "device" and data.get("csr_ids") is not None and self.num_tables < 64:
            import torch
            self._csr_ids = torch.from_numpy(np.ascontiguousarray(data["csr_ids"], dtype=np.int32)).to(self._reranker.dev)
        self._prepared = True

    def memory_bytes(self) -> int:
        total = self._reranker.memory_bytes() if self._prepared else 0
        total += self._csr_ids.numel() * 4 if getattr(self, "_csr_ids", None) is not None else 0
        return total + (self._flat.memory_bytes() if self._flat is not None else 0)

    # ---- host side: hashing, bucket lookup, vote ordering ------------------------------------
    def _prepare_queries(self, queries: np.ndarray) -> np.ndarray:
        q = np.asarray(queries)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        q = q.astype(np.float32, copy=True)
        return _normalize_rows(q) if self.normalize_queries else q

    def _hash_queries(self, q: np.ndarray) -> List[List[Any]]:
        proj = _project(q, self.projections.reshape(self.num_tables, self.hash_size, self.dimension))
        if self.metric == "cosine":
            return _cosine_keys(proj, self.bit_weights[: self.hash_size]).tolist()
        codes = _l2_codes(proj, self.offsets, self.bucket_width)
        return [[tuple(codes[r, t].tolist()) for t in range(self.num_tables)] for r in range(q.shape[0])]

    def _ordered_candidates(self, keys: List[Any]) -> np.ndarray:
        """Union of the T buckets ordered like ``Counter.most_common()``: votes descending, first
        seen first among equal votes (lsh.py:219-229)."""
        hits = [b for b in (self.tables[t].get(key) for t, key in enumerate(keys)) if b is not None and len(b)]
        if not hits:
            return np.empty(0, dtype=np.int64)
        seen = np.concatenate(hits)
        uniq, first, votes = np.unique(seen, return_index=True, return_counts=True)
        return uniq[np.lexsort((first, -votes))].astype(np.int64)

    def _cap(self, k: int) -> int:
        if self.max_candidates is not None:
            return int(self.max_candidates)
        return max(k, int(math.ceil(self.candidate_multiplier * k)))

    # ---- device side ---------------------------------------------------------------------------
    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        from .. import engine
        import torch
        if not self._prepared:
            raise RuntimeError("LSHSearcher not attached to an index")
        k = int(k)
        q = self._prepare_queries(queries)
        nq = q.shape[0]
        cap = self._cap(k)
        if self._csr_ids is not None:
            return self._batch_search_device(q, k, cap)
        lists = [self._ordered_candidates(keys)[:cap] for keys in self._hash_queries(q)]
        empty = np.array([c.size == 0 for c in lists], dtype=bool)
        out_d = np.full((nq, k), np.inf, dtype=np.float32)
        out_i = np.full((nq, k), -1, dtype=np.int64)
        rr = self._reranker
        with torch.cuda.device(rr.dev):
            if not empty.all():
                rows = np.nonzero(~empty)[0]
                width = max(lists[r].size for r in rows)
                cand = np.full((rows.size, width), -1, dtype=np.int64)
                for j, r in enumerate(rows):
                    cand[j, : lists[r].size] = lists[r]
                qd = engine.queries_to_device(q[rows], rr.dev, self.dimension)
                d, i = engine.results_to_host(*rr.search(qd, torch.from_numpy(cand).to(rr.dev), k, self._flags, float("inf")))
                out_d[rows], out_i[rows] = d, i
            if empty.any() and self.fallback_to_bruteforce:          # no bucket hit: score every row (lsh.py:232-233)
                if self._flat is None:
                    self._flat = engine.FlatShard(self.vector_store, "l2" if self.metric == "l2" else "ip", rr.dev)
                rows = np.nonzero(empty)[0]
                qd = engine.queries_to_device(q[rows], rr.dev, self.dimension)
                d, i = engine.results_to_host(*self._flat.search(qd, k, self._flags, float("inf")))
                out_d[rows], out_i[rows] = d, i
        return out_d, out_i

    def _bucket_segments(self, q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """(seg_off [nq, T] int64, seg_len [nq, T] int32): every query's bucket per table inside ``csr_ids``.
        Hashing is the host's NumPy arithmetic (bit-identical keys); the lookup is a binary search per table for
        the integer keys of the cosine hash, a dictionary lookup for the integer tuples of the L2 hash."""
        nq, T, n = q.shape[0], self.num_tables, int(self._csr_ids.shape[1])
        proj = _project(q, self.projections.reshape(T, self.hash_size, self.dimension))
        seg_off = np.zeros((nq, T), dtype=np.int64)
        seg_len = np.zeros((nq, T), dtype=np.int32)
        if self.metric == "cosine":
            keys = _cosine_keys(proj, self.bit_weights[: self.hash_size]).astype(np.uint64)          # [nq, T]
            for t in range(T):
                ukeys, starts, lens = self._csr_slots[t]
                pos = np.minimum(np.searchsorted(ukeys, keys[:, t]), ukeys.shape[0] - 1)
                hit = ukeys[pos] == keys[:, t]
                seg_off[:, t] = t * n + starts[pos]
                seg_len[:, t] = np.where(hit, lens[pos], 0)
        else:
            codes = _l2_codes(proj, self.offsets, self.bucket_width)
            for t in range(T):
                slots = self._csr_slots[t]
                for r in range(nq):
                    a, m = slots.get(tuple(codes[r, t].tolist()), (0, 0))
                    seg_off[r, t], seg_len[r, t] = t * n + a, m
        return seg_off, seg_len

    def _batch_search_device(self, q: np.ndarray, k: int, cap: int, max_entries: int = 1 << 27) -> Tuple[np.ndarray, np.ndarray]:
        """Candidate union, vote order, budget cut and rerank on the device; the host only hashes and looks buckets up."""
        from .. import engine
        import torch
        lib, rr = engine._lib.load(), self._reranker
        nq, T = q.shape[0], self.num_tables
        seg_off, seg_len = self._bucket_segments(q)
        per_query = seg_len.sum(axis=1, dtype=np.int64)
        out_d = np.full((nq, k), np.inf, dtype=np.float32)
        out_i = np.full((nq, k), -1, dtype=np.int64)
        empty = per_query == 0
        with torch.cuda.device(rr.dev):
            lo = 0
            while lo < nq:                                   # query chunks of at most `max_entries` bucket entries
                hi = lo + 1
                total = int(per_query[lo])
                while hi < nq and total + int(per_query[hi]) <= max_entries:
                    total += int(per_query[hi])
                    hi += 1
                if total >= (1 << 31):
                    raise RuntimeError(f"one query unions {total} bucket entries; the device path handles < 2^31")
                n_c = hi - lo
                elem_off = np.zeros(n_c + 1, dtype=np.int64)
                np.cumsum(per_query[lo:hi], out=elem_off[1:])
                dev = rr.dev
                so = torch.from_numpy(np.ascontiguousarray(seg_off[lo:hi])).to(dev)
                sl = torch.from_numpy(np.ascontiguousarray(seg_len[lo:hi])).to(dev)
                eo = torch.from_numpy(elem_off).to(dev)
                cand = torch.empty((n_c, cap), dtype=torch.int64, device=dev)
                ws = None
                if total > 0:
                    nbytes = lib.vdb_lsh_candidates_workspace_bytes(total, n_c)
                    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                engine.check(lib.vdb_lsh_candidates(engine.ptr(self._csr_ids), int(self._csr_ids.shape[1]), engine.ptr(so), engine.ptr(sl),
                                                    engine.ptr(eo), n_c, T, total, int(per_query[lo:hi].max()), cap, engine.ptr(cand), None,
                                                    engine.ptr(ws), 0 if ws is None else ws.numel(), engine._stream(dev)),
                             "vdb_lsh_candidates")
                qd = engine.queries_to_device(q[lo:hi], dev, self.dimension)
                d, i = engine.results_to_host(*rr.search(qd, cand, k, self._flags, float("inf")))
                out_d[lo:hi], out_i[lo:hi] = d, i
                lo = hi
            if empty.any() and self.fallback_to_bruteforce:          # no bucket hit: score every row (lsh.py:232-233)
                if self._flat is None:
                    self._flat = engine.FlatShard(self.vector_store, "l2" if self.metric == "l2" else "ip", rr.dev)
                rows = np.nonzero(empty)[0]
                qd = engine.queries_to_device(q[rows], rr.dev, self.dimension)
                d, i = engine.results_to_host(*self._flat.search(qd, k, self._flags, float("inf")))
                out_d[rows], out_i[rows] = d, i
        return out_d, out_i

    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        distances, indices = self.batch_search(np.asarray(query).reshape(1, -1), k)
        return distances[0], indices[0]


register_searcher("LSHSearcher", LSHSearcher)


class LSH(BaseAlgorithm):
    """Indexer + searcher pair as one algorithm (lsh.py:304-359)."""

    def __init__(self, name: str, dimension: int, metric: str = "cosine", num_tables: int = 8, hash_size: int = 16,
                 bucket_width: float = 4.0, candidate_multiplier: float = 4.0, max_candidates: Optional[int] = None,
                 fallback_to_bruteforce: bool = True, seed: int = 42, **kwargs: Any) -> None:
        super().__init__(name, dimension, metric=metric, num_tables=num_tables, hash_size=hash_size,
                         bucket_width=bucket_width, candidate_multiplier=candidate_multiplier,
                         max_candidates=max_candidates, fallback_to_bruteforce=fallback_to_bruteforce, seed=seed, **kwargs)
        self.metric = metric
        self.indexer = LSHIndexer(name=f"{name}_indexer", dimension=dimension, metric=metric, num_tables=num_tables,
                                  hash_size=hash_size, bucket_width=bucket_width, seed=seed)
        self.searcher = LSHSearcher(name=f"{name}_searcher", dimension=dimension, metric=metric,
                                    candidate_multiplier=candidate_multiplier, max_candidates=max_candidates,
                                    fallback_to_bruteforce=fallback_to_bruteforce)

    def build_index(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        artifact = self.indexer.build(vectors, metadata)
        self.searcher.attach(artifact, vectors, metadata)
        self.index_built = True

    def get_memory_usage(self) -> int:
        return self.searcher.memory_bytes()

    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self.index_built:
            raise RuntimeError("Index has not been built for LSH algorithm")
        return self.searcher.search(query, k)

    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self.index_built:
            raise RuntimeError("Index has not been built for LSH algorithm")
        return self.searcher.batch_search(queries, k)


__all__ = ["LSHIndexer", "LSHSearcher", "LSH"]
