"""Indexer / searcher plug-in framework of the reference (src/algorithms/modular.py), with the
scan + top-k arithmetic on the B200 kernels.

Kept byte-compatible with the reference's YAML surface: the same class names are registered
under the same ``type`` strings (``BruteForceIndexer``, ``FaissFactoryIndexer``,
``FaissIVFIndexer``, ``FaissLSHIndexer``, ``LinearSearcher``, ``FaissSearcher``), artifacts are
``IndexArtifact(kind, data, metadata)`` with the same ``kind`` tags and metadata keys, and
``CompositeAlgorithm`` wires an indexer to a searcher exactly as modular.py:554-622 does.
What changed is *where the arithmetic runs*:

=====================  ==========================================  ===============================
reference              arithmetic there                            here
=====================  ==========================================  ===============================
LinearSearcher         NumPy broadcast / ``@`` + argpartition      ``engine.FlatShard`` (tcgen05)
FaissFactoryIndexer    ``faiss.index_factory`` Flat / IVFn,Flat    ``indexes.GpuIndexFlat/IVFFlat``
FaissLSHIndexer        ``faiss.IndexLSH``                          ``indexes.GpuIndexLSH``
FaissSearcher rerank   per-query Python loop over candidates       ``engine.Reranker`` (one kernel)
=====================  ==========================================  ===============================

``HNSWIndexer`` (graph traversal) is outside this build."""
from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Type

import numpy as np

from .base_algorithm import BaseAlgorithm


@dataclass
class IndexArtifact:
    """What an indexer hands to a searcher (modular.py:19-25)."""

    kind: str
    data: Any
    metadata: Dict[str, Any] = field(default_factory=dict)


class _Component(ABC):
    def __init__(self, name: str, dimension: int, metric: str = "l2", **kwargs: Any) -> None:
        self.name = name
        self.dimension = dimension
        self.metric = metric
        self.params = kwargs

    def describe(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {"name": self.name, "type": self.__class__.__name__, "metric": self.metric}
        if self.params:
            out["params"] = copy.deepcopy(self.params)
        return out


class BaseIndexer(_Component):
    """Indexing strategy (modular.py:28-49)."""

    @abstractmethod
    def build(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> IndexArtifact:
        """Build an index artifact from ``vectors`` [n, dimension]."""


class BaseSearcher(_Component):
    """Search strategy over an artifact (modular.py:52-82)."""

    def __init__(self, name: str, dimension: int, metric: str = "l2", **kwargs: Any) -> None:
        super().__init__(name, dimension, metric, **kwargs)
        self._prepared = False

    @abstractmethod
    def attach(self, artifact: IndexArtifact, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        """Bind to an artifact before serving queries."""

    @abstractmethod
    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        """One query."""

    @abstractmethod
    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        """A batch of queries."""


INDEXER_REGISTRY: Dict[str, Type[BaseIndexer]] = {}
SEARCHER_REGISTRY: Dict[str, Type[BaseSearcher]] = {}


def register_indexer(name: str, cls: Type[BaseIndexer]) -> None:
    INDEXER_REGISTRY[name] = cls


def register_searcher(name: str, cls: Type[BaseSearcher]) -> None:
    SEARCHER_REGISTRY[name] = cls


def get_indexer_class(name: str) -> Type[BaseIndexer]:
    if name not in INDEXER_REGISTRY:
        raise ValueError(f"Unknown indexer type '{name}'. Available: {list(INDEXER_REGISTRY.keys())}")
    return INDEXER_REGISTRY[name]


def get_searcher_class(name: str) -> Type[BaseSearcher]:
    if name not in SEARCHER_REGISTRY:
        raise ValueError(f"Unknown searcher type '{name}'. Available: {list(SEARCHER_REGISTRY.keys())}")
    return SEARCHER_REGISTRY[name]


def _ensure_float32(vectors: np.ndarray) -> np.ndarray:
    if vectors.dtype == np.float32 and vectors.flags["C_CONTIGUOUS"]:
        return vectors
    return np.ascontiguousarray(vectors, dtype=np.float32)


def _as_batch(query: np.ndarray) -> np.ndarray:
    query = np.asarray(query)
    return query.reshape(1, -1) if query.ndim == 1 else query


# ------------------------------------------------------------------------------------------ indexers
class BruteForceIndexer(BaseIndexer):
    """Stores the raw vectors (modular.py:121-130).  The artifact stays a host array so any
    searcher can consume it; the upload to HBM happens when a searcher attaches."""

    def build(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> IndexArtifact:
        return IndexArtifact(kind="raw_vectors", data=_ensure_float32(vectors),
                             metadata={"metric": self.metric, "normalize_vectors": self.metric == "cosine"})


register_indexer("BruteForceIndexer", BruteForceIndexer)


class FaissFactoryIndexer(BaseIndexer):
    """``faiss.index_factory`` grammar -> CUDA index (modular.py:224-286).  cosine = inner product
    over L2-normalised rows with ``normalize_queries`` set, as in ``_prepare_data``
    (modular.py:238-262) - the normalisation itself runs on the device."""

    _RESERVED_PARAM_KEYS = {"index_key", "index_type", "device", "devices"}
    _TRAIN_PARAM_KEYS = ("niter", "seed", "max_points_per_centroid")      # shape the centroids: must be set BEFORE train()
    _RUNTIME_PARAM_KEYS = ("nprobe",)                                      # search-time attributes (modular.py:269-275)

    def __init__(self, name: str, dimension: int, metric: str = "l2", index_key: str = "Flat", **kwargs: Any) -> None:
        from ..indexes import _IVF_FLAT, _IVF_PQ, _IVF_SQ8, _PQ
        key = index_key.strip()
        if key != "Flat" and not any(rx.match(key) for rx in (_IVF_FLAT, _IVF_SQ8, _IVF_PQ, _PQ)):
            raise ValueError(f"index_key '{index_key}' is not supported by the CUDA build "
                             "(supported: 'Flat', 'IVF<nlist>,Flat', 'IVF<nlist>,SQ8', 'IVF<nlist>,PQ<m>', 'PQ<m>')")
        self.index_key = index_key
        params = dict(kwargs)
        params.setdefault("index_key", index_key)
        super().__init__(name, dimension, metric, **params)

    def build(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> IndexArtifact:
        from ..indexes import index_factory
        if vectors.shape[1] != self.dimension:
            raise ValueError(f"Expected dimension {self.dimension}, got {vectors.shape[1]}")
        meta: Dict[str, Any] = {"metric": self.metric, "index_key": self.index_key, "faiss_metric": "l2"}
        kind, normalize = "l2", False
        if self.metric == "cosine":
            kind, normalize = "ip", True
            meta.update({"faiss_metric": "ip", "normalize_queries": True, "normalize_vectors": True})
        elif self.metric == "ip":
            kind = "ip"
            meta["faiss_metric"] = "ip"
        train_kwargs = {key: self.params[key] for key in self._TRAIN_PARAM_KEYS
                        if key in self.params and self.index_key.strip().startswith("IVF")}
        index = index_factory(self.dimension, self.index_key, kind, device=self.params.get("device"), normalize=normalize,
                              **train_kwargs)
        meta.update(train_kwargs)
        if not index.is_trained:
            index.train(vectors)
        index.add(vectors)
        for key in self._RUNTIME_PARAM_KEYS:            # runtime knobs such as nprobe (modular.py:269-275)
            if key in self.params and hasattr(index, key):
                setattr(index, key, self.params[key])
                meta[key] = self.params[key]
        return IndexArtifact(kind="faiss", data=index, metadata=meta)


register_indexer("FaissFactoryIndexer", FaissFactoryIndexer)


class FaissIVFIndexer(FaissFactoryIndexer):
    """``index_type`` spelling of the same thing (modular.py:292-309)."""

    def __init__(self, name: str, dimension: int, metric: str = "l2", index_type: str = "IVF100,Flat", **kwargs: Any) -> None:
        params = dict(kwargs)
        params.setdefault("index_type", index_type)
        super().__init__(name, dimension, metric, index_key=index_type, **params)
        self.index_type = index_type


register_indexer("FaissIVFIndexer", FaissIVFIndexer)


class FaissLSHIndexer(BaseIndexer):
    """Random-hyperplane sign codes + Hamming search (modular.py:182-218) on the device."""

    SUPPORTED_METRICS = {"l2", "cosine", "ip"}

    def __init__(self, name: str, dimension: int, metric: str = "l2", num_bits: int = 256, **kwargs: Any) -> None:
        if metric not in self.SUPPORTED_METRICS:
            raise ValueError(f"FaissLSHIndexer supports metrics {self.SUPPORTED_METRICS}, received '{metric}'")
        if num_bits <= 0:
            raise ValueError("num_bits must be positive")
        super().__init__(name, dimension, metric, num_bits=num_bits, **kwargs)
        self.num_bits = num_bits

    def build(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> IndexArtifact:
        from ..indexes import GpuIndexLSH
        if vectors.shape[1] != self.dimension:
            raise ValueError(f"Expected dimension {self.dimension}, got {vectors.shape[1]}")
        meta: Dict[str, Any] = {"metric": self.metric, "num_bits": self.num_bits, "faiss_index_kind": "lsh"}
        if self.metric == "cosine":
            meta["normalize_queries"] = True      # sign codes are scale invariant: no need to normalise the rows
        elif self.metric == "ip":
            meta["faiss_metric"] = "ip"
        index = GpuIndexLSH(self.dimension, self.num_bits, device=self.params.get("device"),
                            seed=int(self.params.get("seed", 1234)))
        index.add(vectors)
        return IndexArtifact(kind="faiss", data=index, metadata=meta)


register_indexer("FaissLSHIndexer", FaissLSHIndexer)


# ------------------------------------------------------------------------------------------ searchers
class LinearSearcher(BaseSearcher):
    """Exact scan over raw vectors (modular.py:312-387): Euclidean (sqrt) distances for l2,
    negated scores for ip / cosine, (+inf, -1) padding when k exceeds the number of rows."""

    def attach(self, artifact: IndexArtifact, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        from .. import engine
        if artifact.kind != "raw_vectors":
            raise ValueError("LinearSearcher requires 'raw_vectors' artifact")
        if self.metric not in {"l2", "ip", "cosine"}:
            raise ValueError(f"Unsupported metric '{self.metric}' for LinearSearcher")
        self._vectors = artifact.data
        if self._vectors.shape[1] != self.dimension:
            raise ValueError("Vector dimension mismatch in LinearSearcher")
        if self._vectors.shape[0] == 0:
            raise RuntimeError("LinearSearcher cannot operate on empty index")
        self._shard = engine.FlatShard(self._vectors, self.metric, self.params.get("device"))
        self._flags = engine._lib.OUT_SQRT if self.metric == "l2" else engine._lib.OUT_NEGATE
        self._prepared = True

    def memory_bytes(self) -> int:
        return self._shard.memory_bytes() if self._prepared else 0

    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        distances, indices = self.batch_search(_as_batch(query), k)
        return distances[0], indices[0]

    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        from .. import engine
        import torch
        if not self._prepared:
            raise RuntimeError("LinearSearcher not attached to an index")
        dev = self._shard.dev
        with torch.cuda.device(dev):
            q = engine.queries_to_device(_as_batch(queries), dev, self.dimension)
            return engine.results_to_host(*self._shard.search(q, int(k), self._flags, float("inf")))


register_searcher("LinearSearcher", LinearSearcher)


class FaissSearcher(BaseSearcher):
    """Searcher over a ``kind == 'faiss'`` artifact (modular.py:393-548).

    Plain indexes: ``index.search`` and, for ip / cosine, negated scores (modular.py:544-546).
    LSH artifacts (``faiss_index_kind == 'lsh'``) with ``lsh_rerank`` (default): ask the index for
    ``candidate_k = min(max(k, int(k * lsh_candidate_multiplier)), lsh_max_candidates?, ntotal)``
    candidates (modular.py:463-468) and re-score them exactly on the device - one kernel for the
    whole batch instead of the reference's per-query loop (modular.py:483-532).  ``artifact.data``
    may be any object with ``search(queries, k) -> (D, I)`` (the reference's own test injects a
    fake, tests/test_composite_algorithm.py:176-187)."""

    def __init__(self, name: str, dimension: int, metric: str = "l2", **kwargs: Any) -> None:
        super().__init__(name, dimension, metric, **kwargs)
        self.index: Any = None
        self.normalize_queries = False
        self.index_kind: Optional[str] = None
        self._reranker = None
        self._lsh_rerank = bool(self.params.get("lsh_rerank", True))
        self._lsh_candidate_multiplier = float(self.params.get("lsh_candidate_multiplier", 8.0))
        cap = self.params.get("lsh_max_candidates")
        self._lsh_max_candidates: Optional[int] = int(cap) if cap is not None else None

    def attach(self, artifact: IndexArtifact, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        from .. import engine
        if artifact.kind != "faiss":
            raise ValueError("FaissSearcher requires 'faiss' artifact")
        self.index = artifact.data
        meta = artifact.metadata or {}
        self.metric = meta.get("metric", self.metric)
        self.normalize_queries = bool(meta.get("normalize_queries", False))
        self.index_kind = meta.get("faiss_index_kind")
        self._native = hasattr(self.index, "search_device")      # one of ours: queries are normalised on the device
        if self.index_kind == "lsh" and self._lsh_rerank:
            if self.metric not in {"l2", "ip", "cosine"}:
                raise ValueError(f"Unsupported metric '{self.metric}' for FaissSearcher LSH rerank")
            self._reranker = engine.Reranker(vectors, self.metric, self.params.get("device"))
        nprobe = self.params.get("nprobe")
        if nprobe is None:
            nprobe = meta.get("nprobe")
        if nprobe is not None and hasattr(self.index, "nprobe"):
            self.index.nprobe = int(nprobe)                       # the searcher's value wins (modular.py:437-441)
        self._prepared = True

    def memory_bytes(self) -> int:
        total = self.index.memory_bytes() if hasattr(self.index, "memory_bytes") else 0
        return total + (self._reranker.memory_bytes() if self._reranker is not None else 0)

    def _host_queries(self, queries: np.ndarray) -> np.ndarray:
        """Query preparation for a foreign index object, as modular.py:443-449."""
        q = _as_batch(queries).astype(np.float32, copy=True)
        if self.normalize_queries:
            norms = np.linalg.norm(q, axis=1, keepdims=True)
            q = np.divide(q, norms, out=np.zeros_like(q), where=norms > 0)
        return q

    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        distances, indices = self.batch_search(_as_batch(query), k)
        return distances[0], indices[0]

    def _candidate_budget(self, k: int) -> int:
        num_db = int(getattr(self.index, "ntotal", 0) or self._reranker.n)
        if num_db <= 0:
            raise RuntimeError("LSH index has no vectors to search")
        c = max(k, 1)
        if self._lsh_candidate_multiplier > 1.0:
            c = int(max(c, k * self._lsh_candidate_multiplier))
        if self._lsh_max_candidates is not None:
            c = min(c, self._lsh_max_candidates)
        return min(c, num_db)

    def _batch_search_lsh_rerank(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        from .. import engine
        import torch
        rr = self._reranker
        candidate_k = self._candidate_budget(k)
        flags = engine._lib.OUT_SQRT if self.metric == "l2" else engine._lib.OUT_NEGATE
        with torch.cuda.device(rr.dev):
            if candidate_k <= 0:                                   # degenerate cap: raw index order (modular.py:470-475)
                d, i = self.index.search(queries if self._native else self._host_queries(queries), k)
                d = np.asarray(d, dtype=np.float32)
                return (-d if self.metric in {"cosine", "ip"} else d), np.asarray(i, dtype=np.int64)
            q = engine.queries_to_device(_as_batch(queries), rr.dev, self.dimension)
            if self._native:
                qc = q.clone()
                if self.normalize_queries:
                    engine.normalize_rows_(qc)
                _, cand = self.index.search_device(qc, candidate_k)
                empty_rows = np.empty(0, dtype=np.int64)           # our Hamming top-k always fills min(candidate_k, ntotal) slots
            else:
                hq = self._host_queries(queries)
                _, cand_host = self.index.search(hq, candidate_k)
                cand_host = np.ascontiguousarray(cand_host, dtype=np.int64)
                empty_rows = np.nonzero((cand_host >= 0).sum(axis=1) == 0)[0]
                cand = torch.from_numpy(cand_host).to(rr.dev)
            dist, idx = engine.results_to_host(*rr.search(q, cand, int(k), flags, float("inf")))
            if empty_rows.size:
                # a query without a single valid candidate takes the raw index order (modular.py:486-493)
                raw_d, raw_i = self.index.search(hq[empty_rows], k)
                raw_d = np.asarray(raw_d, dtype=np.float32)
                dist[empty_rows] = -raw_d if self.metric in {"cosine", "ip"} else raw_d
                idx[empty_rows] = np.asarray(raw_i, dtype=np.int64)
            return dist, idx

    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self._prepared:
            raise RuntimeError("FaissSearcher not attached to an index")
        if self.index_kind == "lsh" and self._reranker is not None:
            return self._batch_search_lsh_rerank(queries, int(k))
        d, i = self.index.search(_as_batch(queries) if self._native else self._host_queries(queries), int(k))
        d = np.asarray(d, dtype=np.float32)
        if self.metric in {"cosine", "ip"}:
            d = -d
        return d, np.asarray(i, dtype=np.int64)


register_searcher("FaissSearcher", FaissSearcher)


# ------------------------------------------------------------------------------------------ composite
class CompositeAlgorithm(BaseAlgorithm):
    """Indexer + searcher as one ``BaseAlgorithm`` (modular.py:554-622)."""

    def __init__(self, name: str, dimension: int, indexer: Dict[str, Any], searcher: Dict[str, Any],
                 metric: str = "l2", **kwargs: Any) -> None:
        super().__init__(name, dimension)
        self.metric = metric
        self.extra_params = kwargs
        self.index_artifact: Optional[IndexArtifact] = None
        if not indexer or not searcher:
            raise ValueError("Both indexer_config and searcher_config must be provided for CompositeAlgorithm")
        self.indexer_config = copy.deepcopy(indexer)
        self.searcher_config = copy.deepcopy(searcher)
        self.indexer = self._instantiate(self.indexer_config, "Indexer", get_indexer_class)
        self.searcher = self._instantiate(self.searcher_config, "Searcher", get_searcher_class)
        self.config = {"metric": self.metric, "indexer": self.indexer.describe(), "searcher": self.searcher.describe()}
        if self.extra_params:
            self.config["params"] = copy.deepcopy(self.extra_params)

    def _instantiate(self, cfg: Dict[str, Any], what: str, lookup):
        cfg = copy.deepcopy(cfg)
        type_name = cfg.pop("type", None)
        if type_name is None:
            raise ValueError(f"{what} configuration must include a 'type' field")
        name = cfg.pop("name", type_name)
        metric = cfg.pop("metric", self.metric)
        return lookup(type_name)(name=name, dimension=self.dimension, metric=metric, **cfg)

    def build_index(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        self.index_artifact = self.indexer.build(vectors, metadata)
        self.searcher.attach(self.index_artifact, vectors, metadata)
        self.index_built = True

    def get_memory_usage(self) -> int:
        fn = getattr(self.searcher, "memory_bytes", None)
        return int(fn()) if fn is not None else 0

    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self.index_built:
            raise RuntimeError("Index has not been built for this algorithm")
        return self.searcher.search(query, k)

    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self.index_built:
            raise RuntimeError("Index has not been built for this algorithm")
        return self.searcher.batch_search(queries, k)


__all__ = ["BaseIndexer", "BaseSearcher", "CompositeAlgorithm", "IndexArtifact", "register_indexer", "register_searcher",
           "INDEXER_REGISTRY", "SEARCHER_REGISTRY", "get_indexer_class", "get_searcher_class", "BruteForceIndexer",
           "FaissFactoryIndexer", "FaissIVFIndexer", "FaissLSHIndexer", "LinearSearcher", "FaissSearcher"]
