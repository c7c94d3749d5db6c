"""``ApproximateSearch`` on the CUDA scan paths (reference: src/algorithms/approximate_search.py:6-87).

The reference hands ``index_type`` to ``faiss.index_factory``; this build accepts the part of
that grammar which is a scan + top-k - ``"Flat"``, ``"IVF<nlist>,Flat"``, ``"IVF<nlist>,SQ8"``,
``"IVF<nlist>,PQ<m>"`` and ``"PQ<m>"`` - and rejects the rest with ValueError at construction (HNSW, OPQ and other
code widths are out of scope).  Values follow
raw FAISS: squared L2 ascending / inner product descending; only ``'l2'`` selects L2
(approximate_search.py:25); ``nprobe`` comes from the constructor kwargs (approximate_search.py:50-51)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .base_algorithm import BaseAlgorithm


class ApproximateSearch(BaseAlgorithm):
    def __init__(self, name: str, dimension: int, index_type: str, metric: str = "l2", **kwargs: Any) -> None:
        super().__init__(name, dimension, **kwargs)
        from ..indexes import _IVF_FLAT, _IVF_PQ, _IVF_SQ8, _PQ
        key = index_type.strip()
        if key != "Flat" and not any(rx.match(key) for rx in (_IVF_FLAT, _IVF_SQ8, _IVF_PQ, _PQ)):
            raise ValueError(f"index_type '{index_type}' is not supported by the CUDA build "
                             "(supported: 'Flat', 'IVF<nlist>,Flat', 'IVF<nlist>,SQ8', 'IVF<nlist>,PQ<m>', 'PQ<m>')")
        self.index_type = index_type
        self.metric = "l2" if metric == "l2" else "ip"
        self.index = None

    def build_index(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        from ..indexes import index_factory
        if vectors.ndim != 2 or vectors.shape[1] != self.dimension:
            raise RuntimeError(f"expected vectors of shape [n, {self.dimension}], got {vectors.shape}")
        self.vectors = vectors
        extra = {"shard": self.config["shard"]} if "shard" in self.config else {}     # torchrun: 'rows' / 'queries' / 'auto'
        self.index = index_factory(self.dimension, self.index_type, self.metric, device=self.config.get("device"), **extra)
        if "nprobe" in self.config:
            self.index.nprobe = int(self.config["nprobe"])
        if not self.index.is_trained:
            self.index.train(vectors)
        self.index.add(vectors)
        self.index_built = True
        if "nprobe" in self.config:
            self.index.nprobe = int(self.config["nprobe"])

    def get_memory_usage(self) -> int:
        return 0 if self.index is None else self.index.memory_bytes()

    def save_index(self, artifact_dir: str, context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        if not self.index_built:
            raise RuntimeError("Index has not been built yet.")
        return self.index.save(artifact_dir, context)

    def load_index(self, artifact_dir: str, context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        from ..indexes import index_factory
        self.index = index_factory(self.dimension, self.index_type, self.metric, device=self.config.get("device"))
        manifest = self.index.load(artifact_dir, context)
        if "nprobe" in self.config:
            self.index.nprobe = int(self.config["nprobe"])
        elif "nprobe" in manifest["meta"]:
            self.index.nprobe = int(manifest["meta"]["nprobe"])
        self.index_built = True
        return {"build_time_s": float(manifest.get("build_metrics", {}).get("build_time_s", 0.0) or 0.0), "manifest": manifest}

    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self.index_built:
            raise RuntimeError("Index has not been built yet.")
        distances, indices = self.index.search(np.asarray(query).reshape(1, -1), k)
        return distances[0], indices[0]

    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if not self.index_built:
            raise RuntimeError("Index has not been built yet.")
        return self.index.search(queries, k)
