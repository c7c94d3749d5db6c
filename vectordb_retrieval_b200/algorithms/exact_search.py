"""``ExactSearch`` on the B200 scan kernel (reference: src/algorithms/exact_search.py:6-78).

Same constructor, same value conventions as the FAISS-backed original:
* only the literal ``'l2'`` selects L2; every other metric string - including ``'cosine'`` -
  means inner product on the raw, un-normalised vectors (exact_search.py:23);
* L2 returns *squared* distances ascending, inner product returns raw scores descending;
* missing results are id -1 with distance +FLT_MAX (L2) / -FLT_MAX (IP);
* not built -> RuntimeError (exact_search.py:53-54).
``faiss.IndexFlat`` becomes :class:`GpuIndexFlat` (tcgen05 3xTF32 contraction fused with the
per-query top-k bound, exact re-scoring of the winners).  Extra keyword arguments: ``device``
(one GPU) or ``devices`` (list: row-sharded single-process search); under ``torchrun`` the work is
split over the ranks: ``shard='rows'`` (row shards + top-k allgather + merge kernel), ``'queries'``
(replicated base, each rank searches a slice of the batch) or ``'auto'`` (queries while the base is small)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .base_algorithm import BaseAlgorithm


class ExactSearch(BaseAlgorithm):
    def __init__(self, name: str, dimension: int, metric: str = "l2", **kwargs: Any) -> None:
        super().__init__(name, dimension, **kwargs)
        self.metric = "l2" if metric == "l2" else "ip"
        self.index = None

    def build_index(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        from ..indexes import GpuIndexFlat
        if vectors.ndim != 2 or vectors.shape[1] != self.dimension:
            raise RuntimeError(f"expected vectors of shape [n, {self.dimension}], got {vectors.shape}")
        self.vectors = vectors          # the harness' array; never mutated (dtype/layout fixed on upload)
        self.index = GpuIndexFlat(self.dimension, self.metric, device=self.config.get("device"),
                                  devices=self.config.get("devices"), shard=self.config.get("shard", "auto"),
                                  exchange=self.config.get("exchange", "alltoall"))
        self.index.add(vectors)
        self.index_built = True

    def get_memory_usage(self) -> int:
        """Bytes held in HBM (honoured by the harness' memory estimate, experiment_runner.py:493)."""
        return 0 if self.index is None else self.index.memory_bytes()

    def save_index(self, artifact_dir: str, context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """Persist the device operands (base_algorithm.py:98-109 protocol; see persist.py)."""
        if not self.index_built:
            raise RuntimeError("Index has not been built yet.")
        return self.index.save(artifact_dir, context)

    def load_index(self, artifact_dir: str, context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        from ..indexes import GpuIndexFlat
        self.index = GpuIndexFlat(self.dimension, self.metric, device=self.config.get("device"))
        manifest = self.index.load(artifact_dir, context)
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
        distances, indices = self.index.search(queries, k)
        self.record_operation("ndis", float(indices.shape[0]) * float(self.index.ntotal))
        return distances, indices
