"""The algorithm contract of the reference, restated (src/algorithms/base_algorithm.py:5-123).

Every searcher the benchmark harness drives is a ``BaseAlgorithm``: constructed as
``cls(name=..., dimension=..., **yaml_kwargs)``, then ``build_index(vectors)`` once and
``batch_search(queries, k)`` / ``search(query, k)`` many times, each returning
``(distances, indices)``.  The attribute names below are the ones the reference's
``ExperimentRunner`` reads (duck-typed memory estimate, parameter dump, op counters)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class BaseAlgorithm(ABC):
    def __init__(self, name: str, dimension: int, **kwargs: Any) -> None:
        self.name, self.dimension = name, dimension
        self.config: Dict[str, Any] = kwargs           # dumped into the result JSON by the harness: keep it serialisable
        self.vectors = self.metadata = None
        self.index_built = False
        self.build_time = self.index_memory_usage = -1.0
        self.operation_counter: Dict[str, Any] = {}

    # ---- the three calls every algorithm implements ------------------------------------------------
    @abstractmethod
    def build_index(self, vectors: np.ndarray, metadata: Optional[List[Dict[str, Any]]] = None) -> None:
        """Index ``vectors`` [n, dimension]."""

    @abstractmethod
    def search(self, query: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        """One query [dimension] -> (distances [k], indices [k])."""

    @abstractmethod
    def batch_search(self, queries: np.ndarray, k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        """Queries [nq, dimension] -> (distances [nq, k] float32, indices [nq, k] int64)."""

    # ---- optional persistence protocol (overridden where supported) --------------------------------
    def _no_persistence(self) -> NotImplementedError:
        return NotImplementedError(f"{self.__class__.__name__} does not support index persistence")

    def save_index(self, artifact_dir: str, context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        raise self._no_persistence()

    def load_index(self, artifact_dir: str, context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        raise self._no_persistence()

    # ---- what the harness reads back -----------------------------------------------------------------
    def get_name(self) -> str:
        return self.name

    def get_parameters(self) -> Dict[str, Any]:
        return self.config

    def get_operations(self) -> Dict[str, Any]:
        return dict(self.operation_counter)

    def record_operation(self, key: str, value: float) -> None:
        self.operation_counter[key] = float(self.operation_counter.get(key, 0.0)) + float(value)

    def __str__(self) -> str:
        return f"{self.name} (dimension={self.dimension}, parameters={self.config})"
