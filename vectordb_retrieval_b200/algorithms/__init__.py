"""Reference-facing algorithm classes (src/algorithms/__init__.py): the ``type`` strings of the
benchmark YAML resolve through ``ALGORITHM_REGISTRY`` / ``get_algorithm_instance`` exactly as in
the reference; the classes behind them run on the CUDA kernels.  ``HNSW`` and ``CoverTreeV2_2``
are outside this build (not scan + top-k) and are simply not registered: asking for them raises
the same ValueError an unknown type raises in the reference (__init__.py:40-43)."""
from typing import Any, Dict, Type

from .approximate_search import ApproximateSearch
from .base_algorithm import BaseAlgorithm
from .exact_search import ExactSearch
from .lsh import LSH, LSHIndexer, LSHSearcher
from .modular import (INDEXER_REGISTRY, SEARCHER_REGISTRY, BaseIndexer, BaseSearcher, BruteForceIndexer, CompositeAlgorithm,
                      FaissFactoryIndexer, FaissIVFIndexer, FaissLSHIndexer, FaissSearcher, IndexArtifact, LinearSearcher,
                      get_indexer_class, get_searcher_class, register_indexer, register_searcher)

# YAML ``type`` -> class (src/algorithms/__init__.py:25-34); the three composite spellings are the reference's aliases
ALGORITHM_REGISTRY: Dict[str, Type[BaseAlgorithm]] = {
    "ExactSearch": ExactSearch,
    "ApproximateSearch": ApproximateSearch,
    "LSH": LSH,
    "Composite": CompositeAlgorithm,
    "CompositeAlgorithm": CompositeAlgorithm,
    "Modular": CompositeAlgorithm,
}


def get_algorithm_instance(algorithm_type: str, dimension: int, **params: Any) -> BaseAlgorithm:
    """``cls(name=..., dimension=..., **params)`` for a registered type (__init__.py:37-47)."""
    cls = ALGORITHM_REGISTRY.get(algorithm_type)
    if cls is None:
        raise ValueError(f"Unknown algorithm type: {algorithm_type}. Available types: {list(ALGORITHM_REGISTRY.keys())}")
    return cls(name=params.pop("name", algorithm_type), dimension=dimension, **params)


__all__ = ["ALGORITHM_REGISTRY", "get_algorithm_instance", "BaseAlgorithm", "ExactSearch", "ApproximateSearch", "LSH",
           "LSHIndexer", "LSHSearcher", "BaseIndexer", "BaseSearcher", "IndexArtifact", "CompositeAlgorithm", "BruteForceIndexer",
           "LinearSearcher", "FaissFactoryIndexer", "FaissIVFIndexer", "FaissLSHIndexer", "FaissSearcher", "INDEXER_REGISTRY",
           "SEARCHER_REGISTRY", "register_indexer", "register_searcher", "get_indexer_class", "get_searcher_class"]
