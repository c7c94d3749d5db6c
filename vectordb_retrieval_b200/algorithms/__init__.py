"""Reference-facing algorithm classes (src/algorithms/__init__.py): the ``type`` strings of the
benchmark YAML resolve through ``ALGORITHM_REGISTRY`` / ``get_algorithm_instance`` exactly as in
the reference; the classes behind them run on the CUDA kernels.  ``HNSW`` and ``CoverTreeV2_2``
are outside this build (not scan + top-k) and are simply not registered: asking for them raises
the same ValueError an unknown type raises in the reference (__init__.py:40-43)."""
from typing import Any, Dict, Type

from . import approximate_search as _approx
from . import base_algorithm as _base
from . import exact_search as _exact
from . import lsh as _lsh
from . import modular as _modular

# public names, per module that defines them (re-exported below; ``__all__`` is derived from this table)
_PUBLIC = {
    _base: ("BaseAlgorithm",),
    _exact: ("ExactSearch",),
    _approx: ("ApproximateSearch",),
    _lsh: ("LSH", "LSHIndexer", "LSHSearcher"),
    _modular: ("BaseIndexer", "BaseSearcher", "IndexArtifact", "CompositeAlgorithm", "BruteForceIndexer", "LinearSearcher",
               "FaissFactoryIndexer", "FaissIVFIndexer", "FaissLSHIndexer", "FaissSearcher", "INDEXER_REGISTRY",
               "SEARCHER_REGISTRY", "register_indexer", "register_searcher", "get_indexer_class", "get_searcher_class"),
}
for _module, _names in _PUBLIC.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)

BaseAlgorithm = _base.BaseAlgorithm          # (explicit for type checkers)
CompositeAlgorithm = _modular.CompositeAlgorithm

# YAML ``type`` -> class; the three composite spellings are the reference's aliases
ALGORITHM_REGISTRY: Dict[str, Type[BaseAlgorithm]] = {cls.__name__: cls for cls in (_exact.ExactSearch, _approx.ApproximateSearch, _lsh.LSH)}
ALGORITHM_REGISTRY.update({alias: CompositeAlgorithm for alias in ("Composite", "CompositeAlgorithm", "Modular")})


def get_algorithm_instance(algorithm_type: str, dimension: int, **params: Any) -> BaseAlgorithm:
    """``cls(name=..., dimension=..., **params)`` for a registered type (__init__.py:37-47)."""
    cls = ALGORITHM_REGISTRY.get(algorithm_type)
    if cls is None:
        raise ValueError(f"Unknown algorithm type: {algorithm_type}. Available types: {list(ALGORITHM_REGISTRY.keys())}")
    return cls(name=params.pop("name", algorithm_type), dimension=dimension, **params)


__all__ = sorted({n for names in _PUBLIC.values() for n in names} | {"ALGORITHM_REGISTRY", "get_algorithm_instance"})
