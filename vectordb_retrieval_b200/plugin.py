"""Drop the CUDA classes into an existing checkout of the reference
(Human-Augment-Analytics/vectordb-retrieval) without editing its files.

The reference resolves every YAML ``type`` string through three dictionaries -
``ALGORITHM_REGISTRY`` (src/algorithms/__init__.py:25-34) and ``INDEXER_REGISTRY`` /
``SEARCHER_REGISTRY`` (src/algorithms/modular.py:85-94) - and its ``ExperimentRunner`` only checks
``isinstance(algo, BaseAlgorithm)`` (src/experiments/experiment_runner.py:54).  ``install``
therefore (1) overwrites the entries for the scan + top-k classes with the CUDA-backed ones and
(2) registers them as virtual subclasses of the reference's abstract bases, so
``scripts/run_full_benchmark.py`` of the reference runs unchanged on the B200 path.

    from vectordb_retrieval_b200 import plugin
    plugin.install()            # after ``import src.algorithms`` of the reference is importable

The reference hard-imports ``faiss`` and ``matplotlib`` at package import
(src/algorithms/__init__.py:5, src/benchmark/evaluation.py:6); where those are absent,
``import_reference`` puts inert stand-ins on ``sys.modules`` first (constants only - no
arithmetic is ever routed to them: every FAISS-backed class is replaced)."""
from __future__ import annotations

import importlib
import sys
import types
from typing import Any, Dict, Optional


def _stub_missing_modules() -> None:
    try:
        importlib.import_module("faiss")
    except ImportError:
        faiss = types.ModuleType("faiss")
        faiss.METRIC_L2, faiss.METRIC_INNER_PRODUCT = 1, 0
        sys.modules["faiss"] = faiss
    try:
        importlib.import_module("matplotlib.pyplot")
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


def import_reference(root: Optional[str] = None) -> Dict[str, Any]:
    """Import the reference's ``src.algorithms`` package (``root`` = checkout directory)."""
    if root is not None and root not in sys.path:
        sys.path.insert(0, root)
    _stub_missing_modules()
    algorithms = importlib.import_module("src.algorithms")
    modular = importlib.import_module("src.algorithms.modular")
    return {"algorithms": algorithms, "modular": modular}


def install(modules: Optional[Dict[str, Any]] = None) -> None:
    from . import algorithms as ours
    mods = modules or import_reference()
    ref_algorithms, ref_modular = mods["algorithms"], mods["modular"]
    for name in ("ExactSearch", "ApproximateSearch", "LSH", "Composite", "CompositeAlgorithm", "Modular"):
        cls = ours.ALGORITHM_REGISTRY[name]
        ref_algorithms.ALGORITHM_REGISTRY[name] = cls
        ref_algorithms.BaseAlgorithm.register(cls)
    for name in ("BruteForceIndexer", "FaissFactoryIndexer", "FaissIVFIndexer", "FaissLSHIndexer", "LSHIndexer"):
        cls = ours.INDEXER_REGISTRY[name]
        ref_modular.register_indexer(name, cls)
        ref_modular.BaseIndexer.register(cls)
    for name in ("LinearSearcher", "FaissSearcher", "LSHSearcher"):
        cls = ours.SEARCHER_REGISTRY[name]
        ref_modular.register_searcher(name, cls)
        ref_modular.BaseSearcher.register(cls)
