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


class _Inert:
    """Stand-in object for the plotting calls of the reference (src/benchmark/evaluation.py:162-276,
    src/experiments/experiment_runner.py:764-780): every attribute is a callable that returns another
    inert object, so ``fig, ax = plt.subplots(); ax.scatter(...); plt.savefig(path)`` runs and draws
    nothing.  Plots are cosmetic; the result JSON / Markdown are written before them."""

    def __call__(self, *args: Any, **kwargs: Any) -> "_Inert":
        return self

    def __getattr__(self, name: str) -> "_Inert":
        if name.startswith("__"):
            raise AttributeError(name)
        return self

    def __iter__(self):
        return iter(())


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
        inert = _Inert()
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        plt.subplots = lambda *a, **k: (inert, inert)

        def _plt_attr(name: str):                      # module-level __getattr__ (PEP 562): figure, savefig, close, ...
            if name.startswith("__"):                  # inspect / importlib probe __file__, __path__, __spec__: absent
                raise AttributeError(name)
            return inert

        plt.__getattr__ = _plt_attr
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


def run_reference_cli(root: str, argv: Optional[list] = None) -> int:
    """Run the reference's own ``scripts/run_full_benchmark.py`` (file untouched, its ``main()`` called as
    ``python scripts/run_full_benchmark.py <argv>`` would) with the CUDA classes installed first.
    ``root`` = the reference checkout (``baseline/_ref`` after ``scripts/stage_reference.py``)."""
    import importlib.util
    import os
    install(import_reference(root))
    path = os.path.join(root, "scripts", "run_full_benchmark.py")
    spec = importlib.util.spec_from_file_location("_reference_run_full_benchmark", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)                    # imports src.benchmark.runner of the reference
    old_argv = sys.argv
    sys.argv = [path] + list(argv or [])
    try:
        return int(module.main() or 0)
    finally:
        sys.argv = old_argv
