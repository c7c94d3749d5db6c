#!/usr/bin/env python
"""The reference's ``scripts/run_full_benchmark.py``, unchanged, on the B200 kernels.

    python scripts/stage_reference.py                      # once, in the build container
    python scripts/run_reference_benchmark.py --config configs/reference_random20k.yaml --output-dir out/

Everything after the optional ``--reference DIR`` goes to the reference's own argument parser
(``--config``, ``--output-dir``, scripts/run_full_benchmark.py:285-288).  The only thing done before
its ``main()`` runs is ``vectordb_retrieval_b200.plugin.install()``: the YAML ``type`` strings
``ExactSearch`` / ``BruteForceIndexer`` / ``LinearSearcher`` / ``Faiss*Indexer`` / ``FaissSearcher`` /
``LSHIndexer`` / ``LSHSearcher`` / ``Composite`` then resolve to the CUDA-backed classes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main() -> int:
    argv = sys.argv[1:]
    reference = os.path.join(ROOT, "baseline", "_ref")
    if argv[:1] == ["--reference"]:
        reference, argv = argv[1], argv[2:]
    if not os.path.isdir(os.path.join(reference, "src", "algorithms")):
        print(f"no reference checkout at {reference}: run scripts/stage_reference.py (build container) first")
        return 2
    from vectordb_retrieval_b200 import plugin
    return plugin.run_reference_cli(reference, argv)


if __name__ == "__main__":
    sys.exit(main())
