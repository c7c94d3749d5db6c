"""vectordb_retrieval_b200 - B200-native scan + top-k path behind the reference's
BaseAlgorithm / indexer / searcher plug-in API (Human-Augment-Analytics/vectordb-retrieval).

Layout: ``csrc/`` CUDA kernels + C ABI (include/vdb_cuda.h), ``_lib`` ctypes binding,
``engine`` device operators, ``algorithms`` the reference-facing classes, ``harness`` the
minimal experiment/benchmark runner that drives them from the reference's YAML schema."""
__version__ = "0.1.0"
