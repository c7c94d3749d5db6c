#!/usr/bin/env python
"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference classes.

TEST INFRASTRUCTURE - run in the build container only (``/root/reference`` does not
exist on the GPU box); the produced fixtures are committed.  The reference package
hard-imports ``faiss`` (src/algorithms/__init__.py:5 -> approximate_search.py:1) and
``matplotlib`` (src/benchmark/evaluation.py:6), neither installed here, so two
throw-away import stubs are put on ``sys.path`` first.  Only NumPy code paths of the
reference are executed: BruteForceIndexer + LinearSearcher, FaissSearcher's LSH-rerank
with an injected candidate generator (the reference's own test does the same,
tests/test_composite_algorithm.py:169-226), LSHIndexer + LSHSearcher, recall_at_k.

Inputs are regenerated from seeds by the tests; only reference OUTPUTS are stored.
    python oracle/gen_golden.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _install_stubs(tmp: str) -> None:
    os.makedirs(os.path.join(tmp, "faiss"))
    with open(os.path.join(tmp, "faiss", "__init__.py"), "w") as f:
        f.write("METRIC_L2 = 1\nMETRIC_INNER_PRODUCT = 0\n")
    os.makedirs(os.path.join(tmp, "matplotlib"))
    with open(os.path.join(tmp, "matplotlib", "__init__.py"), "w") as f:
        f.write("def use(*a, **k):\n    pass\n")
    with open(os.path.join(tmp, "matplotlib", "pyplot.py"), "w") as f:
        f.write("def subplots(*a, **k):\n    raise RuntimeError('stub')\n")
    sys.path.insert(0, tmp)


# ---- shared input generators (tests/golden_inputs.py re-creates exactly these) ----
def linear_inputs():
    base = np.random.RandomState(101).randn(3000, 24).astype(np.float32)
    queries = np.random.RandomState(102).randn(40, 24).astype(np.float32)
    base[17] = 0.0  # a zero row exercises the cosine guard
    return base, queries


def rerank_inputs():
    base = np.random.RandomState(201).randn(500, 16).astype(np.float32)
    queries = np.random.RandomState(202).randn(12, 16).astype(np.float32)
    rng = np.random.RandomState(203)
    cand = np.stack([rng.permutation(500)[:60] for _ in range(12)]).astype(np.int64)
    cand[3, 40:] = -1          # short row
    cand[7, 5:] = -1           # fewer valid candidates than k
    return base, queries, cand


def random20k_inputs():
    """The reference's ``random`` dataset (src/benchmark/dataset.py:491-495) with the
    published options (seed 7, 20000x64, 512 test) and the runner's query subset
    (src/experiments/experiment_runner.py:79,148: seed 42, choice(512, 256))."""
    np.random.seed(7)
    train = np.random.randn(20000, 64).astype(np.float32)
    test = np.random.randn(512, 64).astype(np.float32)
    np.random.seed(42)
    pick = np.random.choice(512, 256, replace=False)
    return train, test[pick], pick


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="vdb_stubs_")
    _install_stubs(tmp)
    sys.path.insert(1, args.reference)
    from src.algorithms import get_algorithm_instance  # noqa: E402  (reference)
    from src.algorithms.modular import FaissSearcher, IndexArtifact  # noqa: E402
    import src.algorithms.modular as ref_modular  # noqa: E402
    from src.benchmark.metrics import recall_at_k  # noqa: E402

    os.makedirs(OUT, exist_ok=True)

    # 1. LinearSearcher, three metrics, plus k > N padding ---------------------------
    base, queries = linear_inputs()
    out = {}
    for metric in ("l2", "ip", "cosine"):
        algo = get_algorithm_instance(
            "Composite", base.shape[1], name=f"exact_{metric}", metric=metric,
            indexer={"type": "BruteForceIndexer", "metric": metric},
            searcher={"type": "LinearSearcher", "metric": metric})
        algo.build_index(base)
        d, i = algo.batch_search(queries, k=10)
        out[f"{metric}_D"], out[f"{metric}_I"] = d, i
        d1, i1 = algo.search(queries[0], k=10)
        assert np.array_equal(i1, i[0])
        algo.build_index(base[:6])
        d, i = algo.batch_search(queries[:3], k=8)
        out[f"{metric}_pad_D"], out[f"{metric}_pad_I"] = d, i
    np.savez_compressed(os.path.join(OUT, "linear_searcher.npz"), **out)

    # 2. FaissSearcher LSH-rerank with an injected candidate generator ---------------
    base, queries, cand = rerank_inputs()

    class FixedCandidates:
        def __init__(self, c):
            self.c, self.ntotal = c, base.shape[0]

        def search(self, q, k):
            assert q.shape[0] == self.c.shape[0]
            return np.zeros((q.shape[0], k), np.float32), self.c[:, :k]

    out = {}
    saved = ref_modular.faiss
    ref_modular.faiss = object()
    try:
        for metric in ("l2", "ip", "cosine"):
            s = FaissSearcher(name="rr", dimension=16, metric=metric, lsh_rerank=True,
                              lsh_candidate_multiplier=6.0)
            meta = {"metric": metric, "faiss_index_kind": "lsh"}
            if metric == "cosine":
                meta["normalize_queries"] = True
            s.attach(IndexArtifact(kind="faiss", data=FixedCandidates(cand), metadata=meta), base)
            d, i = s.batch_search(queries, k=10)   # candidate_k = 60
            out[f"{metric}_D"], out[f"{metric}_I"] = d, i
    finally:
        ref_modular.faiss = saved
    np.savez_compressed(os.path.join(OUT, "faiss_lsh_rerank.npz"), **out)

    # 3. Python LSH: the reference KATs and the published 20k golden recall ----------
    out = {}
    rng = np.random.RandomState(7)
    train = rng.randn(128, 16).astype(np.float32)
    train /= np.linalg.norm(train, axis=1, keepdims=True)
    algo = get_algorithm_instance(
        "Composite", 16, name="lsh_cos", metric="cosine",
        indexer={"type": "LSHIndexer", "metric": "cosine", "num_tables": 12, "hash_size": 16, "seed": 7},
        searcher={"type": "LSHSearcher", "metric": "cosine", "candidate_multiplier": 12.0,
                  "fallback_to_bruteforce": True})
    algo.build_index(train)
    out["kat_cos_D"], out["kat_cos_I"] = algo.batch_search(train[:5].copy(), k=4)
    rng = np.random.RandomState(11)
    train = rng.randn(160, 8).astype(np.float32)
    algo = get_algorithm_instance(
        "Composite", 8, name="lsh_l2", metric="l2",
        indexer={"type": "LSHIndexer", "metric": "l2", "num_tables": 10, "hash_size": 12,
                 "bucket_width": 3.0, "seed": 11},
        searcher={"type": "LSHSearcher", "metric": "l2", "candidate_multiplier": 10.0,
                  "fallback_to_bruteforce": True})
    algo.build_index(train)
    out["kat_l2_D"], out["kat_l2_I"] = algo.batch_search(train[10:20].copy(), k=4)

    train, q, pick = random20k_inputs()
    gt = np.stack([np.argsort(np.linalg.norm(train - qq[None, :], axis=1))[:100] for qq in q]).astype(np.int32)
    algo = get_algorithm_instance(
        "Composite", 64, name="lsh", metric="l2",
        indexer={"type": "LSHIndexer", "metric": "l2", "num_tables": 12, "hash_size": 4,
                 "bucket_width": 20.0, "seed": 42},
        searcher={"type": "LSHSearcher", "metric": "l2", "candidate_multiplier": 64.0,
                  "max_candidates": None, "fallback_to_bruteforce": False})
    algo.build_index(train)
    d, i = algo.batch_search(q, k=20)
    r10, r1 = recall_at_k(gt, i, 10), recall_at_k(gt, i, 1)
    print("python-LSH 20k recall@10", repr(r10), "recall@1", repr(r1))
    assert r10 == 0.31914062499999996 and r1 == 0.34765625, "published golden not reproduced"
    out["r20k_D"], out["r20k_I"] = d, i.astype(np.int32)
    out["r20k_recall10"], out["r20k_recall1"] = np.float64(r10), np.float64(r1)
    out["r20k_gt20"] = gt[:, :20]
    algo = get_algorithm_instance(
        "Composite", 64, name="exact", metric="l2",
        indexer={"type": "BruteForceIndexer", "metric": "l2"},
        searcher={"type": "LinearSearcher", "metric": "l2"})
    algo.build_index(train)
    d, i = algo.batch_search(q[:64], k=20)
    out["r20k_exact_D"], out["r20k_exact_I"] = d, i.astype(np.int32)
    assert recall_at_k(gt[:64], i, 10) == 1.0
    np.savez_compressed(os.path.join(OUT, "python_lsh.npz"), **out)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
