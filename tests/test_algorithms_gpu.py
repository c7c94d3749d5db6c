"""GPU tests of the reference-facing classes: the reference's own boundary tests
(tests/test_composite_algorithm.py, tests/test_benchmark_runner_modular.py) restated against the
CUDA-backed classes, plus parity with the oracle / golden fixtures through the public API."""
import json
import os

import numpy as np
import pytest

from oracle import oracle  # checker only
from oracle.gen_golden import linear_inputs, random20k_inputs, rerank_inputs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def A():
    import vectordb_retrieval_b200.algorithms as algorithms
    from vectordb_retrieval_b200 import _lib
    _lib.load()
    return algorithms


def _check(ref, got, rtol=1e-5, atol=0.0):
    res = oracle.compare_topk(ref[0], ref[1], got[0], got[1], rtol=rtol, atol=atol)
    assert res["ok"], res


def test_composite_bruteforce_linear_matches_numpy(A):
    """reference tests/test_composite_algorithm.py:29-58 (4-point L2 known answer)."""
    vectors = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], dtype=np.float32)
    queries = np.array([[0.1, 0.1], [0.9, 0.8]], dtype=np.float32)
    algo = A.CompositeAlgorithm(name="exact_composite", dimension=2, metric="l2",
                                indexer={"type": "BruteForceIndexer"}, searcher={"type": "LinearSearcher"})
    algo.build_index(vectors)
    dist, idx = algo.batch_search(queries, k=2)
    expected = np.argsort(np.linalg.norm(vectors[None, :, :] - queries[:, None, :], axis=2), axis=1)[:, :2]
    np.testing.assert_array_equal(idx, expected)
    assert dist.shape == (2, 2) and dist.dtype == np.float32 and idx.dtype == np.int64
    d1, i1 = algo.search(queries[0], k=2)
    np.testing.assert_array_equal(i1, expected[0])
    np.testing.assert_allclose(d1, dist[0])


def test_composite_requires_indexer_and_searcher(A):
    """reference tests/test_composite_algorithm.py:88-105."""
    with pytest.raises(ValueError):
        A.CompositeAlgorithm(name="x", dimension=4, indexer={}, searcher={"type": "LinearSearcher"})
    with pytest.raises(ValueError):
        A.CompositeAlgorithm(name="x", dimension=4, indexer={"type": "BruteForceIndexer"}, searcher={})
    with pytest.raises(ValueError):
        A.CompositeAlgorithm(name="x", dimension=4, indexer={"name": "no_type"}, searcher={"type": "LinearSearcher"})
    with pytest.raises(ValueError):
        A.get_algorithm_instance("NoSuchAlgorithm", 4)
    algo = A.get_algorithm_instance("ExactSearch", 4, name="e")
    with pytest.raises(RuntimeError):
        algo.batch_search(np.zeros((1, 4), dtype=np.float32), 1)


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_linear_searcher_matches_reference_golden(A, metric):
    g = np.load(os.path.join(GOLDEN, "linear_searcher.npz"))
    base, queries = linear_inputs()
    algo = A.get_algorithm_instance("Composite", base.shape[1], name="lin", metric=metric,
                                    indexer={"type": "BruteForceIndexer"}, searcher={"type": "LinearSearcher"})
    atol = 0.0 if metric == "l2" else 2e-6 * (1 if metric == "cosine" else 40)
    algo.build_index(base)
    dist, idx = algo.batch_search(queries, 10)
    _check((g[f"{metric}_D"], g[f"{metric}_I"]), (dist, idx), atol=atol)
    d1, i1 = algo.search(queries[0], 10)
    np.testing.assert_array_equal(i1, idx[0])
    algo.build_index(base[:6])                                     # k > n: (+inf, -1) padding
    dist, idx = algo.batch_search(queries[:3], 8)
    np.testing.assert_array_equal(idx, g[f"{metric}_pad_I"])
    _check((g[f"{metric}_pad_D"], g[f"{metric}_pad_I"]), (dist, idx), atol=atol)
    assert np.isinf(dist[:, 6:]).all() and (idx[:, 6:] == -1).all()


def test_exact_search_faiss_conventions(A):
    base, queries = linear_inputs()
    # float64, non-contiguous input must be accepted (memmap / slicing in the harness)
    wide = np.zeros((base.shape[0], base.shape[1] * 2), dtype=np.float64)
    wide[:, ::2] = base
    for metric, ometric in (("l2", "l2"), ("ip", "ip"), ("cosine", "ip")):       # 'cosine' silently means raw IP (exact_search.py:23)
        algo = A.get_algorithm_instance("ExactSearch", base.shape[1], name="exact", metric=metric)
        algo.build_index(wide[:, ::2])
        dist, idx = algo.batch_search(queries.astype(np.float64), 20)
        ref = oracle.faiss_flat_search(base, queries, 20, ometric)
        _check(ref, (dist, idx), atol=0.0 if ometric == "l2" else 1e-4)
        assert algo.get_memory_usage() > 0 and algo.get_operations()["ndis"] == queries.shape[0] * base.shape[0]
    # k > n pads with (-1, FLT_MAX)
    algo = A.get_algorithm_instance("ExactSearch", base.shape[1], name="exact", metric="l2")
    algo.build_index(base[:7])
    dist, idx = algo.batch_search(queries[:3], 10)
    assert (idx[:, 7:] == -1).all() and (dist[:, 7:] == np.finfo(np.float32).max).all()
    json.dumps(algo.get_parameters())


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_faiss_searcher_lsh_rerank_with_injected_candidates(A, metric):
    """reference tests/test_composite_algorithm.py:169-226: a fake index object supplies candidates
    (here the golden generator's candidate matrix), the searcher re-scores them exactly."""
    g = np.load(os.path.join(GOLDEN, "faiss_lsh_rerank.npz"))
    base, queries, cand = rerank_inputs()

    class FakeLSHIndex:
        ntotal = base.shape[0]

        def search(self, q, k):
            assert k == cand.shape[1]
            return np.zeros((q.shape[0], k), dtype=np.float32), cand[: q.shape[0]]

    searcher = A.FaissSearcher("s", base.shape[1], metric, lsh_rerank=True, lsh_candidate_multiplier=6.0)   # candidate_k = 60
    meta = {"metric": metric, "faiss_index_kind": "lsh"}
    if metric == "cosine":
        meta["normalize_queries"] = True
    searcher.attach(A.IndexArtifact(kind="faiss", data=FakeLSHIndex(), metadata=meta), base)
    dist, idx = searcher.batch_search(queries, 10)
    np.testing.assert_array_equal(idx, g[f"{metric}_I"])
    np.testing.assert_allclose(dist, g[f"{metric}_D"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_lsh_rerank_row_without_candidates_takes_the_raw_index_order(A, metric):
    """reference src/algorithms/modular.py:486-493: a query whose candidate row holds no valid id is answered by
    a second ``index.search(query, k)`` (values negated for ip / cosine); every other row is re-scored as usual."""
    rng = np.random.RandomState(4)
    base = rng.randn(200, 12).astype(np.float32)
    queries = rng.randn(6, 12).astype(np.float32)
    cand = np.stack([rng.permutation(200)[:40] for _ in range(6)]).astype(np.int64)
    cand[2] = -1
    raw_d = np.linspace(1.0, 2.0, 5, dtype=np.float32)[None, :]
    raw_i = np.arange(100, 105, dtype=np.int64)[None, :]
    calls = []

    class Fake:
        ntotal = 200

        def search(self, q, k):
            calls.append((q.shape[0], k))
            if k == 40:
                return np.zeros((q.shape[0], k), dtype=np.float32), cand[: q.shape[0]]
            np.testing.assert_allclose(q, queries[2:3], rtol=1e-6)
            return np.repeat(raw_d, q.shape[0], 0), np.repeat(raw_i, q.shape[0], 0)

    s = A.FaissSearcher("s", 12, metric, lsh_candidate_multiplier=8.0)     # candidate_k = 40
    s.attach(A.IndexArtifact(kind="faiss", data=Fake(), metadata={"metric": metric, "faiss_index_kind": "lsh"}), base)
    dist, idx = s.batch_search(queries, 5)
    assert calls == [(6, 40), (1, 5)]
    np.testing.assert_array_equal(idx[2], raw_i[0])
    np.testing.assert_allclose(dist[2], raw_d[0] if metric == "l2" else -raw_d[0])
    ref_d, ref_i = oracle.rerank_search(base, cand, queries, 5, metric)
    keep = [0, 1, 3, 4, 5]
    np.testing.assert_array_equal(idx[keep], ref_i[keep])
    np.testing.assert_allclose(dist[keep], ref_d[keep], rtol=1e-5, atol=2e-6)


def test_reversed_candidates_kat(A):
    """reference tests/test_composite_algorithm.py:169-226: candidates arrive worst-first."""
    rng = np.random.RandomState(3)
    base = rng.randn(50, 8).astype(np.float32)

    class Reversed:
        ntotal = 50

        def search(self, q, k):
            return np.zeros((q.shape[0], k), dtype=np.float32), np.tile(np.arange(k - 1, -1, -1), (q.shape[0], 1))

    s = A.FaissSearcher("s", 8, "l2", lsh_candidate_multiplier=4.0)
    s.attach(A.IndexArtifact(kind="faiss", data=Reversed(), metadata={"metric": "l2", "faiss_index_kind": "lsh"}), base)
    dist, idx = s.batch_search(base[:1], 3)
    assert idx[0, 0] == 0 and abs(dist[0, 0]) < 1e-6


def test_python_lsh_matches_reference_golden_and_published_recall(A):
    """Golden outputs of the unmodified reference LSHIndexer+LSHSearcher, and the published
    recall@10 = 0.31914062499999996 (benchmark_20260305_070532/random/lsh_results.json:46)."""
    g = np.load(os.path.join(GOLDEN, "python_lsh.npz"))
    base, queries, _ = random20k_inputs()
    algo = A.get_algorithm_instance("Composite", 64, name="lsh", metric="l2",
                                    indexer={"type": "LSHIndexer", "num_tables": 12, "hash_size": 4, "bucket_width": 20.0, "seed": 42},
                                    searcher={"type": "LSHSearcher", "candidate_multiplier": 64, "fallback_to_bruteforce": False})
    algo.build_index(base)
    dist, idx = algo.batch_search(queries, 20)
    _check((g["r20k_D"], g["r20k_I"]), (dist, idx), atol=1e-6)
    gt = g["r20k_gt20"]
    assert oracle.recall_at_k(gt, idx, 10) == pytest.approx(0.31914062499999996, abs=1e-12) == float(g["r20k_recall10"])
    assert oracle.recall_at_k(gt, idx, 1) == pytest.approx(0.34765625, abs=1e-12) == float(g["r20k_recall1"])
    # the exact path on the same data reproduces the reference's exact rows (recall 1.0 published)
    exact = A.get_algorithm_instance("Composite", 64, name="exact", metric="l2", indexer={"type": "BruteForceIndexer"},
                                     searcher={"type": "LinearSearcher"})
    exact.build_index(base)
    d, i = exact.batch_search(queries[:64], 20)
    _check((g["r20k_exact_D"], g["r20k_exact_I"]), (d, i))
    assert oracle.recall_at_k(gt[:64], i, 10) == 1.0


def test_python_lsh_reference_kat_outputs(A):
    """Outputs of the unmodified reference on its own two LSH KAT set-ups (tests/golden/python_lsh.npz)."""
    g = np.load(os.path.join(GOLDEN, "python_lsh.npz"))
    rng = np.random.RandomState(7)
    train = rng.randn(128, 16).astype(np.float32)
    train /= np.linalg.norm(train, axis=1, keepdims=True)
    algo = A.get_algorithm_instance("Composite", 16, name="lsh_cos", metric="cosine",
                                    indexer={"type": "LSHIndexer", "num_tables": 12, "hash_size": 16, "seed": 7},
                                    searcher={"type": "LSHSearcher", "candidate_multiplier": 12.0, "fallback_to_bruteforce": True})
    algo.build_index(train)
    d, i = algo.batch_search(train[:5].copy(), 4)
    _check((g["kat_cos_D"], g["kat_cos_I"]), (d, i), atol=2e-6)
    rng = np.random.RandomState(11)
    train = rng.randn(160, 8).astype(np.float32)
    algo = A.get_algorithm_instance("Composite", 8, name="lsh_l2", metric="l2",
                                    indexer={"type": "LSHIndexer", "num_tables": 10, "hash_size": 12, "bucket_width": 3.0, "seed": 11},
                                    searcher={"type": "LSHSearcher", "candidate_multiplier": 10.0, "fallback_to_bruteforce": True})
    algo.build_index(train)
    d, i = algo.batch_search(train[10:20].copy(), 4)
    _check((g["kat_l2_D"], g["kat_l2_I"]), (d, i), atol=2e-6)


@pytest.mark.parametrize("metric,params", [("cosine", dict(num_tables=8, hash_size=6)), ("cosine", dict(num_tables=12, hash_size=14)),
                                           ("l2", dict(num_tables=10, hash_size=3, bucket_width=6.0)),
                                           ("l2", dict(num_tables=4, hash_size=8, bucket_width=2.0))])
@pytest.mark.parametrize("fallback", [True, False])
def test_lsh_candidates_on_device_match_the_host_walk(A, metric, params, fallback):
    """Device-side bucket union + vote order (vdb_lsh_candidates, two radix sorts) against the NumPy restatement of the
    reference's Counter walk (src/algorithms/lsh.py:219-240): same candidates in the same most_common() order, hence
    identical results - with duplicated rows (many votes), budgets above and below the union size, and queries that
    hit no bucket at all."""
    rng = np.random.RandomState(17)
    base = rng.randn(30_000, 16).astype(np.float32)
    base[200:260] = base[5]
    queries = np.vstack([base[5:6], rng.randn(90, 16).astype(np.float32), 50.0 + rng.randn(3, 16).astype(np.float32)])
    for mult in (2.0, 40.0, 4000.0):
        got = {}
        for mode in ("device", "host"):
            algo = A.get_algorithm_instance("Composite", 16, name=f"lsh_{mode}", metric=metric,
                                            indexer=dict(type="LSHIndexer", seed=9, **params),
                                            searcher=dict(type="LSHSearcher", candidate_multiplier=mult, fallback_to_bruteforce=fallback,
                                                          candidate_generation=mode))
            algo.build_index(base)
            assert (algo.searcher._csr_ids is not None) == (mode == "device")
            got[mode] = algo.batch_search(queries, 10)
        np.testing.assert_array_equal(got["device"][1], got["host"][1])
        np.testing.assert_array_equal(got["device"][0], got["host"][0])
    assert (got["device"][1][0, 0] == 5) and got["device"][1].shape == (94, 10)


def test_lsh_self_retrieval_kats(A):
    """reference tests/test_composite_algorithm.py:108-166: identical vectors come back first with distance ~0."""
    rng = np.random.RandomState(7)
    vectors = rng.randn(128, 16).astype(np.float32)
    vectors /= np.linalg.norm(vectors, axis=1, keepdims=True)
    algo = A.get_algorithm_instance("LSH", 16, name="lsh_cos", metric="cosine", num_tables=12, hash_size=16,
                                    candidate_multiplier=12.0, seed=42)
    algo.build_index(vectors)
    dist, idx = algo.batch_search(vectors[:5], 1)
    np.testing.assert_array_equal(idx[:, 0], np.arange(5))
    np.testing.assert_allclose(dist[:, 0], 0.0, atol=1e-6)
    rng = np.random.RandomState(11)
    vectors = rng.randn(160, 8).astype(np.float32)
    comp = A.get_algorithm_instance("Composite", 8, name="lsh_l2", metric="l2",
                                    indexer={"type": "LSHIndexer", "num_tables": 10, "hash_size": 6, "bucket_width": 3.0, "seed": 1},
                                    searcher={"type": "LSHSearcher", "candidate_multiplier": 16.0})
    comp.build_index(vectors)
    dist, idx = comp.batch_search(vectors[10:20], 1)
    np.testing.assert_array_equal(idx[:, 0], np.arange(10, 20))
    np.testing.assert_allclose(dist[:, 0], 0.0, atol=1e-6)
    # no bucket hit and no fallback -> padding; with fallback -> exact scan
    far = np.full((1, 8), 1e4, dtype=np.float32)
    nofb = A.get_algorithm_instance("LSH", 8, name="x", metric="l2", num_tables=2, hash_size=8, bucket_width=0.5,
                                    fallback_to_bruteforce=False)
    nofb.build_index(vectors)
    d, i = nofb.batch_search(far, 3)
    assert (i == -1).all() and np.isinf(d).all()
    fb = A.get_algorithm_instance("LSH", 8, name="x", metric="l2", num_tables=2, hash_size=8, bucket_width=0.5)
    fb.build_index(vectors)
    d, i = fb.batch_search(far, 3)
    _check(oracle.linear_search(vectors, far, 3, "l2"), (d, i))


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_ivf_pipeline_given_same_centroids_and_recall(A, metric):
    rng = np.random.RandomState(5)
    centers = rng.randn(32, 24).astype(np.float32) * 2
    base = (centers[rng.randint(0, 32, 20000)] + 0.4 * rng.randn(20000, 24)).astype(np.float32)
    queries = (centers[rng.randint(0, 32, 128)] + 0.4 * rng.randn(128, 24)).astype(np.float32)
    algo = A.get_algorithm_instance("Composite", 24, name="ivf", metric=metric,
                                    indexer={"type": "FaissIVFIndexer", "index_type": "IVF64,Flat", "nprobe": 4},
                                    searcher={"type": "FaissSearcher", "nprobe": 8})
    algo.build_index(base)
    index = algo.index_artifact.data
    assert index.nprobe == 8 and index.ntotal == 20000 and algo.index_artifact.metadata["nprobe"] == 4
    dist, idx = algo.batch_search(queries, 10)
    # parity GIVEN the trained centroids and the device's assignment
    b, q = (oracle.safe_normalize(base), oracle.safe_normalize(queries)) if metric == "cosine" else (base, queries)
    m = "l2" if metric == "l2" else "ip"
    ref_d, ref_i, _ = oracle.ivf_flat_search(b, index.centroids, index._impl.assign.cpu().numpy(), q, 10, 8, m)
    got_d = dist if m == "l2" else -dist          # FaissSearcher negates ip / cosine scores (modular.py:545-546)
    _check((ref_d, ref_i), (got_d, idx), atol=0.0 if m == "l2" else 1e-5)
    gt = oracle.faiss_flat_search(b, q, 10, m)[1]
    assert oracle.recall_at_k(gt, idx, 10) > 0.8
    # legacy entry point, raw FAISS conventions
    legacy = A.get_algorithm_instance("ApproximateSearch", 24, name="approx", index_type="IVF64,Flat", metric="l2", nprobe=64)
    legacy.build_index(base)
    d2, i2 = legacy.batch_search(queries, 10)
    _check(oracle.faiss_flat_search(base, queries, 10, "l2"), (d2, i2))     # nprobe == nlist is exact
    with pytest.raises(ValueError):
        A.get_algorithm_instance("ApproximateSearch", 24, name="hnsw", index_type="IVF64,HNSW32", metric="l2")


def test_ivf_sq8_through_the_factory_indexer(A):
    """configs/benchmark_config.yaml:51-60 of the reference: FaissFactoryIndexer(index_key="IVF256,SQ8") + FaissSearcher."""
    rng = np.random.RandomState(8)
    base = (rng.randn(40_000, 32) + 3.0 * rng.randn(64, 32)[rng.randint(0, 64, 40_000)]).astype(np.float32)
    queries = base[rng.permutation(40_000)[:200]] + 0.05 * rng.randn(200, 32).astype(np.float32)
    for metric in ("l2", "cosine"):
        algo = A.get_algorithm_instance("Composite", 32, name="ivf_sq8", metric=metric,
                                        indexer={"type": "FaissFactoryIndexer", "index_key": "IVF256,SQ8", "nprobe": 24},
                                        searcher={"type": "FaissSearcher", "nprobe": 24})
        algo.build_index(base)
        assert type(algo.index_artifact.data).__name__ == "GpuIndexIVFSQ8" and algo.index_artifact.data.nprobe == 24
        dist, idx = algo.batch_search(queries, 20)
        assert dist.dtype == np.float32 and idx.dtype == np.int64 and dist.shape == (200, 20)
        assert bool(np.all(np.diff(dist, axis=1) >= 0))                       # ascending: squared L2 / negated cosine
        gt = oracle.linear_search(base, queries, 20, metric)[1]
        assert oracle.recall_at_k(gt, idx, 10) > 0.9, oracle.recall_at_k(gt, idx, 10)
        assert algo.get_memory_usage() < 40_000 * 32 * 4                      # one byte per component, not four


@pytest.mark.parametrize("index_key,min_recall", [("IVF256,PQ32", 0.55), ("PQ32", 0.55), ("IVF256,PQ8", 0.2)])
def test_product_quantiser_indexes_through_the_factory_indexer(A, index_key, min_recall):
    """configs/benchmark_config.yaml:36-50,61-72 of the reference: FaissFactoryIndexer(index_key="IVF256,PQ64" | "PQ64") +
    FaissSearcher.  Codes are lossy: the check is the contract (shapes, order, conventions) and a recall floor."""
    rng = np.random.RandomState(8)
    base = (rng.randn(40_000, 32) + 3.0 * rng.randn(64, 32)[rng.randint(0, 64, 40_000)]).astype(np.float32)
    queries = base[rng.permutation(40_000)[:200]] + 0.05 * rng.randn(200, 32).astype(np.float32)
    for metric in ("l2", "cosine"):
        algo = A.get_algorithm_instance("Composite", 32, name="pq", metric=metric,
                                        indexer={"type": "FaissFactoryIndexer", "index_key": index_key, "nprobe": 24},
                                        searcher={"type": "FaissSearcher", "nprobe": 24})
        algo.build_index(base)
        dist, idx = algo.batch_search(queries, 20)
        assert dist.dtype == np.float32 and idx.dtype == np.int64 and dist.shape == (200, 20)
        assert bool(np.all(np.diff(dist, axis=1) >= 0)) and int(idx.min()) >= 0 and int(idx.max()) < 40_000
        gt = oracle.linear_search(base, queries, 20, metric)[1]
        rec = oracle.recall_at_k(gt, idx, 10)
        assert rec > min_recall, (index_key, metric, rec)
    with pytest.raises(ValueError):
        A.get_algorithm_instance("Composite", 30, name="bad", metric="l2", indexer={"type": "FaissFactoryIndexer", "index_key": "PQ8"},
                                 searcher={"type": "FaissSearcher"}).build_index(base[:, :30])


def test_benchmark_runner_modular_end_to_end(A, tmp_path):
    """reference tests/test_benchmark_runner_modular.py:9-65: a tiny JSON config through
    BenchmarkRunner.run() with indexer_ref / searcher_ref resolution."""
    from vectordb_retrieval_b200.harness import BenchmarkRunner
    config = {
        "output_dir": str(tmp_path / "out"),
        "n_queries": 5, "topk": 5, "seed": 123, "query_batch_size": 2,
        "indexers": {"bf": {"type": "BruteForceIndexer"}},
        "searchers": {"linear": {"type": "LinearSearcher"}},
        "algorithms": {"modular_exact": {"indexer_ref": "bf", "searcher_ref": "linear"},
                       "exact": {"type": "ExactSearch"}},
        "datasets": [{"name": "random", "metric": "l2",
                      "dataset_options": {"dimensions": 3, "train_size": 32, "test_size": 6, "ground_truth_k": 5, "seed": 123}}],
    }
    path = tmp_path / "cfg.json"
    path.write_text(json.dumps(config))
    runner = BenchmarkRunner(str(path))
    results = runner.run()
    res = results["random"]["modular_exact"]
    assert res["n_train"] == 32 and res["n_test"] == 5 and "recall@1" in res and res["recall@1"] == 1.0
    assert results["random"]["exact"]["recall"] == 1.0
    assert res["parameters"]["indexer"]["type"] == "BruteForceIndexer"
    assert os.path.exists(os.path.join(runner.output_dir, "benchmark_summary.md"))
    assert os.path.exists(os.path.join(runner.output_dir, "all_results.json"))


def test_smoke_config_c1_through_cli_entry(A, tmp_path):
    """BASELINE.json configs[0]: 10k x 128, 100 queries, ExactSearch k=10 via the smoke YAML."""
    import yaml
    from vectordb_retrieval_b200.harness import BenchmarkRunner
    root = os.path.dirname(os.path.dirname(__file__))
    cfg = yaml.safe_load(open(os.path.join(root, "configs", "benchmark_config_smoke.yaml")))
    cfg["output_dir"] = str(tmp_path / "out")
    p = tmp_path / "smoke.yaml"
    p.write_text(yaml.dump(cfg))
    results = BenchmarkRunner(str(p)).run()["random"]
    assert results["exact"]["recall@10"] == 1.0 and results["exact"]["recall@1"] == 1.0
    assert results["exact_linear"]["recall@10"] == 1.0
    assert 0.3 < results["ivf_flat"]["recall@10"] <= 1.0
    assert results["faiss_lsh"]["recall@10"] > 0.5 and 0.0 < results["lsh"]["recall@10"] < 1.0
    assert results["exact"]["n_train"] == 10000 and results["exact"]["n_test"] == 100


@pytest.mark.parametrize("kind", ["exact", "ivf", "ivf_sq8", "ivf_pq", "ivf_pq_ip", "pq"])
def test_save_load_index_round_trip_is_bit_identical(A, kind, tmp_path):
    """save_index / load_index (reference base_algorithm.py:98-120; driver experiment_runner.py:308-344):
    a reloaded index answers exactly as the one that was saved."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((3000, 24)).astype(np.float32)
    q = rng.standard_normal((37, 24)).astype(np.float32)
    make = {"exact": lambda: A.ExactSearch("e", 24, metric="l2"),
            "ivf": lambda: A.ApproximateSearch("a", 24, index_type="IVF16,Flat", metric="l2", nprobe=4),
            "ivf_sq8": lambda: A.ApproximateSearch("a", 24, index_type="IVF16,SQ8", metric="l2", nprobe=4),
            "ivf_pq": lambda: A.ApproximateSearch("a", 24, index_type="IVF16,PQ4", metric="l2", nprobe=4),
            "ivf_pq_ip": lambda: A.ApproximateSearch("a", 24, index_type="IVF16,PQ4", metric="ip", nprobe=4),
            "pq": lambda: A.ApproximateSearch("a", 24, index_type="PQ6", metric="l2")}[kind]
    a = make()
    with pytest.raises(RuntimeError):
        a.save_index(str(tmp_path / "x"))
    a.build_index(x)
    d0, i0 = a.batch_search(q, 10)
    ctx = {"dataset_fingerprint": "fp", "config_hash": "c", "build_metrics": {"build_time_s": 0.25}}
    info = a.save_index(str(tmp_path / "art"), context=ctx)
    assert os.path.exists(os.path.join(info["artifact_dir"], "WRITE_COMPLETE"))
    b = make()
    loaded = b.load_index(str(tmp_path / "art"), context={"dataset_fingerprint": "fp"})
    assert loaded["build_time_s"] == 0.25 and b.index_built
    d1, i1 = b.batch_search(q, 10)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)
    with pytest.raises(RuntimeError):
        make().load_index(str(tmp_path / "art"), context={"dataset_fingerprint": "another"})
    with pytest.raises(FileNotFoundError):
        make().load_index(str(tmp_path / "missing"))


def test_lsh_index_save_load_round_trip(A, tmp_path):
    from vectordb_retrieval_b200.indexes import GpuIndexLSH
    rng = np.random.default_rng(12)
    x = rng.standard_normal((3000, 24)).astype(np.float32)
    q = rng.standard_normal((37, 24)).astype(np.float32)
    a = GpuIndexLSH(24, 64)
    a.add(x)
    d0, i0 = a.search(q, 50)
    a.save(str(tmp_path / "lsh"))
    b = GpuIndexLSH(24, 64)
    b.load(str(tmp_path / "lsh"))
    d1, i1 = b.search(q, 50)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)
    with pytest.raises(RuntimeError):
        GpuIndexLSH(24, 128).load(str(tmp_path / "lsh"))


def test_harness_persistence_modes(A, tmp_path):
    """build_only then retrieve_only through the harness reproduces the built run (reference
    experiment_runner.py:308-372)."""
    from vectordb_retrieval_b200.harness.config import ExperimentConfig
    from vectordb_retrieval_b200.harness.experiment_runner import ExperimentRunner
    art = str(tmp_path / "artifacts")

    def run(mode):
        cfg = ExperimentConfig(dataset="random", dataset_options={"train_size": 2000, "test_size": 50, "dimensions": 16},
                               n_queries=50, topk=10,
                               algorithms={"ivf": {"type": "ApproximateSearch", "index_type": "IVF8,Flat", "nprobe": 8,
                                                   "persistence": {"enabled": True, "mode": mode, "artifact_dir": art,
                                                                   "path_policy": "versioned"}}})
        r = ExperimentRunner(cfg, output_dir=str(tmp_path / mode))
        r.register_algorithm(A.get_algorithm_instance("ApproximateSearch", 16, name="ivf", index_type="IVF8,Flat", nprobe=8))
        return r.run()["ivf"]

    with pytest.raises(FileNotFoundError):
        run("retrieve_only")
    built = run("build_only")
    assert built["status"] == "build_only" and built["qps"] == 0.0 and os.path.isdir(built["persist_dir"])
    loaded = run("retrieve_only")
    assert loaded["index_source"] == "loaded" and loaded["index_load_time_s"] > 0 and loaded["recall@10"] == 1.0
