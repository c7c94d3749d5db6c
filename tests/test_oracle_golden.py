"""The oracle is pinned here: every function that restates a NumPy path of the reference
is checked against outputs of the UNMODIFIED reference classes (tests/golden/*.npz,
written by oracle/gen_golden.py), against the reference's own known-answer tests
(tests/test_composite_algorithm.py:29-226) and against the published LSH recall."""
import os

import numpy as np
import pytest

from oracle import oracle
from oracle.gen_golden import linear_inputs, random20k_inputs, rerank_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_linear_search_matches_reference(metric):
    g = _load("linear_searcher.npz")
    base, queries = linear_inputs()
    d, i = oracle.linear_search(base, queries, 10, metric)
    assert d.dtype == np.float32 and i.dtype == np.int64
    np.testing.assert_array_equal(i, g[f"{metric}_I"])
    np.testing.assert_allclose(d, g[f"{metric}_D"], rtol=1e-6, atol=1e-6)
    d, i = oracle.linear_search(base[:6], queries[:3], 8, metric)
    np.testing.assert_array_equal(i, g[f"{metric}_pad_I"])
    assert np.all(np.isinf(d[:, 6:])) and np.all(i[:, 6:] == -1)
    np.testing.assert_allclose(d[:, :6], g[f"{metric}_pad_D"][:, :6], rtol=1e-6, atol=1e-6)


def test_faiss_flat_conventions_against_linear_reference():
    """IndexFlat values are unpinned (no FAISS) but ranking must agree with the pinned
    LinearSearcher output: squared vs sqrt L2, +IP vs -IP."""
    g = _load("linear_searcher.npz")
    base, queries = linear_inputs()
    d, i = oracle.faiss_flat_search(base, queries, 10, "l2")
    np.testing.assert_array_equal(i, g["l2_I"])
    np.testing.assert_allclose(np.sqrt(d), g["l2_D"], rtol=2e-6)
    d, i = oracle.faiss_flat_search(base, queries, 10, "ip")
    np.testing.assert_array_equal(i, g["ip_I"])
    np.testing.assert_allclose(-d, g["ip_D"], rtol=1e-5, atol=1e-5)
    # ExactSearch quirk: any metric string other than 'l2' is raw inner product
    d2, i2 = oracle.faiss_flat_search(base, queries, 10, "cosine")
    np.testing.assert_array_equal(i2, i)
    d, i = oracle.faiss_flat_search(base[:6], queries[:3], 8, "l2")
    assert np.all(i[:, 6:] == -1) and np.all(d[:, 6:] == np.finfo(np.float32).max)
    db, ib = oracle.faiss_flat_search_blas(base, queries, 10, "l2")
    assert oracle.compare_topk(*oracle.faiss_flat_search(base, queries, 10, "l2"), db, ib, rtol=1e-5)["ok"]


def test_four_point_kat():
    """tests/test_composite_algorithm.py:29-58."""
    train = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.float32)
    queries = np.array([[0.1, 0.1], [0.9, 0.2]], dtype=np.float32)
    expected = np.argsort(np.linalg.norm(train[None] - queries[:, None], axis=2), axis=1)[:, :2]
    for fn in (oracle.linear_search, oracle.faiss_flat_search):
        _, i = fn(train, queries, 2, "l2")
        np.testing.assert_array_equal(i, expected)


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_rerank_matches_reference(metric):
    g = _load("faiss_lsh_rerank.npz")
    base, queries, cand = rerank_inputs()
    if metric == "cosine":
        b, q = oracle.safe_normalize(base), oracle.safe_normalize(queries)
    else:
        b, q = base, queries
    assert oracle.candidate_budget(10, 6.0, None, 500) == 60
    d, i = oracle.rerank_search(b, cand, q, 10, metric)
    np.testing.assert_array_equal(i, g[f"{metric}_I"])
    np.testing.assert_allclose(d, g[f"{metric}_D"], rtol=1e-6, atol=1e-6)
    assert (i[7, 5:] == -1).all() and np.isinf(d[7, 5:]).all()


def test_rerank_reversed_candidates_kat():
    """tests/test_composite_algorithm.py:169-226."""
    train = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.float32)
    cand = np.arange(3, -1, -1, dtype=np.int64)[None, :]
    d, i = oracle.rerank_search(train, cand, np.zeros((1, 2), np.float32), 2, "l2")
    assert i[0, 0] == 0 and abs(d[0, 0]) < 1e-6


def test_python_lsh_kats_and_published_recall():
    g = _load("python_lsh.npz")
    rng = np.random.RandomState(7)
    train = rng.randn(128, 16).astype(np.float32)
    train /= np.linalg.norm(train, axis=1, keepdims=True)
    t = oracle.LSHTables(train, "cosine", 12, 16, 4.0, 7)
    d, i = oracle.lsh_search(t, train[:5].copy(), 4, 12.0, None, True)
    np.testing.assert_array_equal(i, g["kat_cos_I"])
    np.testing.assert_allclose(d, g["kat_cos_D"], atol=1e-6)
    np.testing.assert_array_equal(i[:, 0], np.arange(5))
    np.testing.assert_allclose(d[:, 0], 0.0, atol=1e-6)

    rng = np.random.RandomState(11)
    train = rng.randn(160, 8).astype(np.float32)
    t = oracle.LSHTables(train, "l2", 10, 12, 3.0, 11)
    d, i = oracle.lsh_search(t, train[10:20].copy(), 4, 10.0, None, True)
    np.testing.assert_array_equal(i, g["kat_l2_I"])
    np.testing.assert_allclose(d, g["kat_l2_D"], atol=1e-6)
    np.testing.assert_array_equal(i[:, 0], np.arange(10, 20))

    train, q, _ = random20k_inputs()
    t = oracle.LSHTables(train, "l2", 12, 4, 20.0, 42)
    d, i = oracle.lsh_search(t, q, 20, 64.0, None, False)
    np.testing.assert_array_equal(i, g["r20k_I"])
    np.testing.assert_allclose(d, g["r20k_D"], rtol=1e-6)
    gt = g["r20k_gt20"]
    assert oracle.recall_at_k(gt, i, 10) == 0.31914062499999996 == float(g["r20k_recall10"])
    assert oracle.recall_at_k(gt, i, 1) == 0.34765625 == float(g["r20k_recall1"])
    # exact search on the same data: the published recall@10 = recall@1 = 1.0
    d, i = oracle.linear_search(train, q[:64], 20, "l2")
    np.testing.assert_array_equal(i, g["r20k_exact_I"])
    np.testing.assert_allclose(d, g["r20k_exact_D"], rtol=1e-6)
    assert oracle.recall_at_k(gt[:64], i, 10) == 1.0


def test_ivf_and_merge_properties():
    rng = np.random.RandomState(5)
    base = rng.randn(4000, 12).astype(np.float32)
    q = rng.randn(20, 12).astype(np.float32)
    cent = base[rng.permutation(4000)[:32]].copy()
    assign = oracle.ivf_assign(base, cent, "l2")
    # probing every list is exact search
    d, i, _ = oracle.ivf_flat_search(base, cent, assign, q, 10, 32, "l2")
    de, ie = oracle.faiss_flat_search(base, q, 10, "l2")
    np.testing.assert_array_equal(i, ie)
    np.testing.assert_allclose(d, de, rtol=1e-6)
    d1, i1, probes = oracle.ivf_flat_search(base, cent, assign, q, 10, 4, "l2")
    assert probes.shape == (20, 4)
    assert all(set(assign[i1[r][i1[r] >= 0]]) <= set(probes[r]) for r in range(20))
    # shard merge equals the unsharded search
    parts = [oracle.faiss_flat_search(base[s:s + 1000], q, 10, "l2") for s in range(0, 4000, 1000)]
    dm, im = oracle.merge_topk([p[0] for p in parts], [p[1] + 1000 * n for n, p in enumerate(parts)], 10)
    np.testing.assert_array_equal(im, ie)
    np.testing.assert_allclose(dm, de, rtol=1e-6)


def test_comparator_accepts_ties_and_rejects_errors():
    rd = np.array([[1.0, 2.0, 2.0 + 1e-7, 3.0]])
    ri = np.array([[5, 6, 7, 8]])
    assert oracle.compare_topk(rd, ri, rd, ri)["ok"]
    swapped = oracle.compare_topk(rd, ri, rd, np.array([[5, 7, 6, 8]]))
    assert swapped["ok"] and swapped["tie_swaps"] == 2
    assert not oracle.compare_topk(rd, ri, rd, np.array([[6, 5, 7, 8]]))["ok"]
    assert not oracle.compare_topk(rd, ri, rd * 1.001, ri)["ok"]
    # boundary tie: a different id with the same k-th distance is accepted
    assert oracle.compare_topk(rd, ri, rd, np.array([[5, 6, 7, 99]]))["ok"]
