"""Dataset-side pieces next to the hot path: the ``.fvecs`` / ``.ivecs`` readers
(reference src/benchmark/dataset.py:522-574), the exact GPU search as ground-truth builder
(dataset.py:858-964) and a memory-mapped base streamed into row shards (dataset.py:986-1053)."""
import os

import numpy as np
import pytest

from oracle import oracle  # checker only


def _write_vecs(path, arr):
    """TEXMEX layout: every row = int32 dimension followed by the components (float32 / int32)."""
    n, d = arr.shape
    out = np.empty((n, d + 1), dtype=np.int32)
    out[:, 0] = d
    out[:, 1:] = arr.view(np.int32) if arr.dtype == np.float32 else arr.astype(np.int32)
    out.tofile(path)


def test_fvecs_ivecs_round_trip_and_dataset_loader(tmp_path):
    from vectordb_retrieval_b200.harness.dataset import Dataset, read_fvecs, read_ivecs
    rng = np.random.RandomState(0)
    base = rng.randn(300, 24).astype(np.float32)
    queries = rng.randn(17, 24).astype(np.float32)
    gt = np.argsort(((base[None] - queries[:, None]) ** 2).sum(-1), axis=1)[:, :10].astype(np.int32)
    root = tmp_path / "sift"
    root.mkdir()
    _write_vecs(root / "sift_base.fvecs", base)
    _write_vecs(root / "sift_query.fvecs", queries)
    _write_vecs(root / "sift_groundtruth.ivecs", gt)
    got = read_fvecs(str(root / "sift_base.fvecs"))
    assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
    np.testing.assert_array_equal(got, base)                                   # bit for bit
    np.testing.assert_array_equal(read_fvecs(str(root / "sift_base.fvecs"), limit=7), base[:7])
    np.testing.assert_array_equal(read_ivecs(str(root / "sift_groundtruth.ivecs")), gt)
    ds = Dataset("fvecs", options={"path": str(root)})
    ds.load()
    np.testing.assert_array_equal(ds.train_vectors, base)
    np.testing.assert_array_equal(ds.test_vectors, queries)
    np.testing.assert_array_equal(ds.ground_truth, gt)
    assert ds.ground_truth_source == "file"
    # a base limit invalidates the shipped ground truth: it is recomputed (NumPy recipe at this size)
    ds = Dataset("fvecs", options={"path": str(root), "base_limit": 100, "ground_truth_k": 5})
    ds.load()
    assert ds.train_vectors.shape == (100, 24) and ds.ground_truth_source == "numpy"
    ref = np.argsort(np.linalg.norm(base[None, :100] - queries[:, None], axis=2), axis=1, kind="stable")[:, :5]
    np.testing.assert_array_equal(ds.ground_truth, ref)


@pytest.mark.gpu
def test_gpu_ground_truth_builder_matches_the_numpy_recipe():
    """300k x 32 with 700 queries is past the NumPy limit: ``ground_truth_source == "gpu_exact"``; the same
    dataset forced through the reference's recipe (argsort of L2 norms, dataset.py:497-504) must agree."""
    from vectordb_retrieval_b200.harness.dataset import Dataset
    opts = {"dimensions": 32, "train_size": 300_000, "test_size": 700, "ground_truth_k": 50, "seed": 5}
    gpu = Dataset("random", options=dict(opts))
    gpu.load()
    assert gpu.ground_truth_source == "gpu_exact" and gpu.ground_truth.shape == (700, 50)
    cpu = Dataset("random", options=dict(opts, ground_truth="cpu", test_size=700))
    cpu._random(32, 300_000, 700)
    cpu.test_vectors = cpu.test_vectors[:64]
    cpu._compute_ground_truth(50, "l2")
    assert cpu.ground_truth_source == "numpy"
    np.testing.assert_array_equal(gpu.train_vectors, cpu.train_vectors)
    # fp32 norms in the NumPy recipe vs exact re-scoring on the device: ids may swap only inside fp32-level ties
    same = gpu.ground_truth[:64] == cpu.ground_truth
    assert same.mean() > 0.999, same.mean()
    assert oracle.recall_at_k(cpu.ground_truth, gpu.ground_truth[:64], 50) > 0.9995
    cos = Dataset("random", options=dict(opts, metric="cosine", train_size=250_000, test_size=900))
    cos.load()
    assert cos.ground_truth_source == "gpu_exact"
    ref = oracle.linear_search(cos.train_vectors, cos.test_vectors[:16], 50, "cosine")
    assert oracle.recall_at_k(ref[1], cos.ground_truth[:16], 50) > 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_memmap_base_streams_into_row_shards(tmp_path, dtype):
    """A memory-mapped base (the reference keeps large pre-embedded sets as ``np.memmap``, dataset.py:397-414)
    is sliced by row range and uploaded block-wise; two shards + merge == one shard built from RAM."""
    import torch
    from vectordb_retrieval_b200 import engine
    n, d = 70_001, 40
    rng = np.random.RandomState(8)
    ram = rng.randn(n, d).astype(np.float32)
    path = os.path.join(tmp_path, "base.bin")
    mm = np.memmap(path, dtype=dtype, mode="w+", shape=(n, d))
    mm[:] = ram
    mm.flush()
    ro = np.memmap(path, dtype=dtype, mode="r", shape=(n, d))            # read-only, like the harness' arrays
    q = torch.from_numpy(rng.randn(130, d).astype(np.float32)).cuda()
    D1, I1 = engine.FlatShard(ram, "l2", "cuda").search(q, 100)
    cut = 33_333
    parts = [engine.FlatShard(ro[:cut], "l2", "cuda", id_offset=0, upload_rows=8192),
             engine.FlatShard(ro[cut:], "l2", "cuda", id_offset=cut, upload_rows=8192)]
    res = [p.search(q, 100) for p in parts]
    Dm, Im = engine.merge_topk(torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res]))
    assert torch.equal(Im, I1) and torch.equal(Dm, D1)
    ref = oracle.faiss_flat_search(ram, q.cpu().numpy()[:32], 100, "l2")
    res = oracle.compare_topk(ref[0], ref[1], Dm.cpu().numpy()[:32], Im.cpu().numpy()[:32], rtol=1e-5)
    assert res["ok"], res
