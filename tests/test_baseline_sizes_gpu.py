"""Parity at the sizes BASELINE.json names (configs[2..4]) and for the IVF training recipe.

The oracle runs on a SAMPLE of the queries so that every case stays within seconds-to-a-minute of host
time; the GPU side runs a full-width batch.  Everything goes through the C ABI (engine.* are ctypes calls).
  C3  GloVe-50 shape: 1.2M x 50 cosine, IVF nlist = 4096, nprobe 1 / 32 / 128; Hamming top-6400 over 256-bit codes
      (tensor-pipe scan vs popc scan, bit-identical) and the rerank of those candidates
  C4  MS MARCO shape, one GPU's share of 8: 1.1M x 768 inner product, exact top-100
  C5  100M x 128 shape, one GPU's share of 8: a 12.5M-row shard whose ids start beyond 2^32, merged with a
      second shard (int64 ids end to end)
  k-means: ``engine.kmeans_step`` / ``kmeans_train`` against ``oracle.kmeans_lloyd`` (same sample, same start, fp64)"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import oracle  # noqa: E402  (checker only)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from vectordb_retrieval_b200 import _lib, engine
    _lib.load()
    return engine


def _check(ref, got, rtol=1e-5, atol=0.0):
    res = oracle.compare_topk(ref[0], ref[1], got[0], got[1], rtol=rtol, atol=atol)
    assert res["ok"], res
    return res


def _randn(n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)


# ------------------------------------------------------------------------------------------ k-means
def _blobs(n, d, centers, seed, spread=0.15):
    rng = np.random.RandomState(seed)
    c = rng.randn(centers, d).astype(np.float32) * 2.0
    return (c[rng.randint(0, centers, n)] + spread * rng.randn(n, d)).astype(np.float32)


@pytest.mark.parametrize("metric", ["l2", "cosine", "ip"])
def test_kmeans_step_matches_fp64_lloyd_step(eng, metric):
    """One iteration from the same centroids: sizes and means must agree with the fp64 restatement (the device
    assigns with the 3xTF32 scan + exact re-scoring and accumulates the sums with fp32 atomics)."""
    x = np.random.RandomState(5).randn(60000, 24).astype(np.float32)
    rows, init = eng.kmeans_sample(x, 96, seed=1234)
    assert rows.shape[0] == 256 * 96 and np.all(np.diff(rows) > 0) and np.all(np.diff(init) > 0)
    sample = x[rows]
    if metric == "cosine":
        sample = oracle.safe_normalize(sample)
    cent = sample[init].copy()
    spherical = metric != "l2"
    new, sizes = eng.kmeans_step(torch.from_numpy(sample).cuda(), torch.from_numpy(cent).cuda(), spherical)
    ref_new, ref_assign, ref_sizes = oracle.kmeans_lloyd_step(sample, cent, spherical)
    assert int(np.abs(sizes - ref_sizes).sum()) <= 4, (sizes - ref_sizes)[np.nonzero(sizes - ref_sizes)]   # fp ties only
    np.testing.assert_allclose(new.cpu().numpy(), ref_new, atol=2e-3 * float(np.abs(ref_new).max()), rtol=0)
    close = np.abs(new.cpu().numpy() - ref_new).max(axis=1) < 1e-5 * max(1.0, float(np.abs(ref_new).max()))
    assert close.mean() > 0.9, close.mean()        # clusters untouched by a tie flip agree to fp32 rounding


def test_kmeans_empty_cluster_is_split_like_the_oracle(eng):
    x = np.random.RandomState(9).randn(4000, 8).astype(np.float32)
    cent = x[:16].copy()
    cent[7] = cent[3]                                # duplicate centroid: ties go to the lower index, 7 stays empty
    new, sizes = eng.kmeans_step(torch.from_numpy(x).cuda(), torch.from_numpy(cent).cuda(), False)
    ref_new, _, ref_sizes = oracle.kmeans_lloyd_step(x, cent, False)
    assert sizes[7] == 0 and ref_sizes[7] == 0
    np.testing.assert_array_equal(sizes, ref_sizes)
    np.testing.assert_allclose(new.cpu().numpy(), ref_new, atol=1e-5)


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_kmeans_training_matches_fp64_lloyd(eng, metric):
    """Ten iterations end to end (reference src/algorithms/modular.py:281-282 -> index.train): same sample and
    start as the oracle; on clustered data the trajectories stay together, so centroids and the final
    objective agree closely.  Parity vs FAISS's own k-means stays unpinned (its RNG is not reproducible here)."""
    x = _blobs(50000, 16, 48, seed=21)
    cent = eng.kmeans_train(x, 48, metric, "cuda", niter=10, seed=77)
    ref_cent, rows, inertia = oracle.kmeans_lloyd(x, 48, metric, niter=10, seed=77)
    assert inertia[-1] <= inertia[0]
    sample = x[rows] if metric == "l2" else oracle.safe_normalize(x[rows])
    spherical = metric != "l2"

    def objective(c):
        _, assign, _ = oracle.kmeans_lloyd_step(sample, c, spherical)
        return float(((sample.astype(np.float64) - c.astype(np.float64)[assign]) ** 2).sum(axis=1).mean())

    assert objective(cent) == pytest.approx(objective(ref_cent), rel=1e-3)
    np.testing.assert_allclose(cent, ref_cent, atol=5e-3 * float(np.abs(ref_cent).max()), rtol=0)


# ------------------------------------------------------------------------------------------ C3
@pytest.fixture(scope="module")
def c3(eng):
    n, d, nq = 1_200_000, 50, 2048
    base = _randn(n, d, 300)
    q = _randn(nq, d, 301)
    return {"n": n, "d": d, "nq": nq, "base": base, "q": q, "base_host": base.cpu().numpy(), "q_host": q.cpu().numpy(),
            "pick": np.sort(np.random.RandomState(1).choice(nq, 32, replace=False))}


def test_c3_ivf_flat_nlist4096_matches_oracle_at_size(eng, c3):
    """1.2M x 50 cosine, IVF4096,Flat: the device's own centroids / assignments handed to the oracle
    (src/algorithms/modular.py:277-286,536-548).  Checks the list layout at size (4 096 lists, block offsets in
    the tens of thousands, ragged list lengths), the nprobe = 128 probe matrix and the scan."""
    base_host, q_host, pick = c3["base_host"], c3["q_host"], c3["pick"]
    cent = eng.kmeans_train(base_host, 4096, "cosine", "cuda", niter=4)          # 4 iterations: the recipe is pinned above
    ivf = eng.IVFShard(c3["base"], cent, "cosine", "cuda")
    assign = ivf.assign.cpu().numpy()
    counts = np.bincount(assign, minlength=4096)
    assert int(counts.sum()) == c3["n"] and int(ivf.blk_off[-1].item()) == int(((counts + 31) // 32).sum())
    bn, qn = oracle.safe_normalize(base_host), oracle.safe_normalize(q_host[pick])
    ref_assign = oracle.ivf_assign(bn[:20000], cent, "ip")
    assert (ref_assign == assign[:20000]).mean() > 0.999
    for nprobe in (1, 32, 128):
        scanned = torch.zeros(1, dtype=torch.int64, device="cuda")
        D, I = ivf.search(c3["q"], 100, nprobe, 0, -oracle.FLT_MAX, scanned)
        ref_d, ref_i, _ = oracle.ivf_flat_search(bn, cent, assign, qn, 100, nprobe, "ip")
        _check((ref_d, ref_i), (D[pick].cpu().numpy(), I[pick].cpu().numpy()), atol=1e-5)
        assert int(scanned.item()) == int(counts[ivf.last_probes.cpu().numpy()].sum())
        assert bool((D[:, 1:] <= D[:, :-1]).all())
    del ivf


def test_c3_hamming_top6400_tensor_pipe_vs_popc_and_rerank_at_size(eng, c3):
    """256-bit sign codes of the 1.2M rows, C = 6 400 candidates per query (k = 100, multiplier 64,
    src/algorithms/modular.py:463-468): tensor-pipe scan == popc scan bit for bit, then the rerank of those
    candidates against ``oracle.rerank_search`` (modular.py:483-532) on the sampled queries."""
    from vectordb_retrieval_b200 import _lib
    base_host, q_host, pick = c3["base_host"], c3["q_host"], c3["pick"]
    nq, cand_k = 512, 6400
    proj = np.random.RandomState(1234).normal(size=(256, c3["d"])).astype(np.float32)
    shard = eng.HammingShard(c3["base"], proj, "cuda")
    q = c3["q"][:nq]
    shard.tensor_pipe = False
    d0, i0 = shard.search(q, cand_k)
    shard.tensor_pipe = True
    d1, i1 = shard.search(q, cand_k)
    assert torch.equal(d0, d1) and torch.equal(i0, i1)
    assert bool((d1[:, 1:] >= d1[:, :-1]).all()) and int(i1.min()) >= 0 and int(i1.max()) < c3["n"]
    sub = pick[pick < nq][:16]
    sub_t = torch.from_numpy(sub).cuda()
    codes = shard.codes.cpu().numpy().view(np.uint32)                 # the device's own codes: the scan is what is checked
    qcodes = shard.encode(q[sub_t]).cpu().numpy().view(np.uint32)
    ref_hd, ref_hi = oracle.hamming_topk(codes, qcodes, cand_k)
    np.testing.assert_array_equal(d1[sub_t].cpu().numpy(), ref_hd)
    np.testing.assert_array_equal(i1[sub_t].cpu().numpy(), ref_hi)
    # rerank (cosine conventions of FaissSearcher: negated scores)
    rr = eng.Reranker(c3["base"], "cosine", "cuda")
    D, I = rr.search(q, i1, 100, _lib.OUT_NEGATE, float("inf"))
    ref = oracle.rerank_search(oracle.safe_normalize(base_host), i1[sub_t].cpu().numpy(), oracle.safe_normalize(q_host[sub]),
                               100, "cosine")
    _check(ref, (D[sub_t].cpu().numpy(), I[sub_t].cpu().numpy()), atol=2e-6)


# ------------------------------------------------------------------------------------------ C4
def test_c4_one_gpu_share_1p1m_x_768_inner_product(eng):
    """8.8M x 768 inner product over 8 GPUs = 1.1M rows per GPU; d = 768 streams the query tile with the base
    (kpad > 128).  16 sampled queries against fp64 brute force (src/algorithms/exact_search.py:76-78, IP = raw
    scores descending)."""
    n, d, nq = 1_100_000, 768, 1024
    base, q = _randn(n, d, 400), _randn(nq, d, 401)
    shard = eng.FlatShard(base, "ip", "cuda", id_offset=7 * n)
    D, I = shard.search(q, 100, 0, -oracle.FLT_MAX)
    assert bool((D[:, 1:] <= D[:, :-1]).all()) and int(I.min()) >= 7 * n and int(I.max()) < 8 * n
    pick = np.sort(np.random.RandomState(2).choice(nq, 16, replace=False))
    base_host = base.cpu().numpy()
    del base, shard
    torch.cuda.empty_cache()
    ref_d, ref_i = oracle.faiss_flat_search(base_host, q.cpu().numpy()[pick], 100, "ip")
    scale = float(np.linalg.norm(base_host[:4096], axis=1).max()) * float(np.linalg.norm(q.cpu().numpy(), axis=1).max())
    _check((ref_d, ref_i + 7 * n), (D.cpu().numpy()[pick], I.cpu().numpy()[pick]), atol=1e-5 * scale)


# ------------------------------------------------------------------------------------------ C5
def test_c5_shard_ids_beyond_2_to_32_and_merge(eng):
    """One GPU's share of the 100M x 128 base (12.5M rows) with id_offset = 3e9, merged with a second shard
    whose ids follow on: ids travel as int64 from the finalize kernel through the merge kernel, and the merged
    list equals fp64 brute force over the concatenated rows."""
    n0, n1, d, nq, k = 12_500_000, 1_000_000, 128, 1000, 100
    off0 = 3_000_000_000
    off1 = off0 + n0
    b0 = _randn(n0, d, 500)
    q = _randn(nq, d, 502)
    s0 = eng.FlatShard(b0, "l2", "cuda", id_offset=off0)
    D0, I0 = s0.search(q, k)
    assert int(I0.min()) >= off0 and int(I0.max()) < off0 + n0 and int(I0.max()) > (1 << 32) // 2
    pick = np.sort(np.random.RandomState(3).choice(nq, 16, replace=False))
    q_pick = q.cpu().numpy()[pick]
    b0_host = b0.cpu().numpy()
    del b0, s0
    torch.cuda.empty_cache()
    b1 = _randn(n1, d, 501)
    D1, I1 = eng.FlatShard(b1, "l2", "cuda", id_offset=off1).search(q, k)
    Dm, Im = eng.merge_topk(torch.stack([D0, D1]), torch.stack([I0, I1]))
    assert Im.dtype == torch.int64 and bool((Dm[:, 1:] >= Dm[:, :-1]).all())
    assert int((Im >= off1).sum()) > 0 and int((Im < off1).sum()) > 0
    ref0 = oracle.faiss_flat_search(b0_host, q_pick, k, "l2")
    del b0_host
    ref1 = oracle.faiss_flat_search(b1.cpu().numpy(), q_pick, k, "l2")
    ref_d, ref_i = oracle.merge_topk([ref0[0], ref1[0]], [ref0[1] + off0, ref1[1] + off1], k)
    _check((ref_d, ref_i), (Dm.cpu().numpy()[pick], Im.cpu().numpy()[pick]))
    _check((ref0[0], ref0[1] + off0), (D0.cpu().numpy()[pick], I0.cpu().numpy()[pick]))
