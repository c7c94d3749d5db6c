"""GPU parity tests of the CUDA kernels, called through the C ABI, against the CPU oracle.

Tolerances (north star): ids bit-exact except inside distance ties within 1e-5 relative;
distances within 1e-5 relative (inner-product scores additionally get an absolute floor of
1e-5 * |q||x|, since a relative bound is meaningless around zero)."""
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


def _data(n, d, nq, seed=0, scale=1.0):
    rng = np.random.RandomState(seed)
    return (rng.randn(n, d) * scale).astype(np.float32), (rng.randn(nq, d) * scale).astype(np.float32)


def _keys_ref(base, q, metric):
    b, qq = base.astype(np.float64), q.astype(np.float64)
    ip = qq @ b.T
    return (b ** 2).sum(1)[None, :] - 2 * ip if metric == "l2" else -2 * ip


IMPLS = ["tcgen05", "tcgen05_1cta", "simt"]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("n,d,nq", [(1000, 128, 200), (700, 50, 130), (520, 160, 300), (300, 32, 5)])
def test_dense_keys_match_fp64(eng, impl, n, d, nq):
    """The 3xTF32 contraction (hi*hi + hi*lo + lo*hi, fp32 accumulate) must be fp32-accurate."""
    from vectordb_retrieval_b200 import _lib
    base, q = _data(n, d, nq, seed=n + d)
    for metric in ("l2", "ip"):
        shard = eng.FlatShard(base, metric, "cuda")
        keys = shard.dense_keys(torch.from_numpy(q).cuda(), _lib.IMPL_NAMES[impl]).cpu().numpy().astype(np.float64)
        ref = _keys_ref(base, q, metric)
        scale = np.abs(q).astype(np.float64) @ np.abs(base).astype(np.float64).T * 2 + 1.0
        err = np.abs(keys - ref) / scale
        assert err.max() < 2e-6, f"{impl} {metric} max scaled err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


def _check(ref, got, rtol=1e-5, atol=0.0):
    res = oracle.compare_topk(ref[0], ref[1], got[0], got[1], rtol=rtol, atol=atol)
    assert res["ok"], res
    return res


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("n,d,nq,k", [(5000, 64, 300, 10), (20000, 128, 700, 100), (3000, 50, 64, 200), (4, 2, 2, 2),
                                      (2500, 768, 70, 100), (70000, 1, 33, 504), (1, 5, 3, 4)])
def test_flat_topk_matches_faiss_flat_oracle(eng, impl, metric, n, d, nq, k):
    from vectordb_retrieval_b200 import _lib
    base, q = _data(n, d, nq, seed=7 * n + d)
    shard = eng.FlatShard(base, metric, "cuda")
    D, I = shard.search(torch.from_numpy(q).cuda(), k, 0, oracle.FLT_MAX if metric == "l2" else -oracle.FLT_MAX,
                        _lib.IMPL_NAMES[impl])
    ref = oracle.faiss_flat_search(base, q, k, metric)
    atol = 0.0 if metric == "l2" else 1e-5 * float(np.linalg.norm(base, axis=1).max() * np.linalg.norm(q, axis=1).max())
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()), atol=atol)
    assert I.dtype == torch.int64 and D.dtype == torch.float32
    kk = min(k, n)
    assert oracle.recall_at_k(ref[1][:, :kk], I.cpu().numpy()[:, :kk], kk) == 1.0


def test_flat_topk_linear_searcher_conventions_and_padding(eng):
    """sqrt-L2 / negated IP / cosine as LinearSearcher reports them, (+inf, -1) padding for k > n."""
    from vectordb_retrieval_b200 import _lib
    base, q = _data(600, 24, 50, seed=3)
    base[17] = 0.0
    for metric, flags in (("l2", _lib.OUT_SQRT), ("ip", _lib.OUT_NEGATE), ("cosine", _lib.OUT_NEGATE)):
        shard = eng.FlatShard(base, metric, "cuda")
        D, I = shard.search(torch.from_numpy(q.copy()).cuda(), 10, flags, float("inf"))
        ref = oracle.linear_search(base, q, 10, metric)
        _check(ref, (D.cpu().numpy(), I.cpu().numpy()), atol=0.0 if metric == "l2" else 2e-6 * (1 if metric == "cosine" else 150))
    shard = eng.FlatShard(base[:6], "l2", "cuda")
    D, I = shard.search(torch.from_numpy(q[:3]).cuda(), 8, _lib.OUT_SQRT, float("inf"))
    ref = oracle.linear_search(base[:6], q[:3], 8, "l2")
    np.testing.assert_array_equal(I.cpu().numpy(), ref[1])
    assert np.isinf(D.cpu().numpy()[:, 6:]).all()


def test_self_distance_is_zero(eng):
    """Difference-form re-scoring: a query equal to a base row has distance exactly ~0
    (reference tests/test_composite_algorithm.py:134,165 demand atol 1e-6)."""
    from vectordb_retrieval_b200 import _lib
    base, _ = _data(5000, 128, 1, seed=11, scale=10.0)
    q = base[100:140].copy()
    D, I = eng.FlatShard(base, "l2", "cuda").search(torch.from_numpy(q).cuda(), 5, _lib.OUT_SQRT, float("inf"))
    np.testing.assert_array_equal(I.cpu().numpy()[:, 0], np.arange(100, 140))
    np.testing.assert_allclose(D.cpu().numpy()[:, 0], 0.0, atol=1e-6)


def test_integer_valued_data_with_exact_ties(eng):
    """SIFT-like U{0..255} integer vectors and duplicated rows produce exact ties."""
    rng = np.random.RandomState(5)
    base = rng.randint(0, 256, size=(8000, 128)).astype(np.float32)
    base[4000:4100] = base[:100]          # duplicates
    q = rng.randint(0, 256, size=(200, 128)).astype(np.float32)
    q[:20] = base[:20]
    D, I = eng.FlatShard(base, "l2", "cuda").search(torch.from_numpy(q).cuda(), 100)
    ref = oracle.faiss_flat_search(base, q, 100, "l2")
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()))
    # deterministic tie-break: lowest id first
    assert (I.cpu().numpy()[:20, 0] == np.arange(20)).all() and (I.cpu().numpy()[:20, 1] == np.arange(4000, 4020)).all()


@pytest.mark.parametrize("impl", IMPLS)
def test_tie_groups_wider_than_the_spare_slots_keep_the_lowest_ids(eng, impl):
    """400 copies of one vector scattered over 90 000 rows, the query equal to it: 400 rows tie at distance 0, far
    more than the k' - k spare pool slots.  The (distance, id) contract demands the 100 LOWEST ids whatever order
    the scan meets the copies in - one shard, three shards + merge, tensor-core and CUDA-core kernels alike."""
    from vectordb_retrieval_b200 import _lib
    rng = np.random.RandomState(12)
    n, d, k = 90_000, 32, 100
    base = rng.randn(n, d).astype(np.float32)
    dup = np.sort(rng.choice(n, 400, replace=False))
    base[dup] = base[dup[0]]
    q = np.vstack([base[dup[0]][None, :], rng.randn(40, d).astype(np.float32)])
    qd = torch.from_numpy(q).cuda()
    code = _lib.IMPL_NAMES[impl]
    D, I = eng.FlatShard(base, "l2", "cuda").search(qd, k, 0, oracle.FLT_MAX, code)
    np.testing.assert_array_equal(I.cpu().numpy()[0], dup[:k])
    assert float(D[0].abs().max()) == 0.0
    ref = oracle.faiss_flat_search(base, q[1:], k, "l2")
    _check(ref, (D.cpu().numpy()[1:], I.cpu().numpy()[1:]))
    bounds = [0, 29_999, 61_000, n]
    parts = [eng.FlatShard(base[a:b], "l2", "cuda", id_offset=a).search(qd, k, 0, oracle.FLT_MAX, code) for a, b in zip(bounds[:-1], bounds[1:])]
    Dm, Im = eng.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(Im, I) and torch.equal(Dm, D)


def test_id_offset_and_merge_is_shard_count_invariant(eng):
    base, q = _data(9000, 64, 300, seed=9)
    ref = oracle.faiss_flat_search(base, q, 50, "l2")
    qd = torch.from_numpy(q).cuda()
    for parts in (2, 3, 8):
        bounds = np.linspace(0, 9000, parts + 1).astype(int)
        ds, is_ = [], []
        for p in range(parts):
            sh = eng.FlatShard(base[bounds[p]:bounds[p + 1]], "l2", "cuda", id_offset=int(bounds[p]))
            d, i = sh.search(qd.clone(), 50)
            ds.append(d), is_.append(i)
        D, I = eng.merge_topk(torch.stack(ds), torch.stack(is_))
        np.testing.assert_array_equal(I.cpu().numpy(), ref[1])
        _check(ref, (D.cpu().numpy(), I.cpu().numpy()))
    # descending (+IP) merge
    ref = oracle.faiss_flat_search(base, q, 20, "ip")
    ds, is_ = [], []
    for p in range(2):
        sh = eng.FlatShard(base[p * 4500:(p + 1) * 4500], "ip", "cuda", id_offset=p * 4500)
        d, i = sh.search(qd.clone(), 20, 0, -oracle.FLT_MAX)
        ds.append(d), is_.append(i)
    D, I = eng.merge_topk(torch.stack(ds), torch.stack(is_), descending=True, pad_value=-oracle.FLT_MAX)
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()), atol=1e-3)


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_rerank_matches_reference_golden_and_oracle(eng, metric):
    import os
    from oracle.gen_golden import rerank_inputs
    from vectordb_retrieval_b200 import _lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "faiss_lsh_rerank.npz"))
    base, q, cand = rerank_inputs()
    rr = eng.Reranker(base, metric, "cuda")
    flags = _lib.OUT_SQRT if metric == "l2" else _lib.OUT_NEGATE
    D, I = rr.search(torch.from_numpy(q.copy()).cuda(), torch.from_numpy(cand).cuda(), 10, flags)
    np.testing.assert_array_equal(I.cpu().numpy(), g[f"{metric}_I"])
    np.testing.assert_allclose(D.cpu().numpy(), g[f"{metric}_D"], rtol=1e-5, atol=2e-6)
    # larger random case, d not a multiple of 4, many candidates
    rng = np.random.RandomState(1)
    base, q = _data(20000, 50, 64, seed=21)
    cand = np.stack([rng.permutation(20000)[:1500] for _ in range(64)]).astype(np.int64)
    cand[5, 100:] = -1
    b, qq = (oracle.safe_normalize(base), oracle.safe_normalize(q)) if metric == "cosine" else (base, q)
    ref = oracle.rerank_search(b, cand, qq, 100, metric)
    D, I = eng.Reranker(base, metric, "cuda").search(torch.from_numpy(q.copy()).cuda(), torch.from_numpy(cand).cuda(), 100, flags)
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()), atol=0.0 if metric == "l2" else 1e-5)
    # one-minus convention of the Python LSH searcher (lsh.py:246)
    if metric == "cosine":
        D1, I1 = eng.Reranker(base, metric, "cuda").search(torch.from_numpy(q.copy()).cuda(), torch.from_numpy(cand).cuda(), 100, _lib.OUT_ONE_MINUS)
        np.testing.assert_allclose(D1.cpu().numpy()[np.isfinite(ref[0])], 1.0 + ref[0][np.isfinite(ref[0])], atol=2e-6)


@pytest.mark.parametrize("c,k", [(40, 10), (800, 100), (5000, 100), (9000, 200), (1100, 500), (17000, 64)])
def test_rerank_launch_shapes_match_oracle(eng, c, k):
    """1, 2, 4 and 8 warps per query (about 4 096 candidates per warp; several queries share a CTA below 8):
    every shape must give the oracle's rerank, including short and empty candidate rows."""
    from vectordb_retrieval_b200 import _lib
    base, q = _data(20000, 50, 77, seed=c)
    rng = np.random.RandomState(c + 1)
    cand = np.stack([rng.permutation(20000)[:c] for _ in range(77)]).astype(np.int64)
    cand[5, c // 2:] = -1
    cand[9, :] = -1
    for metric, flags in (("l2", _lib.OUT_SQRT), ("ip", _lib.OUT_NEGATE)):
        rr = eng.Reranker(base, metric, "cuda")
        D, I = rr.search(torch.from_numpy(q).cuda(), torch.from_numpy(cand).cuda(), k, flags, float("inf"))
        ref = oracle.rerank_search(base, cand, q, k, metric)
        _check(ref, (D.cpu().numpy(), I.cpu().numpy()), atol=0.0 if metric == "l2" else 1e-5)
        assert (I.cpu().numpy()[9] == -1).all() and np.isinf(D.cpu().numpy()[9]).all()


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_ivf_scan_matches_oracle_given_same_centroids(eng, metric):
    base, q = _data(30000, 50, 200, seed=31)
    rng = np.random.RandomState(2)
    cent = base[rng.permutation(30000)[:64]].copy()
    if metric == "cosine":
        cent = oracle.safe_normalize(cent)
    ivf = eng.IVFShard(base, cent, metric, "cuda")
    b, qq = (oracle.safe_normalize(base), oracle.safe_normalize(q)) if metric == "cosine" else (base, q)
    m = "l2" if metric == "l2" else "ip"
    assign = oracle.ivf_assign(b, cent, m)
    got_assign = ivf.assign.cpu().numpy()
    assert (got_assign == assign).mean() > 0.9999   # fp ties between two centroids are the only freedom
    assert int(ivf.counts.sum().item()) == 30000
    for nprobe in (1, 8, 12, 24, 64):         # 1 / 1 / 2 / 4 / 8 warps per query (about 4 096 expected rows per warp)
        scanned = torch.zeros(1, dtype=torch.int64, device="cuda")
        D, I = ivf.search(torch.from_numpy(q.copy()).cuda(), 100, nprobe, 0,
                          oracle.FLT_MAX if m == "l2" else -oracle.FLT_MAX, scanned)
        ref_d, ref_i, probes = oracle.ivf_flat_search(b, cent, got_assign, qq, 100, nprobe, m)
        _check((ref_d, ref_i), (D.cpu().numpy(), I.cpu().numpy()), atol=0.0 if m == "l2" else 1e-5)
        counts = np.bincount(got_assign, minlength=64)
        assert int(scanned.item()) == int(counts[ivf.last_probes.cpu().numpy()].sum())
    # nprobe == nlist is exact search
    ref = oracle.faiss_flat_search(b, qq, 100, m)
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()), atol=0.0 if m == "l2" else 1e-5)


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("n,d,nlist,k", [(30000, 50, 64, 100), (9000, 128, 32, 10), (5000, 7, 16, 200)])
def test_ivf_sq8_matches_oracle_given_the_same_index(eng, metric, n, d, nlist, k):
    """"IVF<n>,SQ8" (8-bit scalar-quantised residuals, by_residual): ranges and codes against the oracle's restatement of
    the FAISS codec, then the list scan against the oracle GIVEN the device's centroids, assignments, ranges and codes,
    for 1 / 2 / 4 / 8 warps per query."""
    base, q = _data(n, d, 150, seed=n + d)
    rng = np.random.RandomState(5)
    cent = base[rng.permutation(n)[:nlist]].copy()
    if metric == "cosine":
        cent = oracle.safe_normalize(cent)
    shard = eng.IVFSQ8Shard(base, cent, metric, "cuda")
    b, qq = (oracle.safe_normalize(base), oracle.safe_normalize(q)) if metric == "cosine" else (base, q)
    m = "l2" if metric == "l2" else "ip"
    assign = shard.assign.cpu().numpy()
    resid = b - cent[assign]
    vmin, vdiff = oracle.sq8_train(resid)
    if metric == "cosine":        # rows are normalised on the device: 1 ulp away from NumPy's normalisation here and there
        np.testing.assert_allclose(shard.vmin.cpu().numpy(), vmin, atol=5e-7)
        np.testing.assert_allclose(shard.vdiff.cpu().numpy(), vdiff, atol=1e-6)
    else:
        np.testing.assert_array_equal(shard.vmin.cpu().numpy(), vmin)
        np.testing.assert_allclose(shard.vdiff.cpu().numpy(), vdiff, rtol=1e-6)
    codes = shard.codes_by_row()
    ref_codes = oracle.sq8_encode(resid, shard.vmin.cpu().numpy(), shard.vdiff.cpu().numpy())
    diff = codes.astype(np.int32) - ref_codes.astype(np.int32)
    assert np.abs(diff).max() <= 1 and (diff != 0).mean() < 5e-3        # a component within rounding of a code boundary may flip
    for nprobe in (1, 8, nlist // 2, nlist):
        D, I = shard.search(torch.from_numpy(q.copy()).cuda(), k, nprobe, 0, oracle.FLT_MAX if m == "l2" else -oracle.FLT_MAX)
        ref_d, ref_i = oracle.ivf_sq8_search(codes, cent, assign, shard.vmin.cpu().numpy(), shard.vdiff.cpu().numpy(), qq, k, nprobe, m)
        scale = float(np.linalg.norm(b, axis=1).max() * np.linalg.norm(qq, axis=1).max())
        _check((ref_d, ref_i), (D.cpu().numpy(), I.cpu().numpy()), rtol=2e-5, atol=2e-6 * scale)
    # quantisation costs little recall against the exact search (nprobe = nlist scans every row)
    exact = oracle.faiss_flat_search(b, qq, min(k, 10), m)
    assert oracle.recall_at_k(exact[1], I.cpu().numpy()[:, : min(k, 10)], min(k, 10)) > 0.85


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("n,d,m,nlist,k", [(20000, 64, 64, 32, 100), (12000, 48, 8, 16, 10), (9000, 50, 50, 0, 20), (6000, 96, 16, 0, 200)])
def test_pq_matches_oracle_given_the_same_index(eng, metric, n, d, m, nlist, k):
    """"IVF<n>,PQ<m>" (nlist > 0) and "PQ<m>" (nlist = 0): codes against the oracle's encoder given the device's
    codebooks, then the look-up-table scan against the oracle scoring the RECONSTRUCTED vectors, given the device's
    centroids, assignments, codebooks and codes."""
    base, q = _data(n, d, 120, seed=n + d + m)
    rng = np.random.RandomState(6)
    cent = None
    if nlist:
        cent = base[rng.permutation(n)[:nlist]].copy()
        if metric == "cosine":
            cent = oracle.safe_normalize(cent)
    shard = eng.IVFPQShard(base, cent, m, metric, "cuda", niter=4)
    b, qq = (oracle.safe_normalize(base), oracle.safe_normalize(q)) if metric == "cosine" else (base, q)
    mm = "l2" if metric == "l2" else "ip"
    books = shard.codebooks.cpu().numpy()
    assert books.shape == (m, 256, d // m) and np.isfinite(books).all()
    assign = shard.assign.cpu().numpy().astype(np.int64)
    resid = b - cent[assign] if nlist else b
    codes = shard.codes.cpu().numpy()
    ref_codes = oracle.pq_encode(resid, books)
    differ = codes != ref_codes
    assert differ.mean() < 2e-3, differ.mean()              # near-ties between two sub-centroids (fp32 vs fp64) only
    if differ.any():                                        # ... and where they differ the two choices are equally good
        rr, ss = np.nonzero(differ)
        dsub = d // m
        sub = np.stack([resid[r, s_ * dsub:(s_ + 1) * dsub] for r, s_ in zip(rr, ss)]).astype(np.float64)
        da = ((sub - books[ss, codes[rr, ss]].astype(np.float64)) ** 2).sum(1)
        db = ((sub - books[ss, ref_codes[rr, ss]].astype(np.float64)) ** 2).sum(1)
        np.testing.assert_allclose(da, db, rtol=1e-4, atol=1e-6)
    for nprobe in ((1, 4, nlist) if nlist else (1,)):
        D, I = shard.search(torch.from_numpy(q.copy()).cuda(), k, nprobe, 0, oracle.FLT_MAX if mm == "l2" else -oracle.FLT_MAX)
        ref_d, ref_i = oracle.ivf_pq_search(codes, cent, assign, books, qq, k, nprobe, mm)
        scale = float(np.linalg.norm(b, axis=1).max() * np.linalg.norm(qq, axis=1).max())
        _check((ref_d, ref_i), (D.cpu().numpy(), I.cpu().numpy()), rtol=2e-5, atol=2e-6 * scale)
    if not nlist:
        # "PQ<m>" with a query batch: the flat tensor-pipe scan over the DECODED rows must return what the table scan returns
        assert shard._flat is None                       # 120 queries: the table scan ran above
        shard.decoded_scan = "always"
        D2, I2 = shard.search(torch.from_numpy(q.copy()).cuda(), k, 1, 0, oracle.FLT_MAX if mm == "l2" else -oracle.FLT_MAX)
        assert shard._flat is not None and shard._flat.n == n
        _check((ref_d, ref_i), (D2.cpu().numpy(), I2.cpu().numpy()), rtol=2e-5, atol=2e-6 * scale)
        decoded = oracle.pq_decode(codes, books)
        got = shard._flat.hi[:n, :d].cpu().numpy().astype(np.float64) + shard._flat.lo[:n, :d].cpu().numpy()
        np.testing.assert_array_equal(got.astype(np.float32), decoded.astype(np.float32))       # operands = the decoded rows, exactly
        shard.decoded_max_bytes = 1024                   # too large to keep: back to the table scan
        shard._flat = None
        D3, I3 = shard.search(torch.from_numpy(q.copy()).cuda(), k, 1, 0, oracle.FLT_MAX if mm == "l2" else -oracle.FLT_MAX)
        assert shard._flat is None
        np.testing.assert_array_equal(I3.cpu().numpy(), I.cpu().numpy())


def test_row_utilities(eng):
    base, _ = _data(1000, 50, 1, seed=4)
    base[3] = 0
    x = torch.from_numpy(base).cuda()
    np.testing.assert_allclose(eng.row_norms(x).cpu().numpy(), (base.astype(np.float64) ** 2).sum(1), rtol=1e-6)
    y = eng.normalize_rows_(x.clone()).cpu().numpy()
    np.testing.assert_allclose(y, oracle.safe_normalize(base), rtol=1e-6, atol=1e-7)
    assert (y[3] == 0).all()


@pytest.mark.parametrize("n,d,nbits,nq,k", [(5000, 50, 256, 70, 100), (100000, 32, 256, 33, 1000), (3000, 16, 64, 9, 3000),
                                            (70000, 24, 320, 17, 500), (400, 8, 1000, 5, 64)])
def test_lsh_codes_and_hamming_topk_match_oracle(eng, n, d, nbits, nq, k):
    """Sign codes equal the oracle's except on bits whose projection is within rounding of 0; the
    Hamming top-k over the device's own codes is exactly the oracle's (distance, id) order."""
    base, q = _data(n, d, nq, seed=n + nbits)
    proj = np.random.RandomState(1234).normal(size=(nbits, d)).astype(np.float32)
    shard = eng.HammingShard(base, proj, "cuda")
    native = (nbits + 127) // 128 * 4
    codes = shard.codes.cpu().numpy().view(np.uint32)
    ref_codes, dots = oracle.lsh_sign_codes(base, proj)
    assert (codes[:, native:] == 0).all()
    diff = np.unpackbits((codes[:, :native] ^ ref_codes).view(np.uint8), axis=1, bitorder="little")[:, :nbits]
    scale = np.abs(base).astype(np.float64) @ np.abs(proj).astype(np.float64).T
    assert (np.abs(dots[diff.astype(bool)]) <= 1e-6 * scale[diff.astype(bool)]).all()
    assert diff.mean() < 1e-4
    qd = torch.from_numpy(q).cuda()
    D, I = shard.search(qd, k)
    qcodes = shard.encode(qd).cpu().numpy().view(np.uint32)
    ref_d, ref_i = oracle.hamming_topk(codes, qcodes, k)
    np.testing.assert_array_equal(I.cpu().numpy(), ref_i)
    np.testing.assert_array_equal(D.cpu().numpy(), ref_d)


def test_faiss_lsh_pipeline_recall(eng):
    """FaissLSHIndexer + FaissSearcher rerank through the public classes: rerank over Hamming
    candidates equals the oracle's rerank over the same candidates; recall rises with the budget."""
    import vectordb_retrieval_b200.algorithms as A
    rng = np.random.RandomState(8)
    centers = rng.randn(40, 32).astype(np.float32)
    base = (centers[rng.randint(0, 40, 30000)] + 0.5 * rng.randn(30000, 32)).astype(np.float32)
    queries = (centers[rng.randint(0, 40, 64)] + 0.5 * rng.randn(64, 32)).astype(np.float32)
    gt = oracle.linear_search(base, queries, 10, "l2")[1]
    recalls = []
    for mult in (2.0, 32.0):
        algo = A.get_algorithm_instance("Composite", 32, name="flsh", metric="l2",
                                        indexer={"type": "FaissLSHIndexer", "num_bits": 256},
                                        searcher={"type": "FaissSearcher", "lsh_candidate_multiplier": mult})
        algo.build_index(base)
        dist, idx = algo.batch_search(queries, 10)
        index = algo.index_artifact.data
        _, cand = index.search(queries, int(10 * mult))
        ref = oracle.rerank_search(base, cand, queries, 10, "l2")
        _check(ref, (dist, idx))
        recalls.append(oracle.recall_at_k(gt, idx, 10))
    assert recalls[1] > recalls[0] and recalls[1] > 0.6


def test_full_size_c2_properties(eng):
    """BASELINE.json configs[1] at full size (1M x 128, 10k queries, k=100, L2) through size-independent
    properties: sorted distances, unique in-range ids, shard-count invariance (3 row shards + merge
    kernel == 1 shard, bit for bit), and oracle parity on a sample of the queries."""
    n, d, nq, k = 1_000_000, 128, 10_000, 100
    g = torch.Generator(device="cuda").manual_seed(42)
    base = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
    q = torch.randn((nq, d), generator=g, device="cuda", dtype=torch.float32)
    D, I = eng.FlatShard(base, "l2", "cuda").search(q.clone(), k)
    assert bool((D[:, 1:] >= D[:, :-1]).all()) and int(I.min()) >= 0 and int(I.max()) < n
    srt = torch.sort(I, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "duplicate ids inside a result row"
    # shard-count invariance
    bounds = [0, 333_312, 700_000, n]
    ds, is_ = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        dd, ii = eng.FlatShard(base[a:b], "l2", "cuda", id_offset=a).search(q.clone(), k)
        ds.append(dd), is_.append(ii)
    Dm, Im = eng.merge_topk(torch.stack(ds), torch.stack(is_))
    assert torch.equal(Im, I) and torch.equal(Dm, D)
    # oracle on a query sample (fp64 brute force on the host)
    pick = np.random.RandomState(0).choice(nq, 24, replace=False)
    ref = oracle.faiss_flat_search(base.cpu().numpy(), q[pick].cpu().numpy(), k, "l2")
    _check(ref, (D[pick].cpu().numpy(), I[pick].cpu().numpy()))
    assert oracle.recall_at_k(ref[1], I[pick].cpu().numpy(), k) == 1.0


# ---- seeded bounds: pre-pass guess, verification, redo (flat.cu) -------------------------------
def _redo_count():
    import ctypes
    from vectordb_retrieval_b200 import _lib
    out = ctypes.c_uint64(0)
    assert _lib.load().vdb_debug_redo_queries(ctypes.byref(out)) == 0
    return int(out.value)


@pytest.mark.parametrize("metric,n,d,nq,k", [("l2", 140000, 32, 300, 100), ("ip", 40000, 48, 70, 10), ("l2", 300000, 16, 40, 200)])
def test_seeded_bounds_do_not_change_results(eng, metric, n, d, nq, k):
    """Shards large enough for the seeding pre-pass: same answer with the pre-pass (default), without
    it (mode 7) and with every query forced through the redo pass (mode 6); all equal the oracle."""
    from vectordb_retrieval_b200 import _lib
    lib = _lib.load()
    base, q = _data(n, d, nq, seed=n + k)
    shard = eng.FlatShard(base, metric, "cuda")
    pad = oracle.FLT_MAX if metric == "l2" else -oracle.FLT_MAX
    ref = oracle.faiss_flat_search(base, q, k, metric)
    atol = 0.0 if metric == "l2" else 1e-5 * float(np.linalg.norm(base, axis=1).max() * np.linalg.norm(q, axis=1).max())
    outs = {}
    _redo_count()
    try:
        for mode in (0, 7, 6):
            lib.vdb_set_debug_mode(mode)
            D, I = shard.search(torch.from_numpy(q).cuda(), k, 0, pad)
            torch.cuda.synchronize()
            outs[mode] = (D.cpu().numpy(), I.cpu().numpy())
            redo = _redo_count()
            if mode == 6:
                assert redo == nq, f"forced redo touched {redo} of {nq} queries"
            if mode == 7:
                assert redo == 0
    finally:
        lib.vdb_set_debug_mode(0)
    for mode, got in outs.items():
        _check(ref, got, atol=atol)
        np.testing.assert_array_equal(got[1], outs[0][1])
        np.testing.assert_array_equal(got[0], outs[0][0])


def test_seeded_bounds_misleading_sample_takes_the_redo_pass(eng):
    """Adversarial row order: the sampled tiles hold only rows close to the queries, so the guessed
    bound is far too tight for the rest of the base.  The verify kernel must catch those queries and
    the redo pass must still return the exact answer."""
    n, d, nq, k = 262144, 32, 64, 100          # 1024 tiles; k' = 128, rank 16, margin 4 -> 32 sampled tiles, stride 32
    lib = eng._lib.load()
    assert lib.vdb_flat_set_seeding(64, 16) == 0 and lib.vdb_flat_set_seeding_margin(4) == 0   # the layout below is built for this sample
    rng = np.random.RandomState(3)
    q = rng.randn(nq, d).astype(np.float32)
    # (a) guesses far too loose: the sampled tiles hold only far rows -> slower, never wrong, no redo
    base = rng.randn(n, d).astype(np.float32)
    for t in range(0, 1024, 64):
        base[t * 256:(t + 1) * 256] += 50.0
    shard = eng.FlatShard(base, "l2", "cuda")
    _redo_count()
    D, I = shard.search(torch.from_numpy(q).cuda(), k)
    torch.cuda.synchronize()
    assert _redo_count() == 0
    ref = oracle.faiss_flat_search(base, q, k, "l2")
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()))
    # (b) guesses far too tight: the only near rows of the base sit in the sampled tiles (5 per tile),
    # so the 16th smallest sampled key has just 15 rows of the whole base below it
    base2 = (rng.randn(n, d) * 0.5 + 20.0).astype(np.float32)
    base2[:, 0] += np.linspace(0.0, 30.0, n, dtype=np.float32)      # far rows drift away with the row number
    for t in range(0, 1024, 64):
        base2[t * 256:t * 256 + 5] = rng.randn(5, d).astype(np.float32)
    shard = eng.FlatShard(base2, "l2", "cuda")
    D, I = shard.search(torch.from_numpy(q).cuda(), k)
    torch.cuda.synchronize()
    redo = _redo_count()
    ref = oracle.faiss_flat_search(base2, q, k, "l2")
    _check(ref, (D.cpu().numpy(), I.cpu().numpy()))
    assert lib.vdb_flat_set_seeding(64, 0) == 0 and lib.vdb_flat_set_seeding_margin(0) == 0
    assert redo > 0, "expected the misleading sample to force at least one query through the redo pass"


def test_seeded_bounds_one_cta_variant_and_streamed_query_tile(eng):
    """The seeding pre-pass has its own kernel instances: cover cta_group::1 and the streamed query
    tile (d > 128) as well, against the oracle."""
    from vectordb_retrieval_b200 import _lib
    # the last two shapes are small enough for the rank-32 pre-pass (32 kept minima) on the one-CTA instances as well
    for impl, n, d, nq, k in (("tcgen05_1cta", 70000, 24, 90, 10), ("tcgen05", 40000, 160, 50, 10), ("tcgen05_1cta", 40000, 160, 50, 10),
                              ("tcgen05_1cta", 36000, 24, 90, 100), ("tcgen05_1cta", 36000, 160, 50, 100)):
        base, q = _data(n, d, nq, seed=n + d)
        shard = eng.FlatShard(base, "l2", "cuda")
        _redo_count()
        D, I = shard.search(torch.from_numpy(q).cuda(), k, 0, oracle.FLT_MAX, _lib.IMPL_NAMES[impl])
        torch.cuda.synchronize()
        assert _redo_count() == 0
        _check(oracle.faiss_flat_search(base, q, k, "l2"), (D.cpu().numpy(), I.cpu().numpy()))


@pytest.mark.parametrize("nbits,d,n,nq,k", [(256, 50, 70000, 300, 700), (128, 32, 66000, 260, 64), (200, 24, 90000, 257, 1500),
                                            (255, 40, 70001, 256, 300), (64, 16, 69999, 300, 40000)])
def test_hamming_topk_tensor_pipe_matches_popc_path(eng, nbits, d, n, nq, k):
    """The bf16 +-1 contraction path must return exactly what the popc path returns: same distances,
    same ids, (distance, id) order - including queries whose sampled bound is too small (duplicated
    base rows put far more than k rows at distance 0 for some queries)."""
    rng = np.random.RandomState(nbits + n)
    base = rng.randn(n, d).astype(np.float32)
    base[5000:5000 + 2 * k] = base[17]                     # a big tie group
    q = rng.randn(nq, d).astype(np.float32)
    q[3] = base[17]
    proj = rng.randn(nbits, d).astype(np.float32)
    shard = eng.HammingShard(base, proj, "cuda")
    qd = torch.from_numpy(q).cuda()
    shard.tensor_pipe = False
    d0, i0 = shard.search(qd.clone(), k)
    shard.tensor_pipe = True
    for f16 in (True, False):        # fp16 operands + accumulators (packed epilogue) / bf16 operands + fp32 accumulators
        shard.tc_accumulate_f16 = f16
        d1, i1 = shard.search(qd.clone(), k)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d0.cpu().numpy(), d1.cpu().numpy(), err_msg=f"f16={f16}")
        np.testing.assert_array_equal(i0.cpu().numpy(), i1.cpu().numpy(), err_msg=f"f16={f16}")
        assert (i1.cpu().numpy()[3, : 2 * k + 1] >= 0).all()
