"""CPU oracle for the scan + top-k hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``vectordb_retrieval_b200``) never imports it and has no CPU fallback.

Every function restates, in NumPy, the semantics of one piece of the reference
(Human-Augment-Analytics/vectordb-retrieval); citations are ``path:line`` relative
to the reference root.

Pinning status
--------------
* NumPy paths (LinearSearcher, FaissSearcher LSH-rerank, Python LSH, recall_at_k):
  PINNED - ``oracle/gen_golden.py`` imports the unmodified reference classes (behind
  throw-away ``faiss`` / ``matplotlib`` import stubs) and writes their outputs to
  ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this module against
  those files, the reference's own KATs (tests/test_composite_algorithm.py:29-226)
  and the published LSH recall 0.31914062499999996
  (benchmark_results/benchmark_20260305_070532/random/lsh_results.json:46).
* FAISS paths (ExactSearch/IndexFlat values, IVF k-means, IndexLSH codes, IVF-SQ8 ranges / codes, PQ codebooks / codes):
  PARITY UNPINNED - faiss-cpu (requirements.txt:9, ``>=1.7.4``, no lock) is not in the
  reference tree and not installed.  ``faiss_flat_search`` / ``ivf_flat_search`` restate
  FAISS's documented conventions (squared L2 ascending, raw inner product descending,
  int64 labels, -1 padding) and are anchored on the reference's call sites
  (src/algorithms/exact_search.py:23,38-39,78; src/algorithms/modular.py:277-286,536-548).
"""
from __future__ import annotations

import math
from collections import Counter, defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

FLT_MAX = np.finfo(np.float32).max


# --------------------------------------------------------------------------- helpers
def safe_normalize(matrix: np.ndarray) -> np.ndarray:
    """Row L2-normalisation, zero rows stay zero (src/algorithms/modular.py:109-111,
    src/algorithms/lsh.py:13-16)."""
    norms = np.linalg.norm(matrix, axis=1, keepdims=True)
    out = np.zeros_like(matrix)
    np.divide(matrix, norms, out=out, where=norms > 0)
    return out


def _as_f32(x: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32)


def _sorted_topk(values: np.ndarray, limit: int) -> Tuple[np.ndarray, np.ndarray]:
    """Ascending top-``limit`` of each row with the deterministic (value, index) order.

    The reference uses argpartition + argsort (modular.py:349-354), whose order among
    exactly equal values is unspecified; (value, index) is one member of that class and
    is what the CUDA path produces."""
    n = values.shape[1]
    if limit >= n:
        part = np.broadcast_to(np.arange(n), values.shape).copy()
    else:
        part = np.argpartition(values, limit - 1, axis=1)[:, :limit]
        # argpartition may cut a tie group at the boundary: pull in every index whose
        # value equals the k-th value and keep the lowest indices.
        kth = np.take_along_axis(values, part, axis=1).max(axis=1)
        for r in np.nonzero((values <= kth[:, None]).sum(axis=1) > limit)[0]:
            cand = np.nonzero(values[r] <= kth[r])[0]
            order = np.lexsort((cand, values[r, cand]))[:limit]
            part[r] = cand[order]
    pv = np.take_along_axis(values, part, axis=1)
    order = np.lexsort((part, pv), axis=1)
    idx = np.take_along_axis(part, order, axis=1)
    val = np.take_along_axis(pv, order, axis=1)
    return val, idx


def _pad(dist: np.ndarray, idx: np.ndarray, k: int, pad_value: float) -> Tuple[np.ndarray, np.ndarray]:
    limit = dist.shape[1]
    if limit < k:
        dist = np.pad(dist, ((0, 0), (0, k - limit)), constant_values=pad_value)
        idx = np.pad(idx, ((0, 0), (0, k - limit)), constant_values=-1)
    return dist.astype(np.float32), idx.astype(np.int64)


# --------------------------------------------------------------------------- exact: LinearSearcher
def linear_search(base: np.ndarray, queries: np.ndarray, k: int, metric: str = "l2",
                  query_block: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    """``LinearSearcher.batch_search`` (src/algorithms/modular.py:336-387).

    l2     : difference-form squared distance in fp32, top-k, **sqrt** (modular.py:343-355)
    ip     : ``Q @ V.T``, distance = -score (modular.py:368,381)
    cosine : both sides ``safe_normalize``-d then as ip (modular.py:322-324,365-366)
    k > N pads with (+inf, -1) (modular.py:357-359,382-384).
    The reference materialises an nq x N x d temporary; this restatement blocks over
    queries (same arithmetic per element, bounded memory)."""
    base = _as_f32(base)
    q = _as_f32(np.atleast_2d(queries))
    n = base.shape[0]
    if n == 0:
        raise RuntimeError("LinearSearcher cannot operate on empty index")
    limit = min(k, n)
    out_d = np.empty((q.shape[0], limit), dtype=np.float32)
    out_i = np.empty((q.shape[0], limit), dtype=np.int64)
    if metric == "l2":
        for s in range(0, q.shape[0], query_block):
            blk = q[s:s + query_block]
            diffs = base[None, :, :] - blk[:, None, :]
            sq = np.sum(diffs ** 2, axis=2)
            val, idx = _sorted_topk(sq, limit)
            out_d[s:s + query_block] = np.sqrt(val)
            out_i[s:s + query_block] = idx
    elif metric in ("cosine", "ip"):
        if metric == "cosine":
            vb, vq = safe_normalize(base), safe_normalize(q)
        else:
            vb, vq = base, q
        for s in range(0, q.shape[0], max(query_block, 256)):
            scores = vq[s:s + max(query_block, 256)] @ vb.T
            val, idx = _sorted_topk(-scores, limit)
            out_d[s:s + max(query_block, 256)] = val
            out_i[s:s + max(query_block, 256)] = idx
    else:
        raise ValueError(f"Unsupported metric '{metric}' for LinearSearcher")
    return _pad(out_d, out_i, k, np.inf)


# --------------------------------------------------------------------------- exact: faiss.IndexFlat
def faiss_flat_search(base: np.ndarray, queries: np.ndarray, k: int, metric: str = "l2",
                      base_block: int = 65536, query_block: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """``ExactSearch.batch_search`` -> ``faiss.IndexFlat.search``
    (src/algorithms/exact_search.py:23,38-39,76-78).  [FAISS-upstream, parity unpinned]

    metric == 'l2'  : squared L2, ascending.
    anything else   : raw inner product on un-normalised vectors, descending (not negated)
                      - this is the quirk at exact_search.py:23.
    Missing results: label -1, distance +FLT_MAX (l2) / -FLT_MAX (ip).
    Scores are formed FAISS-style, ``|x|^2 + |q|^2 - 2 q.x`` with one sgemm per block and
    negatives clamped to 0, but accumulated in fp64 so this is the *exact* value the fp32
    implementations approximate."""
    base = _as_f32(base)
    q = _as_f32(np.atleast_2d(queries))
    n, nq = base.shape[0], q.shape[0]
    limit = min(k, n)
    l2 = metric == "l2"
    best_v = np.full((nq, 0), 0.0)
    best_i = np.full((nq, 0), 0, dtype=np.int64)
    bn = (base.astype(np.float64) ** 2).sum(axis=1) if l2 else None
    for qs in range(0, nq, query_block):
        qb = q[qs:qs + query_block].astype(np.float64)
        qn = (qb ** 2).sum(axis=1) if l2 else None
        cur_v = np.empty((qb.shape[0], 0))
        cur_i = np.empty((qb.shape[0], 0), dtype=np.int64)
        for bs in range(0, n, base_block):
            xb = base[bs:bs + base_block].astype(np.float64)
            ip = qb @ xb.T
            if l2:
                key = np.maximum(qn[:, None] + bn[None, bs:bs + base_block] - 2.0 * ip, 0.0)
            else:
                key = -ip
            v, i = _sorted_topk(key, min(limit, key.shape[1]))
            cur_v = np.concatenate([cur_v, v], axis=1)
            cur_i = np.concatenate([cur_i, i + bs], axis=1)
            if cur_v.shape[1] > limit:
                order = np.lexsort((cur_i, cur_v), axis=1)[:, :limit]
                cur_v = np.take_along_axis(cur_v, order, axis=1)
                cur_i = np.take_along_axis(cur_i, order, axis=1)
        order = np.lexsort((cur_i, cur_v), axis=1)[:, :limit]
        cur_v = np.take_along_axis(cur_v, order, axis=1)
        cur_i = np.take_along_axis(cur_i, order, axis=1)
        best_v = cur_v if qs == 0 else np.concatenate([best_v, cur_v], axis=0)
        best_i = cur_i if qs == 0 else np.concatenate([best_i, cur_i], axis=0)
    dist = best_v if l2 else -best_v
    return _pad(dist, best_i, k, FLT_MAX if l2 else -FLT_MAX)


def faiss_flat_search_blas(base: np.ndarray, queries: np.ndarray, k: int, metric: str = "l2",
                           base_block: int = 131072, query_block: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
    """Speed-oriented variant of :func:`faiss_flat_search` used ONLY as ``bench.py``'s CPU
    baseline: fp32 OpenBLAS sgemm per (query block x base block) + argpartition, all host
    threads - the FAISS decomposition for nq >= 20 [FAISS-upstream]."""
    base = _as_f32(base)
    q = _as_f32(np.atleast_2d(queries))
    n, nq = base.shape[0], q.shape[0]
    limit = min(k, n)
    l2 = metric == "l2"
    bn = np.einsum("ij,ij->i", base, base) if l2 else None
    out_v = np.empty((nq, limit), dtype=np.float32)
    out_i = np.empty((nq, limit), dtype=np.int64)
    for qs in range(0, nq, query_block):
        qb = q[qs:qs + query_block]
        qn = np.einsum("ij,ij->i", qb, qb) if l2 else None
        vs, is_ = [], []
        for bs in range(0, n, base_block):
            ip = qb @ base[bs:bs + base_block].T
            if l2:
                ip *= -2.0
                ip += bn[None, bs:bs + base_block]
                ip += qn[:, None]
                np.maximum(ip, 0.0, out=ip)
            else:
                np.negative(ip, out=ip)
            kk = min(limit, ip.shape[1])
            part = np.argpartition(ip, kk - 1, axis=1)[:, :kk] if kk < ip.shape[1] else \
                np.broadcast_to(np.arange(ip.shape[1]), ip.shape)
            vs.append(np.take_along_axis(ip, part, axis=1))
            is_.append(part + bs)
        v = np.concatenate(vs, axis=1)
        i = np.concatenate(is_, axis=1)
        order = np.argsort(v, axis=1, kind="stable")[:, :limit]
        out_v[qs:qs + query_block] = np.take_along_axis(v, order, axis=1)
        out_i[qs:qs + query_block] = np.take_along_axis(i, order, axis=1)
    dist = out_v if l2 else -out_v
    return _pad(dist, out_i, k, FLT_MAX if l2 else -FLT_MAX)


def _flat_blas_worker(base: np.ndarray, bn: Optional[np.ndarray], qb: np.ndarray, limit: int, l2: bool,
                      base_block: int) -> Tuple[np.ndarray, np.ndarray]:
    """One query sub-block against the whole base: sgemm per base block, then FAISS's heap logic in bulk form -
    a key enters only if it beats the query's current k-th best (``thr``), and the few that do are merged
    with one small sort per block.  Keys are |x|^2 - 2 q.x (the |q|^2 term cannot change a query's order)."""
    m, n = qb.shape[0], base.shape[0]
    cand_v = np.full((m, limit), np.inf, dtype=np.float32)
    cand_i = np.full((m, limit), -1, dtype=np.int64)
    thr = np.full(m, np.inf, dtype=np.float32)
    rows_k = np.repeat(np.arange(m), limit)
    for bs in range(0, n, base_block):
        key = qb @ base[bs:bs + base_block].T
        if l2:
            key *= -2.0
            key += bn[None, bs:bs + base_block]
        else:
            np.negative(key, out=key)
        width = key.shape[1]
        if bs == 0 and width > limit:                            # no bound yet: plain selection on the first block
            part = np.argpartition(key, limit - 1, axis=1)[:, :limit]
            r, c = rows_k, part.ravel()
            flat = r * width + c
        else:
            flat = np.flatnonzero(key < thr[:, None])            # (2-D np.nonzero is 3x slower than this + divmod)
            if flat.size == 0:
                continue
            r, c = np.divmod(flat, width)
        v_all = np.concatenate([cand_v.ravel(), key.ravel()[flat]])
        i_all = np.concatenate([cand_i.ravel(), c.astype(np.int64) + bs])
        r_all = np.concatenate([rows_k, r])
        order = np.lexsort((v_all, r_all))                       # by query, then by key
        r_s = r_all[order]
        start = np.searchsorted(r_s, np.arange(m))
        rank = np.arange(r_s.size) - start[r_s]
        keep = order[rank < limit]                               # every query holds >= limit entries
        cand_v = v_all[keep].reshape(m, limit)
        cand_i = i_all[keep].reshape(m, limit)
        thr = cand_v[:, -1].copy()
    return cand_v, cand_i


def faiss_flat_search_threaded(base: np.ndarray, queries: np.ndarray, k: int, metric: str = "l2", threads: int = 0,
                               query_block: int = 0, base_block: int = 8192) -> Tuple[np.ndarray, np.ndarray]:
    """``bench.py``'s CPU baseline / reference arm (a port: FAISS is not installable): the flat search as FAISS
    runs it for nq >= 20 [FAISS-upstream] - queries split over ``threads`` host threads (OpenMP over queries in
    FAISS; a thread pool here, each thread calling single-threaded OpenBLAS sgemm on cache-sized blocks),
    ``|x|^2 - 2 q.x`` per block, a running k-th-best threshold per query instead of a full selection.
    Same results as :func:`faiss_flat_search` up to fp32 rounding.  ``threads`` = 0 -> os.cpu_count()."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    base = _as_f32(base)
    q = _as_f32(np.atleast_2d(queries))
    n, nq = base.shape[0], q.shape[0]
    limit = min(k, n)
    l2 = metric == "l2"
    threads = threads or (os.cpu_count() or 1)
    if query_block <= 0:                                         # >= 2 blocks per thread, 32..128 queries each
        query_block = int(min(128, max(32, nq // (2 * threads))))
    bn = np.einsum("ij,ij->i", base, base) if l2 else None
    blocks = [(s, min(nq, s + query_block)) for s in range(0, nq, query_block)]
    out_v = np.empty((nq, limit), dtype=np.float32)
    out_i = np.empty((nq, limit), dtype=np.int64)

    def run(span):
        v, i = _flat_blas_worker(base, bn, q[span[0]:span[1]], limit, l2, base_block)
        out_v[span[0]:span[1]], out_i[span[0]:span[1]] = v, i

    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1, user_api="blas")
    except Exception:      # noqa: BLE001 - no threadpoolctl: BLAS keeps its own threading
        limiter = None
    try:
        if threads == 1:
            for span in blocks:
                run(span)
        else:
            with ThreadPoolExecutor(max_workers=threads) as ex:
                list(ex.map(run, blocks))
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    if l2:
        out_v += np.einsum("ij,ij->i", q, q)[:, None]
        np.maximum(out_v, 0.0, out=out_v)
        dist = out_v
    else:
        dist = -out_v
    return _pad(dist, out_i, k, FLT_MAX if l2 else -FLT_MAX)


# --------------------------------------------------------------------------- rerank (FAISS-LSH searcher)
def candidate_budget(k: int, multiplier: float, max_candidates: Optional[int], num_db: int) -> int:
    """Candidate count rule of ``FaissSearcher._batch_search_lsh_rerank``
    (src/algorithms/modular.py:463-468)."""
    c = max(k, 1)
    if multiplier > 1.0:
        c = int(max(c, k * multiplier))
    if max_candidates is not None:
        c = min(c, max_candidates)
    return min(c, num_db)


def rerank_search(base: np.ndarray, candidates: np.ndarray, queries: np.ndarray, k: int,
                  metric: str = "l2") -> Tuple[np.ndarray, np.ndarray]:
    """Exact re-scoring of per-query candidate ids (src/algorithms/modular.py:483-532).

    ``candidates`` is [nq, C] int64 with -1 = invalid.  l2 -> sqrt of difference-form
    squared distance; ip/cosine -> -score (cosine: ``base`` and ``queries`` already
    normalised by the caller, modular.py:434-435,447-448).  Rows with fewer than k valid
    candidates are padded with (+inf, -1) (modular.py:480-481).  A row with *no* valid
    candidate falls back to the raw index order in the reference (modular.py:486-493);
    here it stays fully padded - the CUDA path documents the same."""
    base = _as_f32(base)
    q = _as_f32(np.atleast_2d(queries))
    nq = q.shape[0]
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        valid = candidates[r][candidates[r] >= 0]
        if valid.size == 0:
            continue
        vecs = base[valid]
        if metric == "l2":
            key = np.sum((vecs - q[r:r + 1]) ** 2, axis=1)
        elif metric in ("ip", "cosine"):
            key = -(q[r:r + 1] @ vecs.T).ravel()
        else:
            raise ValueError(f"Unsupported metric '{metric}'")
        limit = min(k, key.shape[0])
        order = np.lexsort((valid, key))[:limit]
        vals = key[order]
        out_d[r, :limit] = np.sqrt(vals) if metric == "l2" else vals
        out_i[r, :limit] = valid[order]
    return out_d, out_i


# --------------------------------------------------------------------------- FAISS IndexLSH (sign codes)
def lsh_sign_codes(x: np.ndarray, projection: np.ndarray) -> np.ndarray:
    """Sign bits of a random projection, packed little-endian into uint32 words padded to a
    multiple of 128 bits: bit b = [ x . P[b] >= 0 ].  Restates ``faiss.IndexLSH.add`` /
    ``sa_encode`` as reached from src/algorithms/modular.py:215-216 [FAISS-upstream: IndexLSH
    with the default ``rotate_data`` applies a random d x nbits matrix and thresholds at 0;
    FAISS's own RNG is not reproduced - parity unpinned, the projection is an input here].
    Returns (codes [n, words] uint32, dots [n, nbits] float64) - the dots let a test skip bits
    that sit on the decision boundary."""
    dots = x.astype(np.float64) @ projection.astype(np.float64).T
    bits = dots >= 0
    nbits = projection.shape[0]
    words = (nbits + 127) // 128 * 4
    padded = np.zeros((x.shape[0], words * 32), dtype=bool)
    padded[:, :nbits] = bits
    codes = np.packbits(padded.reshape(x.shape[0], words, 32), axis=2, bitorder="little").view(np.uint32).reshape(x.shape[0], words)
    return codes, dots


def hamming_topk(codes: np.ndarray, qcodes: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """``faiss.IndexLSH.search`` as called at src/algorithms/modular.py:477: the k smallest
    Hamming distances per query [FAISS-upstream; tie order unspecified there].  Deterministic
    (distance, id) order; k > n pads with (+inf, -1)."""
    n, nq = codes.shape[0], qcodes.shape[0]
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    cb = np.unpackbits(codes.view(np.uint8), axis=1)
    for r in range(nq):
        qb = np.unpackbits(qcodes[r:r + 1].view(np.uint8), axis=1)
        dist = (cb != qb).sum(axis=1)
        order = np.lexsort((np.arange(n), dist))[:k]
        out_d[r, :order.size] = dist[order]
        out_i[r, :order.size] = order
    return out_d, out_i


# --------------------------------------------------------------------------- Python LSH (lsh.py)
class LSHTables:
    """Restatement of ``LSHIndexer.build`` (src/algorithms/lsh.py:95-138).

    Draw order matters for reproducing the reference's buckets: one ``RandomState(seed)``,
    first ``normal(size=(T, H, d))`` projections, then (l2 only) ``uniform(0, w, (T, H))``
    offsets (lsh.py:70-76,99-101).  cosine keys: sign bits weighted ``1 << arange(H)`` into
    a uint64 (lsh.py:78-80,102); l2 keys: tuple of ``floor((P v + b) / w)`` int32
    (lsh.py:82-84)."""

    def __init__(self, vectors: np.ndarray, metric: str, num_tables: int, hash_size: int,
                 bucket_width: float, seed: int):
        self.metric, self.T, self.H, self.w = metric, num_tables, hash_size, bucket_width
        d = vectors.shape[1]
        rng = np.random.RandomState(seed)
        self.projections = rng.normal(size=(num_tables, hash_size, d)).astype(np.float32)
        self.offsets = (rng.uniform(0.0, bucket_width, size=(num_tables, hash_size)).astype(np.float32)
                        if metric == "l2" else None)
        self.bit_weights = (1 << np.arange(hash_size, dtype=np.uint64))
        store = vectors.astype(np.float32, copy=True)
        if metric == "cosine":
            store = safe_normalize(store)
        self.vector_store = store
        self.tables: List[Dict[object, List[int]]] = [defaultdict(list) for _ in range(num_tables)]
        for idx in range(store.shape[0]):
            for t, key in enumerate(self.hash(store[idx])):
                self.tables[t][key].append(idx)

    def hash(self, v: np.ndarray) -> List[object]:
        keys: List[object] = []
        for t in range(self.T):
            proj = self.projections[t] @ v
            if self.metric == "cosine":
                bits = (proj >= 0).astype(np.uint64)
                keys.append(int((bits * self.bit_weights).sum()))
            else:
                keys.append(tuple(np.floor((proj + self.offsets[t]) / self.w).astype(np.int32).tolist()))
        return keys


def lsh_candidates(tables: LSHTables, query: np.ndarray, k: int, candidate_multiplier: float,
                   max_candidates: Optional[int], fallback_to_bruteforce: bool) -> np.ndarray:
    """``LSHSearcher._gather_candidates`` + ``_select_candidates`` (lsh.py:219-240):
    vote-count union of the T buckets in ``Counter.most_common()`` order (count descending,
    first-seen order among equal counts), capped at ``max(k, ceil(mult*k))`` unless
    ``max_candidates`` is given; no hit -> all rows (fallback) or nothing."""
    votes: Counter = Counter()
    for t, key in enumerate(tables.hash(query)):
        bucket = tables.tables[t].get(key)
        if bucket:
            votes.update(bucket)
    ordered = [i for i, _ in votes.most_common()]
    if not ordered:
        if fallback_to_bruteforce:
            return np.arange(tables.vector_store.shape[0], dtype=np.int64)
        return np.empty(0, dtype=np.int64)
    cap = max_candidates if max_candidates is not None else max(k, int(math.ceil(candidate_multiplier * k)))
    return np.asarray(ordered[:cap], dtype=np.int64)


def lsh_search(tables: LSHTables, queries: np.ndarray, k: int, candidate_multiplier: float = 4.0,
               max_candidates: Optional[int] = None, fallback_to_bruteforce: bool = True
               ) -> Tuple[np.ndarray, np.ndarray]:
    """``LSHSearcher.batch_search`` (lsh.py:252-298): per query, hash, gather, score with
    cosine distance ``1 - v.q`` or Euclidean ``|v - q|`` (lsh.py:242-250), full argsort,
    top-k, pad (+inf, -1)."""
    q2 = np.atleast_2d(queries)
    out_d = np.full((q2.shape[0], k), np.inf, dtype=np.float32)
    out_i = np.full((q2.shape[0], k), -1, dtype=np.int64)
    for r in range(q2.shape[0]):
        q = q2[r].astype(np.float32, copy=True)
        if tables.metric == "cosine":
            nrm = np.linalg.norm(q)
            q = np.zeros_like(q) if nrm == 0 else q / nrm
        cand = lsh_candidates(tables, q, k, candidate_multiplier, max_candidates, fallback_to_bruteforce)
        if cand.size == 0:
            continue
        vecs = tables.vector_store[cand]
        if tables.metric == "cosine":
            dist = (1.0 - vecs @ q).astype(np.float32)
        else:
            dist = np.linalg.norm(vecs - q[None, :], axis=1).astype(np.float32)
        order = np.argsort(dist, kind="stable")[:k]
        out_d[r, :order.size] = dist[order]
        out_i[r, :order.size] = cand[order]
    return out_d, out_i


# --------------------------------------------------------------------------- IVF-Flat given centroids
def ivf_assign(vectors: np.ndarray, centroids: np.ndarray, metric: str = "l2") -> np.ndarray:
    """Nearest-centroid assignment used by ``IndexIVFFlat.add`` [FAISS-upstream]; the coarse
    quantiser is flat L2 (l2) or flat inner product (ip / normalised cosine), reached from
    src/algorithms/modular.py:279-283 and src/algorithms/approximate_search.py:39-47."""
    _, idx = faiss_flat_search(centroids, vectors, 1, "l2" if metric == "l2" else "ip")
    return idx[:, 0]


def ivf_flat_search(base: np.ndarray, centroids: np.ndarray, assignments: np.ndarray,
                    queries: np.ndarray, k: int, nprobe: int, metric: str = "l2"
                    ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``IndexIVFFlat.search`` [FAISS-upstream] as called from ``FaissSearcher.batch_search``
    (src/algorithms/modular.py:536-548): per query take the ``nprobe`` best centroids, scan
    exactly those inverted lists, keep the k best.  Values follow the FAISS convention
    (squared L2 ascending / raw IP descending, -1 / +-FLT_MAX padding); the wrapper class
    negates for ip/cosine (modular.py:545-546).  Parity is defined *given identical
    centroids and assignments* (FAISS k-means is not reproducible here).
    Returns (D, I, probes)."""
    base = _as_f32(base)
    q = _as_f32(np.atleast_2d(queries))
    l2 = metric == "l2"
    nlist = centroids.shape[0]
    nprobe = min(nprobe, nlist)
    _, probes = faiss_flat_search(centroids, q, nprobe, "l2" if l2 else "ip")
    order = np.argsort(assignments, kind="stable")
    counts = np.bincount(assignments, minlength=nlist)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    out_d = np.full((q.shape[0], k), FLT_MAX if l2 else -FLT_MAX, dtype=np.float32)
    out_i = np.full((q.shape[0], k), -1, dtype=np.int64)
    b64 = base.astype(np.float64)
    for r in range(q.shape[0]):
        ids = np.concatenate([order[offsets[c]:offsets[c + 1]] for c in probes[r] if c >= 0])
        if ids.size == 0:
            continue
        vec = b64[ids]
        qq = q[r].astype(np.float64)
        key = ((vec - qq) ** 2).sum(axis=1) if l2 else -(vec @ qq)
        limit = min(k, ids.size)
        sel = np.lexsort((ids, key))[:limit]
        out_d[r, :limit] = key[sel] if l2 else -key[sel]
        out_i[r, :limit] = ids[sel]
    return out_d, out_i, probes


# --------------------------------------------------------------------------- IVF + 8-bit scalar quantiser
def sq8_train(residuals: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``ScalarQuantizer`` QT_8bit range training [FAISS-upstream: RS_minmax, rangestat_arg = 0]: per dimension
    vmin = min, vdiff = max - min over the training residuals (reached from ``index.train`` of an
    ``"IVF<n>,SQ8"`` index, src/algorithms/modular.py:277-283; configs/benchmark_config.yaml:51-60)."""
    r = _as_f32(residuals)
    vmin = r.min(axis=0)
    return vmin, (r.max(axis=0) - vmin).astype(np.float32)


def sq8_encode(residuals: np.ndarray, vmin: np.ndarray, vdiff: np.ndarray) -> np.ndarray:
    """code = min(255, int(255 * clamp((r - vmin) / vdiff, 0, 1))) per component, fp32 arithmetic [FAISS-upstream Codec8bit]."""
    r = _as_f32(residuals)
    with np.errstate(divide="ignore", invalid="ignore"):
        xi = np.where(vdiff != 0, (r - vmin[None, :]) / vdiff[None, :], np.float32(0.0)).astype(np.float32)
    xi = np.clip(xi, np.float32(0.0), np.float32(1.0))
    return np.minimum(255, (np.float32(255.0) * xi).astype(np.int32)).astype(np.uint8)


def sq8_decode(codes: np.ndarray, vmin: np.ndarray, vdiff: np.ndarray) -> np.ndarray:
    """r^ = vmin + vdiff * (code + 0.5) / 255 (fp64 here: the exact value the fp32 kernel approximates)."""
    return vmin.astype(np.float64)[None, :] + vdiff.astype(np.float64)[None, :] * (codes.astype(np.float64) + 0.5) / 255.0


def ivf_sq8_search(codes: np.ndarray, centroids: np.ndarray, assignments: np.ndarray, vmin: np.ndarray, vdiff: np.ndarray,
                   queries: np.ndarray, k: int, nprobe: int, metric: str = "l2") -> Tuple[np.ndarray, np.ndarray]:
    """``IndexIVFScalarQuantizer.search`` (by_residual) [FAISS-upstream] given the index's own centroids, assignments,
    ranges and codes: per query the ``nprobe`` best lists, every row of those lists scored on its DECODED vector
    c + r^ - squared L2 ascending / inner product descending, -1 / +-FLT_MAX padding (FAISS conventions)."""
    q = _as_f32(np.atleast_2d(queries))
    l2 = metric == "l2"
    nlist = centroids.shape[0]
    nprobe = min(nprobe, nlist)
    _, probes = faiss_flat_search(centroids, q, nprobe, "l2" if l2 else "ip")
    order = np.argsort(assignments, kind="stable")
    counts = np.bincount(assignments, minlength=nlist)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    out_d = np.full((q.shape[0], k), FLT_MAX if l2 else -FLT_MAX, dtype=np.float32)
    out_i = np.full((q.shape[0], k), -1, dtype=np.int64)
    for r in range(q.shape[0]):
        ids = np.concatenate([order[offsets[c]:offsets[c + 1]] for c in probes[r] if c >= 0])
        if ids.size == 0:
            continue
        vec = centroids.astype(np.float64)[assignments[ids]] + sq8_decode(codes[ids], vmin, vdiff)
        qq = q[r].astype(np.float64)
        key = ((vec - qq) ** 2).sum(axis=1) if l2 else -(vec @ qq)
        limit = min(k, ids.size)
        sel = np.lexsort((ids, key))[:limit]
        out_d[r, :limit] = key[sel] if l2 else -key[sel]
        out_i[r, :limit] = ids[sel]
    return out_d, out_i


# --------------------------------------------------------------------------- product quantisation
def pq_encode(x: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    """``ProductQuantizer::compute_codes`` [FAISS-upstream]: per sub-space the nearest of its 256 centroids (squared L2,
    ties to the lowest index), fp64.  ``codebooks`` [M, 256, dsub]; returns codes [n, M] uint8.  Reached from ``index.add`` of
    a ``"PQ<m>"`` / ``"IVF<n>,PQ<m>"`` index (src/algorithms/modular.py:277-283; configs/benchmark_config.yaml:36-50,61-72)."""
    x = np.asarray(x, dtype=np.float64)
    m, _, dsub = codebooks.shape
    codes = np.empty((x.shape[0], m), dtype=np.uint8)
    for s in range(m):
        sub = x[:, s * dsub:(s + 1) * dsub]
        cb = codebooks[s].astype(np.float64)
        d2 = (sub ** 2).sum(axis=1)[:, None] - 2.0 * sub @ cb.T + (cb ** 2).sum(axis=1)[None, :]
        codes[:, s] = np.argmin(d2, axis=1)
    return codes


def pq_decode(codes: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    """Reconstruction: the concatenation of the selected sub-centroids, fp64 [n, d]."""
    m = codebooks.shape[0]
    return np.concatenate([codebooks[s].astype(np.float64)[codes[:, s]] for s in range(m)], axis=1)


def ivf_pq_search(codes: np.ndarray, centroids: Optional[np.ndarray], assignments: np.ndarray, codebooks: np.ndarray,
                  queries: np.ndarray, k: int, nprobe: int, metric: str = "l2") -> Tuple[np.ndarray, np.ndarray]:
    """``IndexIVFPQ.search`` (by_residual) - or ``IndexPQ.search`` when ``centroids`` is None - [FAISS-upstream] given the
    index's own centroids, assignments, codebooks and codes: every row of the probed lists scored on its RECONSTRUCTED
    vector c + decode(code) (what the look-up-table sum equals); FAISS value conventions."""
    q = _as_f32(np.atleast_2d(queries))
    l2 = metric == "l2"
    n = codes.shape[0]
    recon = pq_decode(codes, codebooks)
    if centroids is None:
        probes = np.zeros((q.shape[0], 1), dtype=np.int64)
        assignments = np.zeros(n, dtype=np.int64)
        nlist = 1
    else:
        nlist = centroids.shape[0]
        _, probes = faiss_flat_search(centroids, q, min(nprobe, nlist), "l2" if l2 else "ip")
        recon = recon + centroids.astype(np.float64)[assignments]
    order = np.argsort(assignments, kind="stable")
    counts = np.bincount(assignments, minlength=nlist)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    out_d = np.full((q.shape[0], k), FLT_MAX if l2 else -FLT_MAX, dtype=np.float32)
    out_i = np.full((q.shape[0], k), -1, dtype=np.int64)
    for r in range(q.shape[0]):
        ids = np.concatenate([order[offsets[c]:offsets[c + 1]] for c in probes[r] if c >= 0])
        if ids.size == 0:
            continue
        qq = q[r].astype(np.float64)
        key = ((recon[ids] - qq) ** 2).sum(axis=1) if l2 else -(recon[ids] @ qq)
        limit = min(k, ids.size)
        sel = np.lexsort((ids, key))[:limit]
        out_d[r, :limit] = key[sel] if l2 else -key[sel]
        out_i[r, :limit] = ids[sel]
    return out_d, out_i


# --------------------------------------------------------------------------- IVF training (k-means recipe)
def kmeans_lloyd_step(sample: np.ndarray, cent: np.ndarray, spherical: bool) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One Lloyd iteration in fp64: nearest centroid (L2, or largest inner product when ``spherical``; ties to
    the lowest centroid index), mean per cluster, empty clusters re-seeded by splitting the most populated one
    with the +-1/1024 sign pattern (the FAISS ``Clustering`` rule [FAISS-upstream], reached from
    ``index.train`` at src/algorithms/modular.py:281-282), spherical centroids re-normalised.
    Returns (new centroids float32, assignment, cluster sizes before the split)."""
    x = np.asarray(sample, dtype=np.float64)
    c = np.asarray(cent, dtype=np.float64)
    nlist, d = c.shape
    assign = np.empty(x.shape[0], dtype=np.int64)
    cn = (c ** 2).sum(axis=1)
    for s in range(0, x.shape[0], 65536):
        ip = x[s:s + 65536] @ c.T
        key = -ip if spherical else cn[None, :] - 2.0 * ip
        assign[s:s + 65536] = np.argmin(key, axis=1)
    sizes = np.bincount(assign, minlength=nlist).astype(np.int64)
    sums = np.zeros((nlist, d))
    np.add.at(sums, assign, x)
    new = (sums / np.maximum(sizes, 1)[:, None]).astype(np.float32)
    cnt = sizes.copy()
    for e in np.nonzero(cnt == 0)[0].tolist():
        big = int(np.argmax(cnt))
        eps = 1.0 / 1024.0
        sign = np.where(np.arange(d) % 2 == 0, 1.0 + eps, 1.0 - eps).astype(np.float32)
        new[e] = new[big] * sign
        new[big] = new[big] * (2.0 - sign)
        cnt[e] = cnt[big] // 2
        cnt[big] -= cnt[e]
    if spherical:
        new = safe_normalize(new)
    return new, assign, sizes


def kmeans_lloyd(vectors: np.ndarray, nlist: int, metric: str = "l2", niter: int = 10, seed: int = 1234,
                 max_points_per_centroid: int = 256) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The IVF training recipe of ``engine.kmeans_train`` restated on the CPU in fp64 - what replaces
    ``index.train(data)`` (src/algorithms/modular.py:281-282; FAISS k-means itself is not reproducible here:
    parity unpinned vs FAISS).  Same draw order: ``RandomState(seed)``; if n > 256 * nlist a sorted sample of
    that many rows without replacement; ``nlist`` sorted distinct sample rows as starting centroids; ``niter``
    Lloyd iterations.  cosine: rows normalised first; ip / cosine: spherical.
    Returns (centroids, sample rows, inertia per iteration = mean squared distance / mean (1 - cos))."""
    v = _as_f32(vectors)
    n = v.shape[0]
    rng = np.random.RandomState(seed)
    limit = max_points_per_centroid * nlist
    rows = np.sort(rng.choice(n, limit, replace=False)) if n > limit else np.arange(n)
    init = np.sort(rng.permutation(rows.shape[0])[:nlist])
    sample = v[rows]
    if metric == "cosine":
        sample = safe_normalize(sample)
    spherical = metric != "l2"
    cent = sample[init].copy()
    inertia = []
    for _ in range(niter):
        new, assign, _ = kmeans_lloyd_step(sample, cent, spherical)
        diff = sample.astype(np.float64) - cent.astype(np.float64)[assign]
        inertia.append(float((diff ** 2).sum(axis=1).mean()))
        cent = new
    return cent, rows, np.asarray(inertia)


# --------------------------------------------------------------------------- multi-shard merge
def merge_topk(dists: Sequence[np.ndarray], ids: Sequence[np.ndarray], k: int,
               ascending: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard sorted top-k lists into one; ties broken on (distance, id) so the
    result is independent of the shard count.  id -1 entries are padding."""
    d = np.concatenate(dists, axis=1).astype(np.float64)
    i = np.concatenate(ids, axis=1)
    key = d if ascending else -d
    key = np.where(i < 0, np.inf, key)
    order = np.lexsort((i, key), axis=1)[:, :k]
    return (np.take_along_axis(np.concatenate(dists, axis=1), order, axis=1).astype(np.float32),
            np.take_along_axis(i, order, axis=1).astype(np.int64))


# --------------------------------------------------------------------------- metrics + comparator
def recall_at_k(ground_truth: np.ndarray, predicted: np.ndarray, k: int) -> float:
    """``recall_at_k`` (src/benchmark/metrics.py:4-34): mean over queries of
    |gt[:k] & pred[:k]| / |gt[:k]| with k clipped to the predicted width."""
    k = min(k, predicted.shape[1])
    rec = np.zeros(ground_truth.shape[0])
    for r in range(ground_truth.shape[0]):
        gt = set(ground_truth[r, :k].tolist()) if ground_truth.shape[1] >= k else set(ground_truth[r].tolist())
        pr = set(predicted[r, :k].tolist())
        rec[r] = len(gt & pr) / len(gt) if gt else 0.0
    return float(np.mean(rec))


def compare_topk(ref_d: np.ndarray, ref_i: np.ndarray, got_d: np.ndarray, got_i: np.ndarray,
                 rtol: float = 1e-5, atol: float = 0.0) -> Dict[str, object]:
    """Tie-tolerant parity check (north star: "IDs bit-exact except where distances tie
    within a stated 1e-5 relative tolerance, distances within that tolerance").

    Position p of a row passes when got_i == ref_i, or when got_i[p] appears in the
    reference row inside a run of reference distances that all lie within
    ``rtol*|d| + atol`` of ref_d[p] (a tie group), or - at the tail - when got_d[p] is
    within tolerance of the reference k-th distance (the boundary tie may pull in an id
    the reference cut off).  Distances must match position-wise within tolerance.
    Returns counts; ``ok`` is True when nothing fails."""
    ref_d = np.asarray(ref_d, dtype=np.float64)
    got_d = np.asarray(got_d, dtype=np.float64)
    assert ref_d.shape == got_d.shape == ref_i.shape == got_i.shape, "shape mismatch"
    finite = np.isfinite(ref_d) & np.isfinite(got_d)
    tol = rtol * np.abs(ref_d) + atol
    dist_bad = np.where(finite, np.abs(ref_d - got_d) > tol, ref_d != got_d)
    # padding must agree exactly
    dist_bad |= (ref_i < 0) != (got_i < 0)
    id_exact = ref_i == got_i
    id_bad = np.zeros_like(id_exact)
    tie_swaps = 0
    rows, cols = np.nonzero(~id_exact)
    for r, c in zip(rows.tolist(), cols.tolist()):
        t = tol[r, c]
        pos = np.nonzero(ref_i[r] == got_i[r, c])[0]
        if pos.size and abs(ref_d[r, pos[0]] - ref_d[r, c]) <= t:
            tie_swaps += 1
            continue
        kth = ref_d[r, ref_i[r] >= 0][-1] if (ref_i[r] >= 0).any() else np.inf
        if not pos.size and abs(got_d[r, c] - kth) <= rtol * abs(kth) + atol:
            tie_swaps += 1
            continue
        id_bad[r, c] = True
    return {
        "ok": not dist_bad.any() and not id_bad.any(),
        "n": int(ref_i.size),
        "id_exact": int(id_exact.sum()),
        "tie_swaps": int(tie_swaps),
        "id_mismatch": int(id_bad.sum()),
        "dist_mismatch": int(dist_bad.sum()),
        "max_rel_err": float(np.max(np.where(finite & (np.abs(ref_d) > 0),
                                             np.abs(ref_d - got_d) / np.maximum(np.abs(ref_d), 1e-30), 0.0),
                                    initial=0.0)),
    }
