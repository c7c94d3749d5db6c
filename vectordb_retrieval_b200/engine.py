"""Device-side operators: torch owns memory and streams, libvdbcuda.so does the arithmetic.

Everything here is plumbing around the C ABI (include/vdb_cuda.h): allocate, copy, call.
No arithmetic of the hot path is done in torch, and nothing falls back to the CPU."""
from __future__ import annotations

import math
import warnings
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import METRIC_IP, METRIC_L2, check, ptr

FLT_MAX = float(np.finfo(np.float32).max)
MAX_FLAT_K = 504        # vdb_flat_topk keeps k' >= k + 8 candidates in pools of at most 512 (common.cuh keep_for_k)
MAX_LIST_K = 512        # IVF scan, rerank and merge kernels


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("vectordb_retrieval_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError(f"device must be a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def metric_code(metric: str) -> int:
    return METRIC_L2 if metric == "l2" else METRIC_IP


def to_device_f32(x, dev: torch.device, chunk_rows: int = 1 << 20) -> torch.Tensor:
    """[n, d] float32 contiguous device tensor from numpy (any dtype/layout, memmap ok) or torch."""
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float32).contiguous()
    x = np.asarray(x) if not isinstance(x, np.ndarray) else x
    if x.ndim != 2:
        raise RuntimeError(f"expected a 2-D array, got shape {x.shape}")
    n, d = x.shape
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    for s in range(0, n, chunk_rows):
        blk = np.ascontiguousarray(x[s:s + chunk_rows], dtype=np.float32)
        with warnings.catch_warnings():                    # a read-only memmap slice is only ever read from here
            warnings.simplefilter("ignore", UserWarning)
            src = torch.from_numpy(blk)
        out[s:s + blk.shape[0]].copy_(src, non_blocking=False)
    return out


def queries_to_device(queries, dev: torch.device, dimension: Optional[int] = None) -> torch.Tensor:
    """Host query batch (numpy, any dtype/layout; 1-D = one query) -> fresh [nq, d] fp32 device tensor.
    The copy is asynchronous when the host array is pinned; the kernels that follow are ordered
    behind it on the current stream.  Raises RuntimeError on a shape mismatch (never ValueError:
    the reference harness swallows that and degrades to per-query search)."""
    if isinstance(queries, torch.Tensor):
        q = queries.to(device=dev, dtype=torch.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        return q.contiguous().clone() if q.data_ptr() == queries.data_ptr() else q.contiguous()
    q = np.asarray(queries)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    if q.ndim != 2 or (dimension is not None and q.shape[1] != dimension):
        raise RuntimeError(f"query batch has shape {q.shape}, expected [nq, {dimension}]")
    if q.dtype != np.float32 or not q.flags["C_CONTIGUOUS"]:
        q = np.ascontiguousarray(q, dtype=np.float32)
    if not q.flags["WRITEABLE"]:
        q = q.copy()
    out = torch.empty(q.shape, dtype=torch.float32, device=dev)
    out.copy_(torch.from_numpy(q), non_blocking=True)
    return out


def results_to_host(dist: torch.Tensor, idx: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """Device results -> fresh numpy arrays owned by the caller.  Lands in pinned memory from
    torch's caching host allocator (no extra host memcpy); the arrays keep their blocks alive."""
    hd = torch.empty(dist.shape, dtype=dist.dtype, pin_memory=True)
    hi = torch.empty(idx.shape, dtype=idx.dtype, pin_memory=True)
    hd.copy_(dist, non_blocking=True)
    hi.copy_(idx, non_blocking=True)
    torch.cuda.current_stream(dist.device).synchronize()
    return hd.numpy(), hi.numpy()


def normalize_rows_(x: torch.Tensor) -> torch.Tensor:
    """In-place row normalisation on device (zero rows stay zero)."""
    lib = _lib.load()
    with torch.cuda.device(x.device):
        check(lib.vdb_normalize_rows(ptr(x), x.shape[0], x.shape[1], x.stride(0), ptr(x), x.stride(0), _stream(x.device)),
              "vdb_normalize_rows")
    return x


def normalized_rows(x: torch.Tensor) -> torch.Tensor:
    """Row-normalised COPY (zero rows stay zero): the caller's tensor is never written.  Used wherever the input
    may alias user data (a base or query batch handed over as a device tensor)."""
    lib = _lib.load()
    y = torch.empty((x.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    if x.shape[0] == 0:
        return y
    with torch.cuda.device(x.device):
        check(lib.vdb_normalize_rows(ptr(x), x.shape[0], x.shape[1], x.stride(0), ptr(y), y.stride(0), _stream(x.device)),
              "vdb_normalize_rows")
    return y


def row_norms(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.vdb_row_norms(ptr(x), x.shape[0], x.shape[1], x.stride(0), ptr(out), _stream(x.device)), "vdb_row_norms")
    return out


class FlatShard:
    """One GPU's row range of a flat (exact) index, in the tcgen05 operand layout.

    HBM layout: ``hi``/``lo`` [n_pad, kpad] fp32 (TF32 split, hi + lo == x exactly) and
    ``norms`` [n_pad]; the original fp32 rows are not kept (2x the base bytes in total)."""

    def __init__(self, vectors, metric: str = "l2", device=None, id_offset: int = 0, upload_rows: int = 1 << 20):
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.id_offset = int(id_offset)
        n, d = int(vectors.shape[0]), int(vectors.shape[1])
        if n <= 0:
            raise RuntimeError("cannot build a flat index over zero vectors")
        self.n, self.d = n, d
        self.kpad = self.lib.vdb_flat_kpad(d)
        self.n_pad = self.lib.vdb_flat_npad(n)
        with torch.cuda.device(self.dev):
            self.hi = torch.empty((self.n_pad, self.kpad), dtype=torch.float32, device=self.dev)
            self.lo = torch.empty((self.n_pad, self.kpad), dtype=torch.float32, device=self.dev)
            self.norms = torch.empty(self.n_pad, dtype=torch.float32, device=self.dev)
            upload_rows = max(256, upload_rows // 256 * 256)
            for s in range(0, n, upload_rows):
                blk = to_device_f32(vectors[s:s + upload_rows], self.dev)
                if metric == "cosine":
                    blk = normalized_rows(blk)             # a copy: `vectors` may be the caller's device tensor
                m = blk.shape[0]
                check(self.lib.vdb_flat_prepare(ptr(blk), m, d, blk.stride(0), metric_code(metric),
                                                self.hi[s:].data_ptr(), self.lo[s:].data_ptr(), self.norms[s:].data_ptr(),
                                                _stream(self.dev)), "vdb_flat_prepare")
                del blk
            torch.cuda.current_stream(self.dev).synchronize()
        self._ws: Dict[Tuple[int, int], torch.Tensor] = {}
        self._qbuf: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def memory_bytes(self) -> int:
        return (self.hi.numel() + self.lo.numel() + self.norms.numel()) * 4

    # ---- persistence: the operands themselves, so a reloaded shard is bit-identical -------------
    def state(self) -> Dict[str, np.ndarray]:
        return {"hi": self.hi.cpu().numpy(), "lo": self.lo.cpu().numpy(), "norms": self.norms.cpu().numpy(),
                "meta": np.array([self.n, self.d, self.id_offset], dtype=np.int64)}

    @classmethod
    def from_state(cls, state: Dict[str, np.ndarray], metric: str, device=None) -> "FlatShard":
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.n, self.d, self.id_offset = (int(v) for v in state["meta"])
        self.kpad = self.lib.vdb_flat_kpad(self.d)
        self.n_pad = self.lib.vdb_flat_npad(self.n)
        if tuple(state["hi"].shape) != (self.n_pad, self.kpad):
            raise RuntimeError(f"persisted operands have shape {state['hi'].shape}, expected {(self.n_pad, self.kpad)}")
        self.hi = torch.from_numpy(np.ascontiguousarray(state["hi"])).to(self.dev)
        self.lo = torch.from_numpy(np.ascontiguousarray(state["lo"])).to(self.dev)
        self.norms = torch.from_numpy(np.ascontiguousarray(state["norms"])).to(self.dev)
        self._ws, self._qbuf = {}, {}
        return self

    def _workspace(self, nq: int, k: int) -> torch.Tensor:
        key = (nq, k)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.lib.vdb_flat_topk_workspace_bytes(nq, k)
            if nbytes == 0:
                raise RuntimeError(f"k={k} is not supported by the flat search (1..{MAX_FLAT_K})")
            self._ws.clear()  # one live workspace per shard
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            self._ws[key] = ws
        return ws

    def _query_operands(self, nq: int) -> Tuple[torch.Tensor, torch.Tensor]:
        nq_pad = self.lib.vdb_flat_nqpad(nq)
        buf = self._qbuf.get(nq_pad)
        if buf is None:
            self._qbuf.clear()
            buf = (torch.empty((nq_pad, self.kpad), dtype=torch.float32, device=self.dev),
                   torch.empty((nq_pad, self.kpad), dtype=torch.float32, device=self.dev))
            self._qbuf[nq_pad] = buf
        return buf

    def prepare_queries(self, q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """q: [nq, d] fp32 device tensor; never written (cosine normalises a copy)."""
        if self.metric == "cosine":
            q = normalized_rows(q)
        q_hi, q_lo = self._query_operands(q.shape[0])
        check(self.lib.vdb_flat_prepare_queries(ptr(q), q.shape[0], self.d, q.stride(0), ptr(q_hi), ptr(q_lo),
                                                _stream(self.dev)), "vdb_flat_prepare_queries")
        return q_hi, q_lo

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: float = FLT_MAX, impl: int = _lib.IMPL_AUTO,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fused scan + top-k for a device query batch; returns device tensors (D [nq,k], I [nq,k])."""
        if q.ndim != 2 or q.shape[1] != self.d:
            raise RuntimeError(f"query shape {tuple(q.shape)} does not match dimension {self.d}")
        nq = q.shape[0]
        with torch.cuda.device(self.dev):
            q_hi, q_lo = self.prepare_queries(q)
            if out is None:
                out = (torch.empty((nq, k), dtype=torch.float32, device=self.dev),
                       torch.empty((nq, k), dtype=torch.int64, device=self.dev))
            ws = self._workspace(nq, k)
            check(self.lib.vdb_flat_topk(metric_code(self.metric), ptr(self.hi), ptr(self.lo), ptr(self.norms),
                                         self.n, self.d, self.id_offset, ptr(q_hi), ptr(q_lo), nq, k, flags,
                                         pad_value, impl, ptr(out[0]), ptr(out[1]), ptr(ws), ws.numel(),
                                         _stream(self.dev)), "vdb_flat_topk")
        return out

    def dense_keys(self, q: torch.Tensor, impl: int) -> torch.Tensor:
        """Test hook: the full key matrix n_j - 2 q.x_j as the scan kernel computes it."""
        nq = q.shape[0]
        with torch.cuda.device(self.dev):
            q_hi, q_lo = self.prepare_queries(q)
            keys = torch.zeros((self.lib.vdb_flat_nqpad(nq), self.n_pad), dtype=torch.float32, device=self.dev)
            check(self.lib.vdb_flat_dense_keys(ptr(self.hi), ptr(self.lo), ptr(self.norms), self.n, self.d,
                                               ptr(q_hi), ptr(q_lo), nq, impl, ptr(keys), _stream(self.dev)),
                  "vdb_flat_dense_keys")
        return keys[:nq, :self.n]


def merge_topk(d_all: torch.Tensor, i_all: torch.Tensor, descending: bool = False, pad_value: float = FLT_MAX,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[parts, nq, k] gathered shard results -> merged [nq, k] (parts in ascending id order).  The parts may be
    strided views (one packed exchange buffer per rank, a slice of the queries): only the [nq, k] block of a
    part has to be dense."""
    lib = _lib.load()
    parts, nq, k = d_all.shape
    for t in (d_all, i_all):
        if t.stride(2) != 1 or t.stride(1) != k:
            raise RuntimeError("merge_topk: every part must be a dense [nq, k] block")
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.float32, device=d_all.device),
               torch.empty((nq, k), dtype=torch.int64, device=d_all.device))
    if nq == 0:
        return out
    with torch.cuda.device(d_all.device):
        check(lib.vdb_merge_topk_strided(ptr(d_all), ptr(i_all), d_all.stride(0) if parts > 1 else nq * k,
                                         i_all.stride(0) if parts > 1 else nq * k, parts, nq, k, int(descending), pad_value,
                                         ptr(out[0]), ptr(out[1]), _stream(d_all.device)), "vdb_merge_topk")
    return out


def pad_cols(x: torch.Tensor, multiple: int = 4) -> torch.Tensor:
    """Zero-pad the row pitch to a multiple of `multiple` floats (16-byte aligned rows)."""
    d = x.shape[1]
    dp = (d + multiple - 1) // multiple * multiple
    if dp == d and x.is_contiguous():
        return x
    out = torch.zeros((x.shape[0], dp), dtype=torch.float32, device=x.device)
    out[:, :d].copy_(x)
    return out


class Reranker:
    """Exact re-scoring of candidate ids against fp32 base rows kept on device (LSH rerank)."""

    def __init__(self, vectors, metric: str = "l2", device=None):
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        base = to_device_f32(vectors, self.dev)
        if metric == "cosine":
            base = normalized_rows(base)
        self.n, self.d = base.shape
        self.base = pad_cols(base)

    def memory_bytes(self) -> int:
        return self.base.numel() * 4

    def search(self, q: torch.Tensor, cand: torch.Tensor, k: int, flags: int, pad_value: float = math.inf
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        nq, c = cand.shape
        with torch.cuda.device(self.dev):
            if self.metric == "cosine":
                q = normalized_rows(q)
            qp = pad_cols(q)
            out_d = torch.empty((nq, k), dtype=torch.float32, device=self.dev)
            out_i = torch.empty((nq, k), dtype=torch.int64, device=self.dev)
            check(self.lib.vdb_rerank_topk(metric_code(self.metric), ptr(self.base), self.n, self.d, self.base.stride(0),
                                           ptr(cand), nq, c, ptr(qp), qp.stride(0), k, flags, pad_value,
                                           ptr(out_d), ptr(out_i), _stream(self.dev)), "vdb_rerank_topk")
        return out_d, out_i


def kmeans_sample(vectors, nlist: int, seed: int = 1234, max_points_per_centroid: int = 256) -> Tuple[np.ndarray, np.ndarray]:
    """(training rows, initial-centroid rows within them), both sorted: at most ``max_points_per_centroid * nlist``
    rows drawn without replacement from ``RandomState(seed)``, then ``nlist`` distinct rows of the sample as the
    starting centroids.  the fp64 Lloyd restatement under oracle/ repeats exactly this draw order."""
    n = int(vectors.shape[0])
    rng = np.random.RandomState(seed)
    limit = max_points_per_centroid * nlist
    rows = np.sort(rng.choice(n, limit, replace=False)) if n > limit else np.arange(n)
    init = np.sort(rng.permutation(rows.shape[0])[:nlist])
    return rows, init


def kmeans_step(sample: torch.Tensor, cent: torch.Tensor, spherical: bool, batch: int = 1 << 18) -> Tuple[torch.Tensor, np.ndarray]:
    """One Lloyd iteration on the device: assignment = the flat scan kernel with base := centroids, k := 1
    (L2, or inner product when ``spherical``), ``vdb_kmeans_accumulate`` for the sums, empty clusters re-seeded
    by splitting the most populated one (FAISS does the same), spherical centroids re-normalised.
    Returns (new centroids [nlist, d] on the device, cluster sizes before the split)."""
    lib = _lib.load()
    dev = sample.device
    nlist, d = cent.shape
    ns = sample.shape[0]
    quant = FlatShard(cent, "ip" if spherical else "l2", dev)
    sums = torch.zeros((nlist, d), dtype=torch.float32, device=dev)
    counts = torch.zeros(nlist, dtype=torch.int32, device=dev)
    for s in range(0, ns, batch):
        blk = sample[s:s + batch]
        _, idx = quant.search(blk, 1)
        check(lib.vdb_kmeans_accumulate(ptr(blk), blk.shape[0], d, blk.stride(0), ptr(idx), nlist, ptr(sums),
                                        ptr(counts), _stream(dev)), "vdb_kmeans_accumulate")
    del quant
    cnt_host = counts.cpu().numpy().astype(np.int64)
    sizes = cnt_host.copy()
    new = sums / counts.clamp(min=1).to(torch.float32)[:, None]
    empty = np.nonzero(cnt_host == 0)[0]
    if empty.size:
        new_host = new.cpu().numpy()
        for e in empty.tolist():
            big = int(np.argmax(cnt_host))
            eps = 1.0 / 1024.0
            sign = np.where(np.arange(d) % 2 == 0, 1.0 + eps, 1.0 - eps).astype(np.float32)
            new_host[e] = new_host[big] * sign
            new_host[big] = new_host[big] * (2.0 - sign)
            cnt_host[e] = cnt_host[big] // 2
            cnt_host[big] -= cnt_host[e]
        new = torch.from_numpy(new_host).to(dev)
    if spherical:
        new = normalized_rows(new)
    return new, sizes


def kmeans_train(vectors, nlist: int, metric: str = "l2", device=None, niter: int = 10, seed: int = 1234,
                 max_points_per_centroid: int = 256, batch: int = 1 << 18) -> np.ndarray:
    """Lloyd k-means on the device for the IVF coarse quantiser; returns centroids [nlist, d] (host).

    Follows the shape of FAISS's ``Clustering`` as reached through ``index.train``
    (reference src/algorithms/modular.py:281-282): at most ``max_points_per_centroid * nlist``
    training points sampled without replacement, ``niter`` iterations, random distinct rows as
    initial centroids, empty clusters re-seeded by splitting the most populated one.  l2: nearest
    centroid by L2; ip / cosine: spherical k-means (assignment by inner product, centroids
    re-normalised).  FAISS's RNG stream is not reproduced (parity unpinned vs FAISS);
    the recipe itself is pinned by the fp64 Lloyd restatement under oracle/ (same sample, same start)."""
    dev = _require_cuda(device)
    n = int(vectors.shape[0])
    if n < nlist:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({nlist})")
    rows, init = kmeans_sample(vectors, nlist, seed, max_points_per_centroid)
    sample_host = vectors if rows.shape[0] == n else np.ascontiguousarray(vectors[rows], dtype=np.float32)
    with torch.cuda.device(dev):
        sample = to_device_f32(sample_host, dev)
        if metric == "cosine":
            sample = normalized_rows(sample)
        cent = sample[torch.from_numpy(init).to(dev)].clone()
        for _ in range(max(niter, 0)):
            cent, _ = kmeans_step(sample, cent, metric != "l2", batch)
        torch.cuda.current_stream(dev).synchronize()
        return cent.cpu().numpy()


class IVFShard:
    """IVF-Flat index of one GPU's rows: centroids as a FlatShard (coarse quantiser) plus the
    inverted lists in the interleaved-32 layout."""

    def __init__(self, vectors, centroids, metric: str = "l2", device=None, id_offset: int = 0,
                 assign_batch: int = 1 << 18):
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.id_offset = int(id_offset)
        base = to_device_f32(vectors, self.dev)
        if metric == "cosine":
            base = normalized_rows(base)
        self.n, self.d = base.shape
        cent = to_device_f32(centroids, self.dev)
        self.nlist = cent.shape[0]
        # coarse quantiser: flat L2 / flat IP over the centroids (cosine data is already normalised)
        self.quantizer = FlatShard(cent, "l2" if metric == "l2" else "ip", self.dev)
        self.centroids = cent
        with torch.cuda.device(self.dev):
            assign = torch.empty(self.n, dtype=torch.int32, device=self.dev)
            for s in range(0, self.n, assign_batch):
                _, idx = self.quantizer.search(base[s:s + assign_batch], 1)
                assign[s:s + assign_batch] = idx[:, 0].to(torch.int32)
            self.assign = assign
            counts = torch.zeros(self.nlist, dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_ivf_count(ptr(assign), self.n, self.nlist, ptr(counts), _stream(self.dev)), "vdb_ivf_count")
            blocks = (counts.to(torch.int64) + 31) // 32
            blk_off = torch.zeros(self.nlist + 1, dtype=torch.int32, device=self.dev)
            blk_off[1:] = torch.cumsum(blocks, 0).to(torch.int32)
            n_blocks = int(blk_off[-1].item())
            self.d4 = self.lib.vdb_ivf_d4(self.d)
            self.list_vecs = torch.zeros(max(n_blocks, 1) * self.d4 * 32 * 4, dtype=torch.float32, device=self.dev)
            self.list_ids = torch.full((max(n_blocks, 1) * 32,), -1, dtype=torch.int32, device=self.dev)
            cursor = torch.zeros(self.nlist, dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_ivf_fill(ptr(base), self.n, self.d, base.stride(0), ptr(assign), ptr(blk_off), self.nlist,
                                        ptr(cursor), ptr(self.list_vecs), ptr(self.list_ids), _stream(self.dev)),
                  "vdb_ivf_fill")
            self.blk_off, self.counts, self.n_blocks = blk_off, counts, n_blocks
            torch.cuda.current_stream(self.dev).synchronize()

    def memory_bytes(self) -> int:
        return self.list_vecs.numel() * 4 + self.list_ids.numel() * 4 + self.quantizer.memory_bytes()

    def state(self) -> Dict[str, np.ndarray]:
        return {"centroids": self.centroids.cpu().numpy(), "list_vecs": self.list_vecs.cpu().numpy(),
                "list_ids": self.list_ids.cpu().numpy(), "blk_off": self.blk_off.cpu().numpy(),
                "counts": self.counts.cpu().numpy(), "assign": self.assign.cpu().numpy(),
                "meta": np.array([self.n, self.d, self.id_offset, self.nlist, self.n_blocks], dtype=np.int64)}

    @classmethod
    def from_state(cls, state: Dict[str, np.ndarray], metric: str, device=None) -> "IVFShard":
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.n, self.d, self.id_offset, self.nlist, self.n_blocks = (int(v) for v in state["meta"])
        self.d4 = self.lib.vdb_ivf_d4(self.d)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)    # noqa: E731
        self.centroids = to(state["centroids"])
        self.list_vecs, self.list_ids, self.blk_off = to(state["list_vecs"]), to(state["list_ids"]), to(state["blk_off"])
        self.counts, self.assign = to(state["counts"]), to(state["assign"])
        self.quantizer = FlatShard(self.centroids, "l2" if metric == "l2" else "ip", self.dev)
        return self

    def search(self, q: torch.Tensor, k: int, nprobe: int, flags: int = 0, pad_value: float = FLT_MAX,
               scanned: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        nq = q.shape[0]
        nprobe = max(1, min(int(nprobe), self.nlist))
        if nprobe > MAX_FLAT_K:
            raise RuntimeError(f"nprobe={nprobe} is not supported: the coarse quantiser selects at most {MAX_FLAT_K} lists "
                               f"per query (flat top-k limit); use nprobe <= {MAX_FLAT_K}")
        with torch.cuda.device(self.dev):
            if self.metric == "cosine":
                q = normalized_rows(q)
            _, probes = self.quantizer.search(q, nprobe)
            out_d = torch.empty((nq, k), dtype=torch.float32, device=self.dev)
            out_i = torch.empty((nq, k), dtype=torch.int64, device=self.dev)
            rows_hint = max(1, nprobe * self.n // max(self.nlist, 1))     # expected rows per query -> warps per query
            check(self.lib.vdb_ivf_scan_topk_ex(metric_code(self.metric), ptr(self.list_vecs), ptr(self.list_ids),
                                                ptr(self.blk_off), self.nlist, self.d, ptr(probes), nprobe, ptr(q),
                                                q.stride(0), nq, k, flags, pad_value, self.id_offset, ptr(out_d), ptr(out_i),
                                                ptr(scanned), rows_hint, _stream(self.dev)), "vdb_ivf_scan_topk")
        self.last_probes = probes
        return out_d, out_i


class IVFSQ8Shard:
    """``IVF<nlist>,SQ8`` of one GPU's rows: the IVF-Flat coarse quantiser plus inverted lists of 8-bit codes of the
    residuals x - centroid (FAISS ``IndexIVFScalarQuantizer``, QT_8bit, by_residual) in the interleaved-32 byte layout.
    HBM: d bytes per row instead of 4 d."""

    def __init__(self, vectors, centroids, metric: str = "l2", device=None, id_offset: int = 0, assign_batch: int = 1 << 18):
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.id_offset = int(id_offset)
        base = to_device_f32(vectors, self.dev)
        if metric == "cosine":
            base = normalized_rows(base)
        self.n, self.d = base.shape
        cent = to_device_f32(centroids, self.dev)
        self.nlist = cent.shape[0]
        self.quantizer = FlatShard(cent, "l2" if metric == "l2" else "ip", self.dev)
        self.centroids = cent
        s = _stream(self.dev)
        with torch.cuda.device(self.dev):
            assign = torch.empty(self.n, dtype=torch.int32, device=self.dev)
            for a in range(0, self.n, assign_batch):
                _, idx = self.quantizer.search(base[a:a + assign_batch], 1)
                assign[a:a + assign_batch] = idx[:, 0].to(torch.int32)
            self.assign = assign
            # the quantiser's ranges come from the residuals of every row handed to train/add (the reference trains on the base)
            resid = torch.empty((self.n, self.d), dtype=torch.float32, device=self.dev)
            check(self.lib.vdb_sq8_residuals(ptr(base), self.n, self.d, base.stride(0), ptr(cent), ptr(assign), ptr(resid), s),
                  "vdb_sq8_residuals")
            self.vmin = torch.empty(self.d, dtype=torch.float32, device=self.dev)
            self.vdiff = torch.empty(self.d, dtype=torch.float32, device=self.dev)
            scratch = torch.empty(2 * self.d, dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_sq8_train(ptr(resid), self.n, self.d, resid.stride(0), ptr(self.vmin), ptr(self.vdiff), ptr(scratch), s),
                  "vdb_sq8_train")
            del resid
            counts = torch.zeros(self.nlist, dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_ivf_count(ptr(assign), self.n, self.nlist, ptr(counts), s), "vdb_ivf_count")
            blocks = (counts.to(torch.int64) + 31) // 32
            blk_off = torch.zeros(self.nlist + 1, dtype=torch.int32, device=self.dev)
            blk_off[1:] = torch.cumsum(blocks, 0).to(torch.int32)
            n_blocks = int(blk_off[-1].item())
            self.d16 = self.lib.vdb_sq8_d16(self.d)
            self.list_codes = torch.zeros(max(n_blocks, 1) * self.d16 * 32 * 16, dtype=torch.uint8, device=self.dev)
            self.list_ids = torch.full((max(n_blocks, 1) * 32,), -1, dtype=torch.int32, device=self.dev)
            cursor = torch.zeros(self.nlist, dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_sq8_fill(ptr(base), self.n, self.d, base.stride(0), ptr(cent), ptr(assign), ptr(blk_off), self.nlist,
                                        ptr(cursor), ptr(self.vmin), ptr(self.vdiff), ptr(self.list_codes), ptr(self.list_ids), s),
                  "vdb_sq8_fill")
            self.blk_off, self.counts, self.n_blocks = blk_off, counts, n_blocks
            torch.cuda.current_stream(self.dev).synchronize()

    def memory_bytes(self) -> int:
        return self.list_codes.numel() + self.list_ids.numel() * 4 + self.quantizer.memory_bytes() + 8 * self.d

    def state(self) -> Dict[str, np.ndarray]:
        return {"centroids": self.centroids.cpu().numpy(), "list_codes": self.list_codes.cpu().numpy(),
                "list_ids": self.list_ids.cpu().numpy(), "blk_off": self.blk_off.cpu().numpy(),
                "counts": self.counts.cpu().numpy(), "assign": self.assign.cpu().numpy(),
                "vmin": self.vmin.cpu().numpy(), "vdiff": self.vdiff.cpu().numpy(),
                "meta": np.array([self.n, self.d, self.id_offset, self.nlist, self.n_blocks], dtype=np.int64)}

    @classmethod
    def from_state(cls, state: Dict[str, np.ndarray], metric: str, device=None) -> "IVFSQ8Shard":
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.n, self.d, self.id_offset, self.nlist, self.n_blocks = (int(v) for v in state["meta"])
        self.d16 = self.lib.vdb_sq8_d16(self.d)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)    # noqa: E731
        self.centroids = to(state["centroids"])
        self.list_codes, self.list_ids, self.blk_off = to(state["list_codes"]), to(state["list_ids"]), to(state["blk_off"])
        self.counts, self.assign = to(state["counts"]), to(state["assign"])
        self.vmin, self.vdiff = to(state["vmin"]), to(state["vdiff"])
        if self.list_codes.numel() != max(self.n_blocks, 1) * self.d16 * 32 * 16 or self.list_ids.numel() != max(self.n_blocks, 1) * 32:
            raise RuntimeError("persisted IVF-SQ8 lists do not match their block count")
        self.quantizer = FlatShard(self.centroids, "l2" if metric == "l2" else "ip", self.dev)
        return self

    def codes_by_row(self) -> np.ndarray:
        """[n, d] uint8 codes in row order (test hook: undoes the interleaved list layout on the host)."""
        ids = self.list_ids.cpu().numpy()
        raw = self.list_codes.cpu().numpy().reshape(-1, self.d16, 32, 16)            # [block, chunk, slot, byte]
        per_slot = raw.transpose(0, 2, 1, 3).reshape(-1, self.d16 * 16)[:, : self.d]    # [block * 32 slots, d]
        out = np.zeros((self.n, self.d), dtype=np.uint8)
        live = ids >= 0
        out[ids[live]] = per_slot[live]
        return out

    def search(self, q: torch.Tensor, k: int, nprobe: int, flags: int = 0, pad_value: float = FLT_MAX
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        nq = q.shape[0]
        nprobe = max(1, min(int(nprobe), self.nlist))
        if nprobe > MAX_FLAT_K:
            raise RuntimeError(f"nprobe={nprobe} is not supported: the coarse quantiser selects at most {MAX_FLAT_K} lists per query")
        with torch.cuda.device(self.dev):
            if self.metric == "cosine":
                q = normalized_rows(q)
            _, probes = self.quantizer.search(q, nprobe)
            out_d = torch.empty((nq, k), dtype=torch.float32, device=self.dev)
            out_i = torch.empty((nq, k), dtype=torch.int64, device=self.dev)
            rows_hint = max(1, nprobe * self.n // max(self.nlist, 1))
            check(self.lib.vdb_ivf_sq8_scan_topk(metric_code(self.metric), ptr(self.list_codes), ptr(self.list_ids), ptr(self.blk_off),
                                                 self.nlist, self.d, ptr(self.centroids), ptr(self.vmin), ptr(self.vdiff), ptr(probes),
                                                 nprobe, ptr(q), q.stride(0), nq, k, flags, pad_value, self.id_offset, ptr(out_d),
                                                 ptr(out_i), rows_hint, _stream(self.dev)), "vdb_ivf_sq8_scan_topk")
        self.last_probes = probes
        return out_d, out_i


def pq_train(residuals: torch.Tensor, m: int, niter: int = 25, seed: int = 1234, max_points_per_centroid: int = 256) -> np.ndarray:
    """Codebooks [m, 256, d / m] of a product quantiser: the IVF k-means recipe (``kmeans_train``) run per sub-space on at
    most 256 * 256 sampled rows, 25 iterations (FAISS ``ProductQuantizer::train`` [FAISS-upstream]; its RNG is not
    reproduced - parity is defined given the codebooks)."""
    n, d = int(residuals.shape[0]), int(residuals.shape[1])
    if d % m != 0:
        raise RuntimeError(f"the dimension {d} is not a multiple of the {m} sub-quantisers")
    if n < 256:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters (256)")
    dsub = d // m
    limit = max_points_per_centroid * 256
    if n > limit:
        pick = np.sort(np.random.RandomState(seed).choice(n, limit, replace=False))
        sample = residuals[torch.from_numpy(pick).to(residuals.device)].cpu().numpy()
    else:
        sample = residuals.cpu().numpy()
    books = np.empty((m, 256, dsub), dtype=np.float32)
    for s in range(m):
        books[s] = kmeans_train(sample[:, s * dsub:(s + 1) * dsub], 256, "l2", residuals.device, niter=niter, seed=seed + 1 + s,
                                max_points_per_centroid=max_points_per_centroid)
    return books


class IVFPQShard:
    """``IVF<nlist>,PQ<m>`` (``centroids`` given) or ``PQ<m>`` (``centroids=None``: one list, zero centroid) of one GPU's
    rows: m code bytes per row - the nearest of 256 sub-centroids per sub-space of the residual x - centroid - in the
    interleaved-32 byte lists; search builds a look-up table per (query, probed list) in shared memory.

    ``PQ<m>`` with a batch of queries (``decoded_scan``: "auto" = 256 queries or more, "always", "never") is searched
    as a FLAT scan over the decoded rows instead: the asymmetric distance |q - x^|^2 (q . x^) IS the flat distance to
    the vector the code stands for, so the tensor-pipe scan + exact re-scoring returns the same neighbours without m
    table look-ups per (query, row) - 1.2M x PQ50, 10k queries: 306 ms -> ~10 ms.  The decoded operands are built on
    first use and kept while they fit ``decoded_max_bytes`` (they cost 8 * kpad bytes per row next to the m code
    bytes, so a base that was quantised to FIT stays on the table scan)."""

    decoded_scan = "auto"
    decoded_min_queries = 256
    decoded_max_bytes = 16 << 30

    def __init__(self, vectors, centroids, m: int, metric: str = "l2", device=None, id_offset: int = 0, codebooks=None,
                 niter: int = 25, seed: int = 1234, assign_batch: int = 1 << 18):
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.id_offset = int(id_offset)
        self.m = int(m)
        self._flat: Optional[FlatShard] = None
        base = to_device_f32(vectors, self.dev)
        if metric == "cosine":
            base = normalized_rows(base)
        self.n, self.d = base.shape
        if self.d % self.m != 0:
            raise RuntimeError(f"the dimension {self.d} is not a multiple of the {self.m} sub-quantisers")
        s = _stream(self.dev)
        with torch.cuda.device(self.dev):
            if centroids is None:
                self.nlist, self.centroids, self.quantizer = 1, None, None
                assign = torch.zeros(self.n, dtype=torch.int32, device=self.dev)
                resid = base
            else:
                cent = to_device_f32(centroids, self.dev)
                self.nlist, self.centroids = cent.shape[0], cent
                self.quantizer = FlatShard(cent, "l2" if metric == "l2" else "ip", self.dev)
                assign = torch.empty(self.n, dtype=torch.int32, device=self.dev)
                for a in range(0, self.n, assign_batch):
                    _, idx = self.quantizer.search(base[a:a + assign_batch], 1)
                    assign[a:a + assign_batch] = idx[:, 0].to(torch.int32)
                resid = torch.empty((self.n, self.d), dtype=torch.float32, device=self.dev)
                check(self.lib.vdb_sq8_residuals(ptr(base), self.n, self.d, base.stride(0), ptr(cent), ptr(assign), ptr(resid), s),
                      "vdb_sq8_residuals")
            self.assign = assign
            books = pq_train(resid, self.m, niter=niter, seed=seed) if codebooks is None else np.ascontiguousarray(codebooks, dtype=np.float32)
            if books.shape != (self.m, 256, self.d // self.m):
                raise RuntimeError(f"codebooks have shape {books.shape}, expected {(self.m, 256, self.d // self.m)}")
            self.codebooks = torch.from_numpy(books).to(self.dev)
            codes = torch.empty((self.n, self.m), dtype=torch.uint8, device=self.dev)
            check(self.lib.vdb_pq_encode(ptr(resid), self.n, self.d, resid.stride(0), ptr(self.codebooks), self.m, ptr(codes), s),
                  "vdb_pq_encode")
            self.codes = codes                                   # [n, m] in row order (kept: 1/4..1/32 of the fp32 rows)
            counts = torch.zeros(self.nlist, dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_ivf_count(ptr(assign), self.n, self.nlist, ptr(counts), s), "vdb_ivf_count")
            blocks = (counts.to(torch.int64) + 31) // 32
            blk_off = torch.zeros(self.nlist + 1, dtype=torch.int32, device=self.dev)
            blk_off[1:] = torch.cumsum(blocks, 0).to(torch.int32)
            n_blocks = int(blk_off[-1].item())
            self.m16 = self.lib.vdb_sq8_d16(self.m)
            self.lists = torch.zeros(max(n_blocks, 1) * self.m16 * 32 * 16, dtype=torch.uint8, device=self.dev)
            self.list_ids = torch.full((max(n_blocks, 1) * 32,), -1, dtype=torch.int32, device=self.dev)
            cursor = torch.zeros(self.nlist, dtype=torch.int32, device=self.dev)
            # per-row bias |r^|^2 + 2 c.r^ (L2): with it the scan needs one look-up table per query, not one per probed list
            self.list_bias = None
            row_bias = None
            if metric == "l2":
                row_bias = torch.empty(self.n, dtype=torch.float32, device=self.dev)
                check(self.lib.vdb_pq_bias(ptr(codes), self.n, self.d, self.m, ptr(self.codebooks), ptr(self.centroids), ptr(assign),
                                           ptr(row_bias), s), "vdb_pq_bias")
                self.list_bias = torch.zeros(max(n_blocks, 1) * 32, dtype=torch.float32, device=self.dev)
            check(self.lib.vdb_bytes_fill(ptr(codes), self.n, self.m, ptr(assign), ptr(blk_off), self.nlist, ptr(cursor),
                                          ptr(self.lists), ptr(self.list_ids), ptr(row_bias), ptr(self.list_bias), s), "vdb_bytes_fill")
            self.blk_off, self.counts = blk_off, counts
            torch.cuda.current_stream(self.dev).synchronize()

    def state(self) -> Dict[str, np.ndarray]:
        out = {"codebooks": self.codebooks.cpu().numpy(), "codes": self.codes.cpu().numpy(), "lists": self.lists.cpu().numpy(),
               "list_ids": self.list_ids.cpu().numpy(), "blk_off": self.blk_off.cpu().numpy(), "counts": self.counts.cpu().numpy(),
               "assign": self.assign.cpu().numpy(),
               "meta": np.array([self.n, self.d, self.id_offset, self.nlist, self.m], dtype=np.int64)}
        if self.centroids is not None:
            out["centroids"] = self.centroids.cpu().numpy()
        if self.list_bias is not None:
            out["list_bias"] = self.list_bias.cpu().numpy()
        return out

    @classmethod
    def from_state(cls, state: Dict[str, np.ndarray], metric: str, device=None) -> "IVFPQShard":
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.metric = metric
        self.n, self.d, self.id_offset, self.nlist, self.m = (int(v) for v in state["meta"])
        self.m16 = self.lib.vdb_sq8_d16(self.m)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)    # noqa: E731
        self.codebooks, self.codes = to(state["codebooks"]), to(state["codes"])
        self.lists, self.list_ids, self.blk_off = to(state["lists"]), to(state["list_ids"]), to(state["blk_off"])
        self.counts, self.assign = to(state["counts"]), to(state["assign"])
        self.centroids = to(state["centroids"]) if "centroids" in state else None
        self.list_bias = to(state["list_bias"]) if "list_bias" in state else None
        if tuple(self.codebooks.shape) != (self.m, 256, self.d // self.m) or self.lists.numel() != self.list_ids.numel() * self.m16 * 16:
            raise RuntimeError("persisted PQ codebooks / lists do not match their header")
        if (metric == "l2") != (self.list_bias is not None):
            raise RuntimeError(f"persisted PQ index was not built for metric '{metric}'")
        self.quantizer = None if self.centroids is None else FlatShard(self.centroids, "l2" if metric == "l2" else "ip", self.dev)
        self._flat = None
        return self

    def _decoded_flat(self) -> Optional[FlatShard]:
        """Flat operands of the decoded rows (``PQ<m>`` only), or None when they would not fit ``decoded_max_bytes``."""
        if self._flat is not None or self.quantizer is not None:
            return self._flat
        if self.lib.vdb_flat_npad(self.n) * self.lib.vdb_flat_kpad(self.d) * 8 > self.decoded_max_bytes:
            return None
        shard = self

        class Decoded:                                    # rows materialised block-wise by FlatShard's upload loop
            shape = (shard.n, shard.d)

            def __getitem__(self, sl):
                a, z = sl.start or 0, min(sl.stop, shard.n)
                out = torch.empty((z - a, shard.d), dtype=torch.float32, device=shard.dev)
                check(shard.lib.vdb_pq_decode(shard.codes[a:].data_ptr(), z - a, shard.d, shard.m, ptr(shard.codebooks), None, None,
                                              ptr(out), shard.d, _stream(shard.dev)), "vdb_pq_decode")
                return out

        # the rows were normalised before they were encoded (cosine): the decoded vectors are scored as they are
        self._flat = FlatShard(Decoded(), "l2" if self.metric == "l2" else "ip", self.dev, id_offset=self.id_offset)
        return self._flat

    def memory_bytes(self) -> int:
        q = self.quantizer.memory_bytes() if self.quantizer is not None else 0
        bias = self.list_bias.numel() * 4 if self.list_bias is not None else 0
        flat = self._flat.memory_bytes() if self._flat is not None else 0
        return self.lists.numel() + self.list_ids.numel() * 4 + bias + self.codes.numel() + self.codebooks.numel() * 4 + q + flat

    def search(self, q: torch.Tensor, k: int, nprobe: int = 1, flags: int = 0, pad_value: float = FLT_MAX
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        nq = q.shape[0]
        with torch.cuda.device(self.dev):
            if self.metric == "cosine":
                q = normalized_rows(q)
            q = q.contiguous()
            probes = None
            if (self.quantizer is None and k <= MAX_FLAT_K and self.decoded_scan != "never"
                    and (self.decoded_scan == "always" or nq >= self.decoded_min_queries)):
                flat = self._decoded_flat()
                if flat is not None:
                    return flat.search(q, k, flags, pad_value)
            if self.quantizer is None:
                nprobe = 1                                   # IndexPQ: the single list
            else:
                nprobe = max(1, min(int(nprobe), self.nlist))
                if nprobe > MAX_FLAT_K:
                    raise RuntimeError(f"nprobe={nprobe} is not supported: the coarse quantiser selects at most {MAX_FLAT_K} lists per query")
                _, probes = self.quantizer.search(q, nprobe)
            out_d = torch.empty((nq, k), dtype=torch.float32, device=self.dev)
            out_i = torch.empty((nq, k), dtype=torch.int64, device=self.dev)
            check(self.lib.vdb_ivf_pq_scan_topk(metric_code(self.metric), ptr(self.lists), ptr(self.list_ids), ptr(self.list_bias),
                                                ptr(self.blk_off), self.nlist, self.d, self.m, ptr(self.centroids), ptr(self.codebooks), ptr(probes), nprobe,
                                                ptr(q), q.stride(0), nq, k, flags, pad_value, self.id_offset, ptr(out_d), ptr(out_i),
                                                _stream(self.dev)), "vdb_ivf_pq_scan_topk")
        return out_d, out_i


class HammingShard:
    """Sign-projection codes of one GPU's rows + the Hamming top-k scan (``faiss.IndexLSH``).

    HBM layout: ``codes`` [n, words] int32 (words = 4 * ceil(nbits / 128), 32 bytes per row at 256
    bits) and the transposed projection [d, words * 32]; large bases with nbits <= 256 also keep the
    codes as bf16 +-1 rows (512 bytes per row at 256 bits) for the tensor-pipe scan, built on first use."""

    def __init__(self, vectors, projection: np.ndarray, device=None, id_offset: int = 0, upload_rows: int = 1 << 20):
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.id_offset = int(id_offset)
        self.nbits, self.d = int(projection.shape[0]), int(projection.shape[1])
        if vectors.shape[1] != self.d:
            raise RuntimeError(f"expected dimension {self.d}, got {vectors.shape[1]}")
        if self.nbits > 1024:
            raise RuntimeError("at most 1024 hash bits are supported")
        self.n = int(vectors.shape[0])
        self.words = self.lib.vdb_lsh_code_words(self.nbits)
        if self.words == 12:
            self.words = 16            # (256, 512] bits share the 512-bit kernel
        elif self.words > 16:
            self.words = 32
        with torch.cuda.device(self.dev):
            pt = np.zeros((self.d, self.lib.vdb_lsh_code_words(self.nbits) * 32), dtype=np.float32)
            pt[:, : self.nbits] = np.ascontiguousarray(projection, dtype=np.float32).T
            self.proj_t = torch.from_numpy(pt).to(self.dev)
            self.codes = torch.zeros((self.n, self.words), dtype=torch.int32, device=self.dev)
            for s in range(0, self.n, upload_rows):
                blk = to_device_f32(vectors[s:s + upload_rows], self.dev)
                self._encode(blk, self.codes[s:s + blk.shape[0]])
                del blk
            torch.cuda.current_stream(self.dev).synchronize()

    def _encode(self, x: torch.Tensor, out: torch.Tensor) -> None:
        native = self.lib.vdb_lsh_code_words(self.nbits)
        if native == self.words:
            check(self.lib.vdb_lsh_encode(ptr(x), x.shape[0], self.d, x.stride(0), ptr(self.proj_t), self.nbits, ptr(out),
                                          _stream(self.dev)), "vdb_lsh_encode")
        else:                                   # widen to the kernel's code width, padding words stay zero
            tmp = torch.empty((x.shape[0], native), dtype=torch.int32, device=self.dev)
            check(self.lib.vdb_lsh_encode(ptr(x), x.shape[0], self.d, x.stride(0), ptr(self.proj_t), self.nbits, ptr(tmp),
                                          _stream(self.dev)), "vdb_lsh_encode")
            out.zero_()
            out[:, :native].copy_(tmp)

    def memory_bytes(self) -> int:
        expanded = getattr(self, "_bf16", None)           # bf16 copies for the tensor-pipe scan, built on first use
        extra = 0 if expanded is None else expanded[0].numel() + expanded[1].numel() * 4
        return self.codes.numel() * 4 + self.proj_t.numel() * 4 + extra

    def state(self) -> Dict[str, np.ndarray]:
        return {"codes": self.codes.cpu().numpy(), "proj_t": self.proj_t.cpu().numpy(),
                "meta": np.array([self.n, self.d, self.id_offset, self.nbits, self.words], dtype=np.int64)}

    @classmethod
    def from_state(cls, state: Dict[str, np.ndarray], device=None) -> "HammingShard":
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.dev = _require_cuda(device)
        self.n, self.d, self.id_offset, self.nbits, self.words = (int(v) for v in state["meta"])
        self.codes = torch.from_numpy(np.ascontiguousarray(state["codes"])).to(self.dev)
        self.proj_t = torch.from_numpy(np.ascontiguousarray(state["proj_t"])).to(self.dev)
        return self

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        out = torch.zeros((x.shape[0], self.words), dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            self._encode(x.contiguous(), out)
        return out

    def search(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(Hamming distance as float32 [nq,k], ids int64 [nq,k]) ordered by (distance, id); k > n pads (inf, -1)."""
        nq = q.shape[0]
        nbits_kernel = self.words * 32 if self.words != self.lib.vdb_lsh_code_words(self.nbits) else self.nbits
        with torch.cuda.device(self.dev):
            qc = self.encode(q)
            out_d = torch.full((nq, k), float("inf"), dtype=torch.float32, device=self.dev)
            out_i = torch.full((nq, k), -1, dtype=torch.int64, device=self.dev)
            if self._use_tensor_pipe(nq, k):
                # fp16 operands + accumulators (packed epilogue, half the filter instructions, a bigger append path):
                # measured on 1.2M x 256 bits, 10k queries - faster from k ~ 2 000 up (k = 6 400: 10.7 vs 12.5 ms),
                # slower below (k = 800: 8.7 vs 7.8 ms); "auto" switches there
                f16 = self.nbits % 2 == 0 and (k >= 2048 if self.tc_accumulate_f16 == "auto" else bool(self.tc_accumulate_f16))
                base16, norms = self._expanded(f16)
                row_bytes = self.lib.vdb_hamming_tc_row_bytes(self.nbits)
                nq_pad = self.lib.vdb_flat_nqpad(nq)
                q16 = torch.empty(nq_pad * row_bytes, dtype=torch.uint8, device=self.dev)
                check(self.lib.vdb_hamming_tc_expand(ptr(qc), nq, self.words, self.nbits, 1 | (2 if f16 else 0), ptr(q16), None,
                                                     nq_pad, _stream(self.dev)), "vdb_hamming_tc_expand")
                nbytes = self.lib.vdb_hamming_tc_workspace_bytes(nq, self.nbits, k, self.n)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
                scan = self.lib.vdb_hamming_topk_tc_f16 if f16 else self.lib.vdb_hamming_topk_tc
                check(scan(ptr(base16), ptr(norms), ptr(self.codes), self.n, ptr(q16), ptr(qc), nq, self.nbits, k, self.id_offset,
                           ptr(out_d), ptr(out_i), ptr(ws), nbytes, _stream(self.dev)), "vdb_hamming_topk_tc")
                return out_d, out_i
            nbytes = self.lib.vdb_hamming_topk_workspace_bytes(nq, nbits_kernel)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            check(self.lib.vdb_hamming_topk(ptr(self.codes), self.n, ptr(qc), nq, nbits_kernel, k, self.id_offset,
                                            ptr(out_d), ptr(out_i), ptr(ws), nbytes, _stream(self.dev)), "vdb_hamming_topk")
        return out_d, out_i

    # ---- tensor-pipe path: bf16 +-1 copies of the codes (512 bytes per row at 256 bits), built on first use
    tensor_pipe = "auto"      # "auto" | True | False
    tc_accumulate_f16 = "auto"  # tensor-pipe scan: True = fp16 operands and accumulators (even nbits), False = bf16 operands and
    #                             fp32 accumulators, "auto" = by k

    def _use_tensor_pipe(self, nq: int, k: int) -> bool:
        if self.tensor_pipe is False or self.lib.vdb_hamming_tc_row_bytes(self.nbits) == 0:
            return False
        if self.words != self.lib.vdb_lsh_code_words(self.nbits) or self.n <= 65536 or k > self.n:
            return False
        if self.lib.vdb_flat_npad(self.n) > (1 << 23):
            return False                                   # list entries hold a 23-bit row
        if self.lib.vdb_hamming_tc_workspace_bytes(nq, self.nbits, k, self.n) > (8 << 30):
            return False                                   # candidate lists of huge batches: stay on the popc kernels
        return True if self.tensor_pipe is True else (nq >= 256 and self.n <= 8_000_000)

    def _expanded(self, f16: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        cached = getattr(self, "_bf16", None)
        if cached is None or cached[2] != f16:
            row_bytes = self.lib.vdb_hamming_tc_row_bytes(self.nbits)
            n_pad = self.lib.vdb_flat_npad(self.n)
            base16 = torch.empty(n_pad * row_bytes, dtype=torch.uint8, device=self.dev)
            norms = torch.empty(n_pad, dtype=torch.float32, device=self.dev)
            check(self.lib.vdb_hamming_tc_expand(ptr(self.codes), self.n, self.words, self.nbits, 2 if f16 else 0, ptr(base16),
                                                 ptr(norms), n_pad, _stream(self.dev)), "vdb_hamming_tc_expand")
            cached = self._bf16 = (base16, norms, f16)
        return cached[0], cached[1]
