"""FAISS-shaped index objects backed by the CUDA kernels.

The reference talks to FAISS through a small duck type - ``index.train(x)``, ``index.add(x)``,
``index.search(q, k) -> (D, I)``, ``index.ntotal``, ``index.nprobe``, ``index.is_trained`` - from
``faiss.IndexFlat`` (src/algorithms/exact_search.py:38-39,78), ``faiss.index_factory``
(src/algorithms/modular.py:277-286, src/algorithms/approximate_search.py:39-51) and
``faiss.IndexLSH`` (src/algorithms/modular.py:215-216,477).  These classes offer that duck type
with FAISS's value conventions (squared L2 ascending, raw inner product descending, int64
labels, -1 / +-FLT_MAX padding) so the reference-facing searchers can stay thin.

Host arrays in, host arrays out; ``search_device`` keeps everything on the GPU for callers that
chain kernels.  No CPU fallback: constructing any of these without a CUDA device raises."""
from __future__ import annotations

import re
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, engine

METRIC_L2, METRIC_INNER_PRODUCT = 1, 0     # numeric values of faiss.METRIC_* (only used as tags)


def _metric_name(metric) -> str:
    if isinstance(metric, str):
        return "l2" if metric == "l2" else "ip"
    return "l2" if metric == METRIC_L2 else "ip"


class GpuIndexFlat:
    """``faiss.IndexFlat(d, metric)``: exact search on one GPU, or row-sharded over several."""

    def __init__(self, d: int, metric="l2", device=None, devices=None, normalize: bool = False, shard: str = "auto",
                 exchange: str = "alltoall"):
        self.shard = shard                    # under torchrun: 'rows', 'queries' (replicated base) or 'auto'
        self.exchange = exchange              # rows layout: 'alltoall' or 'allgather' (sharded.TopKExchange)
        self.d = int(d)
        self.metric = _metric_name(metric)
        self.normalize = bool(normalize)      # cosine: rows and queries are L2-normalised on the device
        self.device, self.devices = device, devices
        self.is_trained = True
        self.ntotal = 0
        self._impl = None

    def train(self, x) -> None:   # nothing to train
        return None

    def add(self, x) -> None:
        if self._impl is not None:
            raise RuntimeError("GpuIndexFlat.add may be called once (the shard layout is built in one pass)")
        if x.shape[1] != self.d:
            raise RuntimeError(f"expected dimension {self.d}, got {x.shape[1]}")
        from . import sharded
        rank, world = sharded.dist_info()
        metric = "cosine" if self.normalize and self.metric == "ip" else self.metric
        if world > 1 and sharded.choose_sharding(int(x.shape[0]), (self.d + 31) // 32 * 32, world, self.shard) == "queries":
            self._impl = sharded.ReplicatedFlatIndex(x, metric, self.device)
        elif world > 1:
            self._impl = sharded.DistributedFlatIndex.from_global(x, metric, self.device, exchange=self.exchange)
        elif self.devices is not None and len(self.devices) > 1:
            self._impl = sharded.MultiDeviceFlatIndex(x, metric, self.devices)
        else:
            dev = self.device if self.device is not None else (self.devices[0] if self.devices else None)
            self._impl = engine.FlatShard(x, metric, dev)
        self.ntotal = int(x.shape[0])

    @property
    def home(self) -> torch.device:
        impl = self._impl
        return impl.home if hasattr(impl, "home") else (impl.shard.dev if hasattr(impl, "shard") else impl.dev)

    def memory_bytes(self) -> int:
        return 0 if self._impl is None else self._impl.memory_bytes()

    def save(self, artifact_dir: str, context=None):
        from . import persist
        if not isinstance(self._impl, engine.FlatShard):
            raise NotImplementedError("only a single-GPU flat index can be persisted")
        meta = {"d": self.d, "metric": self.metric, "normalize": self.normalize, "ntotal": self.ntotal}
        return persist.write_artifact(artifact_dir, "flat", self._impl.state(), meta, context)

    def load(self, artifact_dir: str, context=None):
        from . import persist
        arrays, manifest = persist.read_artifact(artifact_dir, "flat", context)
        meta = manifest["meta"]
        if int(meta["d"]) != self.d or meta["metric"] != self.metric or bool(meta["normalize"]) != self.normalize:
            raise RuntimeError(f"persisted flat index {meta} does not match d={self.d} metric={self.metric}")
        dev = self.device if self.device is not None else (self.devices[0] if self.devices else None)
        self._impl = engine.FlatShard.from_state(arrays, "cosine" if self.normalize and self.metric == "ip" else self.metric, dev)
        self.ntotal = int(meta["ntotal"])
        return manifest

    def search_device(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        pad = engine.FLT_MAX if self.metric == "l2" else -engine.FLT_MAX
        return self._impl.search(q, k, 0, pad)

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        with torch.cuda.device(self.home):
            if hasattr(self._impl, "search_host"):      # one process per GPU: each rank moves only its query slice of the result
                pad = engine.FLT_MAX if self.metric == "l2" else -engine.FLT_MAX
                return self._impl.search_host(queries, int(k), 0, pad)
            q = engine.queries_to_device(queries, self.home, self.d)
            return engine.results_to_host(*self.search_device(q, int(k)))


class GpuIndexIVFFlat:
    """``faiss.index_factory(d, "IVF<nlist>,Flat", metric)``: k-means coarse quantiser + inverted lists.

    train  k-means (Lloyd, ``niter`` = 10 like FAISS's IVF default, at most 256 points per
           centroid sampled with seed 1234), assignment by the flat kernel with base := centroids;
           inner-product indexes use spherical k-means.  FAISS's own RNG / subsampling cannot be
           reproduced without FAISS, so parity with FAISS is defined GIVEN the same centroids
           (pass ``centroids=`` to skip training).
    add    nearest-centroid assignment + scatter into the interleaved-32 list layout
    search top-``nprobe`` centroids per query, scan of those lists with exact scoring."""

    def __init__(self, d: int, nlist: int, metric="l2", device=None, niter: int = 10, seed: int = 1234,
                 max_points_per_centroid: int = 256, centroids: Optional[np.ndarray] = None, normalize: bool = False,
                 shard: str = "auto"):
        if nlist <= 0:
            raise ValueError("nlist must be positive")
        if shard not in ("auto", "rows", "queries"):
            raise ValueError(f"shard must be 'auto', 'rows' or 'queries', got '{shard}'")
        self.shard = shard                    # under torchrun (IVF-Flat): lists cut by row range, or every list on every rank
        self._dist = None
        self.d, self.nlist = int(d), int(nlist)
        self.metric = _metric_name(metric)
        self.normalize = bool(normalize)
        self.device = device
        self.niter, self.seed, self.max_points_per_centroid = int(niter), int(seed), int(max_points_per_centroid)
        self.nprobe = 1
        self.ntotal = 0
        self.centroids = None if centroids is None else np.ascontiguousarray(centroids, dtype=np.float32)
        self.is_trained = self.centroids is not None
        self._impl: Optional[engine.IVFShard] = None

    def train(self, x) -> None:
        if self.is_trained:
            return
        from . import sharded
        rank, world = sharded.dist_info()
        if world > 1:
            # one process per GPU: rank 0 trains, everyone gets ITS centroids (k-means accumulates with float atomics,
            # so two ranks training on the same sample would not agree bit for bit)
            import torch.distributed as dist
            dev = engine._require_cuda(self.device)
            if rank == 0:
                cent = torch.from_numpy(engine.kmeans_train(x, self.nlist, self._engine_metric(), dev, niter=self.niter, seed=self.seed,
                                                            max_points_per_centroid=self.max_points_per_centroid)).to(dev)
            else:
                cent = torch.empty((self.nlist, self.d), dtype=torch.float32, device=dev)
            dist.broadcast(cent, src=0)
            self.centroids = cent.cpu().numpy()
        else:
            self.centroids = engine.kmeans_train(x, self.nlist, self._engine_metric(), self.device, niter=self.niter, seed=self.seed,
                                                 max_points_per_centroid=self.max_points_per_centroid)
        self.is_trained = True

    def _engine_metric(self) -> str:
        return "cosine" if self.normalize and self.metric == "ip" else self.metric

    def add(self, x) -> None:
        if not self.is_trained:
            raise RuntimeError("GpuIndexIVFFlat.add before train")
        if self._impl is not None:
            raise RuntimeError("GpuIndexIVFFlat.add may be called once")
        from . import sharded
        rank, world = sharded.dist_info()
        self.ntotal = int(x.shape[0])
        if world > 1:
            # 'queries' (auto while the lists stay under 8 GB): every list on every rank, each rank scans for its slice of
            # the batch; 'rows': the lists cut by row range + the packed top-k exchange + merge kernel
            layout = self.shard if self.shard != "auto" else ("queries" if 4.0 * self.ntotal * self.d <= 8e9 else "rows")
            if layout == "queries":
                self._dist = sharded.ReplicatedIVFIndex(x, self.centroids, self._engine_metric(), self.device, nprobe=self.nprobe)
            else:
                plan = sharded.ShardPlan(self.ntotal, world)
                lo, hi = plan.start(rank), plan.stop(rank)
                if hi <= lo:
                    raise RuntimeError(f"rank {rank} of {world} owns no rows of a {self.ntotal}-row base")
                self._dist = sharded.DistributedIVFIndex(x[lo:hi], self.centroids, self._engine_metric(), self.device, id_offset=lo,
                                                         nprobe=self.nprobe)
            self._impl = self._dist.shard
            return
        self._impl = engine.IVFShard(x, self.centroids, self._engine_metric(), self.device)

    @property
    def home(self) -> torch.device:
        return self._impl.dev

    def memory_bytes(self) -> int:
        return 0 if self._impl is None else self._impl.memory_bytes()

    def save(self, artifact_dir: str, context=None):
        from . import persist
        if self._impl is None:
            raise RuntimeError("index is empty")
        if self._dist is not None:
            raise RuntimeError("a sharded IVF index is not persisted: save from a single-process index")
        meta = {"d": self.d, "nlist": self.nlist, "metric": self.metric, "normalize": self.normalize,
                "nprobe": int(self.nprobe), "ntotal": self.ntotal}
        return persist.write_artifact(artifact_dir, "ivf_flat", self._impl.state(), meta, context)

    def load(self, artifact_dir: str, context=None):
        from . import persist
        arrays, manifest = persist.read_artifact(artifact_dir, "ivf_flat", context)
        meta = manifest["meta"]
        if int(meta["d"]) != self.d or int(meta["nlist"]) != self.nlist or meta["metric"] != self.metric:
            raise RuntimeError(f"persisted IVF index {meta} does not match d={self.d} nlist={self.nlist} metric={self.metric}")
        self._impl = engine.IVFShard.from_state(arrays, self._engine_metric(), self.device)
        self.centroids = arrays["centroids"]
        self.is_trained, self.ntotal = True, int(meta["ntotal"])
        return manifest

    def search_device(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        pad = engine.FLT_MAX if self.metric == "l2" else -engine.FLT_MAX
        if self._dist is not None:
            self._dist.nprobe = int(self.nprobe)
            return self._dist.search(q, k, 0, pad)
        return self._impl.search(q, k, int(self.nprobe), 0, pad)

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        with torch.cuda.device(self.home):
            if self._dist is not None and hasattr(self._dist, "search_host"):     # query slices: only this rank's slice moves
                self._dist.nprobe = int(self.nprobe)
                pad = engine.FLT_MAX if self.metric == "l2" else -engine.FLT_MAX
                return self._dist.search_host(queries, int(k), 0, pad)
            q = engine.queries_to_device(queries, self.home, self.d)
            return engine.results_to_host(*self.search_device(q, int(k)))


class GpuIndexIVFSQ8(GpuIndexIVFFlat):
    """``faiss.index_factory(d, "IVF<nlist>,SQ8", metric)``: the IVF-Flat coarse quantiser with inverted lists of 8-bit
    scalar-quantised residuals (FAISS ``IndexIVFScalarQuantizer``, QT_8bit, by_residual; value conventions of FAISS:
    squared L2 ascending / inner product descending, of the DECODED vectors).  save / load keep the code lists, the
    quantiser's ranges and the centroids bit for bit."""

    def add(self, x) -> None:
        if not self.is_trained:
            raise RuntimeError("GpuIndexIVFSQ8.add before train")
        if self._impl is not None:
            raise RuntimeError("GpuIndexIVFSQ8.add may be called once")
        self._impl = engine.IVFSQ8Shard(x, self.centroids, self._engine_metric(), self.device)
        self.ntotal = int(x.shape[0])

    _KIND, _SHARD = "ivf_sq8", engine.IVFSQ8Shard

    def save(self, artifact_dir: str, context=None):
        from . import persist
        if self._impl is None:
            raise RuntimeError("index is empty")
        return persist.write_artifact(artifact_dir, self._KIND, self._impl.state(), self._meta(), context)

    def _meta(self):
        return {"d": self.d, "nlist": self.nlist, "metric": self.metric, "normalize": self.normalize,
                "nprobe": int(self.nprobe), "ntotal": self.ntotal}

    def load(self, artifact_dir: str, context=None):
        from . import persist
        arrays, manifest = persist.read_artifact(artifact_dir, self._KIND, context)
        meta, mine = manifest["meta"], self._meta()
        for key in mine:
            if key not in ("nprobe", "ntotal", "normalize") and meta.get(key) != mine[key]:
                raise RuntimeError(f"persisted {self._KIND} index {meta} does not match this index ({key}={mine[key]})")
        self._impl = self._SHARD.from_state(arrays, self._engine_metric(), self.device)
        self.centroids = arrays["centroids"]
        self.is_trained, self.ntotal = True, int(meta["ntotal"])
        return manifest


class GpuIndexIVFPQ(GpuIndexIVFFlat):
    """``faiss.index_factory(d, "IVF<nlist>,PQ<m>", metric)``: IVF coarse quantiser + product-quantised residuals
    (FAISS ``IndexIVFPQ``, 8 bits per sub-quantiser, by_residual); distances are those of the reconstructed vectors,
    FAISS value conventions.  save / load keep codebooks, code lists, biases and centroids bit for bit."""

    def __init__(self, d: int, nlist: int, m: int, metric="l2", **kwargs):
        super().__init__(d, nlist, metric, **kwargs)
        if m <= 0 or d % m != 0:
            raise ValueError(f"PQ{m}: the dimension {d} must be a multiple of the number of sub-quantisers")
        self.m = int(m)

    def add(self, x) -> None:
        if not self.is_trained:
            raise RuntimeError("GpuIndexIVFPQ.add before train")
        if self._impl is not None:
            raise RuntimeError("GpuIndexIVFPQ.add may be called once")
        self._impl = engine.IVFPQShard(x, self.centroids, self.m, self._engine_metric(), self.device, seed=self.seed)
        self.ntotal = int(x.shape[0])

    _KIND, _SHARD = "ivf_pq", engine.IVFPQShard

    def _meta(self):
        return dict(GpuIndexIVFSQ8._meta(self), m=self.m)

    save = GpuIndexIVFSQ8.save
    load = GpuIndexIVFSQ8.load


class GpuIndexPQ:
    """``faiss.index_factory(d, "PQ<m>", metric)``: product quantiser over the rows themselves (FAISS ``IndexPQ``): one
    code list, asymmetric distance computation against every row."""

    def __init__(self, d: int, m: int, metric="l2", device=None, seed: int = 1234, normalize: bool = False):
        if m <= 0 or d % m != 0:
            raise ValueError(f"PQ{m}: the dimension {d} must be a multiple of the number of sub-quantisers")
        self.d, self.m = int(d), int(m)
        self.metric = _metric_name(metric)
        self.normalize = bool(normalize)
        self.device, self.seed = device, int(seed)
        self.is_trained = True          # codebooks are trained inside add (train + add see the same rows in the reference)
        self.ntotal = 0
        self._impl: Optional[engine.IVFPQShard] = None

    def train(self, x) -> None:
        return None

    def add(self, x) -> None:
        if self._impl is not None:
            raise RuntimeError("GpuIndexPQ.add may be called once")
        metric = "cosine" if self.normalize and self.metric == "ip" else self.metric
        self._impl = engine.IVFPQShard(x, None, self.m, metric, self.device, seed=self.seed)
        self.ntotal = int(x.shape[0])

    @property
    def home(self) -> torch.device:
        return self._impl.dev

    def memory_bytes(self) -> int:
        return 0 if self._impl is None else self._impl.memory_bytes()

    def save(self, artifact_dir: str, context=None):
        from . import persist
        if self._impl is None:
            raise RuntimeError("index is empty")
        meta = {"d": self.d, "m": self.m, "metric": self.metric, "normalize": self.normalize, "ntotal": self.ntotal}
        return persist.write_artifact(artifact_dir, "pq", self._impl.state(), meta, context)

    def load(self, artifact_dir: str, context=None):
        from . import persist
        arrays, manifest = persist.read_artifact(artifact_dir, "pq", context)
        meta = manifest["meta"]
        if int(meta["d"]) != self.d or int(meta["m"]) != self.m or meta["metric"] != self.metric:
            raise RuntimeError(f"persisted PQ index {meta} does not match d={self.d} m={self.m} metric={self.metric}")
        metric = "cosine" if self.normalize and self.metric == "ip" else self.metric
        self._impl = engine.IVFPQShard.from_state(arrays, metric, self.device)
        self.ntotal = int(meta["ntotal"])
        return manifest

    def search_device(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        pad = engine.FLT_MAX if self.metric == "l2" else -engine.FLT_MAX
        return self._impl.search(q, k, 1, 0, pad)

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        with torch.cuda.device(self.home):
            q = engine.queries_to_device(queries, self.home, self.d)
            return engine.results_to_host(*self.search_device(q, int(k)))


class GpuIndexLSH:
    """``faiss.IndexLSH(d, nbits)``: sign bits of a random projection, Hamming top-k.

    Codes: bit b of row x = [ (x @ P)[b] >= 0 ], P a d x nbits Gaussian matrix from
    ``RandomState(seed)`` (FAISS draws its own rotation with its own RNG - parity with FAISS's
    codes is unpinned, the candidate *rerank* that follows is what the reference's results hinge
    on).  search returns (Hamming distance as float32, ids) ordered by (distance, id)."""

    def __init__(self, d: int, nbits: int, device=None, seed: int = 1234, projection: Optional[np.ndarray] = None):
        if nbits <= 0:
            raise ValueError("num_bits must be positive")
        self.d, self.nbits = int(d), int(nbits)
        self.device = device
        self.is_trained = True
        self.ntotal = 0
        if projection is None:
            projection = np.random.RandomState(seed).normal(size=(self.nbits, self.d)).astype(np.float32)
        self.projection = np.ascontiguousarray(projection, dtype=np.float32)   # [nbits, d]
        self._impl: Optional[engine.HammingShard] = None

    def train(self, x) -> None:
        return None

    def add(self, x) -> None:
        if self._impl is not None:
            raise RuntimeError("GpuIndexLSH.add may be called once")
        self._impl = engine.HammingShard(x, self.projection, self.device)
        self.ntotal = int(x.shape[0])

    @property
    def home(self) -> torch.device:
        return self._impl.dev

    def memory_bytes(self) -> int:
        return 0 if self._impl is None else self._impl.memory_bytes()

    def save(self, artifact_dir: str, context=None):
        from . import persist
        if self._impl is None:
            raise RuntimeError("index is empty")
        arrays = dict(self._impl.state(), projection=self.projection)
        return persist.write_artifact(artifact_dir, "lsh", arrays, {"d": self.d, "nbits": self.nbits, "ntotal": self.ntotal}, context)

    def load(self, artifact_dir: str, context=None):
        from . import persist
        arrays, manifest = persist.read_artifact(artifact_dir, "lsh", context)
        meta = manifest["meta"]
        if int(meta["d"]) != self.d or int(meta["nbits"]) != self.nbits:
            raise RuntimeError(f"persisted LSH index {meta} does not match d={self.d} nbits={self.nbits}")
        self.projection = arrays.pop("projection")
        self._impl = engine.HammingShard.from_state(arrays, self.device)
        self.ntotal = int(meta["ntotal"])
        return manifest

    def search_device(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        return self._impl.search(q, int(k))

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if self._impl is None:
            raise RuntimeError("index is empty")
        with torch.cuda.device(self.home):
            q = engine.queries_to_device(queries, self.home, self.d)
            d, i = self.search_device(q, int(k))
            return engine.results_to_host(d.to(torch.float32), i)


_IVF_FLAT = re.compile(r"^IVF(\d+),Flat$")
_IVF_SQ8 = re.compile(r"^IVF(\d+),SQ8$")
_IVF_PQ = re.compile(r"^IVF(\d+),PQ(\d+)$")
_PQ = re.compile(r"^PQ(\d+)$")


def index_factory(d: int, key: str, metric="l2", **kwargs):
    """The subset of ``faiss.index_factory`` grammar that reaches the scan + top-k path:
    ``"Flat"``, ``"IVF<nlist>,Flat"``, ``"IVF<nlist>,SQ8"``, ``"IVF<nlist>,PQ<m>"``, ``"PQ<m>"`` and ``"LSH"``.  Anything else
    (HNSW, OPQ, other code widths, ...) is outside
    this build (out of scope) and raises ValueError at construction time."""
    key = key.strip()
    if key == "Flat":
        return GpuIndexFlat(d, metric, **kwargs)
    m = _IVF_FLAT.match(key)
    if m:
        return GpuIndexIVFFlat(d, int(m.group(1)), metric, **kwargs)
    m = _IVF_SQ8.match(key)
    if m:
        return GpuIndexIVFSQ8(d, int(m.group(1)), metric, **kwargs)
    m = _IVF_PQ.match(key)
    if m:
        return GpuIndexIVFPQ(d, int(m.group(1)), int(m.group(2)), metric, **kwargs)
    m = _PQ.match(key)
    if m:
        kwargs = {k_: v for k_, v in kwargs.items() if k_ in ("device", "seed", "normalize")}
        return GpuIndexPQ(d, int(m.group(1)), metric, **kwargs)
    if key == "LSH":
        return GpuIndexLSH(d, kwargs.pop("nbits", 256), **kwargs)
    raise ValueError(f"index key '{key}' is not supported by the CUDA build "
                     "(supported: 'Flat', 'IVF<n>,Flat', 'IVF<n>,SQ8', 'IVF<n>,PQ<m>', 'PQ<m>', 'LSH')")
