"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink): rank r owns a contiguous row
range of the base and scans it with the same fused kernel as the single-GPU path; ids leave the
kernel already offset to global row numbers.  The only exchange on the path is an allgather of
the per-rank sorted top-k lists (nq*k*12 bytes per rank) followed by the merge kernel
(``vdb_merge_topk``), whose (distance, id) order makes the result independent of the number
of shards.  The reference has no multi-GPU path; this replaces nothing but scales
``faiss.IndexFlat.search`` (src/algorithms/exact_search.py:78) past one device.

The communication and compute steps are injected (``local_search`` / ``gather`` / ``merge``) so
the sharding logic itself can be exercised with the gloo backend on CPU in the tests; the product
defaults are the CUDA kernels and NCCL and there is no CPU fallback."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous row ranges: shard r owns rows [start(r), stop(r)) of an n-row base."""
    n: int
    parts: int

    def __post_init__(self) -> None:
        if self.parts < 1 or self.n < 0:
            raise ValueError(f"bad shard plan n={self.n} parts={self.parts}")

    @property
    def rows_per_shard(self) -> int:
        return (self.n + self.parts - 1) // self.parts

    def start(self, r: int) -> int:
        return min(self.n, r * self.rows_per_shard)

    def stop(self, r: int) -> int:
        return min(self.n, (r + 1) * self.rows_per_shard)

    def bounds(self) -> List[Tuple[int, int]]:
        return [(self.start(r), self.stop(r)) for r in range(self.parts)]

    def owner(self, row: int) -> int:
        if not 0 <= row < self.n:
            raise ValueError(f"row {row} outside [0, {self.n})")
        return row // self.rows_per_shard


def dist_info() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allgather_topk(dist_local: torch.Tensor, idx_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[nq, k] per rank -> [world, nq, k] on every rank, rank order == ascending id ranges."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nq, k = dist_local.shape
    d_all = torch.empty((world * nq, k), dtype=dist_local.dtype, device=dist_local.device)   # rank-major concatenation
    i_all = torch.empty((world * nq, k), dtype=idx_local.dtype, device=idx_local.device)
    dist.all_gather_into_tensor(d_all, dist_local.contiguous(), group=group)
    dist.all_gather_into_tensor(i_all, idx_local.contiguous(), group=group)
    return d_all.view(world, nq, k), i_all.view(world, nq, k)


class ShardedTopK:
    """local search -> allgather -> merge, with the three steps injectable.

    local_search(queries, k) -> (D [nq,k], I [nq,k] global ids), sorted best-first by (distance, id)
    gather(D, I)              -> (D_all [parts,nq,k], I_all [parts,nq,k]) in rank order
    merge(D_all, I_all)       -> (D [nq,k], I [nq,k])
    Every rank returns the merged result (the allgather leaves it everywhere)."""

    def __init__(self, local_search: Callable, gather: Callable = allgather_topk, merge: Optional[Callable] = None):
        self.local_search = local_search
        self.gather = gather
        self.merge = merge

    def search(self, queries, k: int):
        d_loc, i_loc = self.local_search(queries, k)
        d_all, i_all = self.gather(d_loc, i_loc)
        if d_all.shape[0] == 1:
            return d_all[0], i_all[0]
        return self.merge(d_all, i_all)


class DistributedFlatIndex:
    """This rank's shard of a flat index plus the exchange step.  Collective: every rank must call
    ``search`` with the same queries (they are replicated; nq*d*4 bytes is small next to the base)."""

    def __init__(self, local_vectors, metric: str = "l2", device=None, id_offset: int = 0, group=None):
        from . import engine
        self.engine = engine
        self.group = group
        self.rank, self.world = dist_info()
        self.shard = engine.FlatShard(local_vectors, metric, device, id_offset=id_offset)
        self.metric = metric

    @classmethod
    def from_global(cls, vectors, metric: str = "l2", device=None, group=None) -> "DistributedFlatIndex":
        """Slice this rank's rows out of the full base (host array or memmap)."""
        rank, world = dist_info()
        plan = ShardPlan(int(vectors.shape[0]), world)
        lo, hi = plan.start(rank), plan.stop(rank)
        if hi <= lo:
            raise RuntimeError(f"rank {rank} of {world} owns no rows of a {vectors.shape[0]}-row base")
        return cls(vectors[lo:hi], metric, device, id_offset=lo, group=group)

    def memory_bytes(self) -> int:
        return self.shard.memory_bytes()

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None,
               impl: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        eng = self.engine
        descending = self.metric != "l2" and not (flags & (eng._lib.OUT_NEGATE | eng._lib.OUT_ONE_MINUS))
        if pad_value is None:
            pad_value = -eng.FLT_MAX if descending else eng.FLT_MAX
        d_loc, i_loc = self.shard.search(q, k, flags, pad_value, impl)
        if self.world == 1:
            return d_loc, i_loc
        d_all, i_all = allgather_topk(d_loc, i_loc, self.group)
        return eng.merge_topk(d_all, i_all, descending=descending, pad_value=pad_value)


class ReplicatedFlatIndex:
    """Small-base layout: every rank holds the WHOLE base and searches its slice of the query batch;
    the per-rank result blocks are concatenated with one allgather.  No merge and no top-k exchange:
    the units that shard are the queries.  Chosen by ``shard="auto"`` when the operands fit a GPU
    with room to spare - a 1M x 128 base is 1 GB, and cutting it by rows over 8 GPUs leaves each scan
    125k rows, dominated by fixed costs (bound warm-up, finalize, exchange)."""

    def __init__(self, vectors, metric: str = "l2", device=None, group=None):
        from . import engine
        self.engine = engine
        self.group = group
        self.rank, self.world = dist_info()
        self.shard = engine.FlatShard(vectors, metric, device)
        self.metric = metric

    def memory_bytes(self) -> int:
        return self.shard.memory_bytes()

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None,
               impl: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        import torch.distributed as dist
        eng = self.engine
        descending = self.metric != "l2" and not (flags & (eng._lib.OUT_NEGATE | eng._lib.OUT_ONE_MINUS))
        if pad_value is None:
            pad_value = -eng.FLT_MAX if descending else eng.FLT_MAX
        nq = q.shape[0]
        if self.world == 1:
            return self.shard.search(q, k, flags, pad_value, impl)
        lo, hi, per = self._slice(nq)
        return self._search_slice(q[lo:hi], nq, per, k, flags, pad_value, impl)

    def _slice(self, nq: int) -> Tuple[int, int, int]:
        per = (nq + self.world - 1) // self.world
        return min(nq, self.rank * per), min(nq, (self.rank + 1) * per), per

    def _search_slice(self, q_loc: torch.Tensor, nq: int, per: int, k: int, flags: int, pad_value: float, impl: int):
        import torch.distributed as dist
        dev = self.shard.dev
        d_loc = torch.full((per, k), pad_value, dtype=torch.float32, device=dev)
        i_loc = torch.full((per, k), -1, dtype=torch.int64, device=dev)
        if q_loc.shape[0] > 0:
            self.shard.search(q_loc, k, flags, pad_value, impl, out=(d_loc[: q_loc.shape[0]], i_loc[: q_loc.shape[0]]))
        d_all = torch.empty((self.world * per, k), dtype=torch.float32, device=dev)
        i_all = torch.empty((self.world * per, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(d_all, d_loc, group=self.group)
        dist.all_gather_into_tensor(i_all, i_loc, group=self.group)
        return d_all[:nq], i_all[:nq]

    def search_host(self, queries, k: int, flags: int = 0, pad_value: Optional[float] = None, impl: int = 0):
        """Host query batch in, device results out: only this rank's slice of the batch crosses PCIe."""
        eng = self.engine
        if self.world == 1 or isinstance(queries, torch.Tensor):
            return self.search(eng.queries_to_device(queries, self.shard.dev, self.shard.d), k, flags, pad_value, impl)
        descending = self.metric != "l2" and not (flags & (eng._lib.OUT_NEGATE | eng._lib.OUT_ONE_MINUS))
        if pad_value is None:
            pad_value = -eng.FLT_MAX if descending else eng.FLT_MAX
        import numpy as np
        qh = np.asarray(queries)
        if qh.ndim == 1:
            qh = qh.reshape(1, -1)
        if qh.ndim != 2 or qh.shape[1] != self.shard.d:
            raise RuntimeError(f"query batch has shape {qh.shape}, expected [nq, {self.shard.d}]")
        nq = qh.shape[0]
        lo, hi, per = self._slice(nq)
        if hi > lo:
            q_loc = eng.queries_to_device(qh[lo:hi], self.shard.dev, self.shard.d)
        else:
            q_loc = torch.empty((0, self.shard.d), dtype=torch.float32, device=self.shard.dev)
        return self._search_slice(q_loc, nq, per, k, flags, pad_value, impl)


class DistributedIVFIndex:
    """IVF-Flat across the GPUs of one box: the centroids are replicated, rank r holds the
    inverted lists of ITS rows (every list is cut by row range), so the union of the ranks' list scans
    is the single-GPU scan and the (distance, id) merge returns the same result for any GPU count.
    Exchange = the same allgather of local top-k lists + ``vdb_merge_topk`` as the flat index.
    Centroids come from rank 0 (k-means accumulates with float atomics, so two ranks training on the
    same sample would not agree bit for bit) and are broadcast once."""

    def __init__(self, local_vectors, centroids, metric: str = "l2", device=None, id_offset: int = 0, nprobe: int = 1,
                 group=None):
        from . import engine
        self.engine = engine
        self.group = group
        self.rank, self.world = dist_info()
        self.shard = engine.IVFShard(local_vectors, centroids, metric, device, id_offset=id_offset)
        self.metric = metric
        self.nprobe = int(nprobe)

    @classmethod
    def from_global(cls, vectors, nlist: int, metric: str = "l2", device=None, nprobe: int = 1, group=None, niter: int = 10,
                    seed: int = 1234) -> "DistributedIVFIndex":
        import torch.distributed as dist
        from . import engine
        rank, world = dist_info()
        dev = engine._require_cuda(device)
        d = int(vectors.shape[1])
        if rank == 0:
            cent = torch.from_numpy(engine.kmeans_train(vectors, nlist, metric, dev, niter=niter, seed=seed)).to(dev)
        else:
            cent = torch.empty((nlist, d), dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(cent, src=0, group=group)
        plan = ShardPlan(int(vectors.shape[0]), world)
        lo, hi = plan.start(rank), plan.stop(rank)
        if hi <= lo:
            raise RuntimeError(f"rank {rank} of {world} owns no rows of a {vectors.shape[0]}-row base")
        return cls(vectors[lo:hi], cent, metric, dev, id_offset=lo, nprobe=nprobe, group=group)

    def memory_bytes(self) -> int:
        return self.shard.memory_bytes()

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        eng = self.engine
        descending = self.metric != "l2" and not (flags & (eng._lib.OUT_NEGATE | eng._lib.OUT_ONE_MINUS))
        if pad_value is None:
            pad_value = -eng.FLT_MAX if descending else eng.FLT_MAX
        d_loc, i_loc = self.shard.search(q, k, self.nprobe, flags, pad_value)
        if self.world == 1:
            return d_loc, i_loc
        d_all, i_all = allgather_topk(d_loc, i_loc, self.group)
        return eng.merge_topk(d_all, i_all, descending=descending, pad_value=pad_value)


def choose_sharding(n_rows: int, kpad: int, world: int, requested: str = "auto") -> str:
    """'rows' (north-star layout: row shards + top-k allgather + merge) or 'queries' (replicated base).
    auto: replicate while the operand set (2 * n * kpad * 4 bytes) stays under 8 GB per GPU."""
    if requested in ("rows", "queries"):
        return requested
    if requested != "auto":
        raise ValueError(f"shard must be 'auto', 'rows' or 'queries', got '{requested}'")
    return "queries" if world > 1 and 8.0 * n_rows * kpad <= 8e9 else "rows"


class MultiDeviceFlatIndex:
    """Single-process variant for the reference's one-process harness: one FlatShard per visible
    device, queries broadcast with peer copies, local top-k lists copied to the first device and
    merged there.  Same kernels, same merge, no NCCL communicator needed."""

    def __init__(self, vectors, metric: str = "l2", devices: Optional[Sequence[int]] = None):
        from . import engine
        self.engine = engine
        if not torch.cuda.is_available():
            raise RuntimeError("MultiDeviceFlatIndex needs CUDA devices; there is no CPU fallback")
        devs = list(devices) if devices is not None else list(range(torch.cuda.device_count()))
        if not devs:
            raise RuntimeError("no CUDA devices selected")
        self.devices = [torch.device("cuda", int(d)) for d in devs]
        self.metric = metric
        plan = ShardPlan(int(vectors.shape[0]), len(self.devices))
        self.shards = []
        for r, dev in enumerate(self.devices):
            lo, hi = plan.start(r), plan.stop(r)
            if hi > lo:
                self.shards.append(engine.FlatShard(vectors[lo:hi], metric, dev, id_offset=lo))
        self.home = self.shards[0].dev

    def memory_bytes(self) -> int:
        return sum(s.memory_bytes() for s in self.shards)

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None,
               impl: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        eng = self.engine
        descending = self.metric != "l2" and not (flags & (eng._lib.OUT_NEGATE | eng._lib.OUT_ONE_MINUS))
        if pad_value is None:
            pad_value = -eng.FLT_MAX if descending else eng.FLT_MAX
        if len(self.shards) == 1:
            return self.shards[0].search(q, k, flags, pad_value, impl)
        parts = len(self.shards)
        nq = q.shape[0]
        d_all = torch.empty((parts, nq, k), dtype=torch.float32, device=self.home)
        i_all = torch.empty((parts, nq, k), dtype=torch.int64, device=self.home)
        home_stream = torch.cuda.current_stream(self.home)
        ready = torch.cuda.Event()
        ready.record(home_stream)
        done = []
        for p, shard in enumerate(self.shards):     # launches are asynchronous: the devices scan concurrently
            with torch.cuda.device(shard.dev):
                s = torch.cuda.current_stream(shard.dev)
                s.wait_event(ready)
                qp = q.to(shard.dev, copy=True)
                d, i = shard.search(qp, k, flags, pad_value, impl)
                d_all[p].copy_(d, non_blocking=True)
                i_all[p].copy_(i, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
                done.append(ev)
        for ev in done:
            home_stream.wait_event(ev)
        with torch.cuda.device(self.home):
            return eng.merge_topk(d_all, i_all, descending=descending, pad_value=pad_value)
