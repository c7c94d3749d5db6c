"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink): rank r owns a contiguous row
range of the base and scans it with the same fused kernel as the single-GPU path; ids leave the
kernel already offset to global row numbers.  The only exchange on the path moves the per-rank
sorted top-k lists (nq*k*12 bytes per rank, distances and ids packed into ONE buffer = one
collective) and feeds the merge kernel (``vdb_merge_topk_strided``), whose (distance, id) order
makes the result independent of the number of shards.  Two exchange plans (``TopKExchange``):
``allgather`` - every rank gathers every list and merges all queries (the plan BASELINE.json's
north star spells out); ``alltoall`` - rank r receives only the lists of ITS nq/world queries, merges
those, and a second allgather of the merged slices leaves the result everywhere: 1/world of the
merge work and (world+1)/world^2 of the bytes, bit-identical output.  The reference has no multi-GPU path; this replaces nothing but scales
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


def _cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


class TopKExchange:
    """Packed exchange buffers for per-rank [nq, k] (distance f32, id i64) lists; one object per (nq, k).

    The rank's own search writes into ``d_loc`` / ``i_loc`` (views of one block, so the finalize kernel's
    output IS the send buffer).  Queries are cut into ``world`` slices of ``per`` (even) queries; rows past
    nq carry id -1 and drop out of every merge.  ``merge(d_parts, i_parts, out)`` is injected: d_parts /
    i_parts are [parts, n, k] views whose parts are strided (engine.merge_topk on the GPU, a NumPy merge in
    the gloo tests).  Device-agnostic on purpose: the gloo CPU tests drive exactly this code."""

    def __init__(self, nq: int, k: int, device, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nq, self.k = int(nq), int(k)
        self.per = (_cdiv(self.nq, self.world) + 1) // 2 * 2
        self.nq_pad = self.per * self.world
        self.lo, self.hi = min(self.nq, self.rank * self.per), min(self.nq, (self.rank + 1) * self.per)
        u8 = dict(dtype=torch.uint8, device=device)
        self.db, self.ib = self.nq_pad * k * 4, self.nq_pad * k * 8            # bytes of the D / I halves of a block
        self.cdb, self.cib = self.per * k * 4, self.per * k * 8               # ... of one query slice
        self.local = torch.empty(self.db + self.ib, **u8)
        self.d_loc = self.local[: self.db].view(torch.float32).view(self.nq_pad, k)
        self.i_loc = self.local[self.db:].view(torch.int64).view(self.nq_pad, k)
        self.i_loc[self.nq:].fill_(-1)
        self.d_loc[self.nq:].fill_(0.0)
        self.merged = torch.empty(self.cdb + self.cib, **u8)                  # this rank's merged slice [D | I]
        self.d_mine = self.merged[: self.cdb].view(torch.float32).view(self.per, k)
        self.i_mine = self.merged[self.cdb:].view(torch.int64).view(self.per, k)
        self._gathered = self._send = self._recv = self._final = None
        self._u8 = u8

    # ---- plan "allgather": every rank gets every list ------------------------------------------------
    def allgather_merge(self, merge: Callable) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._gathered is None:
            self._gathered = torch.empty(self.world * (self.db + self.ib), **self._u8)
        self.dist.all_gather_into_tensor(self._gathered, self.local, group=self.group)
        g = self._gathered.view(self.world, self.db + self.ib)
        d_parts = g[:, : self.db].view(torch.float32).view(self.world, self.nq_pad, self.k)[:, : self.nq]
        i_parts = g[:, self.db:].view(torch.int64).view(self.world, self.nq_pad, self.k)[:, : self.nq]
        return merge(d_parts, i_parts, None)

    # ---- plan "alltoall": rank r merges the queries of slice r ----------------------------------------
    def alltoall_merge_slice(self, merge: Callable) -> Tuple[torch.Tensor, torch.Tensor]:
        """Leaves this rank's merged slice in ``d_mine`` / ``i_mine`` ([per, k]; queries lo..hi are live)."""
        chunk = self.cdb + self.cib
        if self._send is None:
            self._send = torch.empty(self.world * chunk, **self._u8)
            self._recv = torch.empty(self.world * chunk, **self._u8)
        sv = self._send.view(self.world, chunk)
        sv[:, : self.cdb].view(torch.float32).copy_(self.d_loc.view(self.world, self.per * self.k))
        sv[:, self.cdb:].view(torch.int64).copy_(self.i_loc.view(self.world, self.per * self.k))
        self.dist.all_to_all_single(self._recv, self._send, group=self.group)
        rv = self._recv.view(self.world, chunk)
        d_parts = rv[:, : self.cdb].view(torch.float32).view(self.world, self.per, self.k)
        i_parts = rv[:, self.cdb:].view(torch.int64).view(self.world, self.per, self.k)
        return merge(d_parts, i_parts, (self.d_mine, self.i_mine))

    def allgather_slices(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Every rank's merged (or, replicated layout: searched) slice -> the full [nq, k] result everywhere."""
        chunk = self.cdb + self.cib
        if self._final is None:
            self._final = torch.empty(self.world * chunk, **self._u8)
        self.dist.all_gather_into_tensor(self._final, self.merged, group=self.group)
        fv = self._final.view(self.world, chunk)
        d = fv[:, : self.cdb].view(torch.float32).reshape(self.nq_pad, self.k)[: self.nq]
        i = fv[:, self.cdb:].view(torch.int64).reshape(self.nq_pad, self.k)[: self.nq]
        return d, i


class SharedHostResult:
    """One pinned host block shared by the ranks of a box: every rank copies ITS query slice of a result
    device -> host straight into it, so a search returns the full [nq, k] result on the host with nq*k*12/world
    bytes crossing each GPU's PCIe link and no second host copy.  POSIX shared memory, registered with
    ``cudaHostRegister`` in every rank; a ring of ``slots`` result blocks (the arrays a call returns stay
    valid until ``slots`` later calls) plus per-rank step counters the ranks wait on instead of a collective."""

    def __init__(self, nq_pad: int, k: int, group=None, slots: int = 4):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nq_pad, self.k, self.slots = nq_pad, k, slots
        self.slot_bytes = (nq_pad * k * 12 + 4095) // 4096 * 4096
        nbytes = 4096 + slots * self.slot_bytes
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            self.shm.buf[:4096] = bytes(4096)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:       # Python < 3.13 registers attachments with the resource tracker too; only rank 0 owns the segment
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:      # noqa: BLE001
                pass
        self.flags = np.ndarray((self.world,), dtype=np.int64, buffer=self.shm.buf, offset=0)
        self._np = np
        self.step = 0
        self._registered = False
        base = np.ndarray((nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        self._addr = base.ctypes.data
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self._addr, nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister of the shared result block failed ({rc})")
            self._registered = True
        self.views = []
        for s in range(slots):
            off = 4096 + s * self.slot_bytes
            d = np.ndarray((nq_pad, k), dtype=np.float32, buffer=self.shm.buf, offset=off)
            i = np.ndarray((nq_pad, k), dtype=np.int64, buffer=self.shm.buf, offset=off + nq_pad * k * 4)
            self.views.append((d, i, torch.from_numpy(d), torch.from_numpy(i)))
        dist.barrier(group=group)          # every rank attached before rank 0 may unlink on close

    def publish(self, d_slice: torch.Tensor, i_slice: torch.Tensor, lo: int, nq: int, timeout_s: float = 120.0):
        """Copy this rank's rows [lo, lo + len) into the current slot, wait until every rank has done the
        same, return (D [nq, k], I [nq, k]) NumPy views of the slot."""
        import time
        self.step += 1
        d_np, i_np, d_t, i_t = self.views[self.step % self.slots]
        n = d_slice.shape[0]
        if n > 0:
            d_t[lo:lo + n].copy_(d_slice, non_blocking=True)
            i_t[lo:lo + n].copy_(i_slice, non_blocking=True)
        if d_slice.is_cuda:
            torch.cuda.current_stream(d_slice.device).synchronize()
        self.flags[self.rank] = self.step
        t0 = time.perf_counter()
        while int(self.flags.min()) < self.step:
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError(f"rank {self.rank}: peers did not publish step {self.step} within {timeout_s:.0f} s "
                                   f"(flags {self.flags.tolist()})")
        return d_np[:nq], i_np[:nq]

    def close(self) -> None:
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self._addr)
            self._registered = False
        self.views, self.flags = [], None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:      # noqa: BLE001 - interpreter shutdown order
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass


def allgather_topk(dist_local: torch.Tensor, idx_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[nq, k] per rank -> [world, nq, k] on every rank, rank order == ascending id ranges (one collective:
    distances and ids travel in one packed block per rank)."""
    nq, k = dist_local.shape
    ex = TopKExchange(nq, k, dist_local.device, group)
    ex.d_loc[:nq].copy_(dist_local)
    ex.i_loc[:nq].copy_(idx_local)
    return ex.allgather_merge(lambda d, i, out: (d, i))


class ShardedTopK:
    """local search -> exchange -> merge, with the steps injectable (the gloo tests run this on CPU tensors).

    local_search(queries, k, out=(D [nq,k], I [nq,k])) writes sorted (distance, id) lists with global ids
    merge(d_parts [parts,n,k], i_parts, out)          -> (D [n,k], I [n,k]); parts may be strided views
    Every rank returns the full merged result."""

    def __init__(self, local_search: Callable, merge: Callable, exchange: str = "alltoall", device="cpu", group=None):
        if exchange not in ("allgather", "alltoall"):
            raise ValueError(f"exchange must be 'allgather' or 'alltoall', got '{exchange}'")
        self.local_search, self.merge, self.exchange = local_search, merge, exchange
        self.device, self.group = device, group
        self._ex: dict = {}
        self._host: dict = {}

    def buffers(self, nq: int, k: int) -> TopKExchange:
        ex = self._ex.get((nq, k))
        if ex is None:
            if len(self._ex) >= 2:
                self._ex.clear()
            ex = self._ex[(nq, k)] = TopKExchange(nq, k, self.device, self.group)
        return ex

    def host_block(self, ex: TopKExchange) -> SharedHostResult:
        hb = self._host.get((ex.nq_pad, ex.k))
        if hb is None:
            for old in self._host.values():
                old.close()
            self._host.clear()
            hb = self._host[(ex.nq_pad, ex.k)] = SharedHostResult(ex.nq_pad, ex.k, self.group)
        return hb

    def search(self, queries, k: int):
        nq = int(queries.shape[0])
        ex = self.buffers(nq, k)
        self.local_search(queries, k, (ex.d_loc[:nq], ex.i_loc[:nq]))
        if ex.world == 1:
            return ex.d_loc[:nq].clone(), ex.i_loc[:nq].clone()      # the buffers are reused by the next call
        if self.exchange == "allgather":
            return ex.allgather_merge(self.merge)
        ex.alltoall_merge_slice(self.merge)
        return ex.allgather_slices()

    def search_to_host(self, queries, k: int):
        """Result on the HOST (NumPy views of the shared pinned block): each rank merges and copies only its
        query slice, so the full-result allgather never happens on this path."""
        nq = int(queries.shape[0])
        ex = self.buffers(nq, k)
        self.local_search(queries, k, (ex.d_loc[:nq], ex.i_loc[:nq]))
        ex.alltoall_merge_slice(self.merge)
        n_live = ex.hi - ex.lo
        return self.host_block(ex).publish(ex.d_mine[:n_live], ex.i_mine[:n_live], ex.lo, nq)


def _conventions(engine, metric: str, flags: int, pad_value: Optional[float]) -> Tuple[bool, float]:
    descending = metric != "l2" and not (flags & (engine._lib.OUT_NEGATE | engine._lib.OUT_ONE_MINUS))
    if pad_value is None:
        pad_value = -engine.FLT_MAX if descending else engine.FLT_MAX
    return descending, pad_value


class DistributedFlatIndex:
    """This rank's shard of a flat index plus the exchange step.  Collective: every rank must call
    ``search`` with the same queries (they are replicated; nq*d*4 bytes is small next to the base).
    ``exchange``: 'alltoall' (default) or 'allgather', see the module docstring."""

    def __init__(self, local_vectors, metric: str = "l2", device=None, id_offset: int = 0, group=None,
                 exchange: str = "alltoall"):
        from . import engine
        self.engine = engine
        self.group = group
        self.rank, self.world = dist_info()
        self.shard = engine.FlatShard(local_vectors, metric, device, id_offset=id_offset)
        self.metric = metric
        self.exchange = exchange
        self._plans: dict = {}
        self._qstage: dict = {}

    @classmethod
    def from_global(cls, vectors, metric: str = "l2", device=None, group=None, exchange: str = "alltoall") -> "DistributedFlatIndex":
        """Slice this rank's rows out of the full base (host array or memmap)."""
        rank, world = dist_info()
        plan = ShardPlan(int(vectors.shape[0]), world)
        lo, hi = plan.start(rank), plan.stop(rank)
        if hi <= lo:
            raise RuntimeError(f"rank {rank} of {world} owns no rows of a {vectors.shape[0]}-row base")
        return cls(vectors[lo:hi], metric, device, id_offset=lo, group=group, exchange=exchange)

    def memory_bytes(self) -> int:
        return self.shard.memory_bytes()

    def _plan(self, flags: int, pad_value: Optional[float], impl: int) -> ShardedTopK:
        eng = self.engine
        descending, pad = _conventions(eng, self.metric, flags, pad_value)
        key = (flags, pad, impl, self.exchange)
        plan = self._plans.get(key)
        if plan is None:
            def local(q, k, out):
                return self.shard.search(q, k, flags, pad, impl, out=out)

            def merge(d_parts, i_parts, out):
                return eng.merge_topk(d_parts, i_parts, descending=descending, pad_value=pad, out=out)

            plan = self._plans[key] = ShardedTopK(local, merge, self.exchange, self.shard.dev, self.group)
        return plan

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None,
               impl: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.world == 1:
            _, pad = _conventions(self.engine, self.metric, flags, pad_value)
            return self.shard.search(q, k, flags, pad, impl)
        return self._plan(flags, pad_value, impl).search(q, k)

    def search_host(self, queries, k: int, flags: int = 0, pad_value: Optional[float] = None, impl: int = 0):
        """Host queries in, HOST result out (NumPy views of the ranks' shared pinned block, valid for the next
        three calls).  Every rank uploads only ITS slice of the batch and an NVLink allgather replicates the
        queries (nq*d*4 / world bytes over each PCIe link instead of the whole batch over all of them); then
        every rank scans its rows, receives the lists of its query slice, merges them and copies that slice of
        the result device -> host."""
        import numpy as np
        import torch.distributed as dist
        eng = self.engine
        if self.world == 1 or isinstance(queries, torch.Tensor):
            q = eng.queries_to_device(queries, self.shard.dev, self.shard.d)
            if self.world == 1:
                _, pad = _conventions(eng, self.metric, flags, pad_value)
                return eng.results_to_host(*self.shard.search(q, k, flags, pad, impl))
            return self._plan(flags, pad_value, impl).search_to_host(q, k)
        qh = np.asarray(queries)
        if qh.ndim == 1:
            qh = qh.reshape(1, -1)
        if qh.ndim != 2 or qh.shape[1] != self.shard.d:
            raise RuntimeError(f"query batch has shape {qh.shape}, expected [nq, {self.shard.d}]")
        nq, d = qh.shape
        per = _cdiv(nq, self.world)
        buf = self._qstage.get(nq)
        if buf is None:
            self._qstage.clear()
            buf = self._qstage[nq] = (torch.empty((per * self.world, d), dtype=torch.float32, device=self.shard.dev),
                                      torch.zeros((per, d), dtype=torch.float32, device=self.shard.dev))
        q_full, q_slice = buf
        lo, hi = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
        if hi > lo:
            part = qh[lo:hi]
            if part.dtype != np.float32 or not part.flags["C_CONTIGUOUS"]:
                part = np.ascontiguousarray(part, dtype=np.float32)
            q_slice[: hi - lo].copy_(torch.from_numpy(part) if part.flags["WRITEABLE"] else torch.from_numpy(part.copy()),
                                     non_blocking=True)
        dist.all_gather_into_tensor(q_full, q_slice, group=self.group)
        return self._plan(flags, pad_value, impl).search_to_host(q_full[:nq], k)


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
        self._ex: dict = {}
        self._host: dict = {}

    def memory_bytes(self) -> int:
        return self.shard.memory_bytes()

    def _buffers(self, nq: int, k: int) -> TopKExchange:
        ex = self._ex.get((nq, k))
        if ex is None:
            if len(self._ex) >= 2:
                self._ex.clear()
            ex = self._ex[(nq, k)] = TopKExchange(nq, k, self.shard.dev, self.group)
        return ex

    def _search_slice(self, q_loc: torch.Tensor, ex: TopKExchange, flags: int, pad_value: float, impl: int) -> None:
        n_live = ex.hi - ex.lo
        if n_live > 0:
            self.shard.search(q_loc, ex.k, flags, pad_value, impl, out=(ex.d_mine[:n_live], ex.i_mine[:n_live]))

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None,
               impl: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        _, pad = _conventions(self.engine, self.metric, flags, pad_value)
        if self.world == 1:
            return self.shard.search(q, k, flags, pad, impl)
        ex = self._buffers(int(q.shape[0]), k)
        self._search_slice(q[ex.lo:ex.hi], ex, flags, pad, impl)
        return ex.allgather_slices()              # one collective: [D | I] block per rank

    def search_host(self, queries, k: int, flags: int = 0, pad_value: Optional[float] = None, impl: int = 0):
        """Host query batch in, HOST result out: only this rank's slice of the batch goes host -> device and
        only its slice of the result comes back (into the ranks' shared pinned block); no collective at all."""
        eng = self.engine
        _, pad = _conventions(eng, self.metric, flags, pad_value)
        if self.world == 1 or isinstance(queries, torch.Tensor):
            q = eng.queries_to_device(queries, self.shard.dev, self.shard.d)
            return eng.results_to_host(*self.search(q, k, flags, pad, impl))
        import numpy as np
        qh = np.asarray(queries)
        if qh.ndim == 1:
            qh = qh.reshape(1, -1)
        if qh.ndim != 2 or qh.shape[1] != self.shard.d:
            raise RuntimeError(f"query batch has shape {qh.shape}, expected [nq, {self.shard.d}]")
        nq = qh.shape[0]
        ex = self._buffers(nq, k)
        n_live = ex.hi - ex.lo
        if n_live > 0:
            self._search_slice(eng.queries_to_device(qh[ex.lo:ex.hi], self.shard.dev, self.shard.d), ex, flags, pad, impl)
        hb = self._host.get((ex.nq_pad, k))
        if hb is None:
            for old in self._host.values():
                old.close()
            self._host.clear()
            hb = self._host[(ex.nq_pad, k)] = SharedHostResult(ex.nq_pad, k, self.group)
        return hb.publish(ex.d_mine[:n_live], ex.i_mine[:n_live], ex.lo, nq)


class ReplicatedIVFIndex(ReplicatedFlatIndex):
    """IVF-Flat in the query-slice layout: every rank holds ALL the inverted lists (a C3-sized index is 0.3 GB) and
    scans for its slice of the query batch; one allgather of the packed result blocks, or none on the host path.
    Unlike the row layout the coarse quantiser is not repeated on every rank, so small indexes scale with the GPU
    count.  The centroids come from rank 0 (see ``DistributedIVFIndex``); results do not depend on the order in
    which a rank filled its lists, so they equal the one-GPU index bit for bit."""

    def __init__(self, vectors, centroids, metric: str = "l2", device=None, nprobe: int = 1, group=None):
        from . import engine
        self.engine = engine
        self.group = group
        self.rank, self.world = dist_info()
        self.shard = engine.IVFShard(vectors, centroids, metric, device)
        self.metric = metric
        self.nprobe = int(nprobe)
        self._ex = {}
        self._host = {}

    @classmethod
    def from_global(cls, vectors, nlist: int, metric: str = "l2", device=None, nprobe: int = 1, group=None, niter: int = 10,
                    seed: int = 1234) -> "ReplicatedIVFIndex":
        import torch.distributed as dist
        from . import engine
        rank, world = dist_info()
        dev = engine._require_cuda(device)
        if rank == 0:
            cent = torch.from_numpy(engine.kmeans_train(vectors, nlist, metric, dev, niter=niter, seed=seed)).to(dev)
        else:
            cent = torch.empty((nlist, int(vectors.shape[1])), dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(cent, src=0, group=group)
        return cls(vectors, cent, metric, dev, nprobe=nprobe, group=group)

    def _search_slice(self, q_loc: torch.Tensor, ex: TopKExchange, flags: int, pad_value: float, impl: int) -> None:
        n_live = ex.hi - ex.lo
        if n_live > 0:
            d, i = self.shard.search(q_loc, ex.k, self.nprobe, flags, pad_value)
            ex.d_mine[:n_live].copy_(d)
            ex.i_mine[:n_live].copy_(i)

    def search(self, q: torch.Tensor, k: int, flags: int = 0, pad_value: Optional[float] = None,
               impl: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.world == 1:
            _, pad = _conventions(self.engine, self.metric, flags, pad_value)
            return self.shard.search(q, k, self.nprobe, flags, pad)
        return super().search(q, k, flags, pad_value, impl)


class DistributedIVFIndex:
    """IVF-Flat across the GPUs of one box: the centroids are replicated, rank r holds the
    inverted lists of ITS rows (every list is cut by row range), so the union of the ranks' list scans
    is the single-GPU scan and the (distance, id) merge returns the same result for any GPU count.
    Exchange = the same packed top-k exchange + ``vdb_merge_topk_strided`` as the flat index.
    Centroids come from rank 0 (k-means accumulates with float atomics, so two ranks training on the
    same sample would not agree bit for bit) and are broadcast once."""

    def __init__(self, local_vectors, centroids, metric: str = "l2", device=None, id_offset: int = 0, nprobe: int = 1,
                 group=None, exchange: str = "alltoall"):
        from . import engine
        self.engine = engine
        self.group = group
        self.rank, self.world = dist_info()
        self.shard = engine.IVFShard(local_vectors, centroids, metric, device, id_offset=id_offset)
        self.metric = metric
        self.nprobe = int(nprobe)
        self.exchange = exchange

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
        descending, pad = _conventions(eng, self.metric, flags, pad_value)
        if self.world == 1:
            return self.shard.search(q, k, self.nprobe, flags, pad)

        def local(qq, kk, out):
            d, i = self.shard.search(qq, kk, self.nprobe, flags, pad)
            out[0].copy_(d)
            out[1].copy_(i)

        def merge(d_parts, i_parts, out):
            return eng.merge_topk(d_parts, i_parts, descending=descending, pad_value=pad, out=out)

        plan = getattr(self, "_plan", None)
        if plan is None or plan[0] != (flags, pad):
            plan = self._plan = ((flags, pad), ShardedTopK(local, merge, self.exchange, self.shard.dev, self.group))
        plan[1].local_search = local
        return plan[1].search(q, k)


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
