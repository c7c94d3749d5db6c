#!/usr/bin/env python
"""Headline benchmark: exact L2 top-100 QPS on the SIFT1M shape (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one query batch: 10 000 queries against the
1 000 000 x 128 fp32 base, k = 100 (query operand split, seeding pre-pass over a strided sample,
tcgen05 main scan with the fused top-k bound, verification + redo launch, exact re-scoring + sort).
At N > 1 (strong scaling: the SIFT1M shape is fixed) the headline `value` is the north-star layout -
base rows sharded over the ranks, every rank scans its rows for all queries, ONE packed NCCL exchange of
the local top-k lists, merge kernel (``config.sharding`` = "rows/N: ...").  The same invocation also
measures the replicated-base layout (every rank holds the 1 GB base and searches nq / N queries, one
allgather of the result blocks) and reports it under ``"replicated"``.

Printed JSON (rank 0, one line):
  value      QPS with the queries already resident in HBM (device time, CUDA events, max over ranks)
  e2e        QPS through the reference-facing API ``ExactSearch.batch_search`` with HOST query /
             result buffers: pinned-host -> device copy of the queries and device -> host copy of
             (distances, ids) inside the timed region; at N > 1 every rank copies only its query slice
             of the result into one pinned host block shared by the ranks
  parity     rank 0, after the timed loops: 32 sampled queries of the LAST device result and of the last
             end-to-end result against ``oracle.faiss_flat_search`` (fp64) on the same base
  roofline   dominant kernel (flat_scan_tc_kernel): tensor-pipe TFLOP/s = 3 * 2*nq*N*d / t
             (3xTF32 issues three MMAs per product) against TF32 peak = measured bf16 / 2
  cpu_baseline  the FAISS-flat port (per-thread sgemm blocks + running k-th-best threshold, all host
             cores) and the UNMODIFIED reference NumPy ``LinearSearcher`` (from baseline/_ref, where it
             fits) timed on this box's host cores on a bounded sample of the same workload
``--impl reference`` times those CPU implementations alone (FAISS itself is not installable here; see
DESIGN.md) and prints the same line shape, the same ``config``, with ``"impl": "reference"``."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BASE, DIM, NQ, TOPK = 1_000_000, 128, 10_000, 100
WORKLOAD = "sift1m_shape_exact_l2: 1M x 128 fp32 base, 10k queries, k=100 (BASELINE.json configs[1])"
METRIC = "qps_exact_l2_top100_sift1m_shape"
L2_BYTES = 126 * 1024 * 1024


# ------------------------------------------------------------------------------------ helpers
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": float(p["bf16_tflops"]), "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (recipe's clocks line)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:      # noqa: BLE001 - no nvidia-smi: report nulls
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:      # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:      # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def _host_data(nq_sample: int):
    import numpy as np
    rng = np.random.default_rng(42)
    base = rng.standard_normal((N_BASE, DIM), dtype=np.float32)
    queries = np.random.default_rng(4242).standard_normal((nq_sample, DIM), dtype=np.float32)
    return base, queries


def _host_cores() -> int:
    """Host threads the CPU legs use: every core of the box, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its ranks; the port runs its own thread pool over single-threaded BLAS calls)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:      # noqa: BLE001
        return os.cpu_count() or 1


def _cpu_port_qps(base, queries, reps: int, cores: int):
    """FAISS-flat restatement on the host cores (oracle port); returns (qps, seconds per pass)."""
    from oracle import oracle
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        oracle.faiss_flat_search_threaded(base, queries, TOPK, "l2", threads=cores)
        best = min(best, time.perf_counter() - t)
    return queries.shape[0] / best, best


def _reference_numpy(base, queries):
    """The UNMODIFIED reference ``BruteForceIndexer`` + ``LinearSearcher`` (src/algorithms/modular.py:121-130,
    312-387) imported from the staged copy in baseline/_ref, on the same base.  Its inner-product path is
    ``Q @ V.T`` + argpartition (what every published cosine/IP "exact" row ran) and fits at the harness'
    default batch of 128 queries; its L2 path materialises an nq x N x d temporary (512 MB per query at this
    shape), so it is timed on 2 queries.  Returns None when baseline/_ref is not staged."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "src", "algorithms")):
        return None
    try:
        from vectordb_retrieval_b200 import plugin
        try:                                              # torchrun exports OMP_NUM_THREADS=1: give NumPy's BLAS the box back
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=_host_cores(), user_api="blas")
        except Exception:      # noqa: BLE001
            pass
        mods = plugin.import_reference(ref_root)          # import stubs for faiss / matplotlib; NO plugin.install()
        out = {"kind": "reference-numpy", "unit": "queries/s", "cores": _host_cores(),
               "impl": "src.algorithms.modular.LinearSearcher (unmodified, baseline/_ref), BLAS on all host cores"}
        for metric, nq, key in (("ip", 128, "ip"), ("l2", 2, "l2")):
            algo = mods["algorithms"].get_algorithm_instance(
                "Composite", DIM, name=f"ref_{metric}", metric=metric, indexer={"type": "BruteForceIndexer"},
                searcher={"type": "LinearSearcher"})
            algo.build_index(base)
            q = queries[:nq]
            best = float("inf")
            for _ in range(2):
                t = time.perf_counter()
                algo.batch_search(q, TOPK)
                best = min(best, time.perf_counter() - t)
            out[f"{key}_qps"] = nq / best
            out[f"{key}_sample"] = f"{nq} queries per call against the full 1M x 128 base, best of 2 ({best:.2f} s)"
        return out
    except Exception as exc:      # noqa: BLE001 - a baseline only; say why it is missing
        return {"kind": "reference-numpy", "unavailable": f"{type(exc).__name__}: {exc}"[:200]}


def _config(world: int, mode: str) -> dict:
    """The workload description BOTH arms print (identical keys and values for one invocation)."""
    shard_gb = 8.0 * N_BASE * DIM / 1e9
    if world == 1:
        sharding = "none"
    elif mode == "queries":
        sharding = (f"queries/{world}: base replicated ({shard_gb:.2f} GB per GPU), each rank searches nq/N queries, "
                    "one NCCL allgather of the packed result blocks")
    else:
        sharding = (f"rows/{world}: row shards ({shard_gb / world:.2f} GB per GPU), one packed NCCL exchange of the local "
                    "top-k lists, merge kernel")
    per_gpu = shard_gb * 1e9 / (1 if mode == "queries" else world)
    return {"workload": WORKLOAD, "n": N_BASE, "d": DIM, "nq": NQ, "k": TOPK, "metric": "l2", "sharding": sharding,
            "l2": ("flushed between steps (256 MB write, outside the per-step events)" if per_gpu < 2 * L2_BYTES else
                   f"operands ({per_gpu / 1e9:.2f} GB per GPU) exceed the 126 MB L2"),
            "timing": "per-step CUDA events on the launching stream, summed; max over ranks"}


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    cores = _host_cores()
    nq_sample = 1000
    base, queries = _host_data(nq_sample)
    for _ in range(max(min(args.warmup, 2), 1)):
        _cpu_port_qps(base, queries[:128], 1, cores)
    times = []
    budget = time.perf_counter() + 150.0              # the whole run stays within a few minutes
    for _ in range(args.steps):
        t = time.perf_counter()
        _cpu_port_qps(base, queries, 1, cores)
        times.append(time.perf_counter() - t)
        if time.perf_counter() > budget:
            break
    sec = sum(times) / len(times)
    qps = nq_sample / sec
    sample = (f"{nq_sample} of the 10k queries per step against the full 1M x 128 base, {len(times)} timed steps "
              "(QPS is per query, so it carries over)")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": _config(world, "rows" if args.shard != "queries" else "queries"),
        "reference_impl": "oracle port of faiss.IndexFlat.search (per-thread fp32 sgemm blocks + running k-th-best "
                          "threshold, OpenBLAS); faiss-cpu itself is not installed / installable here",
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "reference_numpy": _reference_numpy(base, queries),
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ our arm
def _device_rows(lo: int, hi: int, dev):
    """Rows [lo, hi) of the synthetic base, generated on the device in 65 536-row blocks seeded by
    block number - the same data for every GPU count."""
    import torch
    blk = 65536
    out = torch.empty((hi - lo, DIM), dtype=torch.float32, device=dev)
    b0 = lo // blk
    for b in range(b0, (hi + blk - 1) // blk):
        g = torch.Generator(device=dev).manual_seed(42_000 + b)
        rows = torch.randn((blk, DIM), generator=g, device=dev, dtype=torch.float32)
        s, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        out[s - lo:e - lo] = rows[s - b * blk:e - b * blk]
    return out


def _measure(label, index, algo, q_dev, q_host_np, args, lib, dev, world, rank, barrier, sampler=None):
    """Warm-up, K device-timed steps (per-step CUDA events, L2 flushed between steps when the shard is small),
    then K end-to-end steps through ``ExactSearch.batch_search`` with host buffers.  Returns a dict of timings
    plus the last device and host results (for the parity check)."""
    import ctypes
    import torch
    import torch.distributed as dist
    shard_bytes = index.memory_bytes()
    flush = shard_bytes < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev) if flush else None
    # (results are held like in the timed loops: with the previous step's arrays still alive the second
    # call needs a second set of pinned host blocks, and a fresh cudaHostAlloc costs 10-30 ms once)
    if sampler is not None:          # started before the warm-up: nvidia-smi needs ~0.3 s before its first sample
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        d_dev, i_dev = index.search(q_dev, TOPK)
        d_host, i_host = algo.batch_search(q_host_np, TOPK)
    barrier()
    lib.vdb_flat_timing_enable(1)
    launches0 = lib.vdb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for e0, e1 in ev:
        if flush:
            flush_buf.fill_(1.0)
        e0.record()
        d_dev, i_dev = index.search(q_dev, TOPK)
        e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = lib.vdb_launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    buf = (ctypes.c_float * 512)()
    n_rec = ctypes.c_int(0)
    lib.vdb_flat_timing_read(buf, 512, ctypes.byref(n_rec))
    scan_ms = [buf[i] for i in range(n_rec.value)]
    lib.vdb_flat_timing_enable(0)
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    scan_mean = torch.tensor([sum(scan_ms) / max(len(scan_ms), 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(scan_mean, op=dist.ReduceOp.MAX)
    # the clock sampler covers the device-timed region; it stops here so that nvidia-smi's driver
    # queries cannot stall the synchronous host calls of the end-to-end leg
    clocks = sampler.stop() if sampler is not None else None
    d_dev, i_dev = d_dev.clone(), i_dev.clone()          # the exchange buffers are reused by the end-to-end leg
    barrier()
    t0 = time.perf_counter()
    e2e_steps = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        d_host, i_host = algo.batch_search(q_host_np, TOPK)
        e2e_steps.append(time.perf_counter() - ts)        # the call returns host arrays: every step is complete when it returns
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    del flush_buf
    return {"label": label, "ms_per_step": float(total_ms.item()) / args.steps, "scan_ms": float(scan_mean.item()),
            "wall_ms_per_step": t_wall * 1e3 / args.steps, "launches": int(launches), "flush": flush,
            "e2e_ms": float(e2e_s.item()) * 1e3 / args.steps, "clocks": clocks, "shard_bytes": shard_bytes,
            "e2e_step_ms": [min(e2e_steps) * 1e3, statistics.median(e2e_steps) * 1e3, max(e2e_steps) * 1e3],
            "device_result": (d_dev, i_dev), "host_result": (d_host, i_host)}


def _parity(base_host, q_host_np, results, sample_idx):
    """32 sampled queries of each result against the fp64 oracle on the same base (rank 0, outside the timed
    loops; tolerance = the north star's: ids exact except inside distance ties within 1e-5 relative)."""
    import numpy as np
    from oracle import oracle
    ref_d, ref_i = oracle.faiss_flat_search(base_host, q_host_np[sample_idx], TOPK, "l2")
    out = {"queries": int(len(sample_idx)), "k": TOPK, "rtol": 1e-5, "oracle": "oracle.faiss_flat_search (fp64)", "ok": True,
           "id_mismatch": 0, "max_rel_err": 0.0, "checked": []}
    for name, (d, i) in results.items():
        d = d[sample_idx] if isinstance(d, np.ndarray) else d.cpu().numpy()[sample_idx]
        i = i[sample_idx] if isinstance(i, np.ndarray) else i.cpu().numpy()[sample_idx]
        res = oracle.compare_topk(ref_d, ref_i, d, i, rtol=1e-5)
        out["checked"].append({"result": name, "ok": bool(res["ok"]), "id_exact": int(res["id_exact"]),
                               "tie_swaps": int(res["tie_swaps"]), "id_mismatch": int(res["id_mismatch"]),
                               "max_rel_err": float(res["max_rel_err"]),
                               "recall_at_k": float(oracle.recall_at_k(ref_i, i, TOPK))})
        out["ok"] = out["ok"] and bool(res["ok"])
        out["id_mismatch"] += int(res["id_mismatch"])
        out["max_rel_err"] = max(out["max_rel_err"], float(res["max_rel_err"]))
    return out


def run_ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a CUDA device (sm_100a); there is no CPU fallback"}))
        return 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: anything libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from vectordb_retrieval_b200 import _lib, sharded
    from vectordb_retrieval_b200.algorithms import ExactSearch
    from vectordb_retrieval_b200.indexes import GpuIndexFlat
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    modes = ["none"] if world == 1 else (["rows", "queries"] if args.shard == "both" else [args.shard])
    gq = torch.Generator(device=dev).manual_seed(4242)
    q_dev = torch.randn((NQ, DIM), generator=gq, device=dev, dtype=torch.float32)
    q_host = torch.empty((NQ, DIM), dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    q_host_np = q_host.numpy()

    measured = {}
    for mode in modes:
        if mode == "queries":        # small base: replicate it, every rank searches nq / world queries
            lo, hi = 0, N_BASE
            index = sharded.ReplicatedFlatIndex(_device_rows(lo, hi, dev), "l2", dev)
        else:                        # north-star layout: row shards, packed top-k exchange, merge kernel
            plan = sharded.ShardPlan(N_BASE, world)
            lo, hi = plan.start(rank), plan.stop(rank)
            index = sharded.DistributedFlatIndex(_device_rows(lo, hi, dev), "l2", dev, id_offset=lo, exchange=args.exchange)
        torch.cuda.empty_cache()
        # the reference-facing object for the e2e leg shares this rank's shard (no second copy of the base)
        algo = ExactSearch("exact", DIM, metric="l2")
        algo.index = GpuIndexFlat(DIM, "l2", device=dev)
        algo.index._impl, algo.index.ntotal, algo.index_built = index, N_BASE, True
        sampler = ClockSampler(local_rank) if (rank == 0 and mode == modes[0]) else None
        m = _measure(mode, index, algo, q_dev, q_host_np, args, lib, dev, world, rank, barrier, sampler)
        m["rows_per_rank"] = hi - lo
        measured[mode] = m
        d_host, i_host = m["host_result"]
        assert i_host.shape == (NQ, TOPK) and int(i_host.min()) >= 0 and int(i_host.max()) < N_BASE
        assert bool(np.all(np.diff(d_host, axis=1) >= 0)), "distances are not sorted"
        m["host_result"] = (np.array(d_host), np.array(i_host))      # out of the shared block before it is closed
        del index, algo
        barrier()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    head = measured[modes[0]]
    # ---- parity of what was just timed (oracle = checker; outside every timed region)
    parity = None
    if not args.no_parity:
        base_host = _device_rows(0, N_BASE, dev).cpu().numpy()
        sample_idx = np.random.default_rng(7).choice(NQ, 32, replace=False)
        sample_idx.sort()
        results = {}
        for mode, m in measured.items():
            results[f"{mode}: device result"] = m["device_result"]
            results[f"{mode}: end-to-end host result"] = m["host_result"]
        parity = _parity(base_host, q_host_np, results, sample_idx)
        del base_host
    torch.cuda.empty_cache()

    peaks = _peaks()
    # live yardstick on this box: cuBLAS TF32 GEMM 8192^3 (best of 10 after 3 warm-ups, CUDA events)
    cublas_tf32 = None
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((8192, 8192), device=dev)
        b = torch.randn((8192, 8192), device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(10):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(); torch.matmul(a, b); g1.record()
            torch.cuda.synchronize(dev)
            best = min(best, g0.elapsed_time(g1))
        cublas_tf32 = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    except Exception:      # noqa: BLE001 - yardstick only
        cublas_tf32 = None

    def roofline(m, mode):
        nq_launch = (NQ + world - 1) // world if mode == "queries" else NQ
        flops = 2.0 * nq_launch * m["rows_per_rank"] * DIM           # per launch on one rank (2 nq N d)
        pipe = 3.0 * flops / (m["scan_ms"] * 1e-3) / 1e12
        burst, sustained = peaks["bf16_burst"] / 2.0, peaks["bf16_sustained"] / 2.0
        traffic, traffic_note = None, "not captured for this launch shape"
        prof = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
        if world == 1 and os.path.exists(prof):                        # the capture is of THIS launch shape only
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
                traffic_note = "ncu --set full capture of the C2 main-scan launch (profiles/scan_kernel_traffic.json)"
            except Exception:      # noqa: BLE001
                traffic = None
        return {"bound": "tensor", "kernel": "flat_scan_tc_kernel", "achieved": pipe, "peak": burst, "unit": "TFLOP/s",
                "frac": pipe / burst, "frac_burst": pipe / burst, "frac_sustained": pipe / sustained,
                "peak_sustained": sustained, "traffic": traffic, "traffic_note": traffic_note,
                "kernel_ms": m["scan_ms"], "kernel_share_of_step": m["scan_ms"] / m["ms_per_step"],
                "algorithmic_tflops": flops / (m["scan_ms"] * 1e-3) / 1e12,
                "frac_of_nominal_tf32": pipe / 1125.0, "cublas_tf32_live_tflops": cublas_tf32,
                "note": f"achieved = 3 * 2*nq*N_rank*d / t of the main scan launch (3xTF32 issues 3 MMAs per product; the "
                        f"seeding pre-pass and the empty redo launch are separate launches inside the step); peak = "
                        f"{peaks['source']} BURST bf16 cuBLAS / 2 (TF32 runs at half the bf16 rate; the timed region is "
                        "a fraction of a second), frac_sustained uses the sustained figure; both are cuBLAS-derived proxies, "
                        "so the fraction of the nominal dense TF32 rate (1125 TFLOP/s) and a live cuBLAS TF32 8192^3 GEMM "
                        "are given too; algorithmic_tflops is the fp32-equivalent 2*nq*N_rank*d / t"}

    def e2e_block(m, mode):
        h2d = NQ * DIM * 4 if world == 1 else ((NQ + world - 1) // world) * DIM * 4     # every rank uploads its query slice
        d2h = NQ * TOPK * 12 if world == 1 else ((NQ + world - 1) // world) * TOPK * 12
        return {"value": NQ / (m["e2e_ms"] * 1e-3), "unit": "queries/s", "ms_per_step": m["e2e_ms"],
                "step_ms_min_median_max_rank0": m["e2e_step_ms"],
                "h2d_bytes_per_step": h2d * (world if world > 1 else 1), "d2h_bytes_per_step": d2h * (world if world > 1 else 1),
                "h2d_bytes_per_rank": h2d, "d2h_bytes_per_rank": d2h,
                "api": "ExactSearch.batch_search(numpy pinned queries) -> (numpy distances, numpy ids)" +
                       ("" if world == 1 else "; every rank copies its query slice of the result into one pinned host "
                                              "block shared by the ranks (sharded.SharedHostResult)")}

    cpu = None
    ref_numpy = None
    if world == 1 and not args.no_cpu_baseline:
        nq_sample = 1000
        cores = _host_cores()
        base_h, q_h = _host_data(nq_sample)
        _cpu_port_qps(base_h, q_h[:128], 1, cores)
        qps_cpu, sec = _cpu_port_qps(base_h, q_h, 2, cores)
        cpu = {"value": qps_cpu, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"first {nq_sample} queries against the full 1M x 128 base, best of 2 passes ({sec:.1f} s each); "
                         "oracle FAISS-flat restatement (per-thread OpenBLAS sgemm blocks + running k-th-best threshold)"}
        ref_numpy = _reference_numpy(base_h, q_h)

    line = {
        "metric": METRIC, "value": NQ / (head["ms_per_step"] * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32x3 (fp32-accurate split) + f64 re-score", "data": "synthetic",
        "config": _config(world, modes[0]),
        "exchange": None if world == 1 or modes[0] != "rows" else args.exchange,
        "wall_ms_per_step": head["wall_ms_per_step"],
        "e2e": e2e_block(head, modes[0]),
        "gpu_launches": head["launches"],
        "parity": parity,
        "roofline": roofline(head, modes[0]),
        "cpu_baseline": cpu,
        "reference_numpy": ref_numpy,
        "clocks": head["clocks"],
    }
    if "queries" in measured and modes[0] != "queries":
        m = measured["queries"]
        line["replicated"] = {"value": NQ / (m["ms_per_step"] * 1e-3), "unit": "queries/s", "ms_per_step": m["ms_per_step"],
                              "sharding": _config(world, "queries")["sharding"], "e2e": e2e_block(m, "queries"),
                              "gpu_launches": m["launches"], "kernel_ms": m["scan_ms"],
                              "kernel_share_of_step": m["scan_ms"] / m["ms_per_step"]}
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--shard", choices=["both", "rows", "queries"], default="both",
                    help="N > 1: 'rows' = north-star row sharding (the headline); 'queries' = replicated base; "
                         "both = measure the two, report rows as `value` and the other under `replicated`")
    ap.add_argument("--exchange", choices=["alltoall", "allgather"], default="alltoall",
                    help="rows layout: plan of the packed top-k exchange (sharded.TopKExchange)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the last results (profiling runs)")
    args = ap.parse_args()
    if args.steps > 500:
        args.steps = 500
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # convenience: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl,
               "--shard", args.shard, "--exchange", args.exchange]
        return subprocess.call(cmd + (["--no-cpu-baseline"] if args.no_cpu_baseline else []) +
                               (["--no-parity"] if args.no_parity else []))
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
